set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(time python -m pytest tests/test_gpu_ntt.py tests/test_gpu_dist.py tests/test_gpu_vs_reference.py tests/test_gpu_vecops.py -x -q) > gpurun_out/r02b_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02b_pytest.log
B381_NTT_GENERIC=1 python -m pytest tests/test_gpu_ntt.py -x -q -k "vs_oracle or golden or large" > gpurun_out/r02b_pytest_generic.log 2>&1
echo "generic rc=$?" >> gpurun_out/r02b_pytest_generic.log
for s in 0 1 2 3; do
  B381_NTT_SHAPE=$s python tools/gpu_ntt_bench.py 24,22,20,16:256,13:8192 10 >> gpurun_out/r02b_ntt_shapes.txt 2>&1
done
B381_NTT_GENERIC=1 python tools/gpu_ntt_bench.py 24,20 10 >> gpurun_out/r02b_ntt_shapes.txt 2>&1
tail -3 gpurun_out/r02b_pytest.log gpurun_out/r02b_pytest_generic.log
cat gpurun_out/r02b_ntt_shapes.txt
