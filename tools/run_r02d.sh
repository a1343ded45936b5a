set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m pytest tests/test_gpu_ntt.py tests/test_gpu_dist.py -x -q > gpurun_out/r02d_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02d_pytest.log
for s in 0 1; do
  B381_NTT_SHAPE=$s python tools/gpu_ntt_bench.py 24,22,20,16:256 10 >> gpurun_out/r02d_ntt.txt 2>&1
done
tail -n 3 gpurun_out/r02d_pytest.log
cat gpurun_out/r02d_ntt.txt
