cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
rm -f gpurun_out/r02e_ntt.txt
for s in 0 1 2 3 4; do
  B381_NTT_SHAPE=$s python tools/gpu_ntt_bench.py 24,22 10 >> gpurun_out/r02e_ntt.txt 2>&1
done
B381_NTT_SHAPE=2 python -m pytest tests/test_gpu_ntt.py -x -q -k "large or golden" > gpurun_out/r02e_pytest.log 2>&1
tail -n 2 gpurun_out/r02e_pytest.log
grep -E "shape|dir=0" gpurun_out/r02e_ntt.txt
