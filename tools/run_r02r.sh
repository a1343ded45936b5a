cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m pytest tests/test_gpu_ntt.py -x -q -m gpu > gpurun_out/r02r_pytest.log 2>&1; tail -2 gpurun_out/r02r_pytest.log
python tools/gpu_ntt_bench.py 24,22,20 10 > gpurun_out/r02r_ntt.txt 2>&1
cat gpurun_out/r02r_ntt.txt
python tools/gpu_sweep_g2.py 20 16,15,14,13 d,3,4,5 > gpurun_out/r02r_g2.txt 2>&1
cat gpurun_out/r02r_g2.txt
