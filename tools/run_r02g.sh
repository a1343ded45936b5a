cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(time python -m pytest tests/test_gpu_msm.py tests/test_gpu_dist.py tests/test_gpu_vs_reference.py -x -q) > gpurun_out/r02g_pytest.log 2>&1
tail -n 5 gpurun_out/r02g_pytest.log
python tools/gpu_sweep_c.py 24 16 d > gpurun_out/r02g_sweep.txt 2>&1
python tools/gpu_sweep_c.py 21 16 d >> gpurun_out/r02g_sweep.txt 2>&1
python tools/gpu_sweep_c.py 20 16 d >> gpurun_out/r02g_sweep.txt 2>&1
python tools/gpu_sweep_c.py 16 13 d >> gpurun_out/r02g_sweep.txt 2>&1
python tools/gpu_sweep_c.py 12 13 d >> gpurun_out/r02g_sweep.txt 2>&1
cat gpurun_out/r02g_sweep.txt
