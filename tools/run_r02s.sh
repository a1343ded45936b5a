cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m pytest tests/test_gpu_msm.py tests/test_icicle_dispatch.py -x -q -m gpu > gpurun_out/r02s_pytest.log 2>&1; tail -2 gpurun_out/r02s_pytest.log
python tools/gpu_sweep_g2.py 20 16 d > gpurun_out/r02s_g2.txt 2>&1
for L in 16 32 64 128; do B381_MSM_L=$L python tools/gpu_sweep_g2.py 20 16 d >> gpurun_out/r02s_g2.txt 2>&1; done
python tools/gpu_sweep_g2.py 16 13,16 d >> gpurun_out/r02s_g2.txt 2>&1
python tools/gpu_sweep_g2.py 12 13 d >> gpurun_out/r02s_g2.txt 2>&1
cat gpurun_out/r02s_g2.txt
python tools/gpu_sweep_c.py 21 16 d
