cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(time python -m pytest tests/test_gpu_msm.py -m gpu -x -q) > gpurun_out/r02k_pytest.log 2>&1
tail -n 4 gpurun_out/r02k_pytest.log | head -2
python tools/gpu_l0_locality.py 21,24 > gpurun_out/r02k_l0.txt 2>&1
python tools/gpu_sweep_c.py 24 16 d > gpurun_out/r02k_sweep.txt 2>&1
python tools/gpu_sweep_c.py 21 16 d >> gpurun_out/r02k_sweep.txt 2>&1
python tools/gpu_sweep_c.py 20 16 d >> gpurun_out/r02k_sweep.txt 2>&1
cat gpurun_out/r02k_l0.txt gpurun_out/r02k_sweep.txt
