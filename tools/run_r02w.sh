cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m pytest tests/test_gpu_msm.py -x -q -m gpu > gpurun_out/r02w_pytest.log 2>&1; tail -2 gpurun_out/r02w_pytest.log
python tools/gpu_l0_locality.py 21,24
python tools/gpu_sweep_c.py 24 16 d
python tools/gpu_sweep_c.py 21 16 d
python tools/gpu_sweep_g2.py 20 16 d
