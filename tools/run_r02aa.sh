cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_msm.py -x -q -m gpu 2>&1 | tail -2
python tools/gpu_sweep_c.py 21 16 d | head -1
python tools/gpu_sweep_c.py 24 16 d | head -1
python tools/gpu_sweep_c.py 16 13 d | head -1
python tools/gpu_sweep_c.py 12 13 d | head -1
