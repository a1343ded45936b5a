cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python tools/gpu_sweep_c.py 24 16 d | head -1
python tools/gpu_sweep_c.py 21 16 d | head -1
python tools/gpu_sweep_c.py 16 13 d | head -1
python tools/gpu_sweep_c.py 12 13 d | head -1
python tools/gpu_sweep_g2.py 20 16 d | head -1
