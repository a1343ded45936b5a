cd $GRAFT_REPO_ROOT
python tools/gpu_sweep_g2.py 20 16 d | head -1 | tee gpurun_out/r02ba_g2.txt
python tools/gpu_sweep_g2.py 16 13 d | head -1 | tee -a gpurun_out/r02ba_g2.txt
python -m pytest tests/test_gpu_msm.py -x -q -m gpu -k "g2 or G2 or chunk or streamed or batch" 2>&1 | tail -2
