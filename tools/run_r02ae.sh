cd $GRAFT_REPO_ROOT
python tools/gpu_sweep_c.py 24 16 d,8,9 | grep -v DONE
python tools/gpu_sweep_c.py 22 16 d,6,7 | grep -v DONE
python tools/gpu_sweep_c.py 21 16 d,5,6,7 | grep -v DONE
python tools/gpu_sweep_c.py 20 16 d,4,5,6 | grep -v DONE
python tools/gpu_sweep_c.py 18 16,13 d,1,2,3,4 | grep -v DONE
python tools/gpu_sweep_g2.py 20 16 d,4,5 | grep -v DONE
