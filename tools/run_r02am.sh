cd $GRAFT_REPO_ROOT
python tools/gpu_sweep_c.py 21 16 4,5,6,7 2>&1 | grep -v DONE | tee gpurun_out/r02am_levels.txt
python tools/gpu_sweep_c.py 20 16 3,4,5,6 2>&1 | grep -v DONE | tee -a gpurun_out/r02am_levels.txt
python tools/gpu_sweep_c.py 22 16 5,6,7 2>&1 | grep -v DONE | tee -a gpurun_out/r02am_levels.txt
python tools/gpu_sweep_c.py 24 16 7,8,9 2>&1 | grep -v DONE | tee -a gpurun_out/r02am_levels.txt
python tools/gpu_sweep_c.py 18 16,13 d,1,2,3 2>&1 | grep -v DONE | tee -a gpurun_out/r02am_levels.txt
