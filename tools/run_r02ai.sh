# (needs tools/experiments/window_split/window_split.patch applied: B381_MSM_SPLIT is not in the product)
cd $GRAFT_REPO_ROOT
python tools/experiments/window_split/gpu_sweep_split.py g1 21 0,1,2,3,4,5,6 2>&1 | tee gpurun_out/r02ai_split.txt
python tools/experiments/window_split/gpu_sweep_split.py g1 20 0,2,4,6,8 2>&1 | tee -a gpurun_out/r02ai_split.txt
python tools/experiments/window_split/gpu_sweep_split.py g1 22 0,1,2,3 2>&1 | tee -a gpurun_out/r02ai_split.txt
python tools/experiments/window_split/gpu_sweep_split.py g1 24 0,1,2 2>&1 | tee -a gpurun_out/r02ai_split.txt
python tools/experiments/window_split/gpu_sweep_split.py g1 24 0,1 host 2>&1 | tee -a gpurun_out/r02ai_split.txt
python tools/experiments/window_split/gpu_sweep_split.py g2 20 0,1,2,3,4,6 2>&1 | tee -a gpurun_out/r02ai_split.txt
