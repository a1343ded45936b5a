cd $GRAFT_REPO_ROOT
for L in 16 32 64 128; do echo "L=$L"; B381_MSM_L=$L python tools/gpu_sweep_g2.py 20 16 d | head -1; done 2>&1 | tee gpurun_out/r02aj_L.txt
for L in 16 32 64; do echo "L=$L"; B381_MSM_L=$L python tools/gpu_sweep_c.py 21 16 d | head -1; done 2>&1 | tee -a gpurun_out/r02aj_L.txt
