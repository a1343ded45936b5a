cd $GRAFT_REPO_ROOT
for v in 3 1 4; do echo "ACC_MINB=$v"; B381_ACC_MINB=$v python tools/gpu_sweep_g2.py 20 16 d | head -1; done 2>&1 | tee gpurun_out/r02az_g2_acc.txt
