cd $GRAFT_REPO_ROOT
python tools/gpu_host_msm.py g1 24 piece21 host | head -1 | tee gpurun_out/r02al_piece.txt
B381_MSM_PIECE_LOG=20 python tools/gpu_host_msm.py g1 24 piece20 host | head -1 | tee -a gpurun_out/r02al_piece.txt
B381_MSM_PIECE_LOG=22 python tools/gpu_host_msm.py g1 24 piece22 host | head -1 | tee -a gpurun_out/r02al_piece.txt
B381_MSM_PIECE_LOG=20 python tools/gpu_host_msm.py g1 23 piece20 host | head -1 | tee -a gpurun_out/r02al_piece.txt
python tools/gpu_host_msm.py g1 23 piece21 host | head -1 | tee -a gpurun_out/r02al_piece.txt
