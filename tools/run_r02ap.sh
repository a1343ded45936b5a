cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_msm.py -x -q -m gpu 2>&1 | tail -2
python tools/gpu_host_msm.py g1 24 a,b host 2>&1 | head -2 | tee gpurun_out/r02ap_host.txt
python tools/gpu_host_msm.py g1 22 a host 2>&1 | head -1 | tee -a gpurun_out/r02ap_host.txt
