cd $GRAFT_REPO_ROOT
for n in 21 22 23; do
  B381_MSM_CHUNK_LOG=31 python tools/gpu_host_msm.py g1 $n plain host | head -1
  python tools/gpu_host_msm.py g1 $n chunked+streamed host | head -1
done 2>&1 | tee gpurun_out/r02ak_host.txt
B381_MSM_CHUNK_LOG=31 python tools/gpu_host_msm.py g2 20 plain host | head -1 | tee -a gpurun_out/r02ak_host.txt
python tools/gpu_host_msm.py g2 20 chunked+streamed host | head -1 | tee -a gpurun_out/r02ak_host.txt
B381_MSM_CHUNK_LOG=31 python tools/gpu_host_msm.py g2 21 plain host | head -1 | tee -a gpurun_out/r02ak_host.txt
python tools/gpu_host_msm.py g2 21 chunked+streamed host | head -1 | tee -a gpurun_out/r02ak_host.txt
python -m pytest tests/test_gpu_msm.py tests/test_gpu_dist.py -x -q -m gpu 2>&1 | tail -2 | tee gpurun_out/r02ak_pytest.log
