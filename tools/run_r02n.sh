cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(time python -m pytest tests/test_gpu_msm.py tests/test_gpu_dist.py tests/test_icicle_dispatch.py -m gpu -x -q) > gpurun_out/r02n_pytest.log 2>&1
tail -n 12 gpurun_out/r02n_pytest.log | head -9
for cl in 31 19 18 20; do
  echo "== B381_MSM_CHUNK_LOG=$cl" >> gpurun_out/r02n_sweep.txt
  B381_MSM_CHUNK_LOG=$cl python tools/gpu_sweep_c.py 24 16 d >> gpurun_out/r02n_sweep.txt 2>&1
  B381_MSM_CHUNK_LOG=$cl python tools/gpu_l0_locality.py 24 >> gpurun_out/r02n_sweep.txt 2>&1
done
for cl in 31 19; do
  echo "== B381_MSM_CHUNK_LOG=$cl" >> gpurun_out/r02n_sweep.txt
  B381_MSM_CHUNK_LOG=$cl python tools/gpu_sweep_c.py 22 16 d >> gpurun_out/r02n_sweep.txt 2>&1
  B381_MSM_CHUNK_LOG=$cl python tools/gpu_sweep_c.py 21 16 d >> gpurun_out/r02n_sweep.txt 2>&1
done
cat gpurun_out/r02n_sweep.txt
