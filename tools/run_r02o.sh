cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
rm -f gpurun_out/r02o_sweep.txt
for cl in 21 20; do
  echo "== B381_MSM_CHUNK_LOG=$cl" >> gpurun_out/r02o_sweep.txt
  B381_MSM_CHUNK_LOG=$cl python tools/gpu_sweep_c.py 24 16 d >> gpurun_out/r02o_sweep.txt 2>&1
done
for cl in 31 20; do
  echo "== B381_MSM_CHUNK_LOG=$cl" >> gpurun_out/r02o_sweep.txt
  B381_MSM_CHUNK_LOG=$cl python tools/gpu_sweep_c.py 23 16 d >> gpurun_out/r02o_sweep.txt 2>&1
done
cat gpurun_out/r02o_sweep.txt
