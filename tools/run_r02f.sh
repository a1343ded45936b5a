cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
./tools/ubench/atomic_probe > gpurun_out/r02f_atomic.txt 2>&1
python tools/gpu_l0_locality.py > gpurun_out/r02f_l0.txt 2>&1
cat gpurun_out/r02f_atomic.txt gpurun_out/r02f_l0.txt
