set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python tools/gpu_ntt_bench.py 24 3 > gpurun_out/r02c_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_ntt -s 9 -c 3 -o gpurun_out/r02c_ntt python tools/gpu_ntt_bench.py 24 3 > gpurun_out/r02c_ncu.log 2>&1
ls -la gpurun_out/
