cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m pytest tests/test_gpu_ntt.py tests/test_gpu_dist.py tests/test_gpu_vs_reference.py tests/test_icicle_dispatch.py -x -q -m gpu 2>&1 | tail -2
python tools/gpu_ntt_bench.py 26,24,22,20,18,16:256,13:8192 10 > gpurun_out/r02y_ntt.txt 2>&1; cat gpurun_out/r02y_ntt.txt
