cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m pytest tests/test_gpu_ntt.py -x -q -m gpu -k "oracle or golden" 2>&1 | tail -2
for s in 0 2 3; do B381_NTT_SHAPE=$s python tools/gpu_ntt_bench.py 24,22,20,16:256 10; done
