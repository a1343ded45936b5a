cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(time python -m pytest tests/test_gpu_msm.py tests/test_gpu_dist.py tests/test_icicle_dispatch.py -m gpu -x -q) > gpurun_out/r02l_pytest.log 2>&1
tail -n 12 gpurun_out/r02l_pytest.log | head -9
python tools/gpu_batch_bench.py 10,16 12,16 14,16 16,8 18,8 20,8 22,8 > gpurun_out/r02l_batch.txt 2>&1
cat gpurun_out/r02l_batch.txt
