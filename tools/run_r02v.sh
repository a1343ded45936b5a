cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m pytest tests/test_gpu_msm.py tests/test_gpu_dist.py -x -q -m gpu > gpurun_out/r02v_pytest.log 2>&1; tail -2 gpurun_out/r02v_pytest.log
python tools/gpu_batch_bench.py 10,16 12,16 16,8 20,8 22,4 22,8 23,2 > gpurun_out/r02v_batch.txt 2>&1
cat gpurun_out/r02v_batch.txt
python tools/gpu_sweep_g2.py 20 16 d
python tools/gpu_sweep_c.py 24 16 d
