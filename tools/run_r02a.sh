set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02a_smi.txt
(time python -m pytest tests -m gpu -x -q) > gpurun_out/r02a_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02a_pytest.log
(time python bench.py --steps 5 --warmup 3) > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err
echo "bench rc=$?" >> gpurun_out/r02a_bench.err
python tools/gpu_sweep_c.py 24 16,17,18,19,20 d,5,6,7 > gpurun_out/r02a_sweep24.txt 2>&1
python tools/gpu_sweep_c.py 22 14,15,16,17,18 d,3,4,5 > gpurun_out/r02a_sweep22.txt 2>&1
python tools/gpu_sweep_c.py 21 13,14,15,16,17 d,3,4 > gpurun_out/r02a_sweep21.txt 2>&1
tail -3 gpurun_out/r02a_pytest.log
