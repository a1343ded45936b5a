cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q) > gpurun_out/r02i_pytest.log 2>&1
tail -n 6 gpurun_out/r02i_pytest.log
(time python bench.py --steps 5 --warmup 3) > gpurun_out/r02i_bench.json 2> gpurun_out/r02i_bench.err
echo "bench rc=$?"
tail -c 600 gpurun_out/r02i_bench.err
head -c 1500 gpurun_out/r02i_bench.json
