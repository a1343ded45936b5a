cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(time python -m pytest tests/test_gpu_msm.py tests/test_gpu_dist.py tests/test_icicle_dispatch.py -m gpu -x -q) > gpurun_out/r02j_pytest.log 2>&1
tail -n 4 gpurun_out/r02j_pytest.log
(time python bench.py --steps 5 --warmup 3) > gpurun_out/r02j_bench.json 2> gpurun_out/r02j_bench.err
echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02j_bench.json'))
print('value ms', d['ms_per_step'], 'e2e ms', d['e2e']['ms_per_step'], 'pipe ms', d['e2e_pipelined']['ms_per_step'], d['result_check'])
print(d['phases_ms'])
PY
