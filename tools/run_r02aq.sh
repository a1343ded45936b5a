cd $GRAFT_REPO_ROOT
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-reference-gpu > gpurun_out/r02aq_bench.json 2> gpurun_out/r02aq_bench.err
echo rc=$?
tail -5 gpurun_out/r02aq_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02aq_bench.json').read().strip().splitlines()[-1])
print('value ms', d['ms_per_step'], 'pipelined', d['value_pipelined'], 'e2e', d['e2e']['ms_per_step'], d['e2e_pipelined']['ms_per_step'])
PY
python bench.py --log-n 21 --steps 5 --warmup 3 --no-cpu-baseline --no-reference-gpu > gpurun_out/r02aq_bench21.json 2> gpurun_out/r02aq_bench21.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02aq_bench21.json').read().strip().splitlines()[-1])
print('2^21: value ms', d['ms_per_step'], 'pipelined', d['value_pipelined']['ms_per_step'], d['value_pipelined']['result_check'], 'e2e', d['e2e']['ms_per_step'])
PY
python -m pytest tests/test_gpu_dist.py -x -q -m gpu 2>&1 | tail -2
