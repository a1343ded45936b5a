cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q) > gpurun_out/r02p_pytest.log 2>&1
tail -n 12 gpurun_out/r02p_pytest.log | head -9
(time python bench.py --steps 5 --warmup 3) > gpurun_out/r02p_bench.json 2> gpurun_out/r02p_bench.err
echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02p_bench.json'))
print('value ms', d['ms_per_step'], 'e2e ms', d['e2e']['ms_per_step'], 'pipe ms', d['e2e_pipelined']['ms_per_step'], d['result_check'])
print(d['phases_ms'])
print('ntt', d['ntt']['ms_per_step'], 'g2', d['g2']['ms'] if d.get('g2') else None)
PY
