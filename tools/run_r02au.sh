cd $GRAFT_REPO_ROOT
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
(time python bench.py) > gpurun_out/r02au_bench.json 2> gpurun_out/r02au_bench.err
echo "bench rc=$?"; tail -4 gpurun_out/r02au_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02au_bench.json').read().strip().splitlines()[-1])
print('value ms', d['ms_per_step'], 'e2e ms', d['e2e']['ms_per_step'], 'e2e pipe', d['e2e_pipelined']['ms_per_step'], 'value_pipelined', d['value_pipelined']['ms_per_step'], d['value_pipelined']['result_check'], d['result_check'], 'ntt', d['ntt']['ms_per_step'], d['ntt']['result_check'], 'g2', d['g2']['ms_per_step'], d['g2']['result_check'])
print(d['steps'], d['warmup'], d['clocks'])
PY
