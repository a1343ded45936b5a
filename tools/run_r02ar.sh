cd $GRAFT_REPO_ROOT
for ln in 20 21 22 24; do
python bench.py --log-n $ln --steps 6 --warmup 3 --no-cpu-baseline --no-reference-gpu > gpurun_out/r02ar_bench$ln.json 2> gpurun_out/r02ar_bench$ln.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02ar_bench$ln.json').read().strip().splitlines()[-1])
print('2^$ln: value ms', d['ms_per_step'], 'pipelined', d['value_pipelined']['ms_per_step'], d['value_pipelined']['result_check'], 'e2e', d['e2e']['ms_per_step'])
PY
done
