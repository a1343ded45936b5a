cd $GRAFT_REPO_ROOT
N=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r02at_bench_${N}gpu.json 2> gpurun_out/r02at_bench_${N}gpu.err
echo "$N-gpu rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r02at_bench_${N}gpu.json').read().strip().splitlines()[-1])
print('value', d['value'], d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['result_check'])
print('pipelined', d['value_pipelined']['ms_per_step'], d['value_pipelined']['result_check'])
PY
