cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r02u_bench_${N}gpu.json 2> gpurun_out/r02u_bench_${N}gpu.err
echo "rc=$? lines=$(wc -l < gpurun_out/r02u_bench_${N}gpu.json)"
python - <<PY
import json
d=json.loads(open('gpurun_out/r02u_bench_${N}gpu.json').read())
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], d['result_check'])
print('ntt4', d['ntt_fourstep']['ms_per_step'], d['ntt_fourstep']['nccl_all_to_all_variant']['ms_per_step'], d['ntt_fourstep']['phases_ms'])
print('commits', d['plonk_commit_round']['ms_per_step'], d['plonk_commit_round']['value'])
PY
if [ "$N" = "2" ]; then python -m pytest tests/test_gpu_nccl.py -x -q -m gpu 2>&1 | tail -3; fi
