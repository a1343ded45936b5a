cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L | head -8
(time python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 5 --warmup 3) > gpurun_out/r02t_bench_8gpu.json 2> gpurun_out/r02t_bench_8gpu.err
echo "rc=$?"
tail -c 400 gpurun_out/r02t_bench_8gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02t_bench_8gpu.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], d['result_check'])
print('phases', d.get('phases_ms'))
print('ntt4', {k: d['ntt_fourstep'][k] for k in ('value','ms_per_step','result_check') if k in d['ntt_fourstep']})
print('commits', d['plonk_commit_round'])
PY
