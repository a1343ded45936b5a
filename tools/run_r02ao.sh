cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r02ao_bench_${N}gpu.json 2> gpurun_out/r02ao_bench_${N}gpu.err
echo "$N-gpu rc=$?"
tail -c 5200 gpurun_out/r02ao_bench_${N}gpu.json
