cd $GRAFT_REPO_ROOT
N=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tools/gpu_pipe_nccl.py ${2:-22} 8 2>&1 | grep -v "^\*\|OMP_NUM\|NCCL version" | tee gpurun_out/r02as_pipe_${N}gpu.txt
