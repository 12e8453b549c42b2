#!/bin/bash
# round 2, call ZH (N GPUs): configs[4] sweep under torchrun on all visible GPUs
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
echo "== sweep on $N GPUs"
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --sweep > gpurun_out/h_sweep_${N}gpus.jsonl 2> gpurun_out/h_sweep_${N}gpus.err; echo "sweep rc=$?"
wc -l gpurun_out/h_sweep_${N}gpus.jsonl; tail -2 gpurun_out/h_sweep_${N}gpus.err
