#!/bin/bash
# round 2, call ZG (2 GPUs): GPU suite on a 2-device box (library-level sharding test runs), configs[4] sweep under torchrun on 2 GPUs
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/g_gpus.txt 2>&1
echo "== pytest gpu (2 devices)"; timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/g_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/g_pytest.log
echo "== sweep on 2 GPUs"
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --sweep > gpurun_out/g_sweep_2gpus.jsonl 2> gpurun_out/g_sweep_2gpus.err; echo "sweep rc=$?"
wc -l gpurun_out/g_sweep_2gpus.jsonl; tail -2 gpurun_out/g_sweep_2gpus.err
