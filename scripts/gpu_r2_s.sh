#!/bin/bash
# round 2, call S (2 GPUs): the library's own multi-device sharding; c2 at N=2 with the bare-copy ceiling; in-library 2-device e2e
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/s_gpus.txt
echo "== pytest multi-device + sequential meshes"; timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_mesh.py -m gpu -x -q -k "distinct_devices or sequential_meshes" > gpurun_out/s_pytest.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/s_pytest.log
echo "== c2 N=2"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/s_c2_n2.json 2> gpurun_out/s_c2_n2.err; echo " rc=$?"
python -c "
import json; l=json.loads(open('gpurun_out/s_c2_n2.json').read().strip().splitlines()[-1]); print(l['n_gpus'], l['ms_per_step'], l['value']); print(json.dumps(l['e2e']))"
echo "== c2 N=1 (same box)"
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/s_c2_n1.json 2> gpurun_out/s_c2_n1.err; echo " rc=$?"
python -c "
import json; l=json.loads(open('gpurun_out/s_c2_n1.json').read().strip().splitlines()[-1]); print(l['n_gpus'], l['ms_per_step'], l['value']); print(json.dumps(l['e2e']))"
