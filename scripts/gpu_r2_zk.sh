#!/bin/bash
# round 2, call ZK: cmp_chain unrolled by three -- mesh GPU tests, c4cmp step, launch list
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
echo "== mesh tests"; timeout 600 python -m pytest tests/test_gpu_mesh.py -m gpu -x -q > gpurun_out/k_mesh.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/k_mesh.log
B="python bench.py --workload c4cmp --steps 3 --warmup 3 --e2e-steps 0 --no-cpu-baseline"
timeout 300 $B > gpurun_out/k_c4cmp.json 2> gpurun_out/k_c4cmp.err; echo "rc=$?"; python -c "
import json; l=json.loads(open('gpurun_out/k_c4cmp.json').read().strip().splitlines()[-1]); print(l['ms_per_step'], l['roofline']['stage_ms'])"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:cmp_chain -c 3 --csv --log-file gpurun_out/k_launches.csv $B > gpurun_out/k_ncu.log 2>&1; grep cmp_chain gpurun_out/k_launches.csv | tail -2
