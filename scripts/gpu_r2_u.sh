#!/bin/bash
# round 2, call U: outlier groups stay small next to a machine-filling group: batches of ranks 1, 4..7 on one GPU; full GPU tests
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/u_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/u_pytest.log
for r in 1 4 5 6 7; do
  echo "== data rank $r"
  DCB_BENCH_DATA_RANK=$r DCB_DEBUG_PLAN=1 timeout 600 python bench.py --steps 5 --warmup 3 --e2e-steps 1 --no-cpu-baseline --no-copy-ceiling > gpurun_out/u_c2_r$r.json 2> gpurun_out/u_c2_r$r.err; echo " rc=$?"
  python -c "
import json; l=json.loads(open('gpurun_out/u_c2_r$r.json').read().strip().splitlines()[-1]); print(l['ms_per_step'], l['e2e']['ms_per_step'], l['roofline']['stage_ms'])"
  grep "dcb plan" gpurun_out/u_c2_r$r.err | sort | uniq -c | grep -v "10000 streams\|125[0-9] streams\|124[0-9] streams" | head -4
done
for w in c2tagged c3; do
  echo "== $w data rank 1"
  DCB_BENCH_DATA_RANK=1 timeout 600 python bench.py --workload $w --steps 5 --warmup 3 --e2e-steps 0 --no-cpu-baseline > gpurun_out/u_${w}_r1.json 2> gpurun_out/u_${w}_r1.err; echo " rc=$?"
  python -c "
import json; l=json.loads(open('gpurun_out/u_${w}_r1.json').read().strip().splitlines()[-1]); print(l['ms_per_step'], l['roofline']['stage_ms'])"
done
