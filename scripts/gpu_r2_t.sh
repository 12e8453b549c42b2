#!/bin/bash
# round 2, call T: why is rank 1 of a 2-GPU run slower?  the batches of ranks 0..3 on one GPU, with the launch plan
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for r in 0 1 2 3; do
  echo "== data rank $r"
  DCB_BENCH_DATA_RANK=$r DCB_DEBUG_PLAN=1 timeout 600 python bench.py --steps 5 --warmup 3 --e2e-steps 0 --no-cpu-baseline > gpurun_out/t_c2_r$r.json 2> gpurun_out/t_c2_r$r.err; echo " rc=$?"
  python -c "
import json; l=json.loads(open('gpurun_out/t_c2_r$r.json').read().strip().splitlines()[-1]); print(l['ms_per_step'], l['roofline']['stage_ms'], l['roofline']['note'])"
  grep "dcb plan" gpurun_out/t_c2_r$r.err | sort | uniq -c | head -3
done
