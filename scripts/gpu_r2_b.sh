#!/bin/bash
# round 2, call B: how many rANS warps per SM sub-partition?  (single-warp kernels, CTAs per SM swept)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for c in 3 4 6 8 10 12; do
  echo "== v1 ctas/sm=$c"
  DCB_RANS_V1=1 DCB_CTAS_PER_SM=$c DCB_DEBUG_PLAN=1 timeout 600 python bench.py --workload c2 --steps 8 --warmup 3 --e2e-steps 0 --no-cpu-baseline > gpurun_out/b_v1_$c.json 2> gpurun_out/b_v1_$c.err
  python - <<PY
import json
try:
    l=json.loads(open("gpurun_out/b_v1_$c.json").read().strip().splitlines()[-1])
    print("ms_per_step", round(l["ms_per_step"],2), "frac", round(l["roofline"]["frac"],4), l["roofline"]["note"])
except Exception as e:
    print("no line", e)
PY
  grep "dcb plan" gpurun_out/b_v1_$c.err | head -1
done
