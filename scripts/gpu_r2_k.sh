#!/bin/bash
# round 2, call K: split layout (chain warps on sub-partitions 0..2, consumers on 3): parity + c2 timings
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
summ() { python - "$1" <<'PY'
import json,sys
try:
    l=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=l.get("roofline",{})
    print(" ms_per_step", round(l["ms_per_step"],3), "frac", r.get("frac"), "stages", r.get("stage_ms"), r.get("kernel"))
except Exception as e:
    print(" no line", e)
PY
}
echo "== pytest gpu (split, rec)"; DCB_SPLIT=1 timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/k_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/k_pytest.log
echo "== pytest gpu (split, pc)"; DCB_SPLIT=1 DCB_RANS_PC=1 timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/k_pytest2.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/k_pytest2.log
for v in "DCB_SPLIT=1" "DCB_SPLIT=1 DCB_RANS_PC=1"; do
  t=$(echo "$v" | tr -c 'A-Za-z0-9\n' '_')
  echo "== c2 $v"
  env $v DCB_DEBUG_PLAN=1 timeout 600 python bench.py --workload c2 --steps 5 --warmup 3 --e2e-steps 0 --no-cpu-baseline > gpurun_out/k_c2_$t.json 2> gpurun_out/k_c2_$t.err
  echo " rc=$?"; summ gpurun_out/k_c2_$t.json; grep "dcb plan" gpurun_out/k_c2_$t.err | sort | uniq -c | sort -rn | head -1 | cut -c1-330
done
