#!/bin/bash
# round 2, call D: parity on every rANS kernel path; direct slot LUT on the low-residency workloads
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
python -c "from draco_sharp_b200 import build as B; B.build_all(); B.build_oracle()" > gpurun_out/d_build.log 2>&1
echo "== pytest gpu"; timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/d_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/d_pytest.log
summ() { python - "$1" <<'PY'
import json,sys
try:
    l=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(" ms_per_step", round(l["ms_per_step"],3), "e2e_ms", l.get("e2e",{}).get("ms_per_step"), "stages", l.get("roofline",{}).get("stage_ms"), l.get("roofline",{}).get("kernel"))
    if "cpu_baseline" in l: print(" cpu", l["cpu_baseline"]["value"], "gpu value", l["value"], "e2e value", l["e2e"]["value"])
except Exception as e:
    print(" no line", e)
PY
}
for w in c4 c4tagged c1; do
  for v in "" "DCB_NO_DIRECT=1"; do
    echo "== $w $v"
    env $v DCB_DEBUG_PLAN=1 timeout 900 python bench.py --workload $w --steps 5 --warmup 3 --e2e-steps 2 > gpurun_out/d_${w}_${v%%=*}.json 2> gpurun_out/d_${w}_${v%%=*}.err
    echo " rc=$?"; summ gpurun_out/d_${w}_${v%%=*}.json; grep "dcb plan" gpurun_out/d_${w}_${v%%=*}.err | sort | uniq -c | head -3
  done
done
echo "== c2 default"; DCB_DEBUG_PLAN=1 timeout 600 python bench.py --workload c2 --steps 8 --warmup 3 --e2e-steps 0 --no-cpu-baseline > gpurun_out/d_c2.json 2> gpurun_out/d_c2.err; summ gpurun_out/d_c2.json; grep "dcb plan" gpurun_out/d_c2.err | head -1
