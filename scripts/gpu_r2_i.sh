#!/bin/bash
# round 2, call I: bucket-record rANS kernels: parity + c2 / c2tagged / c3 timings
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
echo "== smoke" ; timeout 300 python __graft_entry__.py --smoke > gpurun_out/i_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/i_smoke.log
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/i_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/i_pytest.log
for w in c2 c2tagged c3; do
  for v in "" "DCB_NO_REC=1"; do
    t=$(echo "$v" | tr -c 'A-Za-z0-9\n' '_')
    echo "== $w $v"
    env $v DCB_DEBUG_PLAN=1 timeout 600 python bench.py --workload $w --steps 5 --warmup 3 --e2e-steps 0 --no-cpu-baseline > gpurun_out/i_${w}_$t.json 2> gpurun_out/i_${w}_$t.err
    echo " rc=$?"; summ gpurun_out/i_${w}_$t.json; grep "dcb plan" gpurun_out/i_${w}_$t.err | sort | uniq -c | head -4
  done
done
