#!/bin/bash
# round 2, call N: lean main loop of the fused Raw kernel (window byte supply, predicated renormalisation, regular wrap):
# parity + A/B timing on c2 / c3
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/n_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/n_pytest.log
summ() { python - "$1" <<'PY'
import json,sys
try:
    l=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(" ms_per_step", round(l["ms_per_step"],3), "stages", l.get("roofline",{}).get("stage_ms"), l.get("roofline",{}).get("kernel"))
except Exception as e:
    print(" no line", e)
PY
}
for w in c2 c3; do
  for v in "" "DCB_NO_LEAN=1"; do
    echo "== $w $v"
    env $v timeout 600 python bench.py --workload $w --steps 8 --warmup 3 --e2e-steps 0 --no-cpu-baseline > gpurun_out/n_${w}_${v%%=*}.json 2> gpurun_out/n_${w}_${v%%=*}.err
    echo " rc=$?"; summ gpurun_out/n_${w}_${v%%=*}.json
  done
done
