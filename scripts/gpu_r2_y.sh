#!/bin/bash
# round 2, call Y: software-pipelined lean loops for MODE 2 and the tag kernel
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/y_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/y_pytest.log
summ() { python - "$1" <<'PY'
import json,sys
try:
    l=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(" ms_per_step", round(l["ms_per_step"],3), "stages", (l.get("roofline") or {}).get("stage_ms"))
except Exception as ex:
    print(" no line", ex)
PY
}
for w in c4 c4tagged c1 c2tagged; do
    echo "== $w"
    timeout 600 python bench.py --workload $w --steps 5 --warmup 3 --e2e-steps 0 --no-cpu-baseline > gpurun_out/y_${w}.json 2> gpurun_out/y_${w}.err; echo " rc=$?"; summ gpurun_out/y_${w}.json
done
