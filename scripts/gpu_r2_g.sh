#!/bin/bash
# round 2, call G: par_post2 run-length fix: parity + c2tagged timings (whole-stream runs vs chunk runs)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
summ() { python - "$1" <<'PY'
import json,sys
try:
    l=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=l.get("roofline",{})
    print(" ms_per_step", round(l["ms_per_step"],3), "e2e_ms", l.get("e2e",{}).get("ms_per_step"), "frac", r.get("frac"), "stages", r.get("stage_ms"), r.get("kernel"))
except Exception as e:
    print(" no line", e)
PY
}
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/g_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/g_pytest.log
for v in "" "DCB_PAR_RUN=1" "DCB_PAR_RUN=2"; do
  t=$(echo "$v" | tr -c 'A-Za-z0-9\n' '_')
  echo "== c2tagged $v"
  env $v timeout 600 python bench.py --workload c2tagged --steps 5 --warmup 3 --e2e-steps 0 --no-cpu-baseline > gpurun_out/g_c2tagged_$t.json 2> gpurun_out/g_c2tagged_$t.err
  echo " rc=$?"; summ gpurun_out/g_c2tagged_$t.json
done
