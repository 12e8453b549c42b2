#!/bin/bash
# round 2, call E: par_post2 (persistent warps, 1-D TMA) parity + c2tagged / c3 / c4tagged timings, fast Tagged path on/off
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
python -c "from draco_sharp_b200 import build as B; B.build_all(); B.build_oracle()" > gpurun_out/e_build.log 2>&1
echo "== smoke" ; timeout 300 python __graft_entry__.py --smoke > gpurun_out/e_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/e_smoke.log
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/e_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/e_pytest.log
summ() { python - "$1" <<'PY'
import json,sys
try:
    l=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=l["roofline"]
    print(" ms_per_step", round(l["ms_per_step"],3), "stages", r.get("stage_ms"), r.get("kernel"), "frac", r.get("frac"))
except Exception as e:
    print(" no line", e)
PY
}
for w in c2tagged c4tagged c3; do
  for v in "" "DCB_NO_FAST_TAGGED=1"; do
    echo "== $w $v"
    env $v timeout 900 python bench.py --workload $w --steps 5 --warmup 3 --e2e-steps 0 --no-cpu-baseline > gpurun_out/e_${w}_${v%%=*}.json 2> gpurun_out/e_${w}_${v%%=*}.err
    echo " rc=$?"; summ gpurun_out/e_${w}_${v%%=*}.json
  done
done
