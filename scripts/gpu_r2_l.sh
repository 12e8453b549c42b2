#!/bin/bash
# round 2, call L: chain warps in the upper half of the CTA (arbiter prefers the higher warp id)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
summ() { python - "$1" <<'PY'
import json,sys
try:
    l=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=l.get("roofline",{})
    print(" ms_per_step", round(l["ms_per_step"],3), "frac", r.get("frac"), r.get("kernel"))
except Exception as e:
    print(" no line", e)
PY
}
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/l_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/l_pytest.log
for v in "" "DCB_CHAIN_LOW=1" "DCB_RANS_PC=1" "DCB_RANS_PC=1 DCB_CHAIN_LOW=1"; do
  t=$(echo "$v" | tr -c 'A-Za-z0-9\n' '_')
  echo "== c2 $v"
  env $v timeout 600 python bench.py --workload c2 --steps 5 --warmup 3 --e2e-steps 0 --no-cpu-baseline > gpurun_out/l_c2_$t.json 2> gpurun_out/l_c2_$t.err
  echo " rc=$?"; summ gpurun_out/l_c2_$t.json
done
