#!/bin/bash
# round 2, call F: re-measure everything after the container loss: smoke, every workload, planner variants of c2, pytest
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/f_gpu.txt 2>&1
python -c "from draco_sharp_b200 import build as B; B.build_all(); B.build_oracle()" > gpurun_out/f_build.log 2>&1
echo "== smoke" ; timeout 300 python __graft_entry__.py --smoke > gpurun_out/f_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/f_smoke.log
summ() { python - "$1" <<'PY'
import json,sys
try:
    l=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=l.get("roofline",{})
    print(" ms_per_step", round(l["ms_per_step"],3), "e2e_ms", l.get("e2e",{}).get("ms_per_step"), "frac", r.get("frac"), "stages", r.get("stage_ms"), r.get("kernel"))
    if "cpu_baseline" in l: print(" cpu", l["cpu_baseline"]["value"], "gpu value", l["value"], "e2e value", l["e2e"]["value"])
except Exception as e:
    print(" no line", e)
PY
}
for v in "" "DCB_CTAS_PER_SM=3" "DCB_CTAS_PER_SM=6" "DCB_CTAS_PER_SM=8" "DCB_RANS_PC=1" "DCB_RANS_PC=1 DCB_PAIRS=2" "DCB_RANS_PC=1 DCB_PAIRS=8"; do
  t=$(echo "$v" | tr -c 'A-Za-z0-9\n' '_')
  echo "== c2 $v"
  env $v DCB_DEBUG_PLAN=1 timeout 600 python bench.py --workload c2 --steps 8 --warmup 3 --e2e-steps 0 --no-cpu-baseline > gpurun_out/f_c2_$t.json 2> gpurun_out/f_c2_$t.err
  echo " rc=$?"; summ gpurun_out/f_c2_$t.json; grep "dcb plan" gpurun_out/f_c2_$t.err | head -1
done
for w in c2tagged c3 c4 c4tagged c1; do
  echo "== $w"
  DCB_DEBUG_PLAN=1 timeout 900 python bench.py --workload $w --steps 5 --warmup 3 --e2e-steps 2 > gpurun_out/f_$w.json 2> gpurun_out/f_$w.err
  echo " rc=$?"; summ gpurun_out/f_$w.json; grep "dcb plan" gpurun_out/f_$w.err | sort | uniq -c | head -4
done
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/f_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/f_pytest.log
