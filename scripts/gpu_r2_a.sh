#!/bin/bash
# round 2, call A: parity of the warp-pair rANS kernels + first timings
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/a_gpu.txt 2>&1
echo "== smoke" ; timeout 300 python __graft_entry__.py --smoke > gpurun_out/a_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/a_smoke.log
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/a_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/a_pytest.log
for v in "" "DCB_PAIRS=1" "DCB_PAIRS=2" "DCB_RANS_V1=1"; do
  echo "== c2 $v"
  env $v DCB_DEBUG_PLAN=1 timeout 600 python bench.py --workload c2 --steps 10 --warmup 3 --e2e-steps 0 --no-cpu-baseline > gpurun_out/a_c2_${v%%=*}${v##*=}.json 2> gpurun_out/a_c2_${v%%=*}${v##*=}.err
  echo "rc=$?"; python - <<PY
import json,sys
try:
    l=json.loads(open("gpurun_out/a_c2_${v%%=*}${v##*=}.json").read().strip().splitlines()[-1])
    print("ms_per_step", l["ms_per_step"], "frac", l["roofline"]["frac"], l["roofline"]["kernel"])
except Exception as e:
    print("no line", e)
PY
  grep "dcb plan" gpurun_out/a_c2_${v%%=*}${v##*=}.err | head -2
done
for w in c2tagged c3; do
  echo "== $w"
  DCB_DEBUG_PLAN=1 timeout 900 python bench.py --workload $w --steps 5 --warmup 3 --e2e-steps 0 --no-cpu-baseline > gpurun_out/a_$w.json 2> gpurun_out/a_$w.err
  echo "rc=$?"; tail -c 1500 gpurun_out/a_$w.json | python -c "import json,sys; l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(l['ms_per_step'], l['roofline']['stage_ms'])" 
done
