#!/bin/bash
# round 2, call O: TexCoordsPortable kernels on the reference's whole sample; c2tagged e2e after the slice-aware par_post2 plan;
# c3 with fuller warps (3 CTAs per SM)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
echo "== pytest mesh"; timeout 900 python -m pytest tests/test_gpu_mesh.py -x -q > gpurun_out/o_pytest_mesh.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/o_pytest_mesh.log
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/o_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/o_pytest.log
summ() { python - "$1" <<'PY'
import json,sys
try:
    l=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(" ms_per_step", round(l["ms_per_step"],3), "e2e_ms", l.get("e2e",{}).get("ms_per_step"), "stages", (l.get("roofline") or {}).get("stage_ms"), (l.get("roofline") or {}).get("kernel"))
    if "cpu_baseline" in l: print(" cpu", l["cpu_baseline"]["value"], "gpu value", l["value"], "e2e value", l["e2e"]["value"])
except Exception as e:
    print(" no line", e)
PY
}
echo "== c1"; timeout 300 python bench.py --workload c1 > gpurun_out/o_c1.json 2> gpurun_out/o_c1.err; echo " rc=$?"; summ gpurun_out/o_c1.json; tail -3 gpurun_out/o_c1.err
echo "== c2tagged e2e"; timeout 600 python bench.py --workload c2tagged --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/o_c2tagged.json 2> gpurun_out/o_c2tagged.err; echo " rc=$?"; summ gpurun_out/o_c2tagged.json
for v in "" "DCB_CTAS_PER_SM=3" "DCB_CTAS_PER_SM=2"; do
  echo "== c3 $v"
  env $v timeout 600 python bench.py --workload c3 --steps 5 --warmup 3 --e2e-steps 0 --no-cpu-baseline > gpurun_out/o_c3_${v##*=}.json 2> gpurun_out/o_c3_${v##*=}.err
  echo " rc=$?"; summ gpurun_out/o_c3_${v##*=}.json
done
