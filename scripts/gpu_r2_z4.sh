#!/bin/bash
# round 2, call Z4: pre-summed cmp_chain, flag kernels on a side stream, geometric normals through the MODE 3 rANS path
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
echo "== mesh tests"; timeout 900 python -m pytest tests/test_gpu_mesh.py -m gpu -x -q > gpurun_out/z4_mesh.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/z4_mesh.log
summ() { python - "$1" <<'PY'
import json,sys
try:
    l=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    e=l.get("e2e",{})
    print(" ms_per_step", round(l["ms_per_step"],3), "e2e_ms", e.get("ms_per_step"), "stages", (l.get("roofline") or {}).get("stage_ms"))
except Exception as ex:
    print(" no line", ex)
PY
}
echo "== c4cmp"; timeout 900 python bench.py --workload c4cmp --steps 5 --warmup 3 --e2e-steps 2 --no-cpu-baseline > gpurun_out/z4_c4cmp.json 2> gpurun_out/z4_c4cmp.err; echo " rc=$?"; summ gpurun_out/z4_c4cmp.json; tail -3 gpurun_out/z4_c4cmp.err
B="python bench.py --workload c4cmp --steps 1 --warmup 3 --e2e-steps 0 --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/z4_launches_c4cmp.csv $B > gpurun_out/z4_ncu_l.log 2>&1; echo "launches rc=$?"
echo "== pytest gpu (rest)"; timeout 1500 python -m pytest tests -m gpu -x -q -k "not mesh and not parity" > gpurun_out/z4_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/z4_pytest.log
