#!/bin/bash
# round 2, call Z3: the 8f-3 predictors at configs[3] size (c4cmp): bench line, launch list, ncu --set full of cmp_chain_kernel
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
ls -la bench_cache/ 2>&1 | tail -3
summ() { python - "$1" <<'PY'
import json,sys
try:
    l=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    e=l.get("e2e",{})
    print(" ms_per_step", round(l["ms_per_step"],3), "e2e_ms", e.get("ms_per_step"), "stages", (l.get("roofline") or {}).get("stage_ms"))
    if "cpu_baseline" in l: print(" cpu", l["cpu_baseline"]["value"], "gpu value", l["value"], "e2e value", e.get("value"))
except Exception as ex:
    print(" no line", ex)
PY
}
echo "== c4cmps"; timeout 300 python bench.py --workload c4cmps --steps 3 --warmup 3 --e2e-steps 1 --no-cpu-baseline > gpurun_out/z3_c4cmps.json 2> gpurun_out/z3_c4cmps.err; echo " rc=$?"; summ gpurun_out/z3_c4cmps.json; tail -3 gpurun_out/z3_c4cmps.err
echo "== c4cmp"; timeout 900 python bench.py --workload c4cmp --steps 5 --warmup 3 > gpurun_out/z3_c4cmp.json 2> gpurun_out/z3_c4cmp.err; echo " rc=$?"; summ gpurun_out/z3_c4cmp.json; tail -3 gpurun_out/z3_c4cmp.err
B="python bench.py --workload c4cmp --steps 1 --warmup 3 --e2e-steps 0 --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/z3_launches_c4cmp.csv $B > gpurun_out/z3_ncu_l.log 2>&1; echo "launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cmp_chain -s 3 -c 1 -f -o gpurun_out/prof_r2_cmp_chain_c4cmp $B > gpurun_out/z3_ncu_cmp.log 2>&1; echo "ncu cmp rc=$?"
ls -la gpurun_out/*.ncu-rep
