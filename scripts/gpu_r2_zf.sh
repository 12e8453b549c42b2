#!/bin/bash
# round 2, final call: GPU tests, smoke, bench lines of every workload (e2e + bare-copy ceiling + cpu baseline), reference arm,
# launch lists, ncu --set full of the dominant kernel of c2 and of the new 8f-3 kernels (each after the plain run of the same command)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/f_gpu.txt 2>&1
echo "== smoke" ; timeout 300 python __graft_entry__.py --smoke > gpurun_out/f_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/f_smoke.log
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/f_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/f_pytest.log
summ() { python - "$1" <<'PY'
import json,sys
try:
    l=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    e=l.get("e2e",{})
    print(" ms_per_step", round(l["ms_per_step"],3), "e2e_ms", e.get("ms_per_step"), "ceiling_ms", e.get("host_ceiling_ms"), "stages", (l.get("roofline") or {}).get("stage_ms"))
    if "cpu_baseline" in l: print(" cpu", l["cpu_baseline"]["value"], "gpu value", l["value"], "e2e value", e.get("value"))
except Exception as ex:
    print(" no line", ex)
PY
}
for w in c2 c4cmp c2tagged c3 c4 c4tagged c1; do
  echo "== $w"
  timeout 900 python bench.py --workload $w --steps 10 --warmup 3 > gpurun_out/f_bench_$w.json 2> gpurun_out/f_bench_$w.err
  echo " rc=$?"; summ gpurun_out/f_bench_$w.json
done
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/f_bench_reference.json 2> gpurun_out/f_bench_reference.err; echo " rc=$?"; tail -c 400 gpurun_out/f_bench_reference.json
B="python bench.py --steps 1 --warmup 3 --e2e-steps 0 --no-cpu-baseline"
for w in c2 c2tagged c3 c4 c4cmp; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/f_launches_$w.csv $B --workload $w > gpurun_out/f_ncu_l_$w.log 2>&1; echo "launches $w rc=$?"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rans_raw_fused -s 3 -c 1 -f -o gpurun_out/prof_r2_final_raw_c2 $B --workload c2 > gpurun_out/f_ncu_raw.log 2>&1; echo "ncu raw rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cmp_chain -s 3 -c 1 -f -o gpurun_out/prof_r2_final_cmp_chain $B --workload c4cmp > gpurun_out/f_ncu_cmp.log 2>&1; echo "ncu cmp rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:geo_normal_kernel -s 3 -c 1 -f -o gpurun_out/prof_r2_final_geo_normal $B --workload c4cmp > gpurun_out/f_ncu_geo.log 2>&1; echo "ncu geo rc=$?"
ls -la gpurun_out/*.ncu-rep
