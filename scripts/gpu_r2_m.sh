#!/bin/bash
# round 2, call M: state of HEAD -- GPU tests, smoke, full bench lines (e2e + cpu baseline) of every workload,
# ncu summaries of the kernels round 1 left without one (rans_tag, oct_chain, para_deps), single-warp issue-rate probe
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/m_gpu.txt 2>&1
echo "== ubench"; timeout 120 scripts/ubench/issue_rate > gpurun_out/m_issue_rate.txt 2>&1; echo "rc=$?"; tail -3 gpurun_out/m_issue_rate.txt
echo "== smoke" ; timeout 300 python __graft_entry__.py --smoke > gpurun_out/m_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/m_smoke.log
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/m_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/m_pytest.log
summ() { python - "$1" <<'PY'
import json,sys
try:
    l=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(" ms_per_step", round(l["ms_per_step"],3), "e2e_ms", l.get("e2e",{}).get("ms_per_step"), "stages", l.get("roofline",{}).get("stage_ms"), l.get("roofline",{}).get("kernel"))
    if "cpu_baseline" in l: print(" cpu", l["cpu_baseline"]["value"], "gpu value", l["value"], "e2e value", l["e2e"]["value"])
except Exception as e:
    print(" no line", e)
PY
}
for w in c2 c2tagged c3 c4 c4tagged c1; do
  echo "== $w"
  timeout 900 python bench.py --workload $w --steps 10 --warmup 3 > gpurun_out/m_bench_$w.json 2> gpurun_out/m_bench_$w.err
  echo " rc=$?"; summ gpurun_out/m_bench_$w.json
done
B="python bench.py --steps 1 --warmup 3 --e2e-steps 0 --no-cpu-baseline"
ncu --set full --clock-control none --import-source on -k regex:rans_tag -s 3 -c 1 -f -o gpurun_out/prof_r2_tag_c2tagged $B --workload c2tagged > gpurun_out/m_ncu_tag.log 2>&1; echo "tag rc=$?"
ncu --set full --clock-control none --import-source on -k regex:oct_chain -s 3 -c 1 -f -o gpurun_out/prof_r2_octchain_c3 $B --workload c3 > gpurun_out/m_ncu_oct.log 2>&1; echo "oct rc=$?"
ncu --set full --clock-control none --import-source on -k regex:para_deps -s 3 -c 1 -f -o gpurun_out/prof_r2_paradeps_c4 $B --workload c4tagged > gpurun_out/m_ncu_pd.log 2>&1; echo "pd rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/m_launches_c2.csv $B --workload c2 > gpurun_out/m_ncu_l2.log 2>&1; echo "launches c2 rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/m_launches_c2tagged.csv $B --workload c2tagged > gpurun_out/m_ncu_l2t.log 2>&1; echo "launches c2tagged rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/m_launches_c3.csv $B --workload c3 > gpurun_out/m_ncu_l3.log 2>&1; echo "launches c3 rc=$?"
ls -la gpurun_out/*.ncu-rep
