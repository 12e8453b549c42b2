#!/bin/bash
# round 2, call R: par_post2 with round-by-round tickets for look-back runs: parity, Tagged sweep cells, c2tagged / c4tagged
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r_pytest.log
echo "== sweep"; timeout 1500 python bench.py --sweep > gpurun_out/r_sweep.jsonl 2> gpurun_out/r_sweep.err; echo "sweep rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/r_sweep.jsonl'):
    d=json.loads(l)
    print(d["sweep"][:6], d["scheme"], d["points_per_buffer"], d["buffers"], "ms %.2f"%d["ms_per_step"], "frac %.3f"%d["frac_of_hbm_peak"], d["parity_ok"], d["stage_ms"])
PY
for w in c2tagged c4tagged; do
  echo "== $w"; timeout 600 python bench.py --workload $w --steps 5 --warmup 3 --e2e-steps 1 --no-cpu-baseline > gpurun_out/r_$w.json 2> gpurun_out/r_$w.err; echo " rc=$?"
  python -c "
import json; l=json.loads(open('gpurun_out/r_$w.json').read().strip().splitlines()[-1]); print(l['ms_per_step'], l['e2e']['ms_per_step'], l['roofline']['stage_ms'])"
done
