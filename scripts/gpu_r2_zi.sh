#!/bin/bash
# round 2, call ZI: mesh maps staged through pinned memory by host threads -- mesh GPU tests, c4 and c4cmp end to end
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
echo "== mesh tests"; timeout 600 python -m pytest tests/test_gpu_mesh.py -m gpu -x -q > gpurun_out/i_mesh.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/i_mesh.log
summ() { python - "$1" <<'PY'
import json,sys
try:
    l=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    e=l.get("e2e",{})
    print(" ms_per_step", round(l["ms_per_step"],3), "e2e_ms", e.get("ms_per_step"), "e2e value", e.get("value"))
except Exception as ex:
    print(" no line", ex)
PY
}
for w in c4 c4cmp; do
  echo "== $w"; timeout 600 python bench.py --workload $w --steps 3 --warmup 3 --e2e-steps 4 --no-cpu-baseline > gpurun_out/i_$w.json 2> gpurun_out/i_$w.err; echo " rc=$?"; summ gpurun_out/i_$w.json
done
