#!/bin/bash
# round 2, call P: tex-coord chain on two lanes + random-correction tests; c1 timeline
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
echo "== pytest mesh"; timeout 900 python -m pytest tests/test_gpu_mesh.py -x -q > gpurun_out/p_pytest_mesh.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/p_pytest_mesh.log
echo "== c1"; timeout 300 python bench.py --workload c1 > gpurun_out/p_c1.json 2> gpurun_out/p_c1.err; echo " rc=$?"; python -c "
import json; l=json.loads(open('gpurun_out/p_c1.json').read().strip().splitlines()[-1]); print(l['ms_per_step'], l['roofline']['note'], l['cpu_baseline'])"
echo "== c1 timeline"; DCB_DEBUG_TIMING=1 timeout 300 python bench.py --workload c1 > /dev/null 2> gpurun_out/p_c1_timeline.err; tail -30 gpurun_out/p_c1_timeline.err
