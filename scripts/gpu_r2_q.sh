#!/bin/bash
# round 2, call Q: compute-sanitizer (memcheck, racecheck) over the small GPU parity tests; configs[4] sweep on one GPU
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
SEL='not million and not full_size and not 300 and not slices and not 100000'
echo "== memcheck"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --launch-timeout 120 python -m pytest tests/test_gpu_crafted.py tests/test_gpu_mesh.py -m gpu -x -q -k "$SEL" > gpurun_out/q_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -6 gpurun_out/q_memcheck.log
echo "== racecheck"
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 --launch-timeout 120 python -m pytest tests/test_gpu_mesh.py -m gpu -x -q -k "house or texcoords or random_corrections or 17-9" > gpurun_out/q_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -6 gpurun_out/q_racecheck.log
echo "== sweep"
timeout 1500 python bench.py --sweep > gpurun_out/q_sweep.jsonl 2> gpurun_out/q_sweep.err; echo "sweep rc=$?"; wc -l gpurun_out/q_sweep.jsonl; tail -2 gpurun_out/q_sweep.err
