#!/bin/bash
# round 2, call C: ncu --set full of the warp-pair Raw kernel on c2 (after a plain run of the same command)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
CMD="python bench.py --workload c2 --steps 1 --warmup 3 --e2e-steps 0 --no-cpu-baseline"
$CMD > gpurun_out/c_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:rans_raw_pc -s 3 -c 1 -o gpurun_out/prof_r2_pc_c2 $CMD > gpurun_out/c_ncu.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/c_ncu.log
