#!/bin/bash
# round 2, call J: ncu --set full of the bucket-record Raw kernel on c2
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 3 --e2e-steps 0 --no-cpu-baseline"
$B --workload c2 > gpurun_out/j_plain_c2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:rans_raw_rec -s 3 -c 1 -f -o gpurun_out/prof_r2_rec_c2 $B --workload c2 > gpurun_out/j_ncu_rec.log 2>&1
echo "rec rc=$?"
ls -la gpurun_out/*.ncu-rep
