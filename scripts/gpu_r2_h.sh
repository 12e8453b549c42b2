#!/bin/bash
# round 2, call H: ncu --set full (source-level stall sampling) of the c2 Raw kernel (single warp and warp pairs) and of par_post2 on c2tagged
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 3 --e2e-steps 0 --no-cpu-baseline"
$B --workload c2 > gpurun_out/h_plain_c2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:rans_raw_fused -s 3 -c 1 -f -o gpurun_out/prof_r2_raw_c2 $B --workload c2 > gpurun_out/h_ncu_raw.log 2>&1
echo "raw rc=$?"
DCB_RANS_PC=1 $B --workload c2 > gpurun_out/h_plain_c2pc.log 2>&1 && \
DCB_RANS_PC=1 ncu --set full --clock-control none --import-source on -k regex:rans_raw_pc -s 3 -c 1 -f -o gpurun_out/prof_r2_pc_c2 $B --workload c2 > gpurun_out/h_ncu_pc.log 2>&1
echo "pc rc=$?"
$B --workload c2tagged > gpurun_out/h_plain_c2tagged.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:par_post2 -s 3 -c 1 -f -o gpurun_out/prof_r2_parpost_c2tagged $B --workload c2tagged > gpurun_out/h_ncu_pp.log 2>&1
echo "pp rc=$?"
ls -la gpurun_out/*.ncu-rep
