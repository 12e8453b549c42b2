#!/bin/bash
# round 2, call Z5: ncu --set full of cmp_chain_kernel (second version) on c4cmp
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
B="python bench.py --workload c4cmp --steps 1 --warmup 3 --e2e-steps 0 --no-cpu-baseline"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cmp_chain -s 3 -c 1 -f -o gpurun_out/prof_r2_cmp_chain_v2_c4cmp $B > gpurun_out/z5_ncu_cmp.log 2>&1; echo "ncu cmp rc=$?"
ls -la gpurun_out/*.ncu-rep
