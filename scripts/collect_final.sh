#!/bin/bash
# copy the final GPU call's outputs (gpurun_out/f_*) into profiles/ under their round-2 names
cd /root/repo
for w in c1 c2 c2tagged c3 c4 c4tagged c4cmp reference; do
  [ -s gpurun_out/f_bench_$w.json ] && tail -1 gpurun_out/f_bench_$w.json > profiles/r2_bench_$w.json
done
for w in c2 c2tagged c3 c4 c4cmp; do
  [ -s gpurun_out/f_launches_$w.csv ] && cp gpurun_out/f_launches_$w.csv profiles/r2_launches_$w.csv
done
[ -s gpurun_out/prof_r2_final_raw_c2.ncu-rep ] && python scripts/ncu_summary.py gpurun_out/prof_r2_final_raw_c2.ncu-rep "rans_raw_fused_kernel (software-pipelined lean loop), python bench.py (c2), final state of round 2" > profiles/r2_final_rans_raw_fused_c2_ncu_summary.txt
[ -s gpurun_out/prof_r2_final_cmp_chain.ncu-rep ] && python scripts/ncu_summary.py gpurun_out/prof_r2_final_cmp_chain.ncu-rep "cmp_chain_kernel<3> (two warps per stream), python bench.py --workload c4cmp, final state of round 2" > profiles/r2_cmp_chain_c4cmp_ncu_summary.txt
[ -s gpurun_out/prof_r2_final_geo_normal.ncu-rep ] && python scripts/ncu_summary.py gpurun_out/prof_r2_final_geo_normal.ncu-rep "geo_normal_kernel (point-parallel geometric-normal predictor), python bench.py --workload c4cmp, final state of round 2" > profiles/r2_geo_normal_c4cmp_ncu_summary.txt
[ -s gpurun_out/prof_r2_cmp_chain_c4cmp.ncu-rep ] && python scripts/ncu_summary.py gpurun_out/prof_r2_cmp_chain_c4cmp.ncu-rep "cmp_chain_kernel<3>, FIRST version (one warp, 12 operand loads per entry), c4cmp" > profiles/r2_cmp_chain_v1_c4cmp_ncu_summary.txt
[ -s gpurun_out/prof_r2_cmp_chain_v2_c4cmp.ncu-rep ] && python scripts/ncu_summary.py gpurun_out/prof_r2_cmp_chain_v2_c4cmp.ncu-rep "cmp_chain_kernel<3>, SECOND version (one warp, operands pre-summed by the builder lanes), c4cmp" > profiles/r2_cmp_chain_v2_c4cmp_ncu_summary.txt
ls -la profiles | tail -20
