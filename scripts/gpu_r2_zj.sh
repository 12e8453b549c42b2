#!/bin/bash
# round 2, call ZJ: final bench lines of the mesh workloads after the staged maps upload
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for w in c4 c4tagged c4cmp; do
  echo "== $w"; timeout 600 python bench.py --workload $w --steps 10 --warmup 3 > gpurun_out/j_bench_$w.json 2> gpurun_out/j_bench_$w.err; echo " rc=$?"; tail -c 300 gpurun_out/j_bench_$w.json | head -c 300; echo
done
