#!/bin/bash
# round 2, call Z1: constrained multi-parallelogram kernels -- new GPU tests first, then the whole GPU suite, then c2 sanity
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
echo "== cmp tests"; timeout 600 python -m pytest tests/test_gpu_mesh.py -m gpu -x -q -k "constrained" > gpurun_out/z1_cmp.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/z1_cmp.log
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/z1_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/z1_pytest.log
echo "== c2"; timeout 600 python bench.py --steps 5 --warmup 3 --e2e-steps 0 --no-cpu-baseline > gpurun_out/z1_c2.json 2> gpurun_out/z1_c2.err; echo " rc=$?"; tail -c 1500 gpurun_out/z1_c2.json
