#!/bin/bash
# round 2, call Z2: geometric-normal kernels -- new GPU tests first, then the whole GPU suite
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
echo "== geo tests"; timeout 600 python -m pytest tests/test_gpu_mesh.py -m gpu -x -q -k "geometric or constrained" > gpurun_out/z2_geo.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/z2_geo.log
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/z2_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/z2_pytest.log
