#!/bin/bash
# round 2, call W: GPU test suite with the wide-attribute (nc > 4) cases
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/w_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/w_pytest.log
