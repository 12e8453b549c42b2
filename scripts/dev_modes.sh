#!/bin/bash
# development helper: rebuild dcb_kernels.o and print the issue model of the main loops of the hot instantiations
python -c "
from draco_sharp_b200 import build as B
B.build_lib()
" || exit 1
for spec in "3 1 12" "3 2 12" "2 3 8" "4 2 16" "2 1 8" "3 4 12" "2 4 8"; do
  set -- $spec
  echo -n "NCP=$1 MODE=$2: "; python scripts/sass_loops.py draco_sharp_b200/build/dcb_kernels.o rans_raw_fused_kernelILi${1}EtLb0ELb0ELi${2}ELi2E 200 $3 | grep loop | sort -t, -k1 | head -2 | tr '\n' ' '; echo
done
