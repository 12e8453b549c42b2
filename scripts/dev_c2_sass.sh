#!/bin/bash
# development helper: compile only the headline instantiation of rans_raw_fused_kernel and run the issue model on its loops
cd /root/repo/draco_sharp_b200 || exit 1
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -DDCB_DEV_ONLY_C2 "$@" -c -o /tmp/dev_c2.o csrc/dcb_kernels.cu || exit 1
SASS_OUT=/tmp/dev_c2.sass python ../scripts/sass_loops.py /tmp/dev_c2.o rans_raw_fused_kernelILi3EtLb0ELb0ELi1ELi2E 300 12
