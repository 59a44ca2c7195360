#!/bin/bash
# build a tuning variant of the library: tools/build_variant.sh NAME "-DIBS_W16=16 ..."   -> lib/variants/libibs_NAME.so
set -e
cd "$(dirname "$0")/../ideal-ballooning-solver_b200"
mkdir -p lib/variants
F="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -DIBS_BUILD"
nvcc $F $2 -Xptxas=-v -c csrc/ibs_solver.cu -o lib/variants/solver_$1.o 2> lib/variants/ptxas_$1.log
nvcc -shared -o lib/variants/libibs_$1.so lib/ibs_api.o lib/variants/solver_$1.o lib/ibs_geometry.o lib/ibs_geometry_full.o lib/ibs_adjoint.o -gencode arch=compute_100a,code=sm_100a -cudart static
rm -f lib/variants/solver_$1.o
grep -A2 "solve_kernelILi$3ELi$4ELi2ELb0" lib/variants/ptxas_$1.log | tail -2
