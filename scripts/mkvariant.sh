#!/bin/bash
# Development tool: scripts/mkvariant.sh name [-D flags...]
#   scratch/lib_<name>.so = the library with bqa_fast_d3D4.cu compiled with the flags (e.g. -DBQA_BP_WARPS=12
#   -DBQA_BP_MSG_PITCH=128); time several builds against each other with scripts/compare_bp_variants.py.
#   Needs bqa_b200/build/*.o (python -m bqa_b200.build).
set -e
cd "$(dirname "$0")/.."
mkdir -p scratch
name=$1; shift
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC "$@" -Xptxas -v -c bqa_b200/csrc/bqa_fast_d3D4.cu -o scratch/fast_$name.o 2> scratch/fast_$name.log
objs=$(ls bqa_b200/build/*.o | grep -v bqa_fast_d3D4.o)
nvcc -shared -o scratch/lib_$name.so scratch/fast_$name.o $objs 2>/dev/null
grep -A2 "k_bp_run" scratch/fast_$name.log | tail -2
