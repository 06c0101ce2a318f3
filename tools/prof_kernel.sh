#!/bin/bash
# usage: tools/prof_kernel.sh <tag> <lib.so | default> <kernel regex>  ->  gpurun_out/<tag>.ncu-rep: one full-set capture of one kernel of a bench step
tag=$1; lib=$2; k=$3
if [ "$lib" != "default" ]; then export L3B_LIB=$PWD/$lib; fi
ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -o gpurun_out/$tag -f \
    python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --no-fused --parity-sample 1 > gpurun_out/${tag}_prof.log 2>&1
