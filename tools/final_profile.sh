#!/bin/bash
# Round-end measurement set, run on the GPU box from the repo root:  tools/final_profile.sh <tag>
# Writes gpurun_out/<tag>_*: bench lines (ours + reference arm), ncu launch list, one full-set capture per kernel, micro-benchmarks.
tag=${1:-r02}
out=gpurun_out
mkdir -p $out
python bench.py > $out/${tag}_bench_ours.json 2> $out/${tag}_bench_ours.err
python bench.py --impl reference > $out/${tag}_bench_reference.json 2> $out/${tag}_bench_reference.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $out/${tag}_launches_bench.csv \
    python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-fused --parity-sample 1 > $out/${tag}_bench_under_ncu.log 2>&1
# one launch of each kernel of a step (scalefactors, big_values, count1, granule), then the tolerance-mode granule kernel
ncu --set full --clock-control none --import-source on -k regex:l3_ -c 4 -o $out/${tag}_kernels \
    python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --no-fused --parity-sample 1 > $out/${tag}_prof.log 2>&1
ncu --set full --clock-control none -k regex:l3_granule -s 4 -c 1 -o $out/${tag}_granule_fused \
    python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --parity-sample 1 > $out/${tag}_prof_fused.log 2>&1
tools/microbench/fp32_pipes > $out/${tag}_fp32_pipes.txt 2>&1
tools/microbench/dct32_tc > $out/${tag}_dct32_tc.txt 2>&1
tail -c 600 $out/${tag}_bench_ours.json; echo; tail -c 400 $out/${tag}_bench_reference.json
