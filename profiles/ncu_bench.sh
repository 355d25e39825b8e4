#!/bin/bash
# ncu evidence for the bench command itself (run under gpurun from the repo root):
#   1. launch list (gpu__time_duration.sum) of `bench.py --steps 1 --warmup 3`
#   2. one --set full capture of the nn = 64 fused edge kernel on the bench workload (132 417 atoms): the 25th
#      edge-kernel launch of the first forward (layers 24..31 have nn = 64)
# usage: bash profiles/ncu_bench.sh <tag>
tag=${1:-rX}
out=gpurun_out
mkdir -p $out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $out/${tag}_launches.log 2>&1; echo "ncu launches exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:edge_kernel_tc -c 1 -s 27 -f -o $out/${tag}_edge64_bench \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $out/${tag}_ncu_full.log 2>&1; echo "ncu full exit $?"
ncu -i $out/${tag}_edge64_bench.ncu-rep --page details > $out/${tag}_edge64_bench_details.txt 2>&1
ncu -i $out/${tag}_edge64_bench.ncu-rep --page raw --csv > $out/${tag}_edge64_bench_raw.csv 2>&1
