#!/bin/bash
# Round-2 ncu evidence (run under gpurun from the repo root):  bash profiles/ncu_r2.sh <tag>
#   1. launch list (gpu__time_duration.sum) of the bench command
#   2. --set full of the nn = 64 fused edge kernel on the bench workload (28th edge-kernel launch = layer 27 of the first forward)
#   3. --set full of the same kernel at the north_star size (one 8192-atom structure) and of the per-atom kernel
tag=${1:-r2}
out=gpurun_out
mkdir -p $out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > $out/${tag}_launches.log 2>&1; echo "ncu launches exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:edge_kernel_tc -c 1 -s 27 -f -o $out/${tag}_edge64_bench \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > $out/${tag}_ncu_full.log 2>&1; echo "ncu full (bench) exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:edge_kernel_tc -c 1 -s 27 -f -o $out/${tag}_edge64_n8192 \
    python profiles/run_forward.py --atoms 8192 --mode f16x3 > $out/${tag}_ncu_full_8192.log 2>&1; echo "ncu full (8192) exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:node_umma_kernel -c 1 -s 27 -f -o $out/${tag}_node_bench \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > $out/${tag}_ncu_node.log 2>&1; echo "ncu full (node) exit $?"
for f in edge64_bench edge64_n8192 node_bench; do
  ncu -i $out/${tag}_$f.ncu-rep --page details > $out/${tag}_${f}_details.txt 2>&1
  ncu -i $out/${tag}_$f.ncu-rep --page raw --csv > $out/${tag}_${f}_raw.csv 2>&1
done
