#!/bin/bash
# One GPU-box round: parity tests, bench line, ncu launch list of the bench command, ncu --set full of the nn=64 edge kernel.
# usage (from the repo root, under gpurun):  bash profiles/gpu_round.sh <tag> [skip-tests|tests] [no-ncu]
tag=${1:-rX}
out=gpurun_out
mkdir -p $out
if [ "$2" != "skip-tests" ]; then
  PESTO_TC_DEBUG=1 timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> $out/${tag}_pytest.log
  tail -3 $out/${tag}_pytest.log
fi
timeout 600 python bench.py --steps 10 --warmup 3 > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench exit $?"
cat $out/${tag}_bench.json
if [ "$3" == "no-ncu" ]; then exit 0; fi
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $out/${tag}_launches.log 2>&1; echo "ncu launches exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:edge_kernel_tc -c 1 -s 31 -f -o $out/${tag}_edge64 \
    python profiles/run_forward.py --atoms 32768 --mode f16x3 > $out/${tag}_ncu_full.log 2>&1; echo "ncu full exit $?"
ncu -i $out/${tag}_edge64.ncu-rep --page details > $out/${tag}_edge64_details.txt 2>&1
