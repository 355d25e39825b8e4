#!/bin/bash
# compute-sanitizer over one small forward (profiles/run_forward.py, 700 atoms, all four nn variants of the tensor-core
# edge kernel + the tcgen05 per-atom kernel): memcheck, then racecheck (shared-memory hazards between the column groups /
# halves), then synccheck.   usage (under gpurun): bash profiles/sanitize.sh <tag>
tag=${1:-rX}
out=gpurun_out
mkdir -p $out
for tool in memcheck racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 20 python profiles/run_forward.py --atoms 700 --mode f16x3 \
      > $out/${tag}_sanitize_$tool.log 2>&1
  echo "$tool exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|^ok" $out/${tag}_sanitize_$tool.log | tail -3
done
