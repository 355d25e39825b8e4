#!/bin/bash
# SASS opcode histograms of the hot kernels in the built library (no GPU needed):  bash profiles/tools/sass_hist.sh > profiles/r2_sass_histograms.md
lib=pesto_b200/libpesto_b200.so
echo "# SASS opcode histograms (cuobjdump -sass $lib, round 2)"
echo
echo "Tensor-core / TMA mnemonics over the whole library: $(cuobjdump -sass $lib | grep -oE 'UTCHMMA|UTCBAR|LDTM|STTM|UBLKCP|UTMALDG[.A-Z0-9]*|UTCATOMSWS|SYNCS[.A-Z0-9]*' | sort | uniq -c | awk '{printf "%s x %s, ", $2, $1}')"
echo
for pat in 'edge_kernel_tcILi64ELb1' 'edge_kernel_tcILi8ELb1' 'node_umma_kernelILb1ELb1ELb1' 'knn_main_kernel'; do
  fn=$(cuobjdump -sass $lib | grep -oE "Function : [A-Za-z0-9_]*${pat}[A-Za-z0-9_]*" | head -1 | sed 's/Function : //')
  [ -z "$fn" ] && continue
  echo "## \`$(echo $fn | c++filt | cut -c1-110)\`"
  echo
  cuobjdump -sass -fun "$fn" $lib | grep -E '^\s+/\*[0-9a-f]{4}\*/' | sed -E 's/^\s+\/\*[0-9a-f]+\*\/\s+//' | sed -E 's/^@!?U?P[0-9T]+\s+//' | awk '{print $1}' | sed -E 's/;$//' \
    | awk -F. '{print $1}' | sort | uniq -c | sort -rn | awk 'BEGIN{printf "| opcode | count |\n|---|---|\n"} {t+=$1; if (NR<=28) printf "| %s | %d |\n", $2, $1; else r+=$1} END{printf "| (other) | %d |\n| **total** | **%d** |\n", r, t}'
  echo
done
