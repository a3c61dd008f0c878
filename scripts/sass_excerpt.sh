#!/bin/bash
# Counts the TMA / mbarrier / bulk-copy / cluster instructions per kernel in the built library (evidence for profiles/sass_tma_excerpt.txt).
cd "$(dirname "$0")/.."
echo "# cuobjdump -sass phiml_b200/lib/libphb200.so | per-kernel counts of UTMALDG (TMA tensor loads), UBLKCP (cp.async.bulk 1-D), SYNCS (mbarrier),"
echo "# LDGSTS (cp.async), cluster / DSMEM instructions (UCGABAR, LDS with .CLUSTER / MAPA).  Regenerate: bash scripts/sass_excerpt.sh > profiles/sass_tma_excerpt.txt"
cuobjdump -sass phiml_b200/lib/libphb200.so | awk '
/Function :/ {fn=$3}
/UTMALDG|UBLKCP|SYNCS\.|LDGSTS|UCGABAR|MAPA|MEMBAR\.ALL\.SYS|ERRBAR/ {
  op=$0; sub(/^[ \t]*\/\*[0-9a-f]+\*\/[ \t]*/,"",op); n=split(op,a," "); m=a[1]; if (m ~ /^@/) m=a[2]; c[fn" "m]++ }
END {for (k in c) print c[k], k}' | sort -k2,2 -k3,3 | while read cnt fn op; do printf "%5d  %-34s %s\n" "$cnt" "$op" "$(echo "$fn" | c++filt | cut -c1-150)"; done
