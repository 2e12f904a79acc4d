#!/bin/bash
# Counts of the Blackwell-only SASS mnemonics per kernel of the shipped library (cuobjdump -sass):
# UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / .st (tensor memory), UTMALDG = TMA tensor loads, UTCBAR = tcgen05.commit
set -e
LIB=${1:-whisper-mlx_b200/csrc/libb200whisper.so}
echo "# cuobjdump -sass $LIB  ($(date -u +%Y-%m-%dT%H:%MZ), $(sha256sum $LIB | cut -c1-16))"
cuobjdump -sass "$LIB" | awk '
/Function :/ { fn=$3 }
/UTCHMMA/ { a[fn]++; if ($0 ~ /2CTA/) a2[fn]++ }
/LDTM/ { b[fn]++ }
/STTM/ { c[fn]++ }
/UTMALDG/ { d[fn]++; if ($0 ~ /MULTICAST/) d2[fn]++ }
/UTCBAR/ { e[fn]++ }
/SYNCS/ { f[fn]++ }
END {
  printf "%-8s %-6s %-5s %-5s %-8s %-6s %-7s %-6s  %s\n", "UTCHMMA", ".2CTA", "LDTM", "STTM", "UTMALDG", ".MCAST", "UTCBAR", "SYNCS", "kernel"
  for (k in f) all[k]=1; for (k in a) all[k]=1; for (k in d) all[k]=1
  for (k in all) printf "%-8d %-6d %-5d %-5d %-8d %-6d %-7d %-6d  %s\n", a[k], a2[k], b[k], c[k], d[k], d2[k], e[k], f[k], k
}' | (read -r h; echo "$h"; sort -k9)
echo "# totals"
cuobjdump -sass "$LIB" | grep -c UTCHMMA | sed 's/^/UTCHMMA /'
cuobjdump -sass "$LIB" | grep -c LDTM | sed 's/^/LDTM /'
cuobjdump -sass "$LIB" | grep -c STTM | sed 's/^/STTM /'
cuobjdump -sass "$LIB" | grep -c UTMALDG | sed 's/^/UTMALDG /'
