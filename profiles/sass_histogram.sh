#!/usr/bin/env bash
# SASS mnemonic histogram of the shipped library (the instructions that prove TMA / tcgen05 / cluster use; B200_PROFILING.md).
# Usage: bash profiles/sass_histogram.sh > profiles/r02/sass_histogram.txt
LIB="${1:-cdv-slam_b200/lib/libpgba.so}"
S=$(mktemp)
cuobjdump -sass "$LIB" > "$S"
echo "# $LIB  ($(stat -c %s "$LIB") bytes), cuobjdump -sass, $(date -u +%F)"
echo "# per-kernel counts of the Blackwell-specific instructions"
awk '/Function :/ {fn=$3} /UTMALDG|UTMASTG|UTCMMA|UTCHMMA|UTCQMMA|UTCOMMA|UTCIMMA|LDTM|STTM|UTCBAR|UTCATOM|SYNCS|UCGABAR|FFMA2|HMMA|DFMA|LDSM|REDG|RED\./ {
  m=$0; sub(/^[^A-Z@]*/, "", m); split(m, a, /[ ;]/); op=a[1]; if (op ~ /^@/) op=a[2]; gsub(/\..*/, "", op); c[fn" "op]++ }
  END { for (k in c) print c[k], k }' "$S" | sort -k2,2 -k3,3 | awk '{printf "%-90s %-10s %6d\n", $2, $3, $1}'
echo
echo "# library totals"
for op in UTMALDG UTMASTG UTCMMA UTCHMMA UTCQMMA LDTM STTM UTCBAR SYNCS UCGABAR FFMA2 HMMA DFMA LDSM; do
  printf "%-10s %7d\n" $op "$(grep -c "[ @]$op" "$S")"
done
rm -f "$S"
