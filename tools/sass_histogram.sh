#!/usr/bin/env bash
# Opcode histogram of the shipped library (per kernel family): the SASS evidence for tcgen05 / TMEM / TMA use.
#   bash tools/sass_histogram.sh > profiles/r02_sass_opcode_histogram.txt      (no GPU needed)
set -euo pipefail
LIB=alignn_b200/libalignn_b200.so
echo "# cuobjdump -sass $LIB : occurrences of the Blackwell-specific opcodes, whole library"
cuobjdump -sass "$LIB" | grep -oE "\b(UTCHMMA|UTCQMMA|UTCCP|LDTM|STTM|UTMALDG|UTMASTG|UTMAPF|UBLKCP|UBLKPF|UTCBAR|UTCATOMSWS|SYNCS|HMMA|LDGSTS|UCGABAR_ARV|UCGABAR_WAIT|MAPA)[A-Z0-9_.]*" | sed 's/\..*//' | sort | uniq -c | sort -rn
echo
echo "# per kernel (function name : opcode counts)"
cuobjdump -sass "$LIB" | awk '
/Function :/ { fn=$3 }
{ if (match($0, /(UTCHMMA|LDTM|STTM|UTMALDG|UTMASTG|UTMAPF|UBLKCP|UBLKPF|UCGABAR_ARV|MAPA)/)) { k=substr($0, RSTART, RLENGTH); c[fn" "k]++ } }
END { for (x in c) print c[x], x }' | sort -k2,2 -k3,3 | awk '{ printf "%6d  %-14s %s\n", $1, $3, $2 }' | c++filt | cut -c1-200
