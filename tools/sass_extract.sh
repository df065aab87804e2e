#!/bin/bash
# Evidence of the Blackwell-native instructions in the shipped library: per kernel, the counts of the SASS
# mnemonics B200_PROFILING.md lists (tcgen05.mma -> UTC*MMA, tcgen05.ld -> LDTM, TMA -> UTMALDG/UTMASTG/UBLKCP,
# tcgen05.commit -> UTCBAR) and of float atomics (RED/ATOM .F32).   usage: tools/sass_extract.sh > profiles/rN_sass_extract.txt
cd "$(dirname "$0")/.."
LIB=hs-pose_b200/lib/libhspose_b200.so
echo "# cuobjdump -sass $LIB  ($(date -u +%F), nvcc $(nvcc --version | grep -o 'release [0-9.]*'))"
cuobjdump -sass "$LIB" | awk '
  /Function :/ { fn=$3 }
  { if (match($0, /(UTC[A-Z]*MMA[.A-Z0-9_]*|LDTM[.A-Za-z0-9_]*|STTM[.A-Za-z0-9_]*|UTMALDG[.A-Z0-9_]*|UTMASTG[.A-Z0-9_]*|UBLKCP[.A-Z0-9_]*|UTCBAR[.A-Z0-9_]*|UTCATOMSWS[.A-Z0-9_]*|REDG?\.E\.ADD\.F32[.A-Z0-9_]*|ATOMG?\.E\.ADD\.F32[.A-Z0-9_]*|HMMA[.A-Z0-9_]*|FFMA2|FADD2|FMUL2)/)) {
      k = substr($0, RSTART, RLENGTH); c[fn "\t" k]++ } }
  END { for (x in c) print c[x] "\t" x }' | sort -t$'\t' -k2,2 -k3,3 | while IFS=$'\t' read n fn k; do printf "%6d  %-28s %s\n" "$n" "$k" "$(echo $fn | c++filt | cut -c1-110)"; done
