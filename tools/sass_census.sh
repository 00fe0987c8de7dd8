#!/bin/bash
# SASS opcode census of the shipped library (no GPU needed): per kernel, how many tcgen05 MMAs (UTCHMMA[.2CTA]), TMEM loads (LDTM),
# TMA loads / stores (UTMALDG / UTMASTG), tcgen05 commits (UTCBAR), mbarrier waits (SYNCS) and legacy tensor-core ops (HMMA: none expected).
#   usage: tools/sass_census.sh [lib] > profiles/r02_sass_census.txt
LIB="${1:-hulk_keypoints_b200/libhulk_sm100.so}"
echo "# cuobjdump -sass $LIB  ($(date -u +%Y-%m-%dT%H:%MZ), $(nvcc --version | grep release | sed 's/.*release //'))"
echo "# columns: UTCHMMA UTCHMMA.2CTA LDTM UTMALDG UTMASTG UTCBAR SYNCS HMMA instructions kernel"
cuobjdump -sass "$LIB" | awk '
  function flush() { if (name != "") printf "%7d %12d %5d %7d %7d %6d %5d %4d %12d %s\n", mma, mma2, ldtm, ldg, stg, bar, syncs, hmma, n, name }
  /Function :/ { flush(); name = $3; mma = mma2 = ldtm = ldg = stg = bar = syncs = hmma = n = 0; next }
  /^ *\/\*[0-9a-f][0-9a-f]*\*\// {
    n++
    if ($0 ~ /UTCHMMA\.2CTA/) mma2++; else if ($0 ~ /UTCHMMA/) mma++
    if ($0 ~ /LDTM/) ldtm++
    if ($0 ~ /UTMALDG/) ldg++
    if ($0 ~ /UTMASTG/) stg++
    if ($0 ~ /UTCBAR/) bar++
    if ($0 ~ /SYNCS/) syncs++
    if ($0 ~ / HMMA/) hmma++
  }
  END { flush() }' | sort -k10 | c++filt | cut -c1-220
