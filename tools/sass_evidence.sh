#!/bin/sh
# Static evidence that the built library uses the Blackwell tensor / TMA paths: per-kernel counts of the SASS mnemonics
# /opt/skills/guides/B200_PROFILING.md lists (tcgen05.mma -> UTC*MMA, tcgen05.ld -> LDTM, tcgen05.commit -> UTCBAR,
# cp.async.bulk.tensor -> UTMALDG / UTMASTG, cp.async.bulk -> UBLKCP, mbarrier -> SYNCS, cp.async -> LDGSTS).
#   sh tools/sass_evidence.sh > profiles/r1_sass_evidence.txt
SO=${1:-dcase2019_task4_b200/csrc/libdcase_b200.so}
echo "# cuobjdump -sass $SO : per-kernel mnemonic counts (sm_100a)"
cuobjdump -sass "$SO" 2>/dev/null | awk '
/Function :/ { fn = $3 }
{
  for (i = 1; i <= NF; i++)
    if ($i ~ /^(UTC[A-Z]*MMA|UTCBAR|UTCATOMSWS|LDTM|STTM|UTMALDG|UTMASTG|UTMACMDFLUSH|UBLKCP|SYNCS|LDGSTS|REDG|ATOMG)/) {
      split($i, a, "."); c[fn " " a[1]]++
    }
}
END { for (k in c) print k, c[k] }' | sort | c++filt
