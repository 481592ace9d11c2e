#!/bin/sh
# Per-kernel counts of the SASS mnemonics that prove a Blackwell-native kernel (B200_PROFILING.md): UTC*MMA = tcgen05.mma,
# LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = TMA loads/stores, UTCBAR = tcgen05.commit, HMMA = legacy mma.sync (expected 0).
# usage: tools/sass_counts.sh > profiles/rNN_sass_counts.txt
so="$(dirname "$0")/../owl_vit_object_detection_b200/lib/libowl_b200.so"
echo "# cuobjdump -sass $(basename "$so") : instruction counts per kernel (kernels without any of them are omitted)"
cuobjdump -sass "$so" | awk '
  /Function :/ { name=$3 }
  /UTC[A-Z]*MMA/ { mma[name]++ } /LDTM/ { ldtm[name]++ } /STTM/ { sttm[name]++ }
  /UTMALDG/ { tmal[name]++ } /UTMASTG/ { tmas[name]++ } /UTCBAR/ { bar[name]++ } / HMMA/ { hmma[name]++ }
  END { for (n in mma) printf "%-6d UTCxMMA %-5d LDTM %-5d STTM %-5d UTMALDG %-5d UTMASTG %-5d UTCBAR %-3d HMMA  %s\n", mma[n]+0, ldtm[n]+0, sttm[n]+0, tmal[n]+0, tmas[n]+0, bar[n]+0, hmma[n]+0, n }' | sort -k13 | while read line; do
    name=$(echo "$line" | awk '{print $NF}'); dem=$(echo "$name" | c++filt | cut -c1-110); echo "$line" | sed "s|$name|$dem|"; done
echo "# totals:"
cuobjdump -sass "$so" | grep -o "UTC[A-Z]*MMA[.A-Z0-9_]*\|LDTM[.A-Z0-9_]*\|STTM[.A-Z0-9_]*\|UTMALDG[.A-Z0-9_]*\|UTMASTG[.A-Z0-9_]*\|UTCBAR[.A-Z0-9_]*\| HMMA[.A-Z0-9_]*" | sort | uniq -c | sort -rn
