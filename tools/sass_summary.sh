#!/bin/bash
# SASS instruction summary of the shipped library (cuobjdump runs without a GPU): which tcgen05 / TMA / TMEM /
# 256-bit memory instructions the kernels really contain. Usage: tools/sass_summary.sh > profiles/r02_sass_summary.txt
SO=${1:-pytorch_ddp_resnet_b200/libb200resnet.so}
T=$(mktemp)
cuobjdump -sass "$SO" > "$T" 2>/dev/null
echo "# SASS summary of $SO ($(stat -c %s "$SO") bytes), $(cuobjdump -lelf "$SO" | head -1 | sed 's/.*: //')"
echo "# kernels (entry functions): $(grep -c 'Function :' "$T")"
echo
echo "## tensor core (tcgen05.mma -> UTCHMMA), TMEM (tcgen05.ld -> LDTM), commit / barriers (UTCBAR, SYNCS)"
for p in 'UTCHMMA[.A-Z0-9]*' 'UTCQMMA[.A-Z0-9]*' 'LDTM[.A-Z0-9x]*' 'UTCBAR[.A-Z0-9]*' 'UTCATOMSWS[.A-Z0-9]*' 'SYNCS[.A-Z0-9]*'; do
  grep -o -- "$p" "$T" | sort | uniq -c
done
echo
echo "## TMA (cp.async.bulk.tensor -> UTMALDG; no UTMASTG: epilogues store from registers)"
grep -o 'UTMA[A-Z]*[.A-Z0-9]*' "$T" | sort | uniq -c
echo
echo "## global memory: 256-bit and 128-bit accesses, atomics"
for p in 'LDG\.[.A-Za-z0-9]*256[.A-Za-z0-9]*' 'STG\.[.A-Za-z0-9]*256[.A-Za-z0-9]*' 'LDG\.E[.A-Z]*\.128[.A-Z]*' 'STG\.E[.A-Z]*\.128' 'ATOMG\.[.A-Z0-9]*' 'RED\.[.A-Z0-9]*' 'ATOMS\.[.A-Z0-9]*'; do
  grep -o -- "$p" "$T" | sort | uniq -c
done
echo
echo "## legacy tensor instructions (must be zero: no mma.sync / wgmma fallbacks)"
echo "HMMA $(grep -c ' HMMA' "$T")   IMMA $(grep -c ' IMMA' "$T")   WGMMA $(grep -c 'WGMMA' "$T")"
echo
echo "## per-kernel UTCHMMA / UTMALDG / LDTM counts (kernels that contain any)"
awk '/Function :/ {name=$3} /UTCHMMA/ {m[name]++} /UTMALDG/ {t[name]++} /LDTM/ {l[name]++} END {for (k in m) printf "%5d %5d %5d  %s\n", m[k], t[k], l[k], k}' "$T" | sort -k4 | while read a b c n; do printf "%5s %5s %5s  %s\n" "$a" "$b" "$c" "$(echo "$n" | c++filt | sed -E 's/\(.*//' | cut -c1-90)"; done
rm -f "$T"
