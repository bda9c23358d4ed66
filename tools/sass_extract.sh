#!/bin/bash
# SASS evidence of the shipped library: instruction mix over all kernels and per-kernel counts of the instructions that
# matter here (DMMA.8x8x4 = the FP64 tensor path, LDGSTS = cp.async, SYNCS = mbarrier, ATOMG + .STRONG.GPU = the stream-K
# rank counter and partial-tile flags).  Usage: bash tools/sass_extract.sh > profiles/<round>_sass_extract.txt
LIB=tensororder_b200/csrc/libtob200.so
cuobjdump -sass "$LIB" > /tmp/tob_sass.txt
echo "SASS extract of $LIB (cuobjdump -sass, sm_100a); build id (bench.py lib_build_id): $(python -c 'import bench; print(bench.lib_build_id())' 2>/dev/null)"
echo
echo "instruction counts over all kernels:"
grep -oE "^\s+/\*[0-9a-f]{4}\*/\s+(@!?U?P[0-9T] )?[A-Z0-9_.]+" /tmp/tob_sass.txt | awk '{print $NF}' | sed 's/\..*//' | sort | uniq -c | sort -rn | \
  grep -E " (DMMA|DFMA|DADD|DMUL|LDG|STG|LDS|STS|LDGSTS|SHFL|BAR|SYNCS|ARRIVES|ATOMG|MEMBAR|LDGDEPBAR|DEPBAR|UTMALDG|UTMASTG|UBLKCP|UTCHMMA|NANOSLEEP)$"
echo
echo "per kernel (function name, DMMA.8x8x4, LDGSTS, SYNCS (mbarrier), BAR.SYNC, ATOMG, .STRONG.GPU accesses = ld.cg, ld.acquire, st.release):"
awk '/Function : /{name=$3} /DMMA/{d[name]++} /LDGSTS/{l[name]++} /SYNCS/{s[name]++} /BAR\.SYNC/{b[name]++} /ATOMG/{a[name]++} /\.STRONG\.GPU/{f[name]++} /Function : /{seen[name]=1}
     END{for (n in seen) printf "%s DMMA %d LDGSTS %d SYNCS %d BAR %d ATOMG %d STRONG.GPU %d\n", n, d[n], l[n], s[n], b[n], a[n], f[n]}' /tmp/tob_sass.txt | c++filt | sort | awk '{print}' | grep -v "DMMA 0 LDGSTS 0 SYNCS 0 BAR [0-9]* ATOMG 0 STRONG.GPU 0" 
