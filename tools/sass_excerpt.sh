#!/bin/bash
# SASS evidence for profiles/: which instructions the hot kernels are made of (run here, no GPU needed): bash tools/sass_excerpt.sh > profiles/r2_sass_excerpt.txt
cd "$(dirname "$0")/.."
L=bmsparse_spgemm_spmv_b200/lib/libbmsparse_b200.so
PAT='\b(UBLKCP[.A-Z0-9]*|SYNCS[.A-Z0-9_]*|LDGSTS[.A-Z0-9]*|ATOMG[.A-Z0-9]*|ATOMS[.A-Z0-9]*|LDGDEPBAR|DEPBAR[.A-Z]*|BAR[.A-Z]*|STG[.A-Z0-9]*|LDG[.A-Z0-9]*|LDS[.A-Z0-9]*|STS[.A-Z0-9]*|FFMA|HMMA[.A-Z0-9]*|MEMBAR[.A-Z]*|CCTL[.A-Z]*|RED[.A-Z0-9]*|ST\.E[.A-Z0-9]*|LD\.E[.A-Z0-9]*|SHFL[.A-Z]*|POPC|VOTE[.A-Z]*|MATCH[.A-Z]*)\b'
echo "cuobjdump -sass of $L (sm_100a, built by make -C bmsparse_spgemm_spmv_b200/csrc): instruction mnemonics per kernel (count, mnemonic)"
echo "UBLKCP = cp.async.bulk (TMA engine, 1-D bulk copy), SYNCS.* = mbarrier ops, LDGSTS = cp.async, HMMA = mma.sync tensor-core op"
for f in \
  '_ZN4bmsp18spmv_stream_kernelI6__halffLi64ELi1ELi7ENS_6NoHaloEEEvNS_10StreamArgsIT_EEPKT0_PfT4_' \
  '_ZN4bmsp18spmv_stream_kernelI6__halffLi64ELi1ELi7ENS_7HaloDevEEEvNS_10StreamArgsIT_EEPKT0_PfT4_' \
  $(cuobjdump -elf $L 2>/dev/null | grep -oE '_ZN4bmsp19spgemm_dense_kernelILi4EE[A-Za-z0-9_]*' | sort -u | head -1) \
  $(cuobjdump -elf $L 2>/dev/null | grep -oE '_ZN4bmsp18spgemm_pass_kernelILi2ELi1024EE[A-Za-z0-9_]*' | sort -u | head -1) ; do
  echo; echo "== $(echo $f | c++filt)"
  cuobjdump -sass -fun "$f" $L 2>/dev/null > /tmp/_k.sass
  echo "   $(grep -cE '^\s+/\*[0-9a-f]{4}\*/' /tmp/_k.sass) instructions"
  grep -oE "$PAT" /tmp/_k.sass | sort | uniq -c | sort -rn | awk '{printf "   %6d %s\n", $1, $2}'
done
echo; echo "== the producer's copy sequence of the streaming kernel (first UBLKCP block, verbatim)"
cuobjdump -sass -fun '_ZN4bmsp18spmv_stream_kernelI6__halffLi64ELi1ELi7ENS_6NoHaloEEEvNS_10StreamArgsIT_EEPKT0_PfT4_' $L 2>/dev/null | grep -E "UBLKCP|SYNCS|LDGSTS|ATOMG" | sed 's/^ *//' | cut -c1-120
