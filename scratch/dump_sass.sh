#!/bin/bash
# SASS listings of the kernels the profiles talk about (cuobjdump of the in-tree library)
SO=tgm_b200/csrc/libtgm_b200.so
dump() {  # $1 = substring of the mangled name, $2 = output file
  f=$(cuobjdump -sass $SO 2>/dev/null | grep "Function : " | grep "$1" | head -1 | sed 's/.*Function : //')
  { echo "# cuobjdump -sass -fun '$f' $SO  (encodings stripped)"; cuobjdump -sass -fun "$f" $SO 2>/dev/null | grep -v '^\s*/\* 0x' ; } > "$2"
  echo "$2: $(wc -l < $2) lines; $(grep -o 'UBLKCP[.A-Z0-9]*\|SYNCS[.A-Z0-9]*\|UTCHMMA[.A-Z0-9]*\|UTCBAR[.A-Z0-9]*\|LDTM[.A-Zx0-9]*\|UTMALDG[.A-Z0-9]*' $2 | sort | uniq -c | tr '\n' ' ')"
}
dump 'csr_sample_tma_kernelILb1E' profiles/r2_sass_csr_sample_tma_kernel.txt
dump 'tc3_linear_kernel' profiles/r2_sass_tc3_linear_kernel.txt
dump 'attn_warp_kernelILi1ELi6ELi4ELb1E' profiles/r2_sass_attn_warp_kernel.txt
dump 'ring_query_tma_kernel' profiles/r2_sass_ring_query_tma_kernel.txt
dump 'frontier_compact_kernelILb1E' profiles/r2_sass_frontier_compact_kernel.txt
