#!/bin/bash
# SASS evidence for profiles/: per kernel of libtermgpu.so (sm_100a cubins), how often the Blackwell / async mnemonics occur:
# UBLKCP (cp.async.bulk: TMA 1-D bulk copies), SYNCS.* (mbarrier), MATCH.ANY (warp match: radix ranking, group counting),
# ATOMS / RED (shared / global atomics), plus the absence of library kernels. No GPU needed.
LIB=${1:-term_b200/libtermgpu.so}
echo "# cuobjdump -sass $LIB ($(cuobjdump -lelf $LIB 2>/dev/null | grep -c sm_100a) sm_100a cubins)"
cuobjdump -sass "$LIB" 2>/dev/null | awk '
  /Function :/ { fn=$3; next }
  { for (i=1;i<=NF;i++) if ($i ~ /^(UBLKCP|SYNCS|MATCH|ATOMS|ATOMG|RED|UTMALDG|UTCMMA|LDGSTS|REDUX|SHFL|VOTE)/) { split($i,a,";"); m=a[1]; sub(/\..*/,"",m); c[fn" "m]++ } }
  END { for (k in c) print k, c[k] }' | sort | awk '{ if ($1!=last) { if (last!="") print line; line=$1":"; last=$1 } line=line" "$2"="$3 } END { print line }' | c++filt | sed 's/(.*):/:/' 
echo "# library symbols (must be empty): $(cuobjdump -sass "$LIB" 2>/dev/null | grep -c "Function : .*cub")"
