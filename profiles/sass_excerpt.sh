#!/usr/bin/env bash
# Static SASS evidence (B200_PROFILING.md: the mnemonics that prove TMA bulk copies, mbarriers, reductions, warp matches).
# Usage: bash profiles/sass_excerpt.sh > profiles/r02_sass_excerpt.txt      (needs the built libaindex_cuda.so; no GPU)
LIB="$(dirname "$0")/../aindex_b200/libaindex_cuda.so"
echo "# cuobjdump -sass $(basename "$LIB") (sm_100a), instruction counts per kernel: UBLKCP = cp.async.bulk (TMA 1-D bulk copy),"
echo "# SYNCS = mbarrier ops, RED = reduction atomics, ATOMS/ATOMG = shared/global atomics, MATCH = match.any, POPC, SHFL, LDG.E.128 = 128-bit loads"
cuobjdump -sass "$LIB" | awk '
/Function :/ { name=$3; sub(/^_ZN3aix[0-9]+/, "", name); order[++n]=name; next }
name != "" {
  if ($0 ~ /UBLKCP/) c[name,"UBLKCP"]++
  if ($0 ~ /SYNCS/) c[name,"SYNCS"]++
  if ($0 ~ / REDG?\./) c[name,"RED"]++
  if ($0 ~ /ATOMS/) c[name,"ATOMS"]++
  if ($0 ~ /ATOMG|ATOM\.E/) c[name,"ATOMG"]++
  if ($0 ~ /MATCH/) c[name,"MATCH"]++
  if ($0 ~ /POPC/) c[name,"POPC"]++
  if ($0 ~ /SHFL/) c[name,"SHFL"]++
  if ($0 ~ /LDG\.E\.128|LDG\.E\.[A-Z.]*128/) c[name,"LDG128"]++
  if ($0 ~ /STG\.E\.[A-Z.]*128|STG\.E\.128/) c[name,"STG128"]++
  if ($0 ~ /^ +\/\*[0-9a-f]+\*\/ /) c[name,"total"]++
}
END {
  printf "%-70s %6s %6s %6s %5s %5s %5s %5s %5s %5s %6s %6s\n", "kernel", "instr", "UBLKCP", "SYNCS", "RED", "ATOMS", "ATOMG", "MATCH", "POPC", "SHFL", "LDG128", "STG128"
  for (i = 1; i <= n; i++) { k = order[i];
    printf "%-70s %6d %6d %6d %5d %5d %5d %5d %5d %5d %6d %6d\n", substr(k,1,70), c[k,"total"], c[k,"UBLKCP"], c[k,"SYNCS"], c[k,"RED"], c[k,"ATOMS"], c[k,"ATOMG"], c[k,"MATCH"], c[k,"POPC"], c[k,"SHFL"], c[k,"LDG128"], c[k,"STG128"] }
}'
