#!/bin/bash
# SASS evidence for profiles/: which hardware paths the shipped library uses (no GPU needed).
#   bash junction-tree_b200/tools/sass_evidence.sh   -> profiles/r02_sass_summary.txt, r02_sass_dense_kernel.txt,
#                                                        r02_sass_project_tma_kernel.txt, r02_sass_beta_kernel.txt
set -e
ROOT=$(cd "$(dirname "$0")/../.." && pwd)
LIB=$ROOT/junction-tree_b200/junctiontree/libjt_b200.so
OUT=$ROOT/profiles
TMP=$(mktemp)
cuobjdump -sass "$LIB" > "$TMP"
{
  echo "libjt_b200.so: $(cuobjdump -lelf "$LIB" | grep -c sm_100a) sm_100a cubins, $(grep -c 'Function :' "$TMP") kernels"
  echo "arch lines: $(grep -o 'EF_CUDA_SM[0-9a-z]*' "$TMP" | sort | uniq -c | tr '\n' ' ')"
  echo
  echo "mnemonic counts over the whole library:"
  for m in "DMMA.8x8x4" "UBLKCP" "SYNCS" "LDG.E.128" "STG.E.128" "STG.E.EF.128" "LDS.64" "LDS.128" "DFMA" "DMUL" "HMMA" "UTCHMMA" "UTCQMMA" "LDTM" "STTM"; do
    printf "  %-14s %s\n" "$m" "$(grep -c -- "$m" "$TMP" || true)"
  done
  echo
  echo "per kernel (sum-product, float64 instantiations):"
  for k in jt_dense_kernelIdE jt_project_tma_kernelINS_12SrSumProductEdLi2E jt_beta_kernelINS_12SrSumProductEdLi2E jt_scalar_kernelINS_12SrSumProductEdLi2E jt_init_rows_kernelINS_12SrSumProductEdLi2E; do
    f=$(grep 'Function :' "$TMP" | grep "$k" | head -1 | sed 's/.*Function : //')
    [ -z "$f" ] && continue
    cuobjdump -sass -fun "$f" "$LIB" > "$TMP.k" 2>/dev/null || true
    printf "  %s\n     instructions %s  DMMA %s  UBLKCP %s  SYNCS %s  LDG.E.128 %s  STG.E.128 %s  STG.E.EF.128 %s  LDS %s  BAR %s\n" "$k" \
      "$(grep -c '^ *\/\*[0-9a-f]*\*\/' "$TMP.k")" "$(grep -c DMMA "$TMP.k")" "$(grep -c UBLKCP "$TMP.k")" "$(grep -c SYNCS "$TMP.k")" \
      "$(grep -c 'LDG.E.128' "$TMP.k")" "$(grep -c 'STG.E.128' "$TMP.k")" "$(grep -c 'STG.E.EF.128' "$TMP.k")" "$(grep -c ' LDS' "$TMP.k")" "$(grep -c 'BAR.SYNC' "$TMP.k")"
  done
} > "$OUT/r02_sass_summary.txt"
dump() {  # $1 = pattern, $2 = output name
  f=$(grep 'Function :' "$TMP" | grep "$1" | head -1 | sed 's/.*Function : //')
  cuobjdump -sass -fun "$f" "$LIB" 2>/dev/null | grep -E '^\s+/\*[0-9a-f]{4}\*/|Function' | sed 's/ *\/\* 0x[0-9a-f]* \*\/$//' > "$OUT/$2"
}
dump jt_dense_kernelIdE r02_sass_dense_kernel.txt
dump jt_project_tma_kernelINS_12SrSumProductEdLi2E r02_sass_project_tma_kernel.txt
dump jt_beta_kernelINS_12SrSumProductEdLi2E r02_sass_beta_kernel.txt
rm -f "$TMP" "$TMP.k"
cat "$OUT/r02_sass_summary.txt"
