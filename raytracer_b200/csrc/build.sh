#!/usr/bin/env bash
# Build librtb200.so (sm_100a only).  --fmad=false + IEEE div/sqrt + FTZ mirror the exact-math oracle's arithmetic:
# FMA appears only where the reference calls an explicit fused intrinsic (rtb_math.h).
# Two translation units from ONE source file: the library proper, and k_tail alone at ptxas -O1 (rtb_kernels.cu, "compiled TWICE";
# DESIGN.md section 4c).  Extra arguments (-D...) go to both.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="${RTB_OUT:-$HERE/../librtb200.so}"
# (RTB_MATH_FLAGS: only for measuring what the bit-exact bar costs - see DESIGN.md section 6; such a build fails the parity tests)
MATH="${RTB_MATH_FLAGS:---fmad=false -ftz=true -prec-div=true -prec-sqrt=true}"
TAIL_OPT="${RTB_TAIL_PTXAS:--O1}"
TMP="$(mktemp -d)"
trap 'rm -rf "$TMP"' EXIT
COMMON="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 $MATH -Xcompiler -fPIC -Xcompiler -O2"
nvcc $COMMON -Xptxas -v -c -o "$TMP/main.o" "$HERE/rtb_kernels.cu" "$@" &
nvcc $COMMON -Xptxas -v -Xptxas $TAIL_OPT -DRTB_TU_TAIL=1 -c -o "$TMP/tail.o" "$HERE/rtb_kernels.cu" "$@" &
wait -n; wait -n
nvcc -shared -o "$OUT" "$TMP/main.o" "$TMP/tail.o" -ldl
echo "built $OUT"
