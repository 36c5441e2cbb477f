#!/usr/bin/env bash
# Build librtb200.so (sm_100a only).  --fmad=false + IEEE div/sqrt + FTZ mirror the exact-math oracle's arithmetic:
# FMA appears only where the reference calls an explicit fused intrinsic (rtb_math.h).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="${RTB_OUT:-$HERE/../librtb200.so}"
# (RTB_MATH_FLAGS: only for measuring what the bit-exact bar costs - see DESIGN.md section 6; such a build fails the parity tests)
MATH="${RTB_MATH_FLAGS:---fmad=false -ftz=true -prec-div=true -prec-sqrt=true}"
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 \
     $MATH \
     -Xptxas -v -shared -Xcompiler -fPIC -Xcompiler -O2 \
     -o "$OUT" "$HERE/rtb_kernels.cu" -ldl "$@"
echo "built $OUT"
