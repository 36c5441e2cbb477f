#!/usr/bin/env bash
# Build librtb200_host.so: the C++ host object model (Scene / Camera / Material / lights / loaders / ViewportB200).
# -ffp-contract=off: scene construction must round exactly like the exact-math oracle (no implicit FMA).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
PKG="$HERE/.."
g++ -O2 -std=c++17 -fPIC -shared -ffp-contract=off -mfma -Wall -Wno-unused-function \
    -o "$PKG/librtb200_host.so" "$HERE/rt_capi.cpp" -L"$PKG" -lrtb200 -lz -Wl,-rpath,'$ORIGIN'
g++ -O2 -std=c++17 -ffp-contract=off -mfma -Wall -Wno-unused-function \
    -o "$PKG/rtb200_render" "$HERE/rtb200_render.cpp" -L"$PKG" -lrtb200 -lz -Wl,-rpath,'$ORIGIN'
echo "built $PKG/librtb200_host.so and $PKG/rtb200_render"
