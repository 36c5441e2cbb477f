#!/usr/bin/env bash
# Build the REAL reference (Witek902/Raytracer) as the parity oracle and the CPU timing baseline.
#
# TEST INFRASTRUCTURE ONLY — nothing under oracle/ is linked into, imported by or executed from the product.
#
# The reference sources are compiled from a throw-away, patched COPY that lives only under oracle/_ref/
# (git-ignored; travels to the GPU box with the built binaries).  No reference source is committed.
#
#   oracle/_ref/exact/libCore.so + rtref   IEEE div/sqrt, -ffp-contract=off, fixed-seed Entropy  -> parity oracle
#   oracle/_ref/avx/libCore.so   + rtref   as-shipped flags (-O2 -mavx2 -mfma), asserts off       -> CPU timing baseline
#
# Usage: oracle/build_ref.sh [exact|avx|all]     (needs /root/reference; no-op if it is absent)
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${RT_REFERENCE_ROOT:-/root/reference}"
OUT="$HERE/_ref"
WHAT="${1:-all}"
JOBS="${JOBS:-$(nproc)}"

if [ ! -d "$REF/Core" ]; then
    echo "build_ref: $REF not present - keeping prebuilt oracle/_ref as is"
    exit 0
fi

build_variant() {
    local v="$1"
    local root="$OUT/$v"
    local simd="-msse -msse4.1 -mavx -mavx2 -mfma -DRT_USE_SSE -DRT_USE_AVX -DRT_USE_AVX2 -DRT_USE_FMA"
    local flags
    if [ "$v" = "exact" ]; then
        flags="-O2 -ffp-contract=off $simd"
    else
        flags="-O2 $simd -DRT_CONFIGURATION_FINAL"
    fi
    local common="-std=c++14 -fPIC -w -I$root/src/Core -I$HERE/shim -include cstdint -include limits -include map"

    # ---- libCore.so: every Core/**/*.cpp + tinyexr + glue, from a patched throw-away copy ------------------------
    local stamp="$root/.core_built"
    if [ ! -f "$stamp" ] || [ "$HERE/patch_ref.py" -nt "$stamp" ] || [ "$HERE/build_ref.sh" -nt "$stamp" ]; then
        rm -rf "$root"
        mkdir -p "$root/src" "$root/obj"
        cp -r "$REF/Core" "$root/src/Core"
        mkdir -p "$root/src/External" "$root/src/Demo"
        cp -r "$REF/External/rapidjson" "$REF/External/tinyexr" "$REF/External/imgui" "$root/src/External/"
        cp "$REF/External/tiny_obj_loader.h" "$REF/External/tiny_obj_loader.cpp" "$REF/External/cxxopts.hpp" "$root/src/External/"
        cp "$REF/Demo/SceneLoader.cpp" "$REF/Demo/SceneLoader.h" "$REF/Demo/MeshLoader.cpp" "$REF/Demo/MeshLoader.h" "$REF/Demo/PCH.h" "$root/src/Demo/"
        cp "$HERE/shim/Demo.h" "$root/src/Demo/Demo.h"
        mkdir -p "$root/Data"
        cp "$REF/Data/BlueNoise128_RGBA16.dat" "$root/Data/"
        python3 "$HERE/patch_ref.py" "$root/src" "$v"

        ( cd "$root/src" && find Core -name '*.cpp' ; echo External/tinyexr/tinyexr.cc ) > "$root/core_sources.txt"
        ( cd "$root/src" && cat "$root/core_sources.txt" | xargs -P "$JOBS" -I{} sh -c \
            "g++ $flags $common -c {} -o $root/obj/\$(echo {} | tr '/' '_').o 2>/dev/null" )
        g++ $flags $common -c "$HERE/shim/linux_glue.cpp" -o "$root/obj/linux_glue.o"
        g++ -shared -o "$root/libCore.so" "$root"/obj/*.o -lpthread
        rm -rf "$root/obj"
        mkdir -p "$root/obj_cli"
        for f in Demo/SceneLoader.cpp Demo/MeshLoader.cpp External/tiny_obj_loader.cpp; do
            ( cd "$root/src" && g++ $flags $common -I"$root/src/Demo" -I"$root/src/External" -c "$f" -o "$root/obj_cli/$(basename $f).o" ) &
        done
        wait
        touch "$stamp"
    else
        echo "build_ref: $v libCore up to date"
    fi

    # ---- headless harness: reference loader (patched copy) + oracle/rtref.cpp -----------------------------------
    if [ ! -f "$root/rtref" ] || [ "$HERE/rtref.cpp" -nt "$root/rtref" ] || [ "$HERE/ref_flatten.h" -nt "$root/rtref" ] || [ "$HERE/../include/rtb200.h" -nt "$root/rtref" ] || [ "$stamp" -nt "$root/rtref" ]; then
        g++ $flags $common -fno-access-control -I"$root/src" -I"$root/src/Demo" -I"$root/src/External" -I"$HERE/../include" -I"$HERE" \
            -c "$HERE/rtref.cpp" -o "$root/obj_cli/rtref.o"
        g++ -rdynamic -o "$root/rtref" "$root"/obj_cli/*.o -L"$root" -lCore -lpthread -Wl,-rpath,'$ORIGIN'
        mkdir -p "$root/run"   # rtref must run with CWD=<variant>/run so that ../Data/BlueNoise128_RGBA16.dat resolves (GenericSampler.cpp:13)
        echo "build_ref: built $root/rtref"
    else
        echo "build_ref: $v rtref up to date"
    fi

    # ---- the reference-side adapter of INTEGRATION.md, compiled for real (exact variant only; needs librtb200.so) --------
    local lib="$HERE/../raytracer_b200/librtb200.so"
    if [ "$v" = "exact" ] && [ -f "$lib" ]; then
        if [ ! -f "$root/rtadapter" ] || [ "$HERE/rtadapter.cpp" -nt "$root/rtadapter" ] || [ "$HERE/ref_flatten.h" -nt "$root/rtadapter" ] || [ "$HERE/../include/rtb200.h" -nt "$root/rtadapter" ] || [ "$stamp" -nt "$root/rtadapter" ]; then
            g++ $flags $common -fno-access-control -I"$root/src" -I"$root/src/Demo" -I"$root/src/External" -I"$HERE/../include" -I"$HERE" \
                -c "$HERE/rtadapter.cpp" -o "$root/rtadapter.o"
            g++ -rdynamic -o "$root/rtadapter" "$root/rtadapter.o" "$root/obj_cli/SceneLoader.cpp.o" "$root/obj_cli/MeshLoader.cpp.o" "$root/obj_cli/tiny_obj_loader.cpp.o" \
                -L"$root" -lCore -L"$HERE/../raytracer_b200" -lrtb200 -lpthread -Wl,-rpath,'$ORIGIN' -Wl,-rpath,'$ORIGIN/../../../raytracer_b200'
            rm -f "$root/rtadapter.o"
            echo "build_ref: built $root/rtadapter"
        else
            echo "build_ref: $v rtadapter up to date"
        fi
    fi
}

case "$WHAT" in
    exact) build_variant exact ;;
    avx)   build_variant avx ;;
    all)   build_variant exact; build_variant avx ;;
    *) echo "usage: $0 [exact|avx|all]"; exit 2 ;;
esac
