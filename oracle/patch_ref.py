#!/usr/bin/env python3
"""Apply the oracle patch list to a *copy* of the reference (never to /root/reference).

TEST INFRASTRUCTURE ONLY.  Usage: patch_ref.py <copy-root> <variant>   (variant: exact | avx)

Patches (see DESIGN.md "Oracle"):
  P3  Core/Math/VectorInt8.h      anonymous struct with non-trivial members is a hard error on GCC; only the non-AVX2
                                  implementation uses it -> guard with #ifndef RT_USE_AVX2
  P4  Core/Utils/Memory.cpp       Linux branch of SystemAllocator::Allocate drops the pointer (returns nullptr)
  P7  Demo/SceneLoader.cpp        MSVC-only: temporary bound to non-const reference default argument, `return false`
                                  from a function returning ShapePtr
  P8  Core/Utils/Entropy.cpp      deterministic seed stream (splitmix64) when RT_FIXED_SEED is set in the environment
  P9  (exact only) Core/Math/Math.h, Core/Math/Vector4ImplSSE.h: rcpss / rcpps / rsqrtps approximations -> IEEE ops
Edits are byte-level because Vector4ImplSSE.h contains a non-UTF-8 byte.
"""
import sys
from pathlib import Path


def patch(path: Path, old: bytes, new: bytes, count: int = 1) -> None:
    data = path.read_bytes()
    found = data.count(old)
    if found != count:
        raise SystemExit(f"patch_ref: {path}: expected {count} occurrence(s) of {old[:50]!r}, found {found}")
    path.write_bytes(data.replace(old, new))


def main() -> None:
    root = Path(sys.argv[1])
    variant = sys.argv[2]
    assert variant in ("exact", "avx")

    # P3
    patch(root / "Core/Math/VectorInt8.h",
          b"        struct\n        {\n            VectorInt4 low;\n            VectorInt4 high;\n        };\n",
          b"#ifndef RT_USE_AVX2\n        struct\n        {\n            VectorInt4 low;\n            VectorInt4 high;\n        };\n#endif\n")

    # P4
    patch(root / "Core/Utils/Memory.cpp",
          b"    DefaultAllocator::Allocate(size, alignment);\n",
          b"    ptr = DefaultAllocator::Allocate(size, alignment < 4096 ? 4096 : alignment); // oracle: emulate VirtualAlloc page alignment\n")

    # P7
    loader = root / "Demo/SceneLoader.cpp"
    patch(loader,
          b"static ShapePtr ParseShape(const rapidjson::Value& value, Scene& scene, MaterialsMap& materials = MaterialsMap())",
          b"static MaterialsMap gOracleDummyMaterials;\n"
          b"static ShapePtr ParseShape(const rapidjson::Value& value, Scene& scene, MaterialsMap& materials = gOracleDummyMaterials)")
    data = loader.read_bytes()
    start = data.index(b"static MaterialsMap gOracleDummyMaterials;")
    end = data.index(b"static bool ParseLight(")
    body = data[start:end].replace(b"return false;", b"return nullptr;")
    loader.write_bytes(data[:start] + body + data[end:])

    # P8
    patch(root / "Core/Utils/Entropy.cpp",
          b"uint32 Entropy::GetInt()\n{\n    uint32 result = 0;\n",
          b"uint32 Entropy::GetInt()\n{\n"
          b"    // oracle: deterministic seed stream (splitmix64) when RT_FIXED_SEED is set\n"
          b"    static const char* oracleSeedEnv = getenv(\"RT_FIXED_SEED\");\n"
          b"    if (oracleSeedEnv)\n"
          b"    {\n"
          b"        static uint64 state = strtoull(oracleSeedEnv, nullptr, 10);\n"
          b"        state += 0x9E3779B97F4A7C15ull;\n"
          b"        uint64 z = state;\n"
          b"        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;\n"
          b"        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;\n"
          b"        z = z ^ (z >> 31);\n"
          b"        return (uint32)(z >> 32);\n"
          b"    }\n"
          b"    uint32 result = 0;\n")

    if variant == "exact":
        # P9
        patch(root / "Core/Math/Math.h",
              b"    return a * _mm_cvtss_f32(_mm_rcp_ss(_mm_load_ss(&b)));\n",
              b"    return a / b; // oracle exact\n")
        v4 = root / "Core/Math/Vector4ImplSSE.h"
        patch(v4,
              b"    const __m128 rcp = _mm_rcp_ps(v);\n"
              b"    const __m128 rcpSqr = _mm_mul_ps(rcp, rcp);\n"
              b"    const __m128 rcp2 = _mm_add_ps(rcp, rcp);\n"
              b"    return NegMulAndAdd(rcpSqr, v, rcp2);\n",
              b"    return _mm_div_ps(VECTOR_ONE, v); // oracle exact\n")
        patch(v4,
              b"    v = _mm_mul_ps(v, _mm_rsqrt_ps(vDot));\n",
              b"    v = _mm_div_ps(v, _mm_sqrt_ps(vDot)); // oracle exact\n")


if __name__ == "__main__":
    main()
