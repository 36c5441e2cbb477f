// rtb_math.h — scalar restatement of the reference's SSE Vector4 semantics for device code.
//
// Every helper documents the reference operation it mirrors (file:line relative to the reference checkout) and
// keeps the same rounding points: products and sums are rounded individually, FMA appears ONLY where the
// reference calls MulAndAdd/MulAndSub/NegMulAndAdd/NegMulAndSub (Core/Math/Vector4ImplSSE.h:329-363).
// Compile with nvcc --fmad=false -ftz=true -prec-div=true -prec-sqrt=true (host: -ffp-contract=off + FTZ/DAZ),
// the reference runs with FTZ+DAZ (Core/Math/Math.cpp:27-34).
#pragma once

#include <stdint.h>
#include <math.h>
#include <float.h>

#if defined(__CUDACC__)
#define RTB_HD __host__ __device__ __forceinline__
#define RTB_HD_NOINLINE static __host__ __device__ __noinline__        // (static: the header is part of two translation units)
#else
#define RTB_HD inline __attribute__((always_inline))
#define RTB_HD_NOINLINE inline
#endif

namespace rtb {

#define RTB_PI 3.14159265359f
#define RTB_INV_PI 0.31830988618f
#define RTB_2PI 6.28318530718f

struct V3 { float x, y, z; };
struct V4 { float x, y, z, w; };

RTB_HD V3 v3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
RTB_HD V3 v3s(float s) { return v3(s, s, s); }
RTB_HD V4 v4(float x, float y, float z, float w) { V4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
RTB_HD V4 v4s(float s) { return v4(s, s, s, s); }
RTB_HD V3 xyz(const V4& v) { return v3(v.x, v.y, v.z); }

RTB_HD float bits2f(uint32_t u) { union { uint32_t u; float f; } c; c.u = u; return c.f; }
RTB_HD uint32_t f2bits(float f) { union { uint32_t u; float f; } c; c.f = f; return c.u; }

// single-rounding fused multiply-add; the only place FMA is allowed to appear
RTB_HD float fma_(float a, float b, float c) { return fmaf(a, b, c); }

// IEEE division by a value the compiler can see is a constant: under -ftz=true nvcc rewrites `x / 10.0f` into
// `x * 0.1f` (mul.rn.ftz by the rounded reciprocal), which is one ulp off; the intrinsic keeps div.rn.
RTB_HD float fdiv(float a, float b)
{
#if defined(__CUDA_ARCH__)
    return __fdiv_rn(a, b);
#else
    return a / b;
#endif
}

// _mm_min_ps(a, b): a < b ? a : b   (second operand on NaN / equal)   Vector4ImplSSE.h:388-391
RTB_HD float sse_min(float a, float b) { return a < b ? a : b; }
// _mm_max_ps(a, b): a > b ? a : b                                     Vector4ImplSSE.h:393-396
RTB_HD float sse_max(float a, float b) { return a > b ? a : b; }
// math::Min / math::Max (scalar templates)                            Math.h:61-86
RTB_HD float s_min(float a, float b) { return (a < b) ? a : b; }
RTB_HD float s_max(float a, float b) { return (a < b) ? b : a; }
RTB_HD float s_abs(float x) { return (x < 0.0f) ? -x : x; }            // Math.h:88-99 (keeps -0)
RTB_HD float s_clamp(float x, float lo, float hi) { return x > hi ? hi : (x < lo ? lo : x); } // Math.h:172-181
RTB_HD float s_lerp(float a, float b, float w) { return a + w * (b - a); }                      // Math.h:206-210
RTB_HD float s_sqr(float x) { return x * x; }
RTB_HD float s_copysign(float x, float y) { return bits2f((0x7fffffffu & f2bits(x)) | (0x80000000u & f2bits(y))); } // Math.h:137-145
// Vector4::Abs: clears the sign bit                                   Vector4ImplSSE.h:398-401
RTB_HD float bit_abs(float x) { return bits2f(f2bits(x) & 0x7fffffffu); }

// ---- V3 arithmetic (lane-wise _mm_add/sub/mul/div_ps on the xyz lanes) ------------------------------------------
RTB_HD V3 operator+(const V3& a, const V3& b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
RTB_HD V3 operator-(const V3& a, const V3& b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
RTB_HD V3 operator*(const V3& a, const V3& b) { return v3(a.x * b.x, a.y * b.y, a.z * b.z); }
RTB_HD V3 operator/(const V3& a, const V3& b) { return v3(a.x / b.x, a.y / b.y, a.z / b.z); }
RTB_HD V3 operator*(const V3& a, float b) { return v3(a.x * b, a.y * b, a.z * b); }
RTB_HD V3 operator*(float b, const V3& a) { return v3(a.x * b, a.y * b, a.z * b); }
RTB_HD V3 operator/(const V3& a, float b) { return v3(a.x / b, a.y / b, a.z / b); }
// Vector4::operator-() is Zero() - v                                  Vector4ImplSSE.h:236-239
RTB_HD V3 neg(const V3& a) { return v3(0.0f - a.x, 0.0f - a.y, 0.0f - a.z); }

RTB_HD V4 operator+(const V4& a, const V4& b) { return v4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
RTB_HD V4 operator-(const V4& a, const V4& b) { return v4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
RTB_HD V4 operator*(const V4& a, const V4& b) { return v4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
RTB_HD V4 operator*(const V4& a, float b) { return v4(a.x * b, a.y * b, a.z * b, a.w * b); }
RTB_HD V4 operator/(const V4& a, float b) { return v4(a.x / b, a.y / b, a.z / b, a.w / b); }

// MulAndAdd(a, b, c) = fmadd                                          Vector4ImplSSE.h:329-336
RTB_HD V3 mul_add(const V3& a, const V3& b, const V3& c) { return v3(fma_(a.x, b.x, c.x), fma_(a.y, b.y, c.y), fma_(a.z, b.z, c.z)); }
RTB_HD V3 mul_add(const V3& a, float b, const V3& c) { return v3(fma_(a.x, b, c.x), fma_(a.y, b, c.y), fma_(a.z, b, c.z)); }
RTB_HD V4 mul_add(const V4& a, const V4& b, const V4& c) { return v4(fma_(a.x, b.x, c.x), fma_(a.y, b.y, c.y), fma_(a.z, b.z, c.z), fma_(a.w, b.w, c.w)); }
RTB_HD V4 mul_add(const V4& a, float b, const V4& c) { return v4(fma_(a.x, b, c.x), fma_(a.y, b, c.y), fma_(a.z, b, c.z), fma_(a.w, b, c.w)); }
// MulAndSub(a, b, c) = fmsub = a*b - c                                Vector4ImplSSE.h:338-345
RTB_HD V3 mul_sub(const V3& a, const V3& b, const V3& c) { return v3(fma_(a.x, b.x, -c.x), fma_(a.y, b.y, -c.y), fma_(a.z, b.z, -c.z)); }
// NegMulAndAdd(a, b, c) = fnmadd = -(a*b) + c                         Vector4ImplSSE.h:347-354
RTB_HD V3 neg_mul_add(const V3& a, const V3& b, const V3& c) { return v3(fma_(-a.x, b.x, c.x), fma_(-a.y, b.y, c.y), fma_(-a.z, b.z, c.z)); }
RTB_HD V3 neg_mul_add(float a, const V3& b, const V3& c) { return v3(fma_(-a, b.x, c.x), fma_(-a, b.y, c.y), fma_(-a, b.z, c.z)); }

// Dot3V = _mm_dp_ps(a, b, 0x7F): (x*x' + y*y') + (z*z' + 0)           Vector4ImplSSE.h:451-454
RTB_HD float dot3(const V3& a, const V3& b) { return (a.x * b.x + a.y * b.y) + (a.z * b.z + 0.0f); }
// Dot2V = _mm_dp_ps(a, b, 0x3F): (x*x' + y*y') + (0 + 0)              Vector4ImplSSE.h:446-449
RTB_HD float dot2(float ax, float ay, float bx, float by) { return (ax * bx + ay * by) + 0.0f; }

// Cross3: r = a.yzx * b.zxy; result = fnmadd(a.zxy, b.yzx, r)         Vector4ImplSSE.h:476-484
RTB_HD V3 cross3(const V3& a, const V3& b)
{
    return v3(fma_(-a.z, b.y, a.y * b.z), fma_(-a.x, b.z, a.z * b.x), fma_(-a.y, b.x, a.x * b.y));
}

RTB_HD float sqr_length3(const V3& a) { return dot3(a, a); }
RTB_HD float length3(const V3& a) { return sqrtf(dot3(a, a)); }
// Normalize3: v / sqrt(dot3) (IEEE div + sqrt)                        Vector4ImplSSE.h:511-517
// FastNormalize3 is the same in the exact-math oracle (oracle patch P9)
// (device: a ZERO numerator sends nvcc's IEEE division into its slow-path subroutine - 34 instructions to produce a signed
//  zero.  Normals and tangents of axis-aligned geometry have zero components all the time: ncu on k_shade, round 2, showed
//  7 % of its warp instructions in that subroutine at 3 of 32 lanes.  0 / l = +-0 for every l that is not 0 or NaN.)
RTB_HD float div_or_zero(float x, float l)
{
#if defined(__CUDA_ARCH__)
    if (x == 0.0f && l == l && l != 0.0f) return __uint_as_float((__float_as_uint(x) ^ __float_as_uint(l)) & 0x80000000u);
#endif
    return x / l;
}
RTB_HD V3 normalize3(const V3& a) { const float l = sqrtf(dot3(a, a)); return v3(div_or_zero(a.x, l), div_or_zero(a.y, l), div_or_zero(a.z, l)); }

// Reflect3(i, n) = fnmadd(d + d, n, i), d = dot3(i, n)                Vector4Impl.h:114-119
RTB_HD V3 reflect3(const V3& i, const V3& n) { const float d = dot3(i, n); return neg_mul_add(d + d, n, i); }
// Orthogonalize(v, ref) = fnmadd(dot3(v, ref), ref, v)                Vector4ImplSSE.h:587-591
RTB_HD V3 orthogonalize(const V3& v, const V3& r) { return neg_mul_add(dot3(v, r), r, v); }

// Matrix4 rows (row-vector convention)
struct M4 { V4 r0, r1, r2, r3; };

// Matrix4::TransformPoint                                             Matrix4.h:110-117
RTB_HD V3 transform_point(const M4& m, const V3& a)
{
    V3 t = v3(fma_(a.x, m.r0.x, m.r3.x), fma_(a.x, m.r0.y, m.r3.y), fma_(a.x, m.r0.z, m.r3.z));
    t = v3(fma_(a.y, m.r1.x, t.x), fma_(a.y, m.r1.y, t.y), fma_(a.y, m.r1.z, t.z));
    t = v3(fma_(a.z, m.r2.x, t.x), fma_(a.z, m.r2.y, t.y), fma_(a.z, m.r2.z, t.z));
    return t;
}
// Matrix4::TransformVector                                            Matrix4.h:133-139
RTB_HD V3 transform_vector(const M4& m, const V3& a)
{
    V3 t = v3(a.x * m.r0.x, a.x * m.r0.y, a.x * m.r0.z);
    t = v3(fma_(a.y, m.r1.x, t.x), fma_(a.y, m.r1.y, t.y), fma_(a.y, m.r1.z, t.z));
    t = v3(fma_(a.z, m.r2.x, t.x), fma_(a.z, m.r2.y, t.y), fma_(a.z, m.r2.z, t.z));
    return t;
}
// Matrix4::TransformVectorNeg: x*r0, fnmsub(y, r1, t), fnmadd(z, r2, t)   Matrix4.h:143-149
RTB_HD V3 transform_vector_neg(const M4& m, const V3& a)
{
    V3 t = v3(a.x * m.r0.x, a.x * m.r0.y, a.x * m.r0.z);
    t = v3(fma_(-a.y, m.r1.x, -t.x), fma_(-a.y, m.r1.y, -t.y), fma_(-a.y, m.r1.z, -t.z));
    t = v3(fma_(-a.z, m.r2.x, t.x), fma_(-a.z, m.r2.y, t.y), fma_(-a.z, m.r2.z, t.z));
    return t;
}
// Matrix4::FastInverseNoScale: transpose3 + TransformVectorNeg(row3)  Matrix4.h:195-202
RTB_HD M4 fast_inverse_no_scale(const M4& m)
{
    M4 r;
    r.r0 = v4(m.r0.x, m.r1.x, m.r2.x, 0.0f);
    r.r1 = v4(m.r0.y, m.r1.y, m.r2.y, 0.0f);
    r.r2 = v4(m.r0.z, m.r1.z, m.r2.z, 0.0f);
    r.r3 = v4(0.0f, 0.0f, 0.0f, 1.0f);
    const V3 t = transform_vector_neg(r, xyz(m.r3));
    r.r3 = v4(t.x, t.y, t.z, 0.0f);
    return r;
}

// ---- math::Ray ------------------------------------------------------------------------------------------------
struct Ray
{
    V3 origin, dir, invDir, originDivDir;
};

// Ray(origin, direction): Vector4::InvNormalized                      Ray.h:23-28, Vector4Impl.h:90-98
//   len = sqrt(dot3); inv = 1/[x, y, z, len]; invDir = len * inv; dir = v * inv.w; originDivDir = origin * invDir
RTB_HD Ray make_ray(const V3& origin, const V3& direction)
{
    Ray r;
    const float len = sqrtf(dot3(direction, direction));
    const float ix = 1.0f / direction.x, iy = 1.0f / direction.y, iz = 1.0f / direction.z, il = 1.0f / len;
    r.origin = origin;
    r.invDir = v3(len * ix, len * iy, len * iz);
    r.dir = v3(direction.x * il, direction.y * il, direction.z * il);
    r.originDivDir = origin * r.invDir;
    return r;
}
// Ray::BuildUnsafe: invDir = 1/dir                                    Ray.h:31-39
RTB_HD Ray make_ray_unsafe(const V3& origin, const V3& dir)
{
    Ray r;
    r.origin = origin;
    r.dir = dir;
    r.invDir = v3(1.0f / dir.x, 1.0f / dir.y, 1.0f / dir.z);
    r.originDivDir = origin * r.invDir;
    return r;
}
// Matrix4::TransformRay_Unsafe                                        Matrix4.h:245-250
RTB_HD Ray transform_ray_unsafe(const M4& m, const Ray& ray)
{
    return make_ray_unsafe(transform_point(m, ray.origin), transform_vector(m, ray.dir));
}
// Ray::GetAtDistance = fmadd(dir, t, origin)                          Ray.h:41-44
RTB_HD V3 ray_at(const Ray& r, float t) { return mul_add(r.dir, t, r.origin); }

// ---- transcendental approximations (Core/Math/Transcendental.cpp) ---------------------------------------------

// _mm_cvtps_epi32: round to nearest even                              VectorInt4ImplSSE.h:48-51
RTB_HD int32_t cvt_rn(float x)
{
#if defined(__CUDA_ARCH__)
    return __float2int_rn(x);
#else
    return (int32_t)lrintf(x);
#endif
}

// Sin(Vector4), one lane                                              Transcendental.cpp:54-76
RTB_HD float sin_poly(float a)
{
    const float c0 = 9.9999970197e-01f, c1 = -1.6666577756e-01f, c2 = 8.3325579762e-03f;
    const float c3 = -1.9812576647e-04f, c4 = 2.7040521217e-06f, c5 = -2.0532988642e-08f;
    const int32_t i = cvt_rn(a * (1.0f / RTB_PI));
    const float x = fma_(-(float)i, RTB_PI, a);
    const float x2 = x * x;
    float y = fma_(c5, x2, c4);
    y = fma_(y, x2, c3);
    y = fma_(y, x2, c2);
    y = fma_(y, x2, c1);
    y = fma_(y, x2, c0);
    y = y * x;
    return bits2f(f2bits(y) ^ ((uint32_t)i << 31));
}
// SinCos(x): Sin(Vector4(x) + (0, PI/2, 0, 0)) -> (sin, cos)           TranscendentalImpl.h:22-26
RTB_HD void sin_cos(float x, float& s, float& c)
{
    s = sin_poly(x + 0.0f);
    c = sin_poly(x + (RTB_PI / 2.0f));
}

// FastACos                                                            Transcendental.cpp:112-126
RTB_HD float fast_acos(float x)
{
    const float negate = (float)(x < 0);
    x = fabsf(x);
    float ret = -0.0187293f;
    ret = ret * x + 0.0742610f;
    ret = ret * x - 0.2121144f;
    ret = ret * x + 1.5707288f;
    ret = ret * sqrtf(1.0f - x);
    ret = ret - 2.0f * negate * ret;
    return negate * 3.14159265358979f + ret;
}
// FastATan2                                                           Transcendental.cpp:229-256
RTB_HD float fast_atan2(const float y, const float x)
{
    const float ax = s_abs(x), ay = s_abs(y);
    const float mx = s_max(ay, ax), mn = s_min(ay, ax);
    const float a = mn / mx;
    const float s = a * a, c = s * a, q = s * s;
    const float t = -0.094097948f * q - 0.33213072f;
    float r = (0.024840285f * q + 0.18681418f);
    r = r * s + t;
    r = r * c + a;
    if (ay > ax) r = 1.57079637f - r;
    if (x < 0.0f) r = RTB_PI - r;
    if (y < 0.0f) r = -r;
    return r;
}
// CartesianToSphericalCoordinates                                     Geometry.cpp:8-13
RTB_HD void cartesian_to_spherical(const V3& in, float& u, float& v)
{
    const float theta = fast_acos(s_clamp(in.y, -1.0f, 1.0f));
    const float phi = s_abs(in.x) > FLT_EPSILON ? fast_atan2(in.z, in.x) : 0.0f;
    u = fdiv(phi, 2.0f * RTB_PI) + 0.5f;
    v = fdiv(theta, RTB_PI);
}

// BuildOrthonormalBasis                                               Geometry.cpp:15-33
RTB_HD void build_orthonormal_basis(const V3& n, V3& u, V3& v)
{
    const float sign = s_copysign(1.0f, n.z);
    const float a = -1.0f / (sign + n.z);
    u = v3(1.0f + sign * n.x * n.x * a, sign * n.x * n.y * a, -sign * n.x);
    v = v3(n.x * n.y * a, sign + n.y * n.y * a, -n.y);
}

// ---- Math/Utils.cpp ---------------------------------------------------------------------------------------------
// FresnelDielectric                                                   Utils.cpp:9-29
RTB_HD float fresnel_dielectric(float NdV, float eta)
{
    if (NdV > 0.0f) eta = 1.0f / eta;
    const float c = fabsf(NdV);
    float g = eta * eta * (1.0f - NdV * NdV);
    if (g < 1.0f)
    {
        g = sqrtf(1.0f - g);
        const float A = (g - c) / (g + c);
        const float B = (c * (g + c) - 1.0f) / (c * (g - c) + 1.0f);
        return 0.5f * A * A * (1.0f + B * B);
    }
    return 1.0f;
}
// FresnelMetal                                                        Utils.cpp:31-39
RTB_HD float fresnel_metal(const float NdV, const float eta, const float k)
{
    const float NdV2 = NdV * NdV;
    const float a = eta * eta + k * k;
    const float b = a * NdV2;
    const float rs = (b - (2.0f * eta * NdV) + 1.0f) / (b + (2.0f * eta * NdV) + 1.0f);
    const float rp = (a - (2.0f * eta * NdV) + NdV2) / (a + (2.0f * eta * NdV) + NdV2);
    return (rs + rp) * 0.5f;
}
// Vector4::Refract3 (returns zero vector on total internal reflection) Vector4.cpp:7-30
RTB_HD V3 refract3(const V3& i, const V3& n, float eta)
{
    const float NdotV = dot3(i, n);
    if (NdotV < 0.0f) eta = 1.0f / eta;
    const float k = 1.0f - eta * eta * (1.0f - NdotV * NdotV);
    if (k <= 0.0f) return v3(0.0f, 0.0f, 0.0f);
    V3 t = neg_mul_add(eta * NdotV + sqrtf(k), n, i * eta);
    if (NdotV > 0.0f) t.z = -t.z;
    return normalize3(t);
}

// ---- Math/SamplingHelpers.cpp -----------------------------------------------------------------------------------
// GetHemishpereCos                                                    SamplingHelpers.cpp:134-143
RTB_HD V3 hemisphere_cos(float ux, float uy)
{
    const float theta = 2.0f * RTB_PI * uy;
    const float r = sqrtf(ux);
    float s, c;
    sin_cos(theta, s, c);
    return v3(r * s, r * c, sqrtf(1.0f - ux));
}
// GetSphere                                                           SamplingHelpers.cpp:112-124
RTB_HD V3 sample_sphere(float ux, float uy)
{
    const float vx = fma_(ux, 2.0f, -1.0f), vy = fma_(uy, 2.0f, -1.0f);
    const float t = sqrtf(1.0f - vy * vy);
    const float theta = RTB_PI * vx;
    float s, c;
    sin_cos(theta, s, c);
    return v3(t * s, t * c, vy);
}
// GetHemishpere                                                       SamplingHelpers.cpp:126-132
RTB_HD V3 sample_hemisphere(float ux, float uy)
{
    V3 p = sample_sphere(ux, uy);
    p.z = s_abs(p.z);
    return p;
}
// GetCircle                                                           SamplingHelpers.cpp:29-39
RTB_HD void sample_circle(float ux, float uy, float& ox, float& oy)
{
    const float theta = 2.0f * RTB_PI * ux;
    const float r = sqrtf(uy);
    float s, c;
    sin_cos(theta, s, c);
    ox = r * s; oy = r * c;
}
// GetHexagon                                                          SamplingHelpers.cpp:41-58
RTB_HD void sample_hexagon(float ux, float uy, float uz, float& ox, float& oy)
{
    const float hx[4] = { -1.0f, 0.5f, 0.5f, -1.0f };
    const float hy[4] = { 0.0f, 0.8660254f, -0.8660254f, 0.0f };
    const uint32_t x = (uint32_t)(3.0f * uz);
    ox = ux * hx[x] + uy * hx[x + 1];
    oy = ux * hy[x] + uy * hy[x + 1];
}

RTB_HD float uniform_hemisphere_pdf() { return RTB_INV_PI / 2.0f; }                             // Geometry.h:19-22
RTB_HD float sphere_cap_pdf(const float cosTheta) { return 1.0f / (RTB_2PI * (1.0f - cosTheta)); } // Geometry.h:39-42

} // namespace rtb
