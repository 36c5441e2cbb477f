// rt_math.h — host-side value types of the plugin surface (scalar C++, no SIMD).
//
// Only what scene construction needs, with the reference's operation order so that the flattened scene (object
// matrices, bounding boxes, BVHs, preprocessed triangles) is bit-identical to what the reference's own loader
// produces (checked by tests/test_host_scene.py against the oracle's flat dump).  Compile with -ffp-contract=off.
//
//   Vector4 / Box                  Core/Math/Vector4ImplSSE.h, Core/Math/Box.h
//   Matrix4::Inverse/TransformBox  Core/Math/Matrix4.cpp:103-268
//   Quaternion::FromEulerAngles    Core/Math/Quaternion.cpp:156-181, QuaternionImpl.h:35-63
//   Transform::ToMatrix4           Core/Math/Transform.cpp:89-94
//   Sin/Cos (scalar)               Core/Math/Transcendental.cpp:36-52, TranscendentalImpl.h:8-11
#pragma once

#include <cmath>
#include <cstdint>
#include <cfloat>

namespace rt {
namespace math {

constexpr float RT_PI_F = 3.14159265359f;

struct Vector4
{
    float x = 0.0f, y = 0.0f, z = 0.0f, w = 0.0f;
    Vector4() = default;
    Vector4(float x_, float y_, float z_ = 0.0f, float w_ = 0.0f) : x(x_), y(y_), z(z_), w(w_) {}
    explicit Vector4(float s) : x(s), y(s), z(s), w(s) {}
    float operator[](int i) const { return (&x)[i]; }
    float& operator[](int i) { return (&x)[i]; }
    Vector4 operator+(const Vector4& b) const { return { x + b.x, y + b.y, z + b.z, w + b.w }; }
    Vector4 operator-(const Vector4& b) const { return { x - b.x, y - b.y, z - b.z, w - b.w }; }
    Vector4 operator*(const Vector4& b) const { return { x * b.x, y * b.y, z * b.z, w * b.w }; }
    Vector4 operator/(const Vector4& b) const { return { x / b.x, y / b.y, z / b.z, w / b.w }; }
    Vector4 operator*(float b) const { return { x * b, y * b, z * b, w * b }; }
    Vector4 operator/(float b) const { return { x / b, y / b, z / b, w / b }; }
    // _mm_min_ps / _mm_max_ps operand semantics
    static Vector4 Min(const Vector4& a, const Vector4& b) { return { a.x < b.x ? a.x : b.x, a.y < b.y ? a.y : b.y, a.z < b.z ? a.z : b.z, a.w < b.w ? a.w : b.w }; }
    static Vector4 Max(const Vector4& a, const Vector4& b) { return { a.x > b.x ? a.x : b.x, a.y > b.y ? a.y : b.y, a.z > b.z ? a.z : b.z, a.w > b.w ? a.w : b.w }; }
    // dpps 0x7F
    static float Dot3(const Vector4& a, const Vector4& b) { return (a.x * b.x + a.y * b.y) + (a.z * b.z + 0.0f); }
    // Cross3: r = a.yzx * b.zxy; fnmadd(a.zxy, b.yzx, r)
    static Vector4 Cross3(const Vector4& a, const Vector4& b)
    {
        return { std::fma(-a.z, b.y, a.y * b.z), std::fma(-a.x, b.z, a.z * b.x), std::fma(-a.y, b.x, a.x * b.y), std::fma(-a.w, b.w, a.w * b.w) };
    }
    float SqrLength3() const { return Dot3(*this, *this); }
    float Length3() const { return sqrtf(Dot3(*this, *this)); }
    Vector4 Normalized3() const { const float l = sqrtf(Dot3(*this, *this)); return { x / l, y / l, z / l, w / l }; }
    // Orthogonalize(v, ref) = fnmadd(dot3(v, ref), ref, v)
    static Vector4 Orthogonalize(const Vector4& v, const Vector4& r)
    {
        const float d = Dot3(v, r);
        return { std::fma(-d, r.x, v.x), std::fma(-d, r.y, v.y), std::fma(-d, r.z, v.z), std::fma(-d, r.w, v.w) };
    }
};

struct Box
{
    Vector4 min, max;
    static Box Empty() { return { Vector4(FLT_MAX), Vector4(0.0f) - Vector4(FLT_MAX) }; }
    static Box Full() { return { Vector4(0.0f) - Vector4(FLT_MAX), Vector4(FLT_MAX) }; }
    static Box Merge(const Box& a, const Box& b) { return { Vector4::Min(a.min, b.min), Vector4::Max(a.max, b.max) }; }
    // half of the surface area, Box.h:72-76
    float SurfaceArea() const { const Vector4 s = max - min; return s.x * (s.y + s.z) + s.y * s.z; }
    float Volume() const { const Vector4 s = max - min; return s.x * s.y * s.z; }
};

struct Matrix4
{
    Vector4 rows[4];
    static Matrix4 Identity() { Matrix4 m; m.rows[0] = { 1, 0, 0, 0 }; m.rows[1] = { 0, 1, 0, 0 }; m.rows[2] = { 0, 0, 1, 0 }; m.rows[3] = { 0, 0, 0, 1 }; return m; }
    float f(int i) const { return rows[i >> 2][i & 3]; }

    // full inverse by cofactors (Matrix4.cpp:103-247): every entry is a signed sum of six triple products evaluated
    // left to right, then scaled by 1 / det
    Matrix4 Inverse() const
    {
        // term = {sign, i, j, k}: sign * f[i] * f[j] * f[k]
        static const int8_t T[16][6][4] = {
            { { 1, 5, 10, 15 }, { -1, 5, 11, 14 }, { -1, 9, 6, 15 }, { 1, 9, 7, 14 }, { 1, 13, 6, 11 }, { -1, 13, 7, 10 } },     // inv[0]
            { { -1, 1, 10, 15 }, { 1, 1, 11, 14 }, { 1, 9, 2, 15 }, { -1, 9, 3, 14 }, { -1, 13, 2, 11 }, { 1, 13, 3, 10 } },     // inv[1]
            { { 1, 1, 6, 15 }, { -1, 1, 7, 14 }, { -1, 5, 2, 15 }, { 1, 5, 3, 14 }, { 1, 13, 2, 7 }, { -1, 13, 3, 6 } },         // inv[2]
            { { -1, 1, 6, 11 }, { 1, 1, 7, 10 }, { 1, 5, 2, 11 }, { -1, 5, 3, 10 }, { -1, 9, 2, 7 }, { 1, 9, 3, 6 } },           // inv[3]
            { { -1, 4, 10, 15 }, { 1, 4, 11, 14 }, { 1, 8, 6, 15 }, { -1, 8, 7, 14 }, { -1, 12, 6, 11 }, { 1, 12, 7, 10 } },     // inv[4]
            { { 1, 0, 10, 15 }, { -1, 0, 11, 14 }, { -1, 8, 2, 15 }, { 1, 8, 3, 14 }, { 1, 12, 2, 11 }, { -1, 12, 3, 10 } },     // inv[5]
            { { -1, 0, 6, 15 }, { 1, 0, 7, 14 }, { 1, 4, 2, 15 }, { -1, 4, 3, 14 }, { -1, 12, 2, 7 }, { 1, 12, 3, 6 } },         // inv[6]
            { { 1, 0, 6, 11 }, { -1, 0, 7, 10 }, { -1, 4, 2, 11 }, { 1, 4, 3, 10 }, { 1, 8, 2, 7 }, { -1, 8, 3, 6 } },           // inv[7]
            { { 1, 4, 9, 15 }, { -1, 4, 11, 13 }, { -1, 8, 5, 15 }, { 1, 8, 7, 13 }, { 1, 12, 5, 11 }, { -1, 12, 7, 9 } },       // inv[8]
            { { -1, 0, 9, 15 }, { 1, 0, 11, 13 }, { 1, 8, 1, 15 }, { -1, 8, 3, 13 }, { -1, 12, 1, 11 }, { 1, 12, 3, 9 } },       // inv[9]
            { { 1, 0, 5, 15 }, { -1, 0, 7, 13 }, { -1, 4, 1, 15 }, { 1, 4, 3, 13 }, { 1, 12, 1, 7 }, { -1, 12, 3, 5 } },         // inv[10]
            { { -1, 0, 5, 11 }, { 1, 0, 7, 9 }, { 1, 4, 1, 11 }, { -1, 4, 3, 9 }, { -1, 8, 1, 7 }, { 1, 8, 3, 5 } },             // inv[11]
            { { -1, 4, 9, 14 }, { 1, 4, 10, 13 }, { 1, 8, 5, 14 }, { -1, 8, 6, 13 }, { -1, 12, 5, 10 }, { 1, 12, 6, 9 } },       // inv[12]
            { { 1, 0, 9, 14 }, { -1, 0, 10, 13 }, { -1, 8, 1, 14 }, { 1, 8, 2, 13 }, { 1, 12, 1, 10 }, { -1, 12, 2, 9 } },       // inv[13]
            { { -1, 0, 5, 14 }, { 1, 0, 6, 13 }, { 1, 4, 1, 14 }, { -1, 4, 2, 13 }, { -1, 12, 1, 6 }, { 1, 12, 2, 5 } },         // inv[14]
            { { 1, 0, 5, 10 }, { -1, 0, 6, 9 }, { -1, 4, 1, 10 }, { 1, 4, 2, 9 }, { 1, 8, 1, 6 }, { -1, 8, 2, 5 } },             // inv[15]
        };
        float inv[16];
        for (int e = 0; e < 16; ++e)
        {
            float acc = 0.0f;
            for (int t = 0; t < 6; ++t)
            {
                const float p = f(T[e][t][1]) * f(T[e][t][2]) * f(T[e][t][3]);
                if (t == 0) acc = T[e][t][0] > 0 ? p : -p;
                else acc = T[e][t][0] > 0 ? acc + p : acc - p;
            }
            inv[e] = acc;
        }
        float det = f(0) * inv[0] + f(1) * inv[4] + f(2) * inv[8] + f(3) * inv[12];
        det = 1.0f / det;
        Matrix4 r;
        for (int i = 0; i < 16; ++i) r.rows[i >> 2][i & 3] = inv[i] * det;
        return r;
    }

    // Matrix4::TransformBox, Matrix4.cpp:249-268
    Box TransformBox(const Box& box) const
    {
        const Vector4 xa = rows[0] * box.min.x, xb = rows[0] * box.max.x;
        const Vector4 ya = rows[1] * box.min.y, yb = rows[1] * box.max.y;
        const Vector4 za = rows[2] * box.min.z, zb = rows[2] * box.max.z;
        return { Vector4::Min(xa, xb) + Vector4::Min(ya, yb) + Vector4::Min(za, zb) + rows[3],
                 Vector4::Max(xa, xb) + Vector4::Max(ya, yb) + Vector4::Max(za, zb) + rows[3] };
    }
};

// scalar Sin / Cos approximations (Transcendental.cpp:36-52): truncating range reduction, Horner form, no FMA
inline float Sin(float x)
{
    const float c0 = 9.9999970197e-01f, c1 = -1.6666577756e-01f, c2 = 8.3325579762e-03f;
    const float c3 = -1.9812576647e-04f, c4 = 2.7040521217e-06f, c5 = -2.0532988642e-08f;
    const int32_t i = static_cast<int32_t>(x * (1.0f / RT_PI_F));
    x -= static_cast<float>(i) * RT_PI_F;
    const float x2 = x * x;
    const float y = x * (c0 + x2 * (c1 + x2 * (c2 + x2 * (c3 + x2 * (c4 + x2 * c5)))));
    return (i & 1) ? -y : y;
}
inline float Cos(float x) { return Sin(x + RT_PI_F / 2.0f); }

// Quaternion::FromEulerAngles (angles in radians: pitch = x, yaw = y, roll = z)
inline Vector4 QuaternionFromEulerAngles(float ax, float ay, float az)
{
    const float pitch = ax * 0.5f, yaw = ay * 0.5f, roll = az * 0.5f;
    const float t0 = Cos(yaw), t1 = Sin(yaw), t2 = Cos(roll), t3 = Sin(roll), t4 = Cos(pitch), t5 = Sin(pitch);
    const Vector4 term0(t0, t1, t0, t0), term1(t2, t2, t3, t2), term2(t5, t4, t4, t4);
    const Vector4 term3(t1, -t0, -t1, t1), term4(t3, t3, t2, t3), term5(t4, t5, t5, t5);
    return term0 * term1 * term2 + term3 * term4 * term5;
}

// Transform(translation, rotation).ToMatrix4(): rows = rotated axes (QuaternionImpl.h:35-63), row 3 = translation
inline Matrix4 TransformToMatrix4(const Vector4& t, const Vector4& q)
{
    Matrix4 m;
    m.rows[0] = { 1.0f - 2.0f * (q.y * q.y + q.z * q.z), 2.0f * (q.x * q.y + q.w * q.z), 2.0f * (q.x * q.z - q.w * q.y), 0.0f };
    m.rows[1] = { 2.0f * (q.x * q.y - q.w * q.z), 1.0f - 2.0f * (q.x * q.x + q.z * q.z), 2.0f * (q.y * q.z + q.w * q.x), 0.0f };
    m.rows[2] = { 2.0f * (q.x * q.z + q.w * q.y), 2.0f * (q.y * q.z - q.w * q.x), 1.0f - 2.0f * (q.x * q.x + q.y * q.y), 0.0f };
    m.rows[3] = { t.x, t.y, t.z, 1.0f };
    return m;
}

inline float DegToRad(float x) { return x / 180.0f * RT_PI_F; }      // Math.h:51-55

// BuildOrthonormalBasis, Geometry.cpp:15-33
inline void BuildOrthonormalBasis(const Vector4& n, Vector4& u, Vector4& v)
{
    union { float f; uint32_t u; } one, nz;
    one.f = 1.0f; nz.f = n.z;
    one.u = (0x7fffffffu & one.u) | (0x80000000u & nz.u);
    const float sign = one.f;
    const float a = -1.0f / (sign + n.z);
    u = Vector4(1.0f + sign * n.x * n.x * a, sign * n.x * n.y * a, -sign * n.x);
    v = Vector4(n.x * n.y * a, sign + n.y * n.y * a, -n.y);
}

} // namespace math
} // namespace rt
