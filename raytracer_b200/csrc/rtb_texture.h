// rtb_texture.h — ITexture::Evaluate for the four texture kinds the reference's scene loader can create
// (Demo/SceneLoader.cpp:269-361), restated operation for operation:
//   BitmapTexture::Evaluate        Core/Textures/BitmapTexture.cpp:32-93   wrap addressing, texel (0,0) at uv (0,0) with no
//                                  half-texel offset, bilinear with smooth-step warped weights (the default filter, :22)
//   CheckerboardTexture::Evaluate  Core/Textures/CheckerboardTexture.cpp:32-41
//   NoiseTexture::Evaluate         Core/Textures/NoiseTexture.cpp:59-162   2-D simplex noise octaves
//   MixTexture::Evaluate           Core/Textures/MixTexture.cpp:24-31
// and the places the integrators consume them:
//   MaterialParameter::Evaluate    Core/Material/MaterialParameter.h:22-32  value = baseValue * texture(uv)
//   Material::GetNormalVector      Core/Material/Material.cpp:120-138
//   BackgroundLight::GetBackgroundColor  Core/Scene/Light/BackgroundLight.cpp:45-61
// Bitmap texels are the float quadruples Bitmap::GetPixel returns (decoded + linearised by the host, rtb200.h), so
// the texture units are NOT used: their 8-bit-fraction interpolation could not reproduce the reference's weights.
#pragma once

#include "rtb_scene.h"

namespace rtb {

// c_permutationTable, NoiseTexture.cpp:10-27 (Ken Perlin's reference permutation)
#define RTB_PERM_TABLE { \
    151, 160, 137, 91, 90, 15, 131, 13, 201, 95, 96, 53, 194, 233, 7, 225, 140, 36, 103, 30, 69, 142, 8, 99, 37, 240, 21, 10, 23, \
    190, 6, 148, 247, 120, 234, 75, 0, 26, 197, 62, 94, 252, 219, 203, 117, 35, 11, 32, 57, 177, 33, \
    88, 237, 149, 56, 87, 174, 20, 125, 136, 171, 168, 68, 175, 74, 165, 71, 134, 139, 48, 27, 166, \
    77, 146, 158, 231, 83, 111, 229, 122, 60, 211, 133, 230, 220, 105, 92, 41, 55, 46, 245, 40, 244, \
    102, 143, 54, 65, 25, 63, 161, 1, 216, 80, 73, 209, 76, 132, 187, 208, 89, 18, 169, 200, 196, \
    135, 130, 116, 188, 159, 86, 164, 100, 109, 198, 173, 186, 3, 64, 52, 217, 226, 250, 124, 123, \
    5, 202, 38, 147, 118, 126, 255, 82, 85, 212, 207, 206, 59, 227, 47, 16, 58, 17, 182, 189, 28, 42, \
    223, 183, 170, 213, 119, 248, 152, 2, 44, 154, 163, 70, 221, 153, 101, 155, 167, 43, 172, 9, \
    129, 22, 39, 253, 19, 98, 108, 110, 79, 113, 224, 232, 178, 185, 112, 104, 218, 246, 97, 228, \
    251, 34, 242, 193, 238, 210, 144, 12, 191, 179, 162, 241, 81, 51, 145, 235, 249, 14, 239, 107, \
    49, 192, 214, 31, 181, 199, 106, 157, 184, 84, 204, 176, 115, 121, 50, 45, 127, 4, 150, 254, \
    138, 236, 205, 93, 222, 114, 67, 29, 24, 72, 243, 141, 128, 195, 78, 66, 215, 61, 156, 180 }

#if defined(__CUDACC__)
static __device__ __constant__ uint8_t kNoisePermDevice[256] = RTB_PERM_TABLE;
#endif
static const uint8_t kNoisePermHost[256] = RTB_PERM_TABLE;

RTB_HD int32_t noise_hash(int32_t i)
{
#if defined(__CUDA_ARCH__)
    return kNoisePermDevice[(uint8_t)i];
#else
    return kNoisePermHost[(uint8_t)i];
#endif
}

// utils::Gradient, NoiseTexture.cpp:34-40
RTB_HD float noise_gradient(int32_t hash, float x, float y)
{
    const int32_t h = hash & 0x3F;
    const float u = h < 4 ? x : y;
    const float v = h < 4 ? y : x;
    return ((h & 1) ? -u : u) + ((h & 2) ? -2.0f * v : 2.0f * v);
}

RTB_HD int32_t floor_int(float fp) { const int32_t i = (int32_t)fp; return fp < (float)i ? (i - 1) : i; }    // Math.h:100-104

// NoiseTexture::EvaluateInternal, NoiseTexture.cpp:59-139
RTB_HD float simplex_noise(float cx, float cy)
{
    const float F2 = 0.366025403f, G2 = 0.211324865f;
    const float s = (cx + cy) * F2;
    const float xs = cx + s, ys = cy + s;
    const int32_t i = floor_int(xs), j = floor_int(ys);
    const float t = (float)(i + j) * G2;
    const float X0 = (float)i - t, Y0 = (float)j - t;
    const float x0 = cx - X0, y0 = cy - Y0;
    int32_t i1, j1;
    if (x0 > y0) { i1 = 1; j1 = 0; } else { i1 = 0; j1 = 1; }
    const float x1 = x0 - (float)i1 + G2, y1 = y0 - (float)j1 + G2;
    const float x2 = x0 - 1.0f + 2.0f * G2, y2 = y0 - 1.0f + 2.0f * G2;
    const int32_t gi0 = noise_hash(i + noise_hash(j));
    const int32_t gi1 = noise_hash(i + i1 + noise_hash(j + j1));
    const int32_t gi2 = noise_hash(i + 1 + noise_hash(j + 1));
    float n0, n1, n2;
    float t0 = 0.5f - x0 * x0 - y0 * y0;
    if (t0 < 0.0f) n0 = 0.0f; else { t0 *= t0; n0 = t0 * t0 * noise_gradient(gi0, x0, y0); }
    float t1 = 0.5f - x1 * x1 - y1 * y1;
    if (t1 < 0.0f) n1 = 0.0f; else { t1 *= t1; n1 = t1 * t1 * noise_gradient(gi1, x1, y1); }
    float t2 = 0.5f - x2 * x2 - y2 * y2;
    if (t2 < 0.0f) n2 = 0.0f; else { t2 *= t2; n2 = t2 * t2 * noise_gradient(gi2, x2, y2); }
    return s_clamp(0.5f + 22.615325f * (n0 + n1 + n2), 0.0f, 1.0f);
}

RTB_HD V4 lerp4(const V4& a, const V4& b, const V4& w) { return mul_add(b - a, w, a); }      // Vector4::Lerp, Vector4Impl.h:58-66
RTB_HD V4 lerp4(const V4& a, const V4& b, float w) { return mul_add(b - a, w, a); }

RTB_HD float smooth_step(float x) { return x * x * (3.0f - x * 2.0f); }                     // Math.h:176-179

RTB_HD V4 load_texel(const DTexture& t, int32_t x, int32_t y)
{
    const float* p = t.texels + 4u * ((size_t)y * t.width + (size_t)x);
#if defined(__CUDA_ARCH__)
    const float4 c = __ldg(reinterpret_cast<const float4*>(p));
    return v4(c.x, c.y, c.z, c.w);
#else
    return v4(p[0], p[1], p[2], p[3]);
#endif
}

// bitmap / checkerboard / noise
RTB_HD_NOINLINE V4 texture_evaluate_leaf(const DTexture& t, float u, float v)
{
    if (t.type == RTB_TEXTURE_BITMAP)
    {
        // BitmapTexture::Evaluate, BitmapTexture.cpp:32-93 (filter = Bilinear_SmoothStep)
        const int32_t w = (int32_t)t.width, h = (int32_t)t.height;
        const float wu = u - floorf(u), wv = v - floorf(v);                 // Vector4::Mod1
        const float su = wu * (float)w, sv = wv * (float)h;                 // * Bitmap::mFloatSize
        const int32_t iu = cvt_rn(floorf(su)), iv = cvt_rn(floorf(sv));     // VectorInt4::Convert(Vector4::Floor(..))
        int32_t x0 = iu, y0 = iv;
        if (!(iu < w)) x0 -= w;
        if (!(iv < h)) y0 -= h;
        if (iu < 0) x0 += w;
        if (iv < 0) y0 += h;
        int32_t x1 = x0 + 1, y1 = y0 + 1;
        if (!(x1 < w)) x1 -= w;
        if (!(y1 < h)) y1 -= h;
        // Bitmap::GetPixelBlock order: (x0,y0) (x1,y0) (x0,y1) (x1,y1)
        const V4 c0 = load_texel(t, x0, y0), c1 = load_texel(t, x1, y0), c2 = load_texel(t, x0, y1), c3 = load_texel(t, x1, y1);
        const float weightX = smooth_step(su - (float)iu), weightY = smooth_step(sv - (float)iv);
        const V4 value0 = lerp4(c0, c2, weightY);
        const V4 value1 = lerp4(c1, c3, weightY);
        return lerp4(value0, value1, weightX);
    }
    if (t.type == RTB_TEXTURE_CHECKERBOARD)
    {
        // CheckerboardTexture::Evaluate, CheckerboardTexture.cpp:32-41
        const float wu = u - floorf(u), wv = v - floorf(v);
        const bool a = (wu > 0.5f) != (wv > 0.5f);
        return a ? t.colorA : t.colorB;
    }
    // NoiseTexture::Evaluate, NoiseTexture.cpp:141-155
    float value = 0.0f;
    float octaveValueScale = 0.5f, octaveCoordScale = 1.0f;
    for (uint32_t i = 0; i < t.numOctaves; ++i)
    {
        value += octaveValueScale * simplex_noise(u * octaveCoordScale, v * octaveCoordScale);
        octaveValueScale *= 0.5f;
        octaveCoordScale *= 2.0f;
    }
    return lerp4(t.colorA, t.colorB, value);
}

// ITexture::Evaluate.  Mix textures (Lerp(A(uv), B(uv), W(uv)), MixTexture.cpp:24-31) are walked with an explicit frame
// stack (no recursion on the device); rtb_upload_scene guarantees nesting <= RTB_MAX_MIX_DEPTH and children < parent.
RTB_HD_NOINLINE V4 texture_evaluate(const SceneView& S, uint32_t textureIndex, float u, float v)
{
    const DTexture& root = S.textures[textureIndex];
    if (root.type != RTB_TEXTURE_MIX) return texture_evaluate_leaf(root, u, v);

    struct Frame { uint32_t tex, stage; V4 a, b; };
    Frame frames[RTB_MAX_MIX_DEPTH + 1];
    uint32_t sp = 0;
    frames[sp].tex = textureIndex; frames[sp].stage = 0; sp++;
    V4 ret = v4s(0.0f);
    bool haveRet = false;
    while (sp > 0)
    {
        Frame& f = frames[sp - 1];
        const DTexture& t = S.textures[f.tex];
        if (t.type != RTB_TEXTURE_MIX) { ret = texture_evaluate_leaf(t, u, v); haveRet = true; sp--; continue; }
        if (haveRet)
        {
            if (f.stage == 1) f.a = ret;
            else if (f.stage == 2) f.b = ret;
            else { ret = lerp4(f.a, f.b, ret); sp--; continue; }
            haveRet = false;
        }
        const uint32_t child = f.stage == 0 ? t.texA : (f.stage == 1 ? t.texB : t.texWeight);
        f.stage++;
        if (sp > RTB_MAX_MIX_DEPTH) { ret = v4s(0.0f); haveRet = true; continue; }    // unreachable after upload validation
        frames[sp].tex = child; frames[sp].stage = 0; sp++;
    }
    return ret;
}

// Material::GetNormalVector, Material.cpp:120-138 (called only when the material has a normal map, Scene.cpp:328)
RTB_HD V3 material_normal_vector(const SceneView& S, const RtbMaterialTextures& mt, float u, float v)
{
    V4 n = texture_evaluate(S, mt.normalMap, u, v);
    n = v4(fma_(n.x, 2.0f, -1.0f), fma_(n.y, 2.0f, -1.0f), fma_(n.z, 2.0f, -1.0f), fma_(n.w, 2.0f, -1.0f));     // UnipolarToBipolar
    n.z = sqrtf(s_max(0.0f, 1.0f - dot2(n.x, n.y, n.x, n.y)));
    const V4 z = v4(0.0f, 0.0f, 1.0f, 0.0f);
    n = lerp4(z, n, mt.normalMapStrength);
    return v3(n.x, n.y, n.z);
}

} // namespace rtb
