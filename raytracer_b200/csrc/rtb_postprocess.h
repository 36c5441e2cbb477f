// rtb_postprocess.h — one film pixel -> one BGRA8 front-buffer pixel (Viewport::PostProcessTile, Rendering/Viewport.cpp:492-552).
//
// Order of operations, as in the reference: [bloom mix 511-525] -> scale by 1 / passes (499, 529) -> saturation (532-533)
// -> contrast through FastLog / FastExp (536) -> exposure and colour filter (539) -> tone-mapping curve + sRGB
// approximation (541; Color/ColorHelpers.h:30-126) -> dither (545) -> Vector4::ToBGR (Math/Vector4ImplSSE.h:89-110).
// Every Vector4::MulAndAdd of the reference is one fused multiply-add here, everything else rounds per operation.
// Compiled by nvcc for k_front and by g++ for the CPU tests (tests/hostsim).
#pragma once

#include "rtb_math.h"
#include "../../include/rtb200.h"

namespace rtb {

struct PostParams
{
    V3       colorScale;          // colorFilter * 2^exposure (Viewport.cpp:453), computed by the caller on the host
    float    pixelScaling;        // 1 / passes in the film
    float    contrast, saturation, ditheringStrength, bloomFactor;
    uint32_t tonemapper;          // RtbTonemapper
    uint32_t ditherSeed;
};

// _mm_cvttps_epi32: truncate; NaN and out-of-range give INT_MIN      VectorInt4ImplSSE.h:53-56
RTB_HD int32_t cvt_rz_sse(float x)
{
    if (!(x > -2147483904.0f && x < 2147483648.0f)) return (int32_t)0x80000000u;
    return (int32_t)x;
}

// _mm_cvtps_epi32 with the same out-of-range rule
RTB_HD int32_t cvt_rn_sse(float x)
{
    if (!(x > -2147483904.0f && x < 2147483648.0f)) return (int32_t)0x80000000u;
    return cvt_rn(x);
}

// FastLog(Vector4), one lane                                         Math/Transcendental.cpp:216-232
RTB_HD float fast_log_lane(float a)
{
    const uint32_t e = (f2bits(a) - 0x3f2aaaabu) & 0xff800000u;
    const float m = bits2f(f2bits(a) - e);
    const float i = (float)(int32_t)e * 1.19209290e-7f;
    const float f = m - 1.0f;
    const float s = f * f;
    float r = fma_(f, 0.230836749f, -0.279208571f);
    const float t = fma_(f, 0.331826031f, -0.498910338f);
    r = fma_(r, s, t);
    r = fma_(r, s, f);
    return fma_(i, 0.693147182f, r);
}

// FastExp(Vector4), one lane                                         Math/Transcendental.cpp:147-165
RTB_HD float fast_exp_lane(float a)
{
    const float t = a * 1.442695041f;
    const float fi = floorf(t);
    const int32_t i = cvt_rn_sse(fi);
    const float f = t - fi;
    float y = fma_(f, 0.3371894346f, 0.657636276f);
    y = fma_(f, y, 1.00172476f);
    y = bits2f(f2bits(y) + ((uint32_t)i << 23));
    if ((0.0f - a) >= 87.0f) y = 0.0f;
    if (a >= 87.0f) y = bits2f(0x7f800000u);
    return y;
}

// Convert_Linear_To_sRGB, one lane                                   Color/ColorHelpers.h:30-45
RTB_HD float linear_to_srgb_lane(float c)
{
    const float s1 = sqrtf(c);
    const float s2 = sqrtf(s1);
    const float s3 = sqrtf(s2);
    float r = 0.585122381f * s1;
    r = fma_(s2, 0.783140355f, r);
    r = fma_(s3, -0.368262736f, r);
    return sse_min(1.0f, sse_max(0.0f, r));
}

// ToneMap<Vector4>, one lane                                         Color/ColorHelpers.h:85-126
RTB_HD float tone_map_lane(float c, uint32_t tonemapper)
{
    switch (tonemapper)
    {
    case RTB_TONEMAP_CLAMPED:
        return linear_to_srgb_lane(c);
    case RTB_TONEMAP_REINHARD:
        return linear_to_srgb_lane(c / (1.0f + c));
    case RTB_TONEMAP_JIM_HEJL:
    {
        const float t0 = c * fma_(c, 6.2f, 0.5f);
        const float t1 = fma_(c, 6.2f, 1.7f);
        const float t2 = fma_(c, t1, 0.06f);
        return t0 * fdiv(1.0f, t2);
    }
    default: // RTB_TONEMAP_ACES
    {
        const float t0 = c * fma_(c, 2.51f, 0.03f);
        const float t1 = fma_(c, 2.43f, 0.59f);
        const float t2 = fma_(c, t1, 0.14f);
        return linear_to_srgb_lane(t0 * fdiv(1.0f, t2));
    }
    }
}

// Random::GetVector4Bipolar's mapping of 23 random bits to [-1, 1)   Math/Random.cpp:128-138
RTB_HD float bipolar_from_bits(uint32_t bits) { return bits2f((bits & 0x007fffffu) | 0x40000000u) - 3.0f; }

// The reference dithers with its per-thread xorshift generators (Viewport.cpp:494, 545), so its dither pattern depends
// on which thread drew which tile; here it is a hash of (pixel, channel, seed): same distribution, reproducible.
RTB_HD uint32_t dither_bits(uint32_t pixel, uint32_t channel, uint32_t seed)
{
    uint32_t h = pixel * 3u + channel;
    h ^= seed * 0x9e3779b9u;
    h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;      // murmur3 finaliser
    return h;
}

// ---- bloom: Bitmap::GaussianBlur (Utils/Bitmap.cpp:917-1012) ------------------------------------------------------
//
// The reference approximates a Gaussian by repeated box blurs with running sums (BoxBlur_Internal, Bitmap.cpp:880-915).
// A running sum is order-dependent in floating point, so a line is blurred by ONE thread, element by element, in the
// reference's order; the parallelism is across lines and colour channels (the lanes of the reference's Vector4 are
// independent).  What the reference's pointer swapping amounts to (Bitmap.cpp:964-977 and 993-1009): of the n = 8 box
// blurs it runs per line, the horizontal pass keeps the result of the 7th, the vertical pass that of the 8th.

#define RTB_BLOOM_LEVELS 5           // Viewport.cpp:25
#define RTB_BLUR_ITERATIONS 8        // Viewport.cpp:442
#define RTB_BLUR_KEPT_HORIZONTAL 7
#define RTB_BLUR_KEPT_VERTICAL 8
#ifndef RTB_BLUR_BATCH
#define RTB_BLUR_BATCH 64u          // loads in flight per thread and batch (k_blur per launch at 1080p: 16: 0.60 ms, 64: 0.52, 128: 0.65)
#endif
#if defined(__CUDACC__)
#define RTB_UNROLL _Pragma("unroll")
#else
#define RTB_UNROLL
#endif

struct BlurSchedule { uint32_t radius[RTB_BLUR_ITERATIONS]; };

// box sizes of GaussianBlur(sigma, n) (Bitmap.cpp:935-944, 969); the reference passes them to BoxBlur_Internal as radii
inline BlurSchedule blur_schedule(const float sigma, const uint32_t n = RTB_BLUR_ITERATIONS)
{
    const float wIdeal = sqrtf((12.0f * sigma * sigma / n) + 1.0f);
    uint32_t wl = (uint32_t)floorf(wIdeal);
    if (wl % 2 == 0) wl--;
    const uint32_t wu = wl + 2;
    const float mIdeal = (12.0f * sigma * sigma - n * wl * wl - 4.0f * n * wl - 3.0f * n) / (-4.0f * wl - 4.0f);
    const float m = roundf(mIdeal);
    BlurSchedule s;
    for (uint32_t i = 0; i < RTB_BLUR_ITERATIONS; ++i) s.radius[i] = i < m ? wl : wu;
    return s;
}

// BoxBlur_Internal for one lane of one line: src(j) reads, dst(j, value) writes; needs width > 2 * radius
template <typename Src, typename Dst>
RTB_HD void box_blur_line(const Src& src, const Dst& dst, const uint32_t radius, const uint32_t width)
{
    const float factor = 1.0f / (float)(2u * radius + 1u);
    const float first = src(0u), last = src(width - 1u);
    float val = first * (float)(radius + 1u);
    uint32_t begin = 0u, end = 0u, out = 0u;
    for (uint32_t j = 0; j < radius; ++j) val += src(begin++);
    for (uint32_t j = 0; j <= radius; ++j) { val += src(begin++) - first; dst(out++, val * factor); }
    // (src and dst never alias: the loads of a batch are issued together, only the running sum is sequential.  Issuing the
    //  NEXT batch's loads before this batch's sums - a software pipeline - was measured slower, 0.73 against 0.60 ms per launch)
    uint32_t j = radius + 1u;
    const uint32_t stop = width - radius;
    for (; j + RTB_BLUR_BATCH <= stop; j += RTB_BLUR_BATCH)
    {
        float diff[RTB_BLUR_BATCH];
RTB_UNROLL
        for (uint32_t k = 0; k < RTB_BLUR_BATCH; ++k) diff[k] = src(begin + k) - src(end + k);
        begin += RTB_BLUR_BATCH; end += RTB_BLUR_BATCH;
RTB_UNROLL
        for (uint32_t k = 0; k < RTB_BLUR_BATCH; ++k) { val += diff[k]; dst(out++, val * factor); }
    }
    for (; j < stop; ++j) { val += src(begin++) - src(end++); dst(out++, val * factor); }
    for (uint32_t j = width - radius; j < width; ++j) { val += last - src(end++); dst(out++, val * factor); }
}

// One line (row or column) of one colour channel through the kept box blurs: element j of the line lives at
// image[j * elementStride] on entry (`in`) and exit (`out`); between the blurs it ping-pongs through two scratch lines
// laid out [j * scratchStride] (scratchStride = number of lines in flight, so that neighbouring threads coalesce).
RTB_HD void blur_line(const float* in, float* out, const size_t elementStride, float* scratchA, float* scratchB, const size_t scratchStride,
                      const BlurSchedule& schedule, const uint32_t kept, const uint32_t width)
{
    for (uint32_t i = 0; i < kept; ++i)
    {
        const bool firstBlur = i == 0u, lastBlur = i + 1u == kept;
        const float* s = firstBlur ? in : ((i & 1u) ? scratchA : scratchB);
        float* d = lastBlur ? out : ((i & 1u) ? scratchB : scratchA);
        const size_t ss = firstBlur ? elementStride : scratchStride;
        const size_t ds = lastBlur ? elementStride : scratchStride;
        box_blur_line([s, ss](uint32_t j) { return s[j * ss]; }, [d, ds](uint32_t j, float v) { d[j * ds] = v; }, schedule.radius[i], width);
    }
}

// the bloom colour of one pixel from the blurred images (Viewport.cpp:517-523); levelStride floats between levels
RTB_HD V3 bloom_color(const float* blurred, const size_t levelStride, const size_t pixel)
{
    const float weights[RTB_BLOOM_LEVELS] = { 0.35f, 0.25f, 0.15f, 0.15f, 0.1f };
    V3 b = v3s(0.0f);
    for (uint32_t i = 0; i < RTB_BLOOM_LEVELS; ++i)
    {
        const float* p = blurred + i * levelStride + pixel * 3;
        b = mul_add(v3(p[0], p[1], p[2]), weights[i], b);
    }
    return b;
}

// ---- adaptive rendering: Viewport::ComputeBlockError (Viewport.cpp:552-582) ----------------------------------------

// the error estimate of one pixel: |image - image of the even passes| weighted (1, 2, 1), relative to the brightness
RTB_HD float pixel_error(const V3& sum, const V3& secondarySum, const float imageScalingFactor)
{
    const V3 a = sum * imageScalingFactor;
    const V3 b = secondarySum * (2.0f * imageScalingFactor);
    const V3 diff = v3(bit_abs(a.x - b.x), bit_abs(a.y - b.y), bit_abs(a.z - b.z));
    return fdiv(diff.x + 2.0f * diff.y + diff.z, sqrtf(0.000001f + a.x + 2.0f * a.y + a.z));
}

// the reference adds the pixel errors of a block row from left to right, then the row sums from top to bottom
RTB_HD float row_error(const float* pixelErrors, const uint32_t minX, const uint32_t maxX)
{
    float rowError = 0.0f;
    for (uint32_t x = minX; x < maxX; ++x) rowError += pixelErrors[x];
    return rowError;
}

RTB_HD float block_error(const float* rowErrors, const uint32_t numRows, const uint32_t blockWidth, const uint32_t totalArea)
{
    float totalError = 0.0f;
    for (uint32_t y = 0; y < numRows; ++y) totalError += rowErrors[y];
    const uint32_t blockArea = blockWidth * numRows;
    return fdiv(totalError * sqrtf(fdiv((float)blockArea, (float)totalArea)), (float)blockArea);
}

// `sum` = the film's accumulated colour; `bloom` = sum of the weighted blurred images (Viewport.cpp:517-523), unused
// when bloomFactor == 0; `pixelScaling` = 1 / the passes this pixel has received (P.pixelScaling unless adaptive
// rendering stopped sampling it earlier).  Returns 0x00RRGGBB.
RTB_HD uint32_t postprocess_pixel(const PostParams& P, V3 sum, const V3& bloom, uint32_t pixel, float pixelScaling)
{
    V3 c = sum;
    if (P.bloomFactor > 0.0f)
    {
        c = c * (1.0f - P.bloomFactor);
        c = mul_add(bloom, P.bloomFactor, c);
    }
    c = c * pixelScaling;

    const float gray = dot3(c, v3(0.2126f, 0.7152f, 0.0722f));
    c = v3(sse_max(0.0f, fma_(c.x - gray, P.saturation, gray)),
           sse_max(0.0f, fma_(c.y - gray, P.saturation, gray)),
           sse_max(0.0f, fma_(c.z - gray, P.saturation, gray)));

    c = v3(fast_exp_lane(fast_log_lane(c.x) * P.contrast),
           fast_exp_lane(fast_log_lane(c.y) * P.contrast),
           fast_exp_lane(fast_log_lane(c.z) * P.contrast));

    c = c * P.colorScale;

    float t[3] = { tone_map_lane(c.x, P.tonemapper), tone_map_lane(c.y, P.tonemapper), tone_map_lane(c.z, P.tonemapper) };

    uint32_t out = 0u;
    for (uint32_t k = 0; k < 3u; ++k)
    {
        // MulAndAdd(random, strength, toneMapped): with strength 0 this is toneMapped itself
        const float d = P.ditheringStrength != 0.0f ? fma_(bipolar_from_bits(dither_bits(pixel, k, P.ditherSeed)), P.ditheringStrength, t[k]) : t[k];
        int32_t q = cvt_rz_sse(d * 255.0f);
        q = q < 0 ? 0 : (q > 255 ? 255 : q);
        out |= (uint32_t)q << (16u - 8u * k);       // x -> bits 16..23, y -> 8..15, z -> 0..7
    }
    return out;
}

} // namespace rtb
