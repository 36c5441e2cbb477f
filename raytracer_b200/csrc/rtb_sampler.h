// rtb_sampler.h — device GenericSampler (integer bit-exact) and primary ray generation.
//
//   GenericSampler::ResetPixel / GetInt / GetFloat        Core/Sampling/GenericSampler.cpp:56-112, GenericSampler.h:29-32
//   Viewport::RenderTile pixel -> film coordinates        Core/Rendering/Viewport.cpp:300-330
//   Camera::GenerateRay / GenerateBokeh                   Core/Scene/Camera.cpp:81-118, 186-210
#pragma once

#include "rtb_scene.h"

namespace rtb {

// per-pass constants shared by every kernel of the pass
struct PassParams
{
    // camera (Camera.cpp:81-118)
    M4       camLocalToWorld;
    float    aspectRatio, tanHalfFoV;
    uint32_t dofEnable, bokehShape;
    float    aperture, focalPlaneDistance;
    float    barrelConst, barrelVar;
    const float* bokehCdf;                     // RTB_BOKEH_TEXTURE: cumulative distribution over the bokeh bitmap's texels
    uint32_t bokehWidth, bokehHeight;
    // film
    uint32_t width, height;
    float    sampleOffsetX, sampleOffsetY;     // AA offset of pass slot 0 (used when sampleOffsets is null)
    const float* sampleOffsets;                // batched passes: 2 floats per pass slot
    // integrator (Context.h:55-90)
    uint32_t renderer, debugMode;
    uint32_t maxRayDepth, minRussianRouletteDepth;
    uint32_t lightSamplingAll, useBlueNoise;
    uint32_t passIndex, numSeedDims;
    const uint32_t* seed;         // numSeedDims Halton integers per pass slot (slot k starts at k * numSeedDims)
};

struct Sampler
{
    uint32_t salt;
    uint32_t dim;          // mSamplesGenerated
    uint32_t bnOffset;     // 4 * (128 * (y & 127) + (x & 127))
    uint32_t seedBase;     // passSlot * numSeedDims: which pass of a batch this pixel sample belongs to
};

// Math.h:269-277 (MurmurHash3 finalizer)
RTB_HD uint64_t hash64(uint64_t h)
{
    h ^= h >> 33; h *= 0xff51afd7ed558ccdULL;
    h ^= h >> 33; h *= 0xc4ceb9fe1a85ec53ULL;
    h ^= h >> 33;
    return h;
}

RTB_HD uint32_t xorshift32(uint32_t x) { x ^= x << 13; x ^= x >> 17; x ^= x << 5; return x; }   // GenericSampler.cpp:56-62

// GenericSampler::ResetPixel, GenericSampler.cpp:75-81
RTB_HD void sampler_reset_pixel(Sampler& s, uint32_t x, uint32_t y)
{
    s.bnOffset = 4u * (128u * (y & 127u) + (x & 127u));
    s.salt = (uint32_t)hash64((uint64_t)(x | (y << 16)));
    s.dim = 0;
    s.seedBase = 0;
}

// GenericSampler::GetInt, GenericSampler.cpp:83-112.  Dimensions beyond numSeedDims fall back to a per-thread RNG in
// the reference (irreproducible); here the salted stream simply continues with seed 0 and the host refuses to
// start a pass whose worst-case dimension count exceeds numSeedDims (see rtb_render_pass).
RTB_HD uint32_t sampler_get_int(Sampler& s, const PassParams& P, const uint16_t* blueNoise)
{
    uint32_t sample = s.dim < P.numSeedDims ? P.seed[s.seedBase + s.dim] : 0u;
    if (s.dim < 4u && P.useBlueNoise && blueNoise != nullptr)
    {
        sample += (uint32_t)blueNoise[s.bnOffset + s.dim] << 16;
    }
    else
    {
        const uint32_t salt = s.salt;
        s.salt = xorshift32(salt);
        sample += salt;
    }
    s.dim++;
    return sample;
}

// GenericSampler::GetFloat, GenericSampler.h:29-32
RTB_HD float sampler_get_float(Sampler& s, const PassParams& P, const uint16_t* blueNoise)
{
    return s_min(0.999999940395f, (float)sampler_get_int(s, P, blueNoise) / 4294967296.0f);
}

// Viewport::RenderTile coordinates + Camera::GenerateRay.  (x, y) is the film pixel, row 0 = top.
RTB_HD Ray generate_primary_ray(const PassParams& P, uint32_t x, uint32_t y, float sampleOffsetX, float sampleOffsetY, Sampler& s, const uint16_t* blueNoise)
{
    // coords = (FromIntegers(x, realY) + sampleOffset) * (1 / filmSize)          Viewport.cpp:297-321
    const uint32_t realY = P.height - 1u - y;
    const float invW = 1.0f / (float)(int32_t)P.width, invH = 1.0f / (float)(int32_t)P.height;
    const float cx = ((float)(int32_t)x + sampleOffsetX) * invW;
    const float cy = ((float)(int32_t)realY + sampleOffsetY) * invH;

    // UnipolarToBipolar = fmsub(x, 2, 1)                                           Vector4ImplSSE.h:598-601
    float ox = fma_(cx, 2.0f, -1.0f), oy = fma_(cy, 2.0f, -1.0f);

    if (P.barrelVar != 0.0f)
    {
        // unsupported: the variable factor is multiplied by a per-thread random number (Camera.cpp:87-92)
    }

    const M4& t = P.camLocalToWorld;
    V3 origin = xyz(t.r3);
    // direction = fmadd(fmadd(t[0], ox * aspect, t[1] * oy), tanHalfFoV, t[2])    Camera.cpp:95-99
    const float sx = ox * P.aspectRatio;
    const V3 inner = mul_add(xyz(t.r0), sx, xyz(t.r1) * oy);
    V3 direction = mul_add(inner, P.tanHalfFoV, xyz(t.r2));

    if (P.dofEnable)
    {
        const V3 focusPoint = mul_add(direction, P.focalPlaneDistance, origin);
        const float u0 = sampler_get_float(s, P, blueNoise);
        const float u1 = sampler_get_float(s, P, blueNoise);
        float bx, by;
        if (P.bokehShape == RTB_BOKEH_TEXTURE)
        {
            // Camera::GenerateBokeh, BokehShape::Texture (Camera.cpp:208-213) -> BitmapTexture::Sample (BitmapTexture.cpp:95-119):
            // a texel drawn from the importance map with the FIRST sample component only (Distribution::SampleDiscrete,
            // Distribution.cpp:86-110), its integer coordinates divided by the bitmap size, UnipolarToBipolar
            uint32_t low = 0u, high = P.bokehWidth * P.bokehHeight;
            while (low < high)
            {
                const uint32_t mid = (low + high) / 2u;
                if (u0 >= P.bokehCdf[mid]) low = mid + 1u; else high = mid;
            }
            const uint32_t pixelIndex = low - 1u;
            const uint32_t px = pixelIndex % P.bokehWidth, py = pixelIndex / P.bokehWidth;
            bx = fma_(fdiv((float)px, (float)P.bokehWidth), 2.0f, -1.0f);
            by = fma_(fdiv((float)py, (float)P.bokehHeight), 2.0f, -1.0f);
        }
        else if (P.bokehShape == RTB_BOKEH_HEXAGON) sample_hexagon(u0, u1, 0.0f, bx, by);
        else if (P.bokehShape == RTB_BOKEH_SQUARE) { bx = fma_(u0, 2.0f, -1.0f); by = fma_(u1, 2.0f, -1.0f); }
        else sample_circle(u0, u1, bx, by);
        bx = bx * P.aperture; by = by * P.aperture;
        origin = mul_add(xyz(t.r0), bx, origin);    // fmadd(splatX, right, origin)
        origin = mul_add(xyz(t.r1), by, origin);
        direction = focusPoint - origin;
    }
    return make_ray(origin, direction);
}

} // namespace rtb
