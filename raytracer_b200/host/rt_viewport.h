// rt_viewport.h — ViewportB200: the reference's Viewport public interface (Core/Rendering/Viewport.h:37-58) re-hosted
// on the CUDA path.  Render() reproduces the host prologue of Viewport::Render in the same order (Halton NextSample ->
// per-pass seed vector, Viewport.cpp:200-205; Gaussian AA offset, 235-242) and then enqueues one rtb_render_pass.
//
//   HaltonSequence   Core/Sampling/HaltonSampler.cpp:84-206 (scrambled Halton, double precision, host only)
//   Random           Core/Math/Random.cpp:33-67 (xoroshiro128+); seeded from a splitmix64 stream of a user seed
//                    instead of /dev/urandom so that renders are reproducible
//   GetFloatNormal2  Core/Math/SamplingHelpers.cpp:145-149 (Box-Muller with FastLog + polynomial sin)
#pragma once

#include <cmath>
#include <cstdio>
#include <limits>
#include <algorithm>
#include <string>
#include <vector>

#include "rt_scene.h"
#include "../csrc/rtb_math.h"

namespace rt {

// The seed source of every generator: Entropy::GetInt (Utils/Entropy.cpp:65-85).  The reference reads /dev/urandom; its
// fixed-seed oracle build (oracle/patch_ref.py P8, RT_FIXED_SEED) hands out the upper halves of a splitmix64 stream
// instead, and so does this class, so that a seed reproduces the oracle's renders.
class EntropyStream
{
public:
    explicit EntropyStream(uint64_t seed) : mState(seed) {}
    uint32_t GetInt()
    {
        mState += 0x9E3779B97F4A7C15ULL;
        uint64_t z = mState;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
        z = z ^ (z >> 31);
        return (uint32_t)(z >> 32);
    }
private:
    uint64_t mState;
};

class Random
{
public:
    Random() { mSeed[0] = 0x853c49e6748fea9bULL; mSeed[1] = 0xda3e39cb94b95bdbULL; }
    // Random::Reset (Math/Random.cpp:19-32): per half of the state, 2 integers for the scalar generator, then 4 + 8
    // for the SIMD generators (which the hot path's host side never uses - drawn and dropped to stay in step)
    void Reset(EntropyStream& entropy)
    {
        for (int i = 0; i < 2; ++i)
        {
            const uint64_t hi = entropy.GetInt();
            const uint64_t lo = entropy.GetInt();
            mSeed[i] = (hi << 32) | lo;
            // mSeedSimd4[i] = VectorInt4(x, y, z, w): GCC evaluates the four GetInt() arguments right to left
            uint32_t lane[4];
            for (int k = 3; k >= 0; --k) lane[k] = entropy.GetInt();
            mSeedSimd4[i][0] = ((uint64_t)lane[1] << 32) | lane[0];
            mSeedSimd4[i][1] = ((uint64_t)lane[3] << 32) | lane[2];
            for (int k = 0; k < 8; ++k) entropy.GetInt();      // mSeedSimd8[i]
        }
    }
    // `count` GetVector4() draws whose values nobody reads (Viewport::Resize draws one per pixel, Viewport.cpp:99-102)
    void SkipVector4(uint64_t count)
    {
        float x, y;
        for (uint64_t i = 0; i < count; ++i) GetFloat2(x, y);
    }
    // Random::GetFloat2 (Random.cpp:65-68, 83-126): lanes x, y of the 4-wide xorshift128+ generator, as floats in [0, 1)
    void GetFloat2(float& x, float& y)
    {
        const uint64_t s0 = mSeedSimd4[1][0], s0b = mSeedSimd4[1][1];
        uint64_t s1 = mSeedSimd4[0][0], s1b = mSeedSimd4[0][1];
        const uint64_t v = s0 + s1;
        s1 <<= 23; s1b <<= 23;
        mSeedSimd4[0][0] = s0; mSeedSimd4[0][1] = s0b;
        mSeedSimd4[1][0] = (s0 ^ s1) ^ ((s0 >> 5) ^ (s1 >> 18));
        mSeedSimd4[1][1] = (s0b ^ s1b) ^ ((s0b >> 5) ^ (s1b >> 18));
        union { uint32_t u; float f; } a, b;
        a.u = ((uint32_t)v & 0x007fffffu) | 0x3f800000u;
        b.u = ((uint32_t)(v >> 32) & 0x007fffffu) | 0x3f800000u;
        x = a.f - 1.0f; y = b.f - 1.0f;
    }
    // xoroshiro128+, Random.cpp:33-47
    uint64_t GetLong()
    {
        const uint64_t s0 = mSeed[0];
        uint64_t s1 = mSeed[1];
        const uint64_t result = s0 + s1;
        s1 ^= s0;
        mSeed[0] = ((s0 << 24) | (s0 >> 40)) ^ s1 ^ (s1 << 16);
        mSeed[1] = (s1 << 37) | (s1 >> 27);
        return result;
    }
    uint32_t GetInt() { return static_cast<uint32_t>(GetLong()); }
    // Random.cpp:54-59
    float GetFloat() { union { uint32_t u; float f; } r; r.u = (GetInt() & 0x007fffffu) | 0x3f800000u; return r.f - 1.0f; }
    double GetDouble() { return static_cast<double>(GetLong()) / static_cast<double>(UINT64_MAX); }
private:
    uint64_t mSeed[2];
    uint64_t mSeedSimd4[2][2] = {};      // two 64-bit lanes per half
};

// scrambled Halton sequence with random start (HaltonSampler.cpp): digit expansion per dimension in base p_d, digits
// permuted by a per-dimension random permutation, radical inverse accumulated in double precision
class HaltonSequence
{
public:
    static constexpr uint32_t Width = 64;

    // HaltonSequence::mRandom (HaltonSampler.h:45) is seeded once, when the sequence is constructed; Initialize keeps
    // drawing from it, so a second Initialize gives a different scramble (as in the reference)
    void Seed(EntropyStream& entropy) { mRandom.Reset(entropy); }

    void Initialize(uint32_t dims)
    {
        Random& rng = mRandom;
        mDims = dims;
        mBase.assign(dims, 0); mStarts.assign(dims, 0);
        mRnd.assign(dims, std::vector<double>(Width + 1, 0.0));
        mDigit.assign(dims, std::vector<uint64_t>(Width + 1, 0));
        mPower.assign(dims, std::vector<uint64_t>(Width, 0));
        mPerm.assign(dims, std::vector<uint64_t>());
        if (dims == 0) return;
        // InitPrimes
        uint32_t prime = 1, m = 0;
        while (m < dims)
        {
            prime++;
            bool isPrime = true;
            for (uint64_t i = 2; i <= (uint64_t)std::sqrt((double)prime); i++) if (prime % i == 0) { isPrime = false; break; }
            if (isPrime) mBase[m++] = prime;
        }
        // InitStart: random starting index, built digit by digit from a uniform double
        for (uint32_t i = 0; i < dims; i++)
        {
            double r = rng.GetDouble();
            const uint64_t base = mBase[i];
            uint64_t z = 0, b = base;
            while (r > 1.0e-16)
            {
                if (r >= 1.0 / b)
                {
                    const uint64_t cnt = (uint64_t)std::floor(r * b);
                    r = r - cnt * 1.0 / b;
                    z += cnt * b / base;
                }
                b *= base;
            }
            mStarts[i] = z;
        }
        // InitPowerBuffer
        for (uint32_t d = 0; d < dims; d++)
            for (uint32_t j = 0; j < Width; j++) mPower[d][j] = j == 0 ? mBase[d] : mPower[d][j - 1] * mBase[d];
        // InitPermutation: random transposition shuffle that keeps digit 0 fixed unless swapped explicitly
        for (uint32_t i = 0; i < dims; i++)
        {
            mPerm[i].resize(mBase[i]);
            for (uint64_t j = 0; j < mBase[i]; j++) mPerm[i][j] = j;
            for (uint64_t j = 1; j < mBase[i]; j++)
            {
                const uint64_t tmp = (uint64_t)std::floor(rng.GetDouble() * mBase[i]);
                if (tmp != 0 && tmp < mBase[i]) std::swap(mPerm[i][j], mPerm[i][tmp]);
            }
        }
        // InitExpansion
        for (uint32_t i = 0; i < dims; i++)
        {
            uint64_t n = mStarts[i] - 1;
            int j = 0;
            while (n > 0 && j < (int)Width) { mDigit[i][j] = n % mBase[i]; n = n / mBase[i]; j++; }
            j--;
            while (j >= 0)
            {
                const uint64_t d = mPerm[i][mDigit[i][j]];
                mRnd[i][j] = mRnd[i][j + 1] + d * 1.0 / mPower[i][j];
                j--;
            }
        }
    }

    uint32_t GetNumDimensions() const { return mDims; }

    void NextSample()
    {
        for (uint32_t i = 0; i < mDims; i++)
        {
            int j = 0;
            while (mDigit[i][j] + 1 >= mBase[i] && j + 1 < (int)Width) j++;
            mDigit[i][j]++;
            uint64_t d = mPerm[i][mDigit[i][j]];
            mRnd[i][j] = mRnd[i][j + 1] + d * 1.0 / mPower[i][j];
            for (j = j - 1; j >= 0; j--)
            {
                mDigit[i][j] = 0;
                d = mPerm[i][0];
                mRnd[i][j] = mRnd[i][j + 1] + d * 1.0 / mPower[i][j];
            }
        }
    }

    // HaltonSampler.h:26
    uint32_t GetInt(uint32_t dimension) const { return uint32_t(mRnd[dimension][0] * (double)UINT32_MAX); }

private:
    Random mRandom;
    uint32_t mDims = 0;
    std::vector<uint64_t> mStarts;
    std::vector<uint32_t> mBase;
    std::vector<std::vector<double>> mRnd;
    std::vector<std::vector<uint64_t>> mDigit, mPower, mPerm;
};

// FastLog, Transcendental.cpp:196-227 (scalar)
inline float FastLog(float x)
{
    union { float f; int32_t i; } a, m;
    a.f = x;
    const int32_t e = (a.i - 0x3f2aaaab) & 0xff800000;
    m.i = a.i - e;
    const float i = 1.19209290e-7f * (float)e;
    const float f = m.f - 1.0f;
    const float s = f * f;
    float r = 0.230836749f * f - 0.279208571f;
    const float t = 0.331826031f * f - 0.498910338f;
    r = r * s + t;
    r = r * s + f;
    r = i * 0.693147182f + r;
    return r;
}

// AdaptiveRenderingSettings, Context.h:35-43
struct AdaptiveRenderingSettings
{
    bool enable = false;
    uint32_t numInitialPasses = 10;
    uint32_t minBlockSize = 4;
    uint32_t maxBlockSize = 256;
    float subdivisionTreshold = 0.005f;
    float convergenceTreshold = 0.0001f;
};

// RenderingProgress, Viewport.h:25-32
struct RenderingProgress
{
    uint32_t passesFinished = 0;
    uint32_t activePixels = 0;
    uint32_t activeBlocks = 0;
    float converged = 0.0f;
    float averageError = std::numeric_limits<float>::infinity();
};

// LightSamplingStrategy / SamplingParams / RenderingParams subset, Context.h:28-90
struct RenderingParams
{
    AdaptiveRenderingSettings adaptiveSettings;
    uint32_t dimensions = 64;
    bool useBlueNoiseDithering = true;
    float antiAliasingSpread = 0.5f;
    uint32_t maxRayDepth = 20;
    uint32_t minRussianRouletteDepth = 1;
    uint32_t lightSamplingStrategy = RTB_LIGHT_SAMPLING_SINGLE;
};

// PostprocessParams, Rendering/PostProcess.h:11-29 with the defaults of PostProcess.cpp:6-14
struct PostprocessParams
{
    float colorFilter[3] = { 1.0f, 1.0f, 1.0f };
    float exposure = 0.0f;             // log2 scale
    float contrast = 0.8f;
    float saturation = 0.98f;
    float ditheringStrength = 0.005f;  // applied after tonemapping
    float bloomFactor = 0.0f;
    uint32_t tonemapper = RTB_TONEMAP_ACES;
};

class ViewportB200
{
public:
    ViewportB200() { SetSeed(12345); }
    ~ViewportB200() { if (mCtx) rtb_destroy(mCtx); }
    ViewportB200(const ViewportB200&) = delete;
    ViewportB200& operator=(const ViewportB200&) = delete;

    const std::string& GetLastError() const { return mError; }
    RtbContext* GetContext() const { return mCtx; }

    bool Create(int deviceId, uint32_t numShards, uint32_t shardIndex, uint32_t flags, const std::string& blueNoisePath, void* cudaStream = nullptr)
    {
        RtbCreateDesc d;
        memset(&d, 0, sizeof(d));
        d.deviceId = deviceId; d.numShards = numShards; d.shardIndex = shardIndex; d.flags = flags; d.stream = cudaStream;
        mNumShards = numShards ? numShards : 1;
        if (rtb_create(&d, &mCtx) != RTB_OK) return Error();
        if (!blueNoisePath.empty())
        {
            std::vector<uint16_t> bn;
            if (!LoadBlueNoise(blueNoisePath, bn)) { mError = "Failed to load blue noise texture: " + blueNoisePath; return false; }
            if (rtb_set_blue_noise(mCtx, bn.data()) != RTB_OK) return Error();
        }
        return true;
    }

    // the scene must outlive nothing here: it is flattened and copied to the device (CreateRenderer + SetRenderer)
    // (`camera`: its texture-shaped bokeh, if any, is uploaded with the scene)
    bool SetScene(const Scene& scene, const Camera* camera = nullptr)
    {
        rtb::SceneFile flat;
        scene.Flatten(flat);
        if (camera) camera->DescribeBokeh(flat);
        const RtbSceneDesc d = flat.View();
        if (rtb_upload_scene(mCtx, &d) != RTB_OK) return Error();
        mHasScene = true;
        return true;
    }

    // CreateRenderer names, Renderer.cpp:45-69
    bool SetRenderer(const std::string& name)
    {
        if (name == "Path Tracer") mRenderer = RTB_RENDERER_PATH_TRACER;
        else if (name == "Path Tracer MIS") mRenderer = RTB_RENDERER_PATH_TRACER_MIS;
        else if (name == "Debug") mRenderer = RTB_RENDERER_DEBUG;
        else if (name == "VCM") mRenderer = RTB_RENDERER_VCM;          // VertexConnectionAndMerging; its public members: SetVcmParams
        else if (name == "Light Tracer") mRenderer = RTB_RENDERER_LIGHT_TRACER;
        else { mError = "Unknown renderer name: '" + name + "' (supported: 'Path Tracer', 'Path Tracer MIS', 'Debug', 'VCM', 'Light Tracer')"; return false; }
        return true;
    }
    // VertexConnectionAndMerging::mUseVertexConnection ... mMergingRadiusMultiplier (VertexConnectionAndMerging.h:48-56)
    bool SetVcmParams(const RtbVcmParams& params)
    {
        if (rtb_set_vcm_params(mCtx, &params) != RTB_OK) return Error();
        return true;
    }
    void SetDebugMode(uint32_t mode) { mDebugMode = mode; }

    bool Resize(uint32_t width, uint32_t height)
    {
        if (rtb_resize(mCtx, width, height) != RTB_OK) return Error();
        mWidth = width; mHeight = height;
        // the reference fills a per-pixel salt table from the viewport's generator here (Viewport.cpp:99-102; the table
        // is not used on this path): the generator moves on by one draw per pixel, and with it every later AA offset
        mRandom.SkipVector4((uint64_t)width * height);
        Reset();
        return true;
    }

    bool SetRenderingParams(const RenderingParams& p) { mParams = p; return true; }
    const RenderingParams& GetRenderingParams() const { return mParams; }
    bool SetPostprocessParams(const PostprocessParams& p) { mPostprocessParams = p; return true; }      // Viewport.cpp:166-177
    // Seeds the generators the way the reference's constructor does (Viewport.h:106-107: mRandomGenerator, then the
    // Halton sequence's own generator, each taking its integers from the entropy source in turn).  With the oracle's
    // RT_FIXED_SEED = seed, a Viewport that is the process's first consumer of entropy draws the same integers, so
    // the same calls (Resize, Reset, Render ...) give the same pass scripts.  Default seed: 12345.
    void SetSeed(uint64_t seed)
    {
        EntropyStream entropy(seed);
        mRandom.Reset(entropy);
        mHalton.Seed(entropy);
    }

    // Viewport::Reset, Viewport.cpp:117-138
    void Reset()
    {
        mHalton.Initialize(mParams.dimensions);      // (continues the sequence's generator: no re-seeding, Viewport.cpp:120-138)
        mPassesFinished = 0;
        mProgress = RenderingProgress();
        BuildInitialBlocksList();
        if (mCtx) rtb_reset(mCtx);      // (also makes every pixel active again)
    }

    // Viewport::Render, Viewport.cpp:185-289 (asynchronous: call Synchronize / GetSumBuffer to wait)
    bool Render(const Camera& camera) { return RenderPasses(camera, 1); }

    // `count` consecutive Render() calls handed to the device at once (rtb_render_passes): the host prologue runs per
    // pass in the reference's order, the device traces the passes in shared wavefronts and adds them to the film in
    // pass order, so the image is the one `count` Render() calls produce.
    bool RenderPasses(const Camera& camera, uint32_t count)
    {
        if (!mParams.adaptiveSettings.enable) return RenderPassesNow(camera, count);
        // adaptive rendering: the block list is revised after every second pass (Viewport.cpp:267-274), so the device
        // gets the passes two at a time
        if (mNumShards > 1) { mError = "adaptive rendering estimates errors over whole image rows: with several shards, all-gather the films and use rtb_block_errors / rtb_set_active_tiles directly"; return false; }
        while (count > 0)
        {
            const uint32_t n = (mPassesFinished & 1u) ? 1u : (count < 2u ? count : 2u);
            if (!RenderPassesNow(camera, n)) return false;
            count -= n;
            if (mPassesFinished > 0 && (mPassesFinished & 1u) == 0u && !UpdateBlocksList()) return false;
        }
        return true;
    }

    // RenderingProgress (Viewport::GetProgress, Viewport.h:55).  averageError (Viewport::ComputeError, Viewport.cpp:179-183)
    // is evaluated here, when asked for, if an even number of passes is in the film - the reference evaluates it
    // after every second pass, which would put a device round trip between the passes; like the reference's it stays
    // at its previous value otherwise, and is not maintained while adaptive rendering is on (Viewport.cpp:269-278).
    bool GetProgress(RenderingProgress& out)
    {
        mProgress.passesFinished = mPassesFinished;
        if (!mParams.adaptiveSettings.enable && mNumShards == 1 && mPassesFinished > 0 && (mPassesFinished & 1u) == 0u && mErrorPasses != mPassesFinished)
        {
            const RtbBlock whole = { 0, mWidth, 0, mHeight };
            if (rtb_block_errors(mCtx, nullptr, nullptr, mPassesFinished, &whole, 1, &mProgress.averageError) != RTB_OK) return Error();
            mErrorPasses = mPassesFinished;
        }
        out = mProgress;
        return true;
    }
    const std::vector<RtbBlock>& GetBlocks() const { return mBlocks; }

    // Viewport::VisualizeActiveBlocks (Viewport.cpp:734-771) on a front-buffer image in host memory (width * height
    // 0x00RRGGBB as GetFrontBuffer delivers it): the blocks still being rendered tinted red (LdrColor Lerp with
    // alpha 64, LdrColor.h:18-26) with solid red borders
    static void VisualizeBlocks(uint32_t* bgra, uint32_t width, const RtbBlock* blocks, size_t numBlocks)
    {
        const uint8_t color[4] = { 0, 0, 255, 255 };        // b, g, r, a
        const uint32_t solid = 0xFFFF0000u, alpha = 64;
        for (size_t i = 0; i < numBlocks; ++i)
        {
            const RtbBlock& b = blocks[i];
            for (uint32_t y = b.minY; y < b.maxY; ++y)
                for (uint32_t x = b.minX; x < b.maxX; ++x)
                {
                    uint32_t& pixel = bgra[(size_t)y * width + x];
                    uint32_t out = 0;
                    for (int k = 0; k < 4; ++k)
                    {
                        const uint32_t a = (pixel >> (8 * k)) & 0xFFu;
                        out |= (((a * (255u - alpha) + color[k] * alpha) / 256u) & 0xFFu) << (8 * k);
                    }
                    pixel = out;
                }
            for (uint32_t y = b.minY; y < b.maxY; ++y) { bgra[(size_t)y * width + b.minX] = solid; bgra[(size_t)y * width + b.maxX - 1] = solid; }
            for (uint32_t x = b.minX; x < b.maxX; ++x) { bgra[(size_t)b.minY * width + x] = solid; bgra[(size_t)(b.maxY - 1) * width + x] = solid; }
        }
    }
    void VisualizeActiveBlocks(uint32_t* bgra) const { VisualizeBlocks(bgra, mWidth, mBlocks.data(), mBlocks.size()); }

private:
    // Viewport::BuildInitialBlocksList, Viewport.cpp:618-645
    void BuildInitialBlocksList()
    {
        mBlocks.clear();
        mErrorPasses = 0;
        if (mWidth == 0 || mHeight == 0) return;
        const uint32_t blockSize = mParams.adaptiveSettings.maxBlockSize;
        const uint32_t rows = 1 + (mHeight - 1) / blockSize, columns = 1 + (mWidth - 1) / blockSize;
        for (uint32_t j = 0; j < rows; ++j)
            for (uint32_t i = 0; i < columns; ++i)
                mBlocks.push_back(RtbBlock{ i * blockSize, std::min(mWidth, (i + 1) * blockSize), j * blockSize, std::min(mHeight, (j + 1) * blockSize) });
        mProgress.activeBlocks = (uint32_t)mBlocks.size();
    }

public:
    // The host prologue of one Viewport::Render: the pass's Halton seed vector (Viewport.cpp:200-205) and its
    // anti-aliasing offset GetFloatNormal2(rng.GetFloat2()) * antiAliasingSpread (Viewport.cpp:235-242), drawn in the
    // reference's order.  `seeds` receives halton.GetNumDimensions() integers.
    static void NextPass(HaltonSequence& halton, Random& random, float antiAliasingSpread, uint32_t* seeds, float sampleOffset[2])
    {
        halton.NextSample();
        for (uint32_t i = 0; i < halton.GetNumDimensions(); ++i) seeds[i] = halton.GetInt(i);
        float u0, u1;
        random.GetFloat2(u0, u1);
        float s, c;
        rtb::sin_cos(2.0f * RTB_PI * u1, s, c);
        const float radius = sqrtf(-2.0f * FastLog(u0));
        sampleOffset[0] = radius * s * antiAliasingSpread;
        sampleOffset[1] = radius * c * antiAliasingSpread;
    }

    // the pass script of `numPasses` Render() calls of a viewport that was seeded with `seed`, resized once to
    // width x height (which resets it) and reset `extraResets` more times - no device involved
    static void PassScript(uint64_t seed, uint32_t dimensions, float antiAliasingSpread, uint32_t width, uint32_t height, uint32_t extraResets,
                           uint32_t numPasses, uint32_t* seeds, float* sampleOffsets)
    {
        EntropyStream entropy(seed);
        Random random;
        HaltonSequence halton;
        random.Reset(entropy);
        halton.Seed(entropy);
        random.SkipVector4((uint64_t)width * height);
        for (uint32_t i = 0; i < 1 + extraResets; ++i) halton.Initialize(dimensions);
        for (uint32_t p = 0; p < numPasses; ++p) NextPass(halton, random, antiAliasingSpread, seeds + (size_t)p * dimensions, sampleOffsets + 2 * (size_t)p);
    }

    // The list revision of Viewport::UpdateBlocksList (Viewport.cpp:659-716) given each block's error: converged blocks
    // are dropped, somewhat converged ones split in two along their longer side.  A dropped or split block is replaced
    // by the last one of the list, which - as in the reference - is therefore not looked at in this round.  Returns
    // whether a block was dropped (the set of active pixels changed).
    static bool ReviseBlocks(std::vector<RtbBlock>& blocks, std::vector<float> errors, const AdaptiveRenderingSettings& settings)
    {
        std::vector<RtbBlock> newBlocks;
        bool removed = false;
        for (size_t i = 0; i < blocks.size(); ++i)
        {
            const RtbBlock block = blocks[i];
            const float blockError = errors[i];
            const uint32_t width = block.maxX - block.minX, height = block.maxY - block.minY;
            if (blockError < settings.convergenceTreshold)
            {
                blocks[i] = blocks.back(); blocks.pop_back();
                errors[i] = errors.back(); errors.pop_back();
                removed = true;
                continue;
            }
            if (blockError < settings.subdivisionTreshold && (width > settings.minBlockSize || height > settings.minBlockSize))
            {
                blocks[i] = blocks.back(); blocks.pop_back();
                errors[i] = errors.back(); errors.pop_back();
                RtbBlock childA = block, childB = block;
                if (width > height) { const uint32_t halfPoint = (block.minX + block.maxX) / 2u; childA.maxX = halfPoint; childB.minX = halfPoint; }
                else { const uint32_t halfPoint = (block.minY + block.maxY) / 2u; childA.maxY = halfPoint; childB.minY = halfPoint; }
                newBlocks.push_back(childA);
                newBlocks.push_back(childB);
            }
        }
        blocks.insert(blocks.end(), newBlocks.begin(), newBlocks.end());
        return removed;
    }

private:
    // Viewport::UpdateBlocksList, Viewport.cpp:647-732.  The error of a block depends on the film only, so the errors of
    // the whole list come from one device call; the list is then revised exactly as the reference revises it: a removed
    // block is replaced by the last one, which is therefore not looked at in this round.
    bool UpdateBlocksList()
    {
        const AdaptiveRenderingSettings& settings = mParams.adaptiveSettings;
        if (mPassesFinished < settings.numInitialPasses) return true;
        std::vector<float> errors(mBlocks.size());
        if (!mBlocks.empty() && rtb_block_errors(mCtx, nullptr, nullptr, mPassesFinished, mBlocks.data(), (uint32_t)mBlocks.size(), errors.data()) != RTB_OK) return Error();
        const bool removed = ReviseBlocks(mBlocks, errors, settings);
        mProgress.activePixels = 0;
        for (const RtbBlock& b : mBlocks) mProgress.activePixels += (b.maxX - b.minX) * (b.maxY - b.minY);
        mProgress.converged = 1.0f - (float)mProgress.activePixels / (float)(mWidth * mHeight);
        mProgress.activeBlocks = (uint32_t)mBlocks.size();
        // Viewport::GenerateRenderingTiles cuts the blocks into tiles for its thread pool; the pixels are the same
        if (removed && rtb_set_active_tiles(mCtx, mBlocks.data(), (uint32_t)mBlocks.size()) != RTB_OK) return Error();
        return true;
    }

    bool RenderPassesNow(const Camera& camera, uint32_t count)
    {
        if (mWidth == 0 || mHeight == 0) return false;
        if (!mHasScene) { mError = "Viewport: Missing renderer"; return false; }
        if (count == 0) return true;
        const uint32_t dims = mHalton.GetNumDimensions();
        std::vector<uint32_t> seeds((size_t)count * dims);
        std::vector<RtbPassDesc> passes(count);
        for (uint32_t p = 0; p < count; ++p)
        {
            float sampleOffset[2];
            NextPass(mHalton, mRandom, mParams.antiAliasingSpread, &seeds[(size_t)p * dims], sampleOffset);
            RtbPassDesc& pass = passes[p];
            memset(&pass, 0, sizeof(pass));
            pass.renderer = mRenderer; pass.debugMode = mDebugMode;
            camera.Describe(pass.camera);
            pass.maxRayDepth = mParams.maxRayDepth; pass.minRussianRouletteDepth = mParams.minRussianRouletteDepth;
            pass.lightSamplingStrategy = mParams.lightSamplingStrategy; pass.useBlueNoise = mParams.useBlueNoiseDithering;
            pass.passIndex = mPassesFinished + p; pass.numSeedDims = dims; pass.seed = dims ? &seeds[(size_t)p * dims] : nullptr;
            pass.sampleOffset[0] = sampleOffset[0];
            pass.sampleOffset[1] = sampleOffset[1];
        }
        if (rtb_render_passes(mCtx, passes.data(), count) != RTB_OK) return Error();
        mPassesFinished += count;
        return true;
    }

public:
    bool Synchronize() { return rtb_synchronize(mCtx) == RTB_OK || Error(); }
    bool GetSumBuffer(float* rgb, int which = 0) { return rtb_read_sum(mCtx, rgb, which) == RTB_OK || Error(); }
    // Viewport::GetFrontBuffer (Viewport.h:49): the tone-mapped image of what has been rendered so far, width * height
    // 0x00RRGGBB, row 0 = top.  The reference refreshes it inside every Render(); here it is computed when asked for.
    bool GetFrontBuffer(uint32_t* bgra)
    {
        if (mPassesFinished == 0) { mError = "Viewport: nothing rendered yet"; return false; }
        RtbPostprocessDesc d;
        memset(&d, 0, sizeof(d));
        for (int i = 0; i < 3; ++i) d.colorFilter[i] = mPostprocessParams.colorFilter[i];
        d.colorFilter[3] = 1.0f;
        d.exposure = mPostprocessParams.exposure; d.contrast = mPostprocessParams.contrast; d.saturation = mPostprocessParams.saturation;
        d.ditheringStrength = mPostprocessParams.ditheringStrength; d.bloomFactor = mPostprocessParams.bloomFactor;
        d.tonemapper = mPostprocessParams.tonemapper; d.numPasses = mPassesFinished; d.ditherSeed = mPassesFinished;
        return rtb_read_front(mCtx, &d, nullptr, bgra) == RTB_OK || Error();
    }
    bool GetCounters(RtbCounters& c) { return rtb_get_counters(mCtx, &c) == RTB_OK || Error(); }
    uint32_t GetWidth() const { return mWidth; }
    uint32_t GetHeight() const { return mHeight; }
    uint32_t GetPassesFinished() const { return mPassesFinished; }

    static bool LoadBlueNoise(const std::string& path, std::vector<uint16_t>& interleaved)
    {
        FILE* f = fopen(path.c_str(), "rb");
        if (!f) return false;
        std::vector<uint16_t> raw(4 + 4 * 128 * 128);
        const bool ok = fread(raw.data(), 2, raw.size(), f) == raw.size() && raw[0] == 0x5442 && raw[1] == 0x4E52;
        fclose(f);
        if (!ok) return false;
        interleaved.resize(128 * 128 * 4);
        for (int l = 0; l < 4; ++l) for (int p = 0; p < 128 * 128; ++p) interleaved[p * 4 + l] = raw[4 + l * 128 * 128 + p];
        return true;
    }

private:
    bool Error() { mError = rtb_last_error(); return false; }

    RtbContext* mCtx = nullptr;
    bool mHasScene = false;
    uint32_t mRenderer = RTB_RENDERER_PATH_TRACER_MIS, mDebugMode = RTB_DEBUG_TRIANGLE_ID;
    uint32_t mWidth = 0, mHeight = 0, mPassesFinished = 0;
    RenderingParams mParams;
    PostprocessParams mPostprocessParams;
    RenderingProgress mProgress;
    std::vector<RtbBlock> mBlocks;        // regions of the image still being rendered (Viewport.h:126)
    uint32_t mNumShards = 1, mErrorPasses = 0;
    Random mRandom;
    HaltonSequence mHalton;
    std::string mError;
};

// One frame on several GPUs of one box from ONE process (the reference spreads its tiles over a thread pool,
// Viewport.cpp:216-268; here tile t goes to device t % N): N ViewportB200 objects, one per device, each rendering its
// shard of every pass.  Every member viewport draws the same pass script from the same seed, so the N-GPU film is the
// one-GPU film bit for bit.  Frames for the host are assembled by every device writing its own pixels straight into one
// page-locked frame (rtb_scatter_shard: N PCIe links, no collective); device-resident frames (FinishFrame) by one NCCL
// all-gather per frame inside the library (rtb_comm_init_all / rtb_finish_frame).
class ViewportGroupB200
{
public:
    ~ViewportGroupB200() { Release(); }
    const std::string& GetLastError() const { return mError; }
    uint32_t GetNumDevices() const { return (uint32_t)mViewports.size(); }
    ViewportB200& GetViewport(uint32_t i) { return *mViewports[i]; }

    bool Create(const std::vector<int>& deviceIds, uint32_t flags, const std::string& blueNoisePath)
    {
        Release();
        if (deviceIds.empty()) { mError = "no devices"; return false; }
        for (size_t i = 0; i < deviceIds.size(); ++i)
        {
            mViewports.push_back(new ViewportB200());
            if (!mViewports.back()->Create(deviceIds[i], (uint32_t)deviceIds.size(), (uint32_t)i, flags, blueNoisePath)) return Fail(i);
        }
        return true;
    }
    bool SetScene(const Scene& scene, const Camera* camera = nullptr) { return All([&](ViewportB200& v) { return v.SetScene(scene, camera); }); }
    bool SetRenderer(const std::string& name) { return All([&](ViewportB200& v) { return v.SetRenderer(name); }); }
    bool SetRenderingParams(const RenderingParams& p) { return All([&](ViewportB200& v) { return v.SetRenderingParams(p); }); }
    void SetSeed(uint64_t seed) { for (ViewportB200* v : mViewports) v->SetSeed(seed); }
    bool Resize(uint32_t width, uint32_t height)
    {
        FreeFrame();
        mWidth = width; mHeight = height;
        return All([&](ViewportB200& v) { return v.Resize(width, height); });
    }
    void Reset() { for (ViewportB200* v : mViewports) v->Reset(); }
    // every device gets its passes enqueued before any is waited for: the launches are asynchronous
    bool RenderPasses(const Camera& camera, uint32_t count) { return All([&](ViewportB200& v) { return v.RenderPasses(camera, count); }); }
    bool Synchronize() { return All([&](ViewportB200& v) { return v.Synchronize(); }); }

    // Viewport::GetSumBuffer for the whole frame
    bool GetSumBuffer(float* rgb, int which = 0)
    {
        const size_t bytes = (size_t)mWidth * mHeight * 3 * sizeof(float);
        if (!mFrame)
        {
            mFrame = static_cast<float*>(rtb_host_alloc(bytes));
            if (!mFrame) { mError = rtb_last_error(); return false; }
        }
        void* deviceView = rtb_host_device_pointer(mFrame);
        if (!deviceView) { mError = rtb_last_error(); return false; }
        for (size_t i = 0; i < mViewports.size(); ++i)
            if (rtb_scatter_shard(mViewports[i]->GetContext(), nullptr, which, deviceView) != RTB_OK) { mError = rtb_last_error(); return false; }
        if (!Synchronize()) return false;
        memcpy(rgb, mFrame, bytes);
        return true;
    }

    // the all-gathered film on every device (NCCL inside the library); `outGathered[i]` is device i's copy
    bool FinishFrame(std::vector<void*>& outGathered, int which = 0)
    {
        if (!mHaveComm && mViewports.size() > 1)
        {
            std::vector<RtbContext*> ctxs;
            for (ViewportB200* v : mViewports) ctxs.push_back(v->GetContext());
            if (rtb_comm_init_all(ctxs.data(), (uint32_t)ctxs.size()) != RTB_OK) { mError = rtb_last_error(); return false; }
            mHaveComm = true;
        }
        outGathered.assign(mViewports.size(), nullptr);
        // (one thread drives all devices: their collectives go into one NCCL group inside rtb_finish_frame_all)
        std::vector<RtbContext*> ctxs;
        for (ViewportB200* v : mViewports) ctxs.push_back(v->GetContext());
        if (rtb_finish_frame_all(ctxs.data(), (uint32_t)ctxs.size(), which, outGathered.data()) != RTB_OK) { mError = rtb_last_error(); return false; }
        return true;
    }

    bool GetCounters(RtbCounters& total)
    {
        memset(&total, 0, sizeof(total));
        for (ViewportB200* v : mViewports)
        {
            RtbCounters c;
            if (!v->GetCounters(c)) { mError = v->GetLastError(); return false; }
            total.numRays += c.numRays; total.numShadowRays += c.numShadowRays; total.numShadowRaysHit += c.numShadowRaysHit;
            total.numPrimaryRays += c.numPrimaryRays; total.numKernelLaunches += c.numKernelLaunches;
            total.numRayBoxTests += c.numRayBoxTests; total.numRayTriangleTests += c.numRayTriangleTests;
            total.numShadowBoxTests += c.numShadowBoxTests; total.numShadowTriangleTests += c.numShadowTriangleTests;
        }
        return true;
    }

private:
    template <typename F> bool All(F f)
    {
        for (size_t i = 0; i < mViewports.size(); ++i) if (!f(*mViewports[i])) return Fail(i);
        return true;
    }
    bool Fail(size_t i) { mError = "device " + std::to_string(i) + ": " + mViewports[i]->GetLastError(); return false; }
    void FreeFrame() { if (mFrame) rtb_host_free(mFrame); mFrame = nullptr; }
    void Release()
    {
        FreeFrame();
        for (ViewportB200* v : mViewports) delete v;
        mViewports.clear();
        mHaveComm = false;
    }

    std::vector<ViewportB200*> mViewports;
    std::string mError;
    uint32_t mWidth = 0, mHeight = 0;
    float* mFrame = nullptr;
    bool mHaveComm = false;
};

} // namespace rt
