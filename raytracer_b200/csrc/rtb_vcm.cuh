// rtb_vcm.cuh — the reference's "VCM" renderer (Rendering/VertexConnectionAndMerging.cpp) on the device.
//
// One pass = one light sub-path and one camera path per pixel, like VertexConnectionAndMerging::RenderPixel (:172-290):
//
//   k_vcm_grid_count / k_vcm_scan_* / k_vcm_grid_fill    HashGrid::Build over the PREVIOUS pass' photons (Utils/HashGrid.h:17-82;
//                                                        vertex merging is delayed by one pass, :95-103, PreRenderGlobal :153-168)
//   k_vcm_light     TraceLightPath (:292-428): emits a light sample, walks the sub-path, stores the pixel's light vertices and
//                   this pass' photons, connects every vertex to the camera (ConnectToCamera :908-970) and splats the
//                   result onto the film with float atomics (Film::AccumulateColor(pos, value, rng), Film.cpp:41-77)
//   k_vcm_camera    the camera path (:176-288): emission / light hits (EvaluateLight), next event estimation (SampleLights),
//                   connection to the pixel's own light vertices (ConnectVertices :738-806), merging with the photons of the
//                   previous pass inside radiusVM (MergeVertices :808-906) -> pass colour, added to the film by k_film
//
// All per-vertex arithmetic is rtb_bidir.h's (bidir_*), written once for host and device and pinned statistically against
// the reference through the CPU model tests/hostsim/hostsim_vcm.h.  The camera path draws from the pixel's Halton / blue
// noise sampler like the reference's (ctx.sampler); the light path draws from a counter-seeded xorshift generator where the
// reference uses its per-thread generator, so - like the reference itself, whose image depends on its thread schedule -
// the result is reproducible here but comparable with the reference only statistically (tests/test_gpu_vcm.py).
//
// Both path kernels give a pixel one thread that runs its sub-path to the end (the shape of k_tail): light vertices make
// one round trip through HBM per pass (160 B each, one thread writes and later reads its own), photons are appended
// with one atomic each, the hash grid is a counting sort by cell.
#pragma once

#include "rtb_bidir.h"

namespace rtb {

#define RTB_VCM_MAX_PATH_LENGTH 32u
#define RTB_VCM_VERTEX_WORDS 40u           // BidirLightVertex padded to 160 B (10 x float4)
#define RTB_VCM_PHOTON_WORDS 12u           // position, dVM | direction, dVCM | throughput, - (3 x float4)

struct VcmDevice
{
    BidirPass pass;
    CameraConnection camera;
    float4*   lightVertices;      // [pixel slot][maxPathLength] x 10 float4
    uint32_t* lightVertexCount;   // per pixel slot
    float4*   newPhotons;         // this pass' photons (3 float4 each) ...
    uint32_t* newPhotonCount;     // ... appended with an atomic
    const float4*   photons;      // the previous pass' photons, merged by this pass' camera paths
    const uint32_t* photonCount;
    const uint32_t* cellStart;    // hashSize + 1 entries: photon indices of cell c are gridIndices[cellStart[c] .. cellStart[c + 1])
    const uint32_t* gridIndices;
    uint32_t  hashMask;
    uint32_t  photonCapacity;
    float*    splat;              // raster W x H x 3: camera connections of this context's light paths (folded into the film by k_vcm_fold)
    float*    splat2;             // the same for the secondary sum (even passes), or null on odd passes
    uint32_t  vertexStride;       // light vertices reserved per pixel (= maxPathLength)
    uint32_t  seedLo, seedHi;     // light-path generator seed
};

struct DeviceRng
{
    uint64_t state;
    __device__ uint64_t next() { state ^= state >> 12; state ^= state << 25; state ^= state >> 27; return state * 0x2545F4914F6CDD1DULL; }
    __device__ uint32_t next_uint() { return (uint32_t)(next() >> 32); }
    __device__ float next_float() { return __uint_as_float(((uint32_t)(next() >> 41)) | 0x3f800000u) - 1.0f; }
};

__device__ __forceinline__ void vcm_store_vertex(float4* dst, const BidirLightVertex& lv)
{
    const Vertex& v = lv.vx;
    dst[0] = make_float4(v.tangent.x, v.tangent.y, v.tangent.z, v.bitangent.x);
    dst[1] = make_float4(v.bitangent.y, v.bitangent.z, v.normal.x, v.normal.y);
    dst[2] = make_float4(v.normal.z, v.position.x, v.position.y, v.position.z);
    dst[3] = make_float4(v.texU, v.texV, __uint_as_float(v.materialIndex), __uint_as_float(v.bsdf));
    dst[4] = make_float4(v.baseColor.x, v.baseColor.y, v.baseColor.z, v.baseColor.w);
    dst[5] = make_float4(v.emission.x, v.emission.y, v.emission.z, v.emission.w);
    dst[6] = make_float4(v.roughness, v.metalness, v.IoR, v.K);
    dst[7] = make_float4(v.outgoingDirWorld.x, v.outgoingDirWorld.y, v.outgoingDirWorld.z, __uint_as_float(lv.pathLength));
    dst[8] = make_float4(lv.throughput.x, lv.throughput.y, lv.throughput.z, lv.throughput.w);
    dst[9] = make_float4(lv.dVC, lv.dVM, lv.dVCM, 0.0f);
}

__device__ __forceinline__ void vcm_load_vertex(const float4* src, BidirLightVertex& lv)
{
    Vertex& v = lv.vx;
    const float4 a = src[0], b = src[1], c = src[2], d = src[3], e = src[4], f = src[5], g = src[6], h = src[7], i = src[8], j = src[9];
    v.tangent = v3(a.x, a.y, a.z); v.bitangent = v3(a.w, b.x, b.y); v.normal = v3(b.z, b.w, c.x); v.position = v3(c.y, c.z, c.w);
    v.texU = d.x; v.texV = d.y; v.materialIndex = __float_as_uint(d.z); v.bsdf = __float_as_uint(d.w);
    v.baseColor = v4(e.x, e.y, e.z, e.w); v.emission = v4(f.x, f.y, f.z, f.w);
    v.roughness = g.x; v.metalness = g.y; v.IoR = g.z; v.K = g.w;
    v.outgoingDirWorld = v3(h.x, h.y, h.z); lv.pathLength = __float_as_uint(h.w);
    lv.throughput = v4(i.x, i.y, i.z, i.w);
    lv.dVC = j.x; lv.dVM = j.y; lv.dVCM = j.z;
}

// HashGrid::GetCellIndex (HashGrid.h:154-176) with cells of 2 * radius; coordinates are taken relative to the world origin
// with a floor (the reference subtracts the photons' bounding box first and truncates - the same cells up to a shift,
// and a range query that filters by distance finds the same photons in either)
__device__ __forceinline__ uint32_t vcm_cell_hash(int32_t x, int32_t y, int32_t z, uint32_t mask)
{
    return (((uint32_t)x * 73856093u) ^ ((uint32_t)y * 19349663u) ^ ((uint32_t)z * 83492791u)) & mask;
}

__global__ void __launch_bounds__(256) k_vcm_grid_count(const float4* __restrict__ photons, const uint32_t* __restrict__ photonCount, uint32_t capacity, float invCellSize, uint32_t mask,
                                                        uint32_t* __restrict__ cellCount)
{
    const uint32_t n = min(*photonCount, capacity);
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    {
        const float4 p = photons[3u * i];
        atomicAdd(&cellCount[vcm_cell_hash((int32_t)floorf(p.x * invCellSize), (int32_t)floorf(p.y * invCellSize), (int32_t)floorf(p.z * invCellSize), mask)], 1u);
    }
}

// exclusive prefix sum over the cell counts in three steps: 1024-element tiles, the tile totals, the tile offsets
__global__ void __launch_bounds__(256) k_vcm_scan_tiles(uint32_t* __restrict__ data, uint32_t n, uint32_t* __restrict__ tileTotals)
{
    __shared__ uint32_t warpSums[8];
    const uint32_t base = blockIdx.x * 1024u + threadIdx.x * 4u;
    uint32_t v[4];
    for (int k = 0; k < 4; ++k) v[k] = base + k < n ? data[base + k] : 0u;
    const uint32_t mine = v[0] + v[1] + v[2] + v[3];
    uint32_t incl = mine;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    for (int o = 1; o < 32; o <<= 1) { const uint32_t up = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= (uint32_t)o) incl += up; }
    if (lane == 31u) warpSums[warp] = incl;
    __syncthreads();
    uint32_t warpOffset = 0;
    for (uint32_t w = 0; w < warp; ++w) warpOffset += warpSums[w];
    uint32_t run = warpOffset + incl - mine;
    for (int k = 0; k < 4; ++k) { if (base + k < n) data[base + k] = run; run += v[k]; }
    if (threadIdx.x == 255u) tileTotals[blockIdx.x] = warpOffset + incl;
}

__global__ void __launch_bounds__(1024) k_vcm_scan_totals(uint32_t* __restrict__ tileTotals, uint32_t numTiles, uint32_t* __restrict__ grandTotal)
{
    __shared__ uint32_t partial[1024];
    const uint32_t per = (numTiles + 1023u) / 1024u;
    const uint32_t first = threadIdx.x * per;
    uint32_t sum = 0;
    for (uint32_t k = 0; k < per; ++k) if (first + k < numTiles) sum += tileTotals[first + k];
    partial[threadIdx.x] = sum;
    __syncthreads();
    if (threadIdx.x == 0)
    {
        uint32_t run = 0;
        for (uint32_t t = 0; t < 1024u; ++t) { const uint32_t s = partial[t]; partial[t] = run; run += s; }
        *grandTotal = run;
    }
    __syncthreads();
    uint32_t run = partial[threadIdx.x];
    for (uint32_t k = 0; k < per; ++k) if (first + k < numTiles) { const uint32_t s = tileTotals[first + k]; tileTotals[first + k] = run; run += s; }
}

__global__ void __launch_bounds__(256) k_vcm_scan_offsets(uint32_t* __restrict__ data, uint32_t n, const uint32_t* __restrict__ tileTotals)
{
    const uint32_t i = blockIdx.x * 256u + threadIdx.x;
    if (i < n) data[i] += tileTotals[i >> 10];
}

// cellCursor starts as a copy of cellStart; every photon takes the next free place of its cell
__global__ void __launch_bounds__(256) k_vcm_grid_fill(const float4* __restrict__ photons, const uint32_t* __restrict__ photonCount, uint32_t capacity, float invCellSize, uint32_t mask,
                                                       uint32_t* __restrict__ cellCursor, uint32_t* __restrict__ gridIndices)
{
    const uint32_t n = min(*photonCount, capacity);
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    {
        const float4 p = photons[3u * i];
        const uint32_t cell = vcm_cell_hash((int32_t)floorf(p.x * invCellSize), (int32_t)floorf(p.y * invCellSize), (int32_t)floorf(p.z * invCellSize), mask);
        gridIndices[atomicAdd(&cellCursor[cell], 1u)] = i;
    }
}

struct VcmOccluded
{
    const SceneView& S;
    SharedStack sstack;
    unsigned long long* shadowRays;
    unsigned long long* visible;
    __device__ bool operator()(const Ray& ray, float tMax) const
    {
        (*shadowRays)++;
        const bool occluded = scene_traverse_shadow<false>(S, ray, tMax, nullptr, sstack);
        if (!occluded) (*visible)++;
        return occluded;
    }
};

// TraceLightPath, VertexConnectionAndMerging.cpp:292-428
template <bool kTex>
__global__ void __launch_bounds__(128, 3) k_vcm_light(const __grid_constant__ SceneView S, const __grid_constant__ PassParams P, const __grid_constant__ VcmDevice V,
                                                      const uint32_t* __restrict__ pixelList, uint32_t numPixels, Totals* tot)
{
    extern __shared__ uint32_t smemStack[];
    SharedStack sstack = (SharedStack)__cvta_generic_to_shared(smemStack) + threadIdx.x * 4u;
    asm volatile("" : "+r"(sstack));
    const uint32_t stride = gridDim.x * blockDim.x;
    const uint32_t lane = threadIdx.x & 31u;
    unsigned long long rays = 0, shadowRays = 0, visible = 0;
    const BidirPass& pass = V.pass;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < numPixels; i += stride)
    {
        const uint32_t pix = pixelList[i];
        DeviceRng rng;
        rng.state = hash64(((uint64_t)pix << 32) ^ (uint64_t)P.passIndex ^ (((uint64_t)V.seedHi << 32 | V.seedLo) * 0x9E3779B97F4A7C15ULL)) | 1ull;
        const VcmOccluded occluded = { S, sstack, &shadowRays, &visible };
        uint32_t numVertices = 0;
        BidirPath path;
        path.throughput = v4(1.0f, 1.0f, 1.0f, 1.0f); path.dVC = path.dVM = path.dVCM = 0.0f;
        path.length = 1; path.lastSpecular = 1; path.isFiniteLight = 0;
        const uint32_t pick = S.numLights ? rng.next_uint() : 0u;
        const float p0 = S.numLights ? rng.next_float() : 0.0f, p1 = S.numLights ? rng.next_float() : 0.0f, p2 = S.numLights ? rng.next_float() : 0.0f;
        const float d0 = S.numLights ? rng.next_float() : 0.0f, d1 = S.numLights ? rng.next_float() : 0.0f;
        bool alive = bidir_generate_light_sample<kTex>(S, pass, pick, p0, p1, p2, d0, d1, path);
        while (alive)
        {
            Hit hit;
            hit.objectId = RTB_INVALID_OBJECT; hit.subObjectId = 0; hit.distance = INFINITY; hit.u = 0.0f; hit.v = 0.0f;
            scene_traverse<false>(S, path.ray, hit, nullptr, sstack);
            rays++;
            if (hit.distance == INFINITY || hit.subObjectId == RTB_LIGHT_OBJECT) break;
            BidirLightVertex lv;
            evaluate_intersection<kTex>(S, path.ray, hit, lv.vx, RTB_NO_MATERIAL);
            lv.vx.outgoingDirWorld = neg(path.ray.dir);
            evaluate_shading_data<kTex>(S, lv.vx);
            bidir_arrive(path, lv.vx, hit.distance, true);
            if (!bidir_is_delta_bsdf(lv.vx))
            {
                if (pass.useVertexConnection && numVertices < V.vertexStride)
                {
                    lv.pathLength = path.length; lv.throughput = path.throughput; lv.dVC = path.dVC; lv.dVM = path.dVM; lv.dVCM = path.dVCM;
                    vcm_store_vertex(V.lightVertices + ((size_t)i * V.vertexStride + numVertices) * (RTB_VCM_VERTEX_WORDS / 4u), lv);
                    numVertices++;
                    float filmX, filmY;
                    V3 value;
                    if (bidir_connect_to_camera(pass, V.camera, lv, occluded, filmX, filmY, value))
                    {
                        // Film::AccumulateColor(pos, value, rng), Film.cpp:41-77: nearest pixel, a random +1 step in x by the fractional part
                        const float cx = filmX * (float)P.width + 0.0f, cy = filmY * (float)P.height + 0.5f;
                        int32_t ix = cvt_rn(cx);
                        const float ux = rng.next_float();
                        rng.next_float(); rng.next_float(); rng.next_float();          // Random::GetVector4 draws four lanes
                        if (ux < cx - (float)ix) ix++;
                        const int32_t y = (int32_t)(P.height - 1u) - (int32_t)cy;
                        if ((uint32_t)ix < P.width && (uint32_t)y < P.height)
                        {
                            const size_t o = ((size_t)y * P.width + (uint32_t)ix) * 3u;
                            atomicAdd(V.splat + o, value.x); atomicAdd(V.splat + o + 1, value.y); atomicAdd(V.splat + o + 2, value.z);
                            if (V.splat2) { atomicAdd(V.splat2 + o, value.x); atomicAdd(V.splat2 + o + 1, value.y); atomicAdd(V.splat2 + o + 2, value.z); }
                        }
                    }
                }
                if (pass.useVertexMerging)
                {
                    const uint32_t slot = atomicAdd(V.newPhotonCount, 1u);
                    if (slot < V.photonCapacity)
                    {
                        float4* ph = V.newPhotons + 3u * (size_t)slot;
                        ph[0] = make_float4(lv.vx.position.x, lv.vx.position.y, lv.vx.position.z, path.dVM);
                        ph[1] = make_float4(lv.vx.outgoingDirWorld.x, lv.vx.outgoingDirWorld.y, lv.vx.outgoingDirWorld.z, path.dVCM);
                        ph[2] = make_float4(path.throughput.x, path.throughput.y, path.throughput.z, 0.0f);
                    }
                }
            }
            if (path.length + 2u > pass.maxPathLength) break;
            const float s0 = rng.next_float(), s1 = rng.next_float(), s2 = rng.next_float();
            alive = bidir_advance_path(pass, path, lv.vx, s0, s1, s2);
        }
        V.lightVertexCount[i] = numVertices;
    }
    for (int o = 16; o > 0; o >>= 1)
    {
        rays += __shfl_down_sync(0xFFFFFFFFu, rays, o);
        shadowRays += __shfl_down_sync(0xFFFFFFFFu, shadowRays, o);
        visible += __shfl_down_sync(0xFFFFFFFFu, visible, o);
    }
    if (lane == 0)
    {
        if (rays) atomicAdd(&tot->numRays, rays);
        if (shadowRays) atomicAdd(&tot->numShadowRays, shadowRays);
        if (visible) atomicAdd(&tot->numShadowRaysHit, visible);
    }
}

// Film::AccumulateColor(pos, value, rng), Film.cpp:41-77, onto the raster splat buffers: nearest pixel, a random +1 step in x by
// the fractional part (Random::GetVector4 draws four lanes)
struct FilmSplat
{
    const PassParams& P;
    const VcmDevice& V;
    DeviceRng& rng;
    __device__ void operator()(float filmX, float filmY, const V3& value) const
    {
        const float cx = filmX * (float)P.width + 0.0f, cy = filmY * (float)P.height + 0.5f;
        int32_t ix = cvt_rn(cx);
        const float ux = rng.next_float();
        rng.next_float(); rng.next_float(); rng.next_float();
        if (ux < cx - (float)ix) ix++;
        const int32_t y = (int32_t)(P.height - 1u) - (int32_t)cy;
        if ((uint32_t)ix < P.width && (uint32_t)y < P.height)
        {
            const size_t o = ((size_t)y * P.width + (uint32_t)ix) * 3u;
            atomicAdd(V.splat + o, value.x); atomicAdd(V.splat + o + 1, value.y); atomicAdd(V.splat + o + 2, value.z);
            if (V.splat2) { atomicAdd(V.splat2 + o, value.x); atomicAdd(V.splat2 + o + 1, value.y); atomicAdd(V.splat2 + o + 2, value.z); }
        }
    }
};

// The reference's "Light Tracer" renderer (LightTracer::RenderPixel, LightTracer.cpp:26-175): one light sub-path per pixel and
// pass, every vertex connected to the camera and splatted; the pixel itself receives nothing.  The loop is rtb_bidir.h's
// light_trace_path (pinned against the reference on the CPU, tests/test_hostsim_parity.py); one thread per sub-path like
// the VCM kernels, numbers from the same counter-seeded generator.
template <bool kTex>
__global__ void __launch_bounds__(128, 3) k_light_tracer(const __grid_constant__ SceneView S, const __grid_constant__ PassParams P, const __grid_constant__ VcmDevice V,
                                                         const uint32_t* __restrict__ pixelList, uint32_t numPixels, Totals* tot)
{
    extern __shared__ uint32_t smemStack[];
    SharedStack sstack = (SharedStack)__cvta_generic_to_shared(smemStack) + threadIdx.x * 4u;
    asm volatile("" : "+r"(sstack));
    const uint32_t stride = gridDim.x * blockDim.x;
    const uint32_t lane = threadIdx.x & 31u;
    unsigned long long rays = 0, shadowRays = 0, visible = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < numPixels; i += stride)
    {
        const uint32_t pix = pixelList[i];
        DeviceRng rng;
        rng.state = hash64(((uint64_t)pix << 32) ^ (uint64_t)P.passIndex ^ (((uint64_t)V.seedHi << 32 | V.seedLo) * 0x9E3779B97F4A7C15ULL)) | 1ull;
        FilmSplat splat = { P, V, rng };
        rays += light_trace_path<kTex>(S, V.camera, P.maxRayDepth, rng, splat, sstack, &shadowRays, &visible);
    }
    for (int o = 16; o > 0; o >>= 1)
    {
        rays += __shfl_down_sync(0xFFFFFFFFu, rays, o);
        shadowRays += __shfl_down_sync(0xFFFFFFFFu, shadowRays, o);
        visible += __shfl_down_sync(0xFFFFFFFFu, visible, o);
    }
    if (lane == 0)
    {
        if (rays) atomicAdd(&tot->numRays, rays);
        if (shadowRays) atomicAdd(&tot->numShadowRays, shadowRays);
        if (visible) atomicAdd(&tot->numShadowRaysHit, visible);
    }
}

// MergeVertices, VertexConnectionAndMerging.cpp:808-906 + HashGrid::Process (HashGrid.h:84-146): the 2 x 2 x 2 block of
// cells around the query point, every photon within radiusVM
__device__ __forceinline__ V4 vcm_merge(const VcmDevice& V, const BidirPath& cameraPath, const Vertex& vx)
{
    V4 contribution = v4s(0.0f);
    const uint32_t numPhotons = min(*V.photonCount, V.photonCapacity);
    if (numPhotons == 0u) return contribution;
    const float radiusSqr = s_sqr(V.pass.radiusVM);
    const float invCell = 1.0f / (V.pass.radiusVM * 2.0f);
    const int32_t cx = (int32_t)floorf(fma_(vx.position.x, invCell, -0.5f)), cy = (int32_t)floorf(fma_(vx.position.y, invCell, -0.5f)), cz = (int32_t)floorf(fma_(vx.position.z, invCell, -0.5f));
    uint32_t visited[8];
    uint32_t numVisited = 0;
    for (uint32_t k = 0; k < 8u; ++k)
    {
        const uint32_t cell = vcm_cell_hash(cx + (int32_t)(k & 1u), cy + (int32_t)((k >> 1) & 1u), cz + (int32_t)(k >> 2), V.hashMask);
        bool seen = false;
        for (uint32_t j = 0; j < numVisited; ++j) seen = seen || visited[j] == cell;
        if (seen) continue;
        visited[numVisited++] = cell;
        const uint32_t first = V.cellStart[cell], last = V.cellStart[cell + 1u];
        for (uint32_t j = first; j < last; ++j)
        {
            const float4* ph = V.photons + 3u * (size_t)V.gridIndices[j];
            const float4 a = ph[0];
            if (sqr_length3(vx.position - v3(a.x, a.y, a.z)) > radiusSqr) continue;
            const float4 b = ph[1], c = ph[2];
            BidirPhoton photon;
            photon.position = v3(a.x, a.y, a.z); photon.direction = v3(b.x, b.y, b.z); photon.throughput = v3(c.x, c.y, c.z);
            photon.dVM = a.w; photon.dVCM = b.w;
            bidir_merge_photon(V.pass, cameraPath, vx, photon, contribution);
        }
    }
    return contribution;
}

__device__ __forceinline__ V4 vcm_mul_add(const V4& a, const V4& b, const V4& c) { return v4(fma_(a.x, b.x, c.x), fma_(a.y, b.y, c.y), fma_(a.z, b.z, c.z), fma_(a.w, b.w, c.w)); }

// the camera path of VertexConnectionAndMerging::RenderPixel, :176-288
// (measured and dropped, round 2: both path kernels as ONE loop over path vertices per warp on a persistent grid, lanes taking
//  their next pixel whenever their path ends - `bench.py --config c5` 39.6 -> 63 ms per pass: the lanes of a warp stop
//  walking neighbouring pixels' paths, and the incoherence costs more than the idle lanes did)
template <bool kTex>
__global__ void __launch_bounds__(128, 3) k_vcm_camera(const __grid_constant__ SceneView S, const __grid_constant__ PassParams P, const __grid_constant__ VcmDevice V,
                                                       const uint32_t* __restrict__ pixelList, uint32_t numPixels, float4* __restrict__ passColor, Totals* tot)
{
    extern __shared__ uint32_t smemStack[];
    SharedStack sstack = (SharedStack)__cvta_generic_to_shared(smemStack) + threadIdx.x * 4u;
    asm volatile("" : "+r"(sstack));
    const uint32_t stride = gridDim.x * blockDim.x;
    const uint32_t lane = threadIdx.x & 31u;
    unsigned long long rays = 0, shadowRays = 0, visible = 0;
    const BidirPass& pass = V.pass;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < numPixels; i += stride)
    {
        const uint32_t pix = pixelList[i];
        PathState ps;
        init_path(P, S.blueNoise, pix % P.width, pix / P.width, 0, ps);
        const VcmOccluded occluded = { S, sstack, &shadowRays, &visible };
        const uint32_t numLightVertices = V.lightVertexCount[i];
        const float4* lightVertices = V.lightVertices + (size_t)i * V.vertexStride * (RTB_VCM_VERTEX_WORDS / 4u);
        V4 result = v4s(0.0f);
        BidirPath path;
        bidir_begin_camera_path(path, ps.ray, V.camera);
        Vertex vx;
        for (;;)
        {
            Hit hit;
            hit.objectId = RTB_INVALID_OBJECT; hit.subObjectId = 0; hit.distance = INFINITY; hit.u = 0.0f; hit.v = 0.0f;
            scene_traverse<false>(S, path.ray, hit, nullptr, sstack);
            rays++;
            if (hit.distance == INFINITY)
            {
                V4 sum = v4s(0.0f);
                for (uint32_t g = 0; g < S.numGlobalLights; ++g) sum = sum + bidir_evaluate_light<kTex>(S, pass, S.lights[S.globalLights[g]], nullptr, path);
                result = vcm_mul_add(path.throughput, sum, result);
                break;
            }
            evaluate_intersection<kTex>(S, path.ray, hit, vx, RTB_NO_MATERIAL);
            bidir_arrive(path, vx, hit.distance, false);
            if (hit.subObjectId == RTB_LIGHT_OBJECT)
            {
                const DLight& light = S.lights[S.objShade[hit.objectId].lightIndex];
                result = vcm_mul_add(path.throughput, bidir_evaluate_light<kTex>(S, pass, light, &vx, path), result);
                break;
            }
            vx.outgoingDirWorld = neg(path.ray.dir);
            evaluate_shading_data<kTex>(S, vx);
            result = vcm_mul_add(path.throughput, vx.emission, result);
            if (path.length >= pass.maxPathLength) break;
            const bool isDeltaBsdf = bidir_is_delta_bsdf(vx);
            if (!isDeltaBsdf && pass.useVertexConnection)
            {
                V4 sum = v4s(0.0f);
                for (uint32_t l = 0; l < S.numLights; ++l)
                {
                    const float u0 = sampler_get_float(ps.sampler, P, S.blueNoise), u1 = sampler_get_float(ps.sampler, P, S.blueNoise), u2 = sampler_get_float(ps.sampler, P, S.blueNoise);
                    sum = sum + bidir_sample_light<kTex>(S, pass, S.lights[l], vx, path, u0, u1, u2, occluded);
                }
                result = vcm_mul_add(path.throughput, sum, result);
            }
            if (!isDeltaBsdf && pass.useVertexConnection && numLightVertices > 0u)
            {
                V4 sum = v4s(0.0f);
                for (uint32_t k = 0; k < numLightVertices; ++k)
                {
                    BidirLightVertex lv;
                    vcm_load_vertex(lightVertices + (size_t)k * (RTB_VCM_VERTEX_WORDS / 4u), lv);
                    if (lv.pathLength + path.length + 1u > pass.maxPathLength) break;
                    sum = vcm_mul_add(lv.throughput, bidir_connect_vertices(pass, path, vx, lv, occluded), sum);
                }
                result = vcm_mul_add(path.throughput, sum, result);
            }
            if (!isDeltaBsdf && pass.useVertexMerging && pass.iteration > 0u)
                result = vcm_mul_add(path.throughput * vcm_merge(V, path, vx), v4s(pass.mergingNormalization), result);
            if (path.length > pass.maxPathLength) break;
            const float s0 = sampler_get_float(ps.sampler, P, S.blueNoise), s1 = sampler_get_float(ps.sampler, P, S.blueNoise), s2 = sampler_get_float(ps.sampler, P, S.blueNoise);
            if (!bidir_advance_path(pass, path, vx, s0, s1, s2)) break;
        }
        passColor[i] = make_float4(result.x, result.y, result.z, 0.0f);
    }
    for (int o = 16; o > 0; o >>= 1)
    {
        rays += __shfl_down_sync(0xFFFFFFFFu, rays, o);
        shadowRays += __shfl_down_sync(0xFFFFFFFFu, shadowRays, o);
        visible += __shfl_down_sync(0xFFFFFFFFu, visible, o);
    }
    if (lane == 0)
    {
        if (rays) atomicAdd(&tot->numRays, rays);
        if (shadowRays) atomicAdd(&tot->numShadowRays, shadowRays);
        if (visible) atomicAdd(&tot->numShadowRaysHit, visible);
    }
}

// the film takes the splats of this shard's pixels (and the raster buffer is cleared for the next frame)
__global__ void __launch_bounds__(256) k_vcm_fold(float* __restrict__ splat, float* __restrict__ sum, const uint32_t* __restrict__ pixelList, uint32_t numPixels)
{
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < numPixels; i += stride)
    {
        const size_t o = (size_t)pixelList[i] * 3u;
        sum[(size_t)i * 3u] += splat[o]; sum[(size_t)i * 3u + 1u] += splat[o + 1u]; sum[(size_t)i * 3u + 2u] += splat[o + 2u];
    }
}

} // namespace rtb
