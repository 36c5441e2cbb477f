// rtb_kernels.cu — sm_100a wavefront kernels + the extern "C" boundary (include/rtb200.h).
//
// One pass (= one sample per pixel of this context's pixel shard) is enqueued as
//
//   k_raygen                       camera rays + path state for every pixel of the shard        (Viewport::RenderTile 300-330)
//   for depth = 0 .. maxRayDepth:
//     k_traverse   closest hit     two-level BVH, one ray per thread, persistent grid           (Scene::Traverse)
//     k_shade      per path vertex emission / light hit / miss / NEE sampling / RR / BSDF sample (PathTracerMIS::RenderPixel body)
//                                  -> compacts surviving paths into the other path buffer (warp ballot + one atomic per warp)
//                                  -> writes this vertex' shadow rays + compacted lists of the rays and of the paths that own some
//     k_shadow     any hit         one shadow ray per thread; an occluded ray zeroes its contribution      (Scene::Traverse_Shadow)
//     k_resolve                    per path: sum of its contributions in light order -> pass colour         (PathTracerMIS::SampleLights)
//   k_film                         sum += pass colour (+ secondarySum on even passes)             (Film::AccumulateColor)
//
// Everything is stream-ordered on one stream with device-resident queue counters: no host round trip inside a pass.
// Per-pixel radiance is accumulated in `passColor` in exactly the reference's order (emission, then the NEE sum of
// that vertex, then the next vertex ...) because shade(d) -> shadow(d) -> shade(d+1) are ordered and every pixel
// has exactly one path per pass.
#include <cuda_runtime.h>
#include <nccl.h>          // types only: the library is loaded with dlopen when a context asks for a communicator
#include <dlfcn.h>
#include <cstdio>
#include <cstdarg>
#include <new>
#include <mutex>
#include <string>
#include <vector>
#include <algorithm>

#include "rtb_integrator.h"
#include "rtb_postprocess.h"
#include "rtb_upload.h"

using namespace rtb;

// ---------------------------------------------------------------------------------------------------------------------
// device data
// ---------------------------------------------------------------------------------------------------------------------

#define RTB_MAX_DEPTH_SLOTS 64
#define RTB_SNAP_RING 12      // raster snapshots in flight for rtb_read_sum_async / rtb_unshard_film_async
// shared-memory traversal stack: at most 227 KB per CTA = 454 entries x 128 threads x 4 B (rtb_trace.h)
#define RTB_MAX_SMEM_STACK_ENTRIES (RTB_SMEM_STACK ? 440u : (uint32_t)RTB_STACK_SIZE)

// 64 registers -> 8 CTAs of 128 threads per SM: the traversal kernels are latency bound, occupancy is worth more than
// the few spilled bytes (measured: 72 regs / 7 CTAs 4.4 ms, 64 regs / 8 CTAs 3.6 ms per pass for k_traverse)
#ifndef RTB_SHADE_MINBLOCKS
#define RTB_SHADE_MINBLOCKS 5     // k_shade per 8 passes, one kernel per integrator: 4 CTAs/SM (116 regs, no spills) 8.91 ms, 5 (96 regs) 8.30, 6 (80 regs) 8.63
                                  // (with the three integrators in one kernel: 4: 9.13, 5: 8.81, 6: 8.20 on the round-1 build, 8.72 now)
#endif
#ifndef RTB_TRAV_MINBLOCKS
#define RTB_TRAV_MINBLOCKS 8
#endif

struct PathBuf
{
    float4* p0;   // origin.xyz, dir.x
    float4* p1;   // dir.yz, invDir.xy
    float4* p2;   // invDir.z, originDivDir.xyz
    float4* p3;   // throughput
    uint4*  p4;   // pixel slot | pass slot << 26, sampler salt, dim (12 bits) | depth << 12 (7) | lastSpecular << 19 | lastMaterial << 20 (12, 0xFFF = none), lastPdfW bits
};

// queue counters of ONE wavefront (every wavefront slot has its own)
struct Control
{
    uint32_t pathCount[RTB_MAX_DEPTH_SLOTS + 2];     // live paths entering depth d
    uint32_t shadowPathCount[RTB_MAX_DEPTH_SLOTS + 2];
    uint32_t shadowRayCount[RTB_MAX_DEPTH_SLOTS + 2];
    uint32_t shadowCursor[RTB_MAX_DEPTH_SLOTS + 2];   // next unfetched entry of rayList (k_shadow's dynamic fetch)
    uint32_t meshItemCount[RTB_MAX_DEPTH_SLOTS + 2];  // k_shadow_top -> k_shadow_mesh: (ray, mesh) visits queued at this depth
    uint32_t meshCursor[RTB_MAX_DEPTH_SLOTS + 2];     // next unfetched item
    uint32_t pathCursor[RTB_MAX_DEPTH_SLOTS + 2];     // next unfetched path of depth d (k_traverse_p's dynamic fetch)
    uint32_t tailTaken;                               // != 0: k_tail finished every remaining path (1 + the depth it started at): later kernels have nothing to do
    uint32_t pad[3];
};

// cumulative ray counters of the context (shared by all wavefront slots; atomics)
struct Totals
{
    unsigned long long numRays, numShadowRays, numShadowRaysHit, numBoxTests, numTriTests, numShadowBoxTests, numShadowTriTests;
};

#if !RTB_TU_TAIL
#include "rtb_vcm.cuh"
#endif

#define RTB_ALL_ACTIVE 0xFFFFFFFFu

struct Wavefront
{
    PathBuf  buf[2];
    uint4*   hit0;        // objectId, subObjectId, distance bits, u bits
    float*   hit1;        // v
    float4*  passColor;   // per pixel slot: radiance of this pass (xyz)
    // shadow rays of the current depth: ray r of the path at input index i lives at i * raysPerPath + r
    float4*  sr0;         // origin.xyz, tMax
    float4*  sr1;         // dir.xyz, contribution.x
    float4*  sr2;         // invDir.xyz, contribution.y
    float4*  sr3;         // originDivDir.xyz, contribution.z
    float4*  shThr;       // per input index: throughput.xyz at that vertex, pixel slot bits
    uint32_t* shCount;    // per input index: number of shadow rays
    uint32_t* shList;     // compacted input indices that own >= 1 shadow ray
    uint32_t* rayList;    // compacted shadow-ray slots (i * raysPerPath + r) of the current depth
    // split any-hit walk (k_shadow_top / k_shadow_mesh): one item per (ray, mesh) visit = the mesh-local ray
    float4*  mi0;         // origin.xyz, tMax
    float4*  mi1;         // dir.xyz, ray slot bits
    float4*  mi2;         // invDir.xyz, mesh index bits
    uint32_t* shFlag;     // per ray slot: 1 = found occluded (set once)
    uint8_t*  shOcc;      // per ray slot: cleared by k_shade when it emits the ray, set by the any-hit kernel when the ray is occluded: ONE byte
                          // per occluded ray instead of zeroing the three contribution words in place (three 32-byte sectors dirtied per ray)
    uint32_t miCapacity;
    Control* ctl;
    Totals*  tot;
    const uint32_t* pixelList;   // pixel slot -> y * width + x
    uint32_t numPixels;          // pixel slots of this shard
    uint32_t raysPerPath;
    uint32_t tailThreshold;      // k_tail takes over when at most this many paths are left (0 = never)
    uint32_t batch;              // passes in flight in this wavefront (pass slot b uses passColor[b * numPixels + pixel])
    uint32_t firstPassIndex;     // Viewport pass index of pass slot 0
    // adaptive rendering (rtb_set_active_tiles): only these pixel slots get a sample in this wavefront
    const uint32_t* activeList;
    uint32_t numActive;          // RTB_ALL_ACTIVE: every slot, activeList unused
};

__device__ __forceinline__ void store_path(const PathBuf& b, uint32_t i, const PathState& ps)
{
    b.p0[i] = make_float4(ps.ray.origin.x, ps.ray.origin.y, ps.ray.origin.z, ps.ray.dir.x);
    b.p1[i] = make_float4(ps.ray.dir.y, ps.ray.dir.z, ps.ray.invDir.x, ps.ray.invDir.y);
    b.p2[i] = make_float4(ps.ray.invDir.z, ps.ray.originDivDir.x, ps.ray.originDivDir.y, ps.ray.originDivDir.z);
    b.p3[i] = make_float4(ps.throughput.x, ps.throughput.y, ps.throughput.z, ps.throughput.w);
    b.p4[i] = make_uint4(ps.pixel | (ps.passSlot << 26), ps.sampler.salt, (ps.sampler.dim & 0xFFFu) | ((ps.depth & 0x7Fu) << 12) | ((ps.lastSpecular & 1u) << 19) | ((ps.lastMaterial & 0xFFFu) << 20), __float_as_uint(ps.lastPdfW));
}

__device__ __forceinline__ void load_ray(const PathBuf& b, uint32_t i, Ray& r)
{
    const float4 a = b.p0[i], c = b.p1[i], d = b.p2[i];
    r.origin = v3(a.x, a.y, a.z);
    r.dir = v3(a.w, c.x, c.y);
    r.invDir = v3(c.z, c.w, d.x);
    r.originDivDir = v3(d.y, d.z, d.w);
}

// world-ray sources for the traversal kernels (rtb_trace.h RayInRegisters): load the ray again from its queue
struct RayInPathQueue
{
    const PathBuf& b;
    uint32_t i;
    __device__ __forceinline__ Ray load() const { Ray r; load_ray(b, i, r); return r; }
};

struct RayInShadowQueue
{
    const Wavefront& W;
    uint32_t j;
    __device__ __forceinline__ Ray load() const
    {
        const float4 a = W.sr0[j], b = W.sr1[j], c = W.sr2[j], d = W.sr3[j];
        Ray r;
        r.origin = v3(a.x, a.y, a.z); r.dir = v3(b.x, b.y, b.z); r.invDir = v3(c.x, c.y, c.z); r.originDivDir = v3(d.x, d.y, d.z);
        return r;
    }
};

__device__ __forceinline__ void load_path(const PathBuf& b, uint32_t i, const Wavefront& W, uint32_t width, uint32_t numSeedDims, PathState& ps, bool needRaster = true)
{
    load_ray(b, i, ps.ray);
    const float4 t = b.p3[i];
    ps.throughput = v4(t.x, t.y, t.z, t.w);
    const uint4 m = b.p4[i];
    ps.pixel = m.x & 0x03FFFFFFu;
    ps.passSlot = m.x >> 26;
    ps.sampler.salt = m.y;
    ps.sampler.dim = m.z & 0xFFFu;
    ps.depth = (m.z >> 12) & 0x7Fu;
    ps.lastSpecular = (m.z >> 19) & 1u;
    ps.lastMaterial = (m.z >> 20) == 0xFFFu ? RTB_NO_MATERIAL : (m.z >> 20);
    ps.lastPdfW = __uint_as_float(m.w);
    ps.sampler.seedBase = ps.passSlot * numSeedDims;
    // The raster position is behind a dependent gather (slot -> pixelList), and what needs it is rare: the blue-noise offset
    // serves sampler dimensions 0..3 only (GenericSampler.cpp:88-96) - a path is past them after its first vertex - and the
    // raster index keys the `Single` light pick.  Paths that need neither skip the gather (ncu, round 2: 10 % of k_shade's
    // stall samples sat on this load and on the division by the width behind it).
    if (needRaster || ps.sampler.dim < 4u)
    {
        const uint32_t pix = W.pixelList[ps.pixel];
        ps.rasterPixel = pix;
        const uint32_t x = pix % width, y = pix / width;
        ps.sampler.bnOffset = 4u * (128u * (y & 127u) + (x & 127u));
    }
    else { ps.rasterPixel = 0u; ps.sampler.bnOffset = 0u; }
}

// ---------------------------------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------------------------------
// This file is compiled TWICE (build.sh): as the library proper, and once more with -DRTB_TU_TAIL=1 -Xptxas -O1 for k_tail alone.
// ptxas' code for k_tail at -O2 and -O3 returns different hits / path lengths for a few hundred paths per frame whenever CTAs of
// other streams' kernels are co-resident on its SMs, from identical inputs (checksummed) and whatever its register budget,
// load flavour or vote exits; the -O1 code of the same source is bit-identical in every repetition.
#ifndef RTB_TU_TAIL
#define RTB_TU_TAIL 0
#endif
void rtb_launch_tail(bool textured, int blocks, size_t smemBytes, cudaStream_t stream, const SceneView& S, const PassParams& P, const struct Wavefront& W, uint32_t depth);
void rtb_configure_tail(int maxDynamicSmem, int carveoutPercent /* < 0: leave it */);

#if !RTB_TU_TAIL

__global__ void __launch_bounds__(256) k_raygen(const __grid_constant__ SceneView S, const __grid_constant__ PassParams P, const __grid_constant__ Wavefront W)
{
    const uint32_t stride = gridDim.x * blockDim.x;
    const bool allActive = W.numActive == RTB_ALL_ACTIVE;
    const uint32_t perPass = allActive ? W.numPixels : W.numActive;
    const uint32_t total = perPass * W.batch;
    if (blockIdx.x == 0)
    {
        // queue counters of this wavefront (the kernels that read them are launched after this one on the same stream)
        uint32_t* words = reinterpret_cast<uint32_t*>(W.ctl);
        for (uint32_t i = threadIdx.x; i < sizeof(Control) / 4u; i += blockDim.x) words[i] = i == 0u ? total : 0u;
        static_assert(offsetof(Control, pathCount) == 0, "pathCount[0] is word 0");
    }
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride)
    {
        const uint32_t passSlot = i / perPass, a = i - passSlot * perPass;
        const uint32_t slot = allActive ? a : W.activeList[a];
        const uint32_t pix = W.pixelList[slot];
        PathState ps;
        init_path(P, S.blueNoise, pix % P.width, pix / P.width, passSlot, ps);
        ps.pixel = slot;
        store_path(W.buf[0], i, ps);
        W.passColor[passSlot * W.numPixels + slot] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);      // (inactive slots: zeroed by the host's memset)
    }
}

template <bool kCount>
__global__ void __launch_bounds__(128, RTB_TRAV_MINBLOCKS) k_traverse(const __grid_constant__ SceneView S, const __grid_constant__ Wavefront W, uint32_t depth)
{
    const uint32_t count = W.ctl->tailTaken ? 0u : W.ctl->pathCount[depth];
    const PathBuf& in = W.buf[depth & 1];
    const uint32_t stride = gridDim.x * blockDim.x;
    TraceCounters cnt; cnt.boxTests = 0; cnt.triTests = 0;
#if RTB_SMEM_STACK > 0
    extern __shared__ uint32_t smemStack[];     // stackEntries x 128 words (dynamic, sized by the scene's BVH depth)
    SharedStack sstack = (SharedStack)__cvta_generic_to_shared(smemStack) + threadIdx.x * 4u;
    asm volatile("" : "+r"(sstack));             // keep it in a register (see stk_put)
#else
    const SharedStack sstack = 0;
#endif
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride)
    {
        const RayInPathQueue src = { in, i };
        Hit hit;
        hit.objectId = RTB_INVALID_OBJECT; hit.subObjectId = 0; hit.distance = INFINITY; hit.u = 0.0f; hit.v = 0.0f;
        scene_traverse_from<kCount>(S, src, hit, &cnt, sstack);
        W.hit0[i] = make_uint4(hit.objectId, hit.subObjectId, __float_as_uint(hit.distance), __float_as_uint(hit.u));
        W.hit1[i] = hit.v;
    }
    if (kCount)
    {
        atomicAdd(&W.tot->numBoxTests, (unsigned long long)cnt.boxTests);
        atomicAdd(&W.tot->numTriTests, (unsigned long long)cnt.triTests);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&W.tot->numRays, (unsigned long long)count);
}

// Closest-hit rays on persistent warps (RTB_TRAV_KERNEL=persistent): RTB_TRAV_MINBLOCKS CTAs per SM stay resident, a warp
// draws paths from the queue with one atomic whenever at least RTB_TP_REFILL of its lanes are idle, and every lane runs
// the resumable TraverseWalk.  Same visits per ray as k_traverse, fuller warps.
// (measured and dropped, round 2: `prefetch.global.L2` of the queue entries 2 048 / 8 192 positions ahead of the cursor in both
//  persistent kernels - k_shadow_wide 18.31 -> 18.78 ms per 8 passes, k_traverse_p unchanged; the same for the path state of
//  k_shade's tile during its sort, 8.63 -> 8.73: the queues stream from HBM but their latency is already covered)
#ifndef RTB_TP_REFILL
#define RTB_TP_REFILL 24      // (k_traverse_p per launch on the bench: 4: 1.387, 8: 1.338, 12: 1.285, 20: 1.197, 24: 1.174, 28: 1.175 ms; k_traverse 1.246)
#endif
template <bool kCount>
__global__ void __launch_bounds__(128, RTB_TRAV_MINBLOCKS) k_traverse_p(const __grid_constant__ SceneView S, const __grid_constant__ Wavefront W, uint32_t depth)
{
    const uint32_t count = W.ctl->tailTaken ? 0u : W.ctl->pathCount[depth];
    const PathBuf& in = W.buf[depth & 1];
    const uint32_t lane = threadIdx.x & 31u;
    TraceCounters cnt; cnt.boxTests = 0; cnt.triTests = 0;
    extern __shared__ uint32_t smemStack[];
    SharedStack sstack = (SharedStack)__cvta_generic_to_shared(smemStack) + threadIdx.x * 4u;
    asm volatile("" : "+r"(sstack));
    uint32_t stack[1];
    TraverseWalk walk;
    bool active = false, exhausted = false;
    uint32_t i = 0;
    for (;;)
    {
        const uint32_t idle = __ballot_sync(0xFFFFFFFFu, !active);
        if (!exhausted && (idle == 0xFFFFFFFFu || __popc(idle) >= RTB_TP_REFILL))
        {
            const uint32_t want = __popc(idle);
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(&W.ctl->pathCursor[depth], want);
            base = __shfl_sync(0xFFFFFFFFu, base, 0);
            exhausted = base + want >= count;
            if (!active)
            {
                const uint32_t k = base + __popc(idle & ((1u << lane) - 1u));
                if (k < count)
                {
                    i = k;
                    const RayInPathQueue src = { in, i };
                    if (walk.begin(S, stack, src, sstack)) active = true;
                    else
                    {
                        W.hit0[i] = make_uint4(walk.hit.objectId, walk.hit.subObjectId, __float_as_uint(walk.hit.distance), __float_as_uint(walk.hit.u));
                        W.hit1[i] = walk.hit.v;
                    }
                }
            }
        }
        if (__ballot_sync(0xFFFFFFFFu, active) == 0u)
        {
            if (exhausted) break;
            continue;
        }
        if (active)
        {
            const RayInPathQueue src = { in, i };
            if (!walk.template step<kCount>(S, stack, src, &cnt, sstack))
            {
                active = false;
                W.hit0[i] = make_uint4(walk.hit.objectId, walk.hit.subObjectId, __float_as_uint(walk.hit.distance), __float_as_uint(walk.hit.u));
                W.hit1[i] = walk.hit.v;
            }
        }
    }
    if (kCount)
    {
        atomicAdd(&W.tot->numBoxTests, (unsigned long long)cnt.boxTests);
        atomicAdd(&W.tot->numTriTests, (unsigned long long)cnt.triTests);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&W.tot->numRays, (unsigned long long)count);
}

struct DeviceSink
{
    V3 result;
    bool dirty;
    const Wavefront& W;
    uint32_t rayBase, numRays;
    V4 lightThr;
    __device__ DeviceSink(const Wavefront& w, uint32_t base) : W(w), rayBase(base), numRays(0) { dirty = false; result = v3s(0.0f); lightThr = v4s(0.0f); }
    __device__ void add(const V4& thr, const V4& c) { result = mul_add(xyz(thr), xyz(c), result); dirty = true; }
    // slots are filled in emission order = light order (a measured alternative, one ray list per light so that rays
    // towards the same light are traced together, did not pay: 40.9 vs 40.4 ms per 8 passes for k_shadow)
    __device__ void shadow(const Ray& ray, float tMax, const V4& c, uint32_t)
    {
        const uint32_t j = rayBase + numRays++;
        W.sr0[j] = make_float4(ray.origin.x, ray.origin.y, ray.origin.z, tMax);
        W.sr1[j] = make_float4(ray.dir.x, ray.dir.y, ray.dir.z, c.x);
        W.sr2[j] = make_float4(ray.invDir.x, ray.invDir.y, ray.invDir.z, c.y);
        W.sr3[j] = make_float4(ray.originDivDir.x, ray.originDivDir.y, ray.originDivDir.z, c.z);
        W.shOcc[j] = 0u;
    }
    __device__ void end_lights(const V4& thr) { lightThr = thr; }
};

// ---- material binning before shading (north_star: "material sorting before shade") ---------------------------------------
// shade_vertex_* is one large switch: miss / light hit / surface hit, then nine BSDFs (sample + evaluate per light), and
// a warp executes every branch any of its lanes takes.  Each CTA therefore takes a TILE of RTB_SHADE_TILE consecutive paths,
// computes a key per path (miss, light, or BSDF of the material at the hit x mesh / analytic shape) and counting-sorts the
// tile's indices by key in shared memory (histogram with shared atomics, 32-bin scan by one warp, scatter); its threads then
// shade the tile in key order, so all but the warps that straddle a bin boundary run ONE branch.  The sort is local to a
// CTA: no extra kernel, no extra pass over HBM (the 16-byte hit records the keys need are read again from L1 by the
// shading itself), the path state of a tile is gathered from a 1024-path window that the CTA consumes entirely.
// A path owns its pixel's pass colour and its own shadow-ray slots, so the order in which paths are shaded cannot change
// the film (the compacted output order is not deterministic to begin with: one atomic per warp).
#ifndef RTB_SHADE_TILE
#define RTB_SHADE_TILE 512u        // a multiple of 128 (k_shade per 8 passes at 5 CTAs / SM: 256: 8.39 ms, 512: 8.16, 1024: 8.30, 2048: 9.16)
#endif
#define RTB_SHADE_BINS 32u
// (measured and dropped, round 2: asking for the NEXT path's state between the issue of the queue atomics and the first use of
//  their results - with `prefetch.global.L1` 7.86 -> 7.89 ms per 8 passes, with real loads into registers 8.35: the waits that
//  ncu's per-line view shows on those loads are covered by the other warps, the kernel does not get faster by moving them)

__device__ __forceinline__ uint32_t shade_key(const SceneView& S, uint32_t objectId, uint32_t subObjectId)
{
    if (objectId == RTB_INVALID_OBJECT) return 0u;                       // miss: background lights
    if (subObjectId == RTB_LIGHT_OBJECT) return 1u;                      // a light hit directly
    const DObjectShade& obj = S.objShade[objectId];
    uint32_t material = obj.materialIndex;
    const RtbShape& shape = S.shapes[obj.shapeIndex];
    const bool mesh = shape.type == RTB_SHAPE_MESH;
    if (mesh)
    {
        const DMesh& m = S.meshes[shape.meshIndex];
        const uint32_t slot = m.indices[subObjectId].materialIndex;
        if (slot != 0xFFFFFFFFu) material = m.materialIds[slot];
    }
    const uint32_t bsdf = material != RTB_NO_MATERIAL ? S.materials[material].bsdf : 14u;
    return 2u + (bsdf < 14u ? bsdf : 14u) + (mesh ? 0u : 15u);       // 2..16 mesh, 17..31 analytic shapes
}

// kRenderer: one instantiation per integrator, so that a launch carries one integrator's code only (the three together
// are 13 600 SASS instructions against a 32 KB instruction cache, and the register allocation of the common part was
// the worst case of the three)
template <bool kTex, bool kSort, uint32_t kRenderer>
__global__ void __launch_bounds__(128, RTB_SHADE_MINBLOCKS) k_shade(const __grid_constant__ SceneView S, const __grid_constant__ PassParams P, const __grid_constant__ Wavefront W, uint32_t depth)
{
    const uint32_t count = W.ctl->tailTaken ? 0u : W.ctl->pathCount[depth];
    const PathBuf& in = W.buf[depth & 1];
    const PathBuf& out = W.buf[(depth + 1) & 1];
    const uint32_t lane = threadIdx.x & 31u;
    unsigned long long shadowRaysLocal = 0;
    __shared__ uint16_t tileOrder[kSort ? RTB_SHADE_TILE : 1u];
    __shared__ uint8_t tileKeys[kSort ? RTB_SHADE_TILE : 1u];
    __shared__ uint32_t binCursor[RTB_SHADE_BINS];
#pragma unroll 1
    for (uint32_t tileBase = blockIdx.x * RTB_SHADE_TILE; tileBase < count; tileBase += gridDim.x * RTB_SHADE_TILE)
    {
    const uint32_t tileCount = min(RTB_SHADE_TILE, count - tileBase);
    if (kSort)
    {
        __syncthreads();                                   // the previous tile's order is no longer read
        if (threadIdx.x < RTB_SHADE_BINS) binCursor[threadIdx.x] = 0u;
        __syncthreads();
        // (a key is the end of a chain of dependent loads - hit, object, shape, mesh, triangle, material: the thread's
        //  RTB_SHADE_TILE / 128 chains are written side by side so that their latencies overlap instead of adding up)
        {
            constexpr uint32_t kPerThread = RTB_SHADE_TILE / 128u;
            uint4 h0[kPerThread];
            uint32_t keys[kPerThread];
#pragma unroll
            for (uint32_t k = 0; k < kPerThread; ++k)
            {
                const uint32_t j = threadIdx.x + 128u * k;
                h0[k] = j < tileCount ? W.hit0[tileBase + j] : make_uint4(RTB_INVALID_OBJECT, 0u, 0u, 0u);
            }
#pragma unroll
            for (uint32_t k = 0; k < kPerThread; ++k) keys[k] = shade_key(S, h0[k].x, h0[k].y);
#pragma unroll
            for (uint32_t k = 0; k < kPerThread; ++k)
            {
                const uint32_t j = threadIdx.x + 128u * k;
                if (j < tileCount) { tileKeys[j] = (uint8_t)keys[k]; atomicAdd(&binCursor[keys[k]], 1u); }
            }
        }
        __syncthreads();
        if (threadIdx.x < 32u)
        {
            const uint32_t v = binCursor[lane];
            uint32_t incl = v;
            for (int o = 1; o < 32; o <<= 1)
            {
                const uint32_t up = __shfl_up_sync(0xFFFFFFFFu, incl, o);
                if (lane >= (uint32_t)o) incl += up;
            }
            binCursor[lane] = incl - v;
        }
        __syncthreads();
        for (uint32_t j = threadIdx.x; j < tileCount; j += 128u) tileOrder[atomicAdd(&binCursor[tileKeys[j]], 1u)] = (uint16_t)j;
        __syncthreads();
    }
    // (loop bound rounded up to a whole warp so that every lane takes part in the ballots)
    const bool needRaster = kRenderer == RTB_RENDERER_PATH_TRACER_MIS && !P.lightSamplingAll;
#pragma unroll 1
    for (uint32_t j = threadIdx.x; j < ((tileCount + 31u) & ~31u); j += 128u)
    {
        const bool valid = j < tileCount;
        const uint32_t i = tileBase + (kSort ? (valid ? (uint32_t)tileOrder[j] : 0u) : j);
        bool alive = false;
        uint32_t numShadow = 0;
        PathState ps;
        if (valid)
        {
            Hit hit;
            load_path(in, i, W, P.width, P.numSeedDims, ps, needRaster);
            const uint4 h0 = W.hit0[i];
            hit.objectId = h0.x; hit.subObjectId = h0.y; hit.distance = __uint_as_float(h0.z); hit.u = __uint_as_float(h0.w);
            hit.v = W.hit1[i];

            const uint32_t colorIndex = ps.passSlot * W.numPixels + ps.pixel;
            const float4 pc = W.passColor[colorIndex];
            DeviceSink sink(W, i * W.raysPerPath);
            sink.result = v3(pc.x, pc.y, pc.z);
            if (kRenderer == RTB_RENDERER_PATH_TRACER_MIS) alive = shade_vertex_mis<kTex>(S, P, ps, hit, sink, numShadow);
            else if (kRenderer == RTB_RENDERER_PATH_TRACER) alive = shade_vertex_pt<kTex>(S, P, ps, hit, sink);
            else shade_debug<kTex>(S, P, ps, hit, sink);
            if (sink.dirty) W.passColor[colorIndex] = make_float4(sink.result.x, sink.result.y, sink.result.z, 0.0f);
            if (numShadow)
            {
                W.shThr[i] = make_float4(sink.lightThr.x, sink.lightThr.y, sink.lightThr.z, __uint_as_float(colorIndex));
                W.shCount[i] = numShadow;
                shadowRaysLocal += numShadow;
            }
        }
        // stream compaction of the survivors, compacted lists of the paths that own shadow rays and of the rays themselves
        // (warp prefix sum): three queue counters, bumped by lanes 0..2 in ONE atomic instruction - one L2 round trip per
        // warp and iteration where three atomics in a row cost two (ncu, round 2: 7 % of k_shade's stall samples)
        const uint32_t aliveMask = __ballot_sync(0xFFFFFFFFu, alive);
        const uint32_t shMask = __ballot_sync(0xFFFFFFFFu, numShadow != 0);
        uint32_t incl = numShadow;
        if (shMask)
        {
            for (int o = 1; o < 32; o <<= 1)
            {
                const uint32_t up = __shfl_up_sync(0xFFFFFFFFu, incl, o);
                if (lane >= (uint32_t)o) incl += up;
            }
        }
        const uint32_t total = __shfl_sync(0xFFFFFFFFu, incl, 31);
        uint32_t got = 0;
        {
            uint32_t* const counter = lane == 0u ? &W.ctl->pathCount[depth + 1] : lane == 1u ? &W.ctl->shadowPathCount[depth] : &W.ctl->shadowRayCount[depth];
            const uint32_t amount = lane == 0u ? (uint32_t)__popc(aliveMask) : lane == 1u ? (uint32_t)__popc(shMask) : total;
            if (lane < 3u && amount != 0u) got = atomicAdd(counter, amount);
        }
        const uint32_t base = __shfl_sync(0xFFFFFFFFu, got, 0);
        const uint32_t pathBase = __shfl_sync(0xFFFFFFFFu, got, 1);
        const uint32_t rayBase = __shfl_sync(0xFFFFFFFFu, got, 2);
        if (alive) store_path(out, base + __popc(aliveMask & ((1u << lane) - 1u)), ps);
        if (shMask)
        {
            if (numShadow)
            {
                W.shList[pathBase + __popc(shMask & ((1u << lane) - 1u))] = i;
                const uint32_t first = rayBase + incl - numShadow;
                for (uint32_t r = 0; r < numShadow; ++r) W.rayList[first + r] = i * W.raysPerPath + r;
            }
        }
    }
    }
    // counters: one atomic per warp
    for (int o = 16; o > 0; o >>= 1) shadowRaysLocal += __shfl_down_sync(0xFFFFFFFFu, shadowRaysLocal, o);
    if (lane == 0 && shadowRaysLocal) atomicAdd(&W.tot->numShadowRays, shadowRaysLocal);
}

// k_shade of the pass' integrator (the wavefront runs "Path Tracer", "Path Tracer MIS" and "Debug")
template <bool kTex, bool kSort>
static void launch_shade(int blocks, cudaStream_t s, const SceneView& S, const PassParams& P, const Wavefront& W, uint32_t depth)
{
    if (P.renderer == RTB_RENDERER_PATH_TRACER_MIS) k_shade<kTex, kSort, RTB_RENDERER_PATH_TRACER_MIS><<<blocks, 128, 0, s>>>(S, P, W, depth);
    else if (P.renderer == RTB_RENDERER_PATH_TRACER) k_shade<kTex, kSort, RTB_RENDERER_PATH_TRACER><<<blocks, 128, 0, s>>>(S, P, W, depth);
    else k_shade<kTex, kSort, RTB_RENDERER_DEBUG><<<blocks, 128, 0, s>>>(S, P, W, depth);
}

// RTB_SHADOW_REFILL = n > 0: persistent warps that fetch new shadow rays when at least n lanes are idle.  Measured on the
// bench scene (k_shadow + k_resolve per 8 passes): no refill 33.2 ms, n = 24: 33.0, 16: 33.6, 8: 37.0, 4: 38.3, 1: 38.4.
// Fuller warps do not help: the kernel is bound by L1 data-pipe wavefronts and L2 latency (ncu: l1tex wavefronts
// 65-70 % of peak at 10-12 active lanes), and mixed-age rays lose the shared top-of-tree fetches.  Kept off.
#ifndef RTB_SHADOW_REFILL
#define RTB_SHADOW_REFILL 0
#endif

template <bool kCount>
__global__ void __launch_bounds__(128, RTB_TRAV_MINBLOCKS) k_shadow(const __grid_constant__ SceneView S, const __grid_constant__ Wavefront W, uint32_t depth)
{
    const uint32_t count = W.ctl->tailTaken ? 0u : W.ctl->shadowRayCount[depth];
    const uint32_t lane = threadIdx.x & 31u;
    unsigned long long visibleLocal = 0;
    TraceCounters cnt; cnt.boxTests = 0; cnt.triTests = 0;
#if RTB_SMEM_STACK > 0
    extern __shared__ uint32_t smemStack[];     // stackEntries x 128 words (dynamic, sized by the scene's BVH depth)
    SharedStack sstack = (SharedStack)__cvta_generic_to_shared(smemStack) + threadIdx.x * 4u;
    asm volatile("" : "+r"(sstack));             // keep it in a register (see stk_put)
#else
    const SharedStack sstack = 0;
#endif
#if RTB_SHADOW_REFILL
    // persistent warps with dynamic fetch (experiment, see above)
    ShadowWalk walk;
    uint32_t stack[RTB_LOCAL_STACK_SIZE];
    bool active = false, exhausted = false;
    uint32_t j = 0;
    for (;;)
    {
        const uint32_t idle = __ballot_sync(0xFFFFFFFFu, !active);
        if (!exhausted && (idle == 0xFFFFFFFFu || __popc(idle) >= RTB_SHADOW_REFILL))
        {
            const uint32_t want = __popc(idle);
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(&W.ctl->shadowCursor[depth], want);
            base = __shfl_sync(0xFFFFFFFFu, base, 0);
            exhausted = base + want >= count;
            if (!active)
            {
                const uint32_t k = base + __popc(idle & ((1u << lane) - 1u));
                if (k < count)
                {
                    j = W.rayList[k];
                    const float4 a = W.sr0[j], b = W.sr1[j], c = W.sr2[j], d = W.sr3[j];
                    Ray ray;
                    ray.origin = v3(a.x, a.y, a.z); ray.dir = v3(b.x, b.y, b.z); ray.invDir = v3(c.x, c.y, c.z); ray.originDivDir = v3(d.x, d.y, d.z);
                    const RayInRegisters src = { ray };
                    if (walk.begin(S, stack, src, a.w, sstack)) active = true;
                    else visibleLocal++;
                }
            }
        }
        if (__ballot_sync(0xFFFFFFFFu, active) == 0u)
        {
            if (exhausted) break;
            continue;
        }
        if (active)
        {
            const RayInShadowQueue src = { W, j };
            const int r = walk.template step<kCount, RayInShadowQueue, false, false>(S, stack, src, &cnt, sstack);
            if (r != ShadowWalk::kContinue)
            {
                active = false;
                // SampleLight returns RayColor::Zero() for an occluded light (PathTracerMIS.cpp:92-96)
                if (r == ShadowWalk::kOccluded) { W.shOcc[j] = 1u; }
                else visibleLocal++;
            }
        }
    }
#else
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < count; k += stride)
    {
        const uint32_t j = W.rayList[k];
        const RayInShadowQueue src = { W, j };
        const bool occluded = scene_traverse_shadow_from<kCount, RayInShadowQueue, false>(S, src, W.sr0[j].w, &cnt, sstack);      // the binary tree
        if (occluded)
        {
            // SampleLight returns RayColor::Zero() for an occluded light (PathTracerMIS.cpp:92-96)
            W.shOcc[j] = 1u;
        }
        else visibleLocal++;
    }
#endif
    for (int o = 16; o > 0; o >>= 1) visibleLocal += __shfl_down_sync(0xFFFFFFFFu, visibleLocal, o);
    if (lane == 0 && visibleLocal) atomicAdd(&W.tot->numShadowRaysHit, visibleLocal);
    if (kCount)
    {
        atomicAdd(&W.tot->numShadowBoxTests, (unsigned long long)cnt.boxTests);
        atomicAdd(&W.tot->numShadowTriTests, (unsigned long long)cnt.triTests);
    }
}


// ---- any-hit rays on the 8-wide tree: persistent warps, top of the tree staged in shared memory ---------------------------
// k_shadow above gives every ray a thread and lets it run the two-level walk on its own.  k_shadow_wide is the kernel
// built around rtb_wide.h:
//   * persistent CTAs (RTB_SW_MINBLOCKS per SM); a warp draws 32 rays at a time from the ray list with one atomic
//     (W.ctl->shadowCursor) until the list is empty;
//   * the first `hotCount` nodes of the scene's largest 8-wide tree (breadth-first order = the top levels every ray
//     goes through) are copied into shared memory once per CTA with ONE bulk copy (cp.async.bulk, completion on an
//     mbarrier) that overlaps the first top-level steps; nodes below come from L1 / L2 with 256-bit loads;
//   * the walk of a warp's 32 rays alternates between two phases.  Top phase: every lane advances its exact top-level
//     walk (rtb_trace.h ShadowWalk: top-level BVH, analytic shapes, meshes without a wide tree) until it is finished or
//     stands in front of a mesh that has a wide tree.  Wide phase: all lanes that stand in front of a mesh walk it
//     TOGETHER - eight child tests per node with no branch in between - until each has its answer.  With one
//     dominant mesh (the bench) that is one wide phase per ray batch.
#ifndef RTB_SW_THREADS
#define RTB_SW_THREADS 384
#endif
#ifndef RTB_SW_MINBLOCKS
#define RTB_SW_MINBLOCKS 2
#endif
#ifndef RTB_SW_DEFER
#define RTB_SW_DEFER 1      // leaf confirmations as a warp-wide phase of their own (WideWalkDeferred) instead of inside the node visit
#endif
#ifndef RTB_SW_LEAF_BIAS
#define RTB_SW_LEAF_BIAS 2  // the confirmation phase runs when (lanes with a queued leaf) * bias >= (lanes with a node to visit) * 2
#endif
#ifndef RTB_SW_REFILL
#define RTB_SW_REFILL 16    // the wide phase is left for a refill when fewer than this many lanes of the warp are still inside a mesh
                            // (k_shadow_wide per 8 bench passes, 384 x 2: never 35.4 / 36.4 / 34.1 / 36.3 ms, 20: 34.1 / 34.1 / 34.3 / 35.6)
#endif

struct WideFetchHot
{
    const WNode* hotNodes;     // the tree whose first hotCount nodes are staged
    uint32_t hotCount;
    uint32_t hotShared;        // shared-space byte address of the staged copy
    __device__ __forceinline__ void operator()(const WNode* nodes, uint32_t index, uint32_t* n) const
    {
        if (nodes == hotNodes && index < hotCount)
        {
            const uint32_t a = hotShared + index * 96u;
            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(n[0]), "=r"(n[1]), "=r"(n[2]), "=r"(n[3]) : "r"(a));
            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4+16];" : "=r"(n[4]), "=r"(n[5]), "=r"(n[6]), "=r"(n[7]) : "r"(a));
            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4+32];" : "=r"(n[8]), "=r"(n[9]), "=r"(n[10]), "=r"(n[11]) : "r"(a));
            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4+48];" : "=r"(n[12]), "=r"(n[13]), "=r"(n[14]), "=r"(n[15]) : "r"(a));
            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4+64];" : "=r"(n[16]), "=r"(n[17]), "=r"(n[18]), "=r"(n[19]) : "r"(a));
        }
        else WideFetchGlobal()(nodes, index, n);
    }
};

template <bool kCount, bool kHot>
__global__ void __launch_bounds__(RTB_SW_THREADS, RTB_SW_MINBLOCKS) k_shadow_wide(const __grid_constant__ SceneView S, const __grid_constant__ Wavefront W, uint32_t depth,
                                                                                  const WNode* __restrict__ hotNodes, uint32_t hotCount, uint32_t panelBytes)
{
    extern __shared__ __align__(128) uint8_t smemRaw[];
    // [0, 8): mbarrier   [128, 128 + hotCount * 96): staged nodes   then one stack panel (rtb_trace.h layout) per 128 threads
    const uint32_t smemBase = (uint32_t)__cvta_generic_to_shared(smemRaw);
    const uint32_t hotShared = smemBase + 128u;
    const uint32_t hotBytes = hotCount * (uint32_t)sizeof(WNode);
    const uint32_t stackBase = hotShared + ((hotBytes + 127u) & ~127u);
    const uint32_t count = W.ctl->tailTaken ? 0u : W.ctl->shadowRayCount[depth];
    if (kHot && hotCount)
    {
        if (threadIdx.x == 0)
        {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smemBase));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smemBase), "r"(hotBytes) : "memory");
            const char* src = reinterpret_cast<const char*>(hotNodes);
            for (uint32_t off = 0; off < hotBytes; off += 32768u)
            {
                const uint32_t chunk = hotBytes - off < 32768u ? hotBytes - off : 32768u;
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             :: "r"(hotShared + off), "l"(src + off), "r"(chunk), "r"(smemBase) : "memory");
            }
        }
        __syncthreads();       // the barrier is initialised before anyone waits on it
    }
    // (the instantiation without staged nodes has none of this in its node visit: the staging is an option that did not pay, §4a)
    bool hotReady = !kHot || hotCount == 0u;
    typename std::conditional<kHot, WideFetchHot, WideFetchGlobal>::type fetch;
    if constexpr (kHot) { fetch.hotNodes = hotNodes; fetch.hotCount = hotCount; fetch.hotShared = hotShared; }

    const uint32_t lane = threadIdx.x & 31u;
    SharedStack sstack = stackBase + (threadIdx.x >> 7) * panelBytes + (threadIdx.x & 127u) * 4u;
    asm volatile("" : "+r"(sstack));
    unsigned long long visibleLocal = 0;
    TraceCounters cnt; cnt.boxTests = 0; cnt.triTests = 0;
    uint32_t stack[1];
    enum { kIdle = 0, kTop = 1, kWide = 2 };
    int state = kIdle;
    bool exhausted = false;
    uint32_t j = 0, wideTop = 0;
    ShadowWalk walk;
#if RTB_SW_DEFER
    WideWalkDeferred wide;
    const uint32_t leafTop = panelBytes / (RTB_STACK_THREADS * 4u);      // the leaf queue hangs from the top of the lane's column
#else
    WideWalk wide;
#endif
    WideView wv; wv.nodes = nullptr; wv.tris = nullptr; wv.refNodes = nullptr; wv.maxAbs = 0.0f;
    for (;;)
    {
        // ---- idle lanes draw new rays (one atomic per warp) ------------------------------------------------------------
        const uint32_t idle = __ballot_sync(0xFFFFFFFFu, state == kIdle);
        if (idle != 0u && !exhausted)
        {
            const uint32_t want = __popc(idle);
            uint32_t k0 = 0;
            if (lane == 0) k0 = atomicAdd(&W.ctl->shadowCursor[depth], want);
            k0 = __shfl_sync(0xFFFFFFFFu, k0, 0);
            exhausted = k0 + want >= count;
            const uint32_t k = k0 + __popc(idle & ((1u << lane) - 1u));
            if (state == kIdle && k < count)
            {
                j = W.rayList[k];
                const RayInShadowQueue src = { W, j };
                if (walk.begin(S, stack, src, W.sr0[j].w, sstack)) state = kTop;
                else visibleLocal++;
            }
        }
        if (__ballot_sync(0xFFFFFFFFu, state != kIdle) == 0u)
        {
            if (exhausted) break;
            continue;
        }
        const RayInShadowQueue src = { W, j };
        // ---- top phase: the exact two-level walk up to the next mesh that has a wide tree ---------------------------------------
        while (state == kTop)
        {
            const int r = walk.template step<kCount, RayInShadowQueue, true>(S, stack, src, &cnt, sstack);
            if (r == ShadowWalk::kEnterWide)
            {
                state = kWide;
                const DMesh& mesh = S.meshes[walk.wideMesh];
                wv.nodes = mesh.wideNodes; wv.tris = mesh.wideTris; wv.refNodes = mesh.nodes;
                wide.begin(walk.ray, mesh.wideMaxAbs);
                wideTop = walk.stackSize;
            }
            else if (r != ShadowWalk::kContinue)
            {
                state = kIdle;
                // SampleLight returns RayColor::Zero() for an occluded light (PathTracerMIS.cpp:92-96)
                if (r == ShadowWalk::kOccluded) { W.shOcc[j] = 1u; }
                else visibleLocal++;
            }
        }
        if (!hotReady)
        {
            // the staged nodes have landed (phase 0 of the barrier completes when all bytes arrived)
            uint32_t done = 0;
            while (!done)
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smemBase) : "memory");
            hotReady = true;
        }
        // ---- wide phase: one node visit per iteration for every lane that is inside a mesh; it ends when those lanes are
        //      done, or - while there are rays left to draw - when so few are left inside that a refill pays ---------------------------
#if RTB_SW_DEFER
        // a node visit only queues the leaves it touches; the confirmations (triangle + leaf box) run as a phase of their
        // own once enough lanes have one queued, so that both halves of the walk run at many lanes (rtb_wide.h)
        for (;;)
        {
            WideStackAbove ws; ws.stack = stack; ws.sstack = sstack; ws.floor = walk.stackSize; ws.top = wideTop;
            const bool visit = state == kWide && wide.has_nodes(ws) && wide.can_visit();
            const bool confirm = state == kWide && wide.leafCount != 0u;
            const uint32_t numVisit = __popc(__ballot_sync(0xFFFFFFFFu, visit));
            const uint32_t numConfirm = __popc(__ballot_sync(0xFFFFFFFFu, confirm));
            if (numConfirm != 0u && (numVisit == 0u || numConfirm * RTB_SW_LEAF_BIAS >= numVisit * 2u))
            {
                if (confirm && wide.template leaf_step<kCount>(wv, walk.ray, walk.maxDistance, ws, sstack, leafTop, &cnt))
                {
                    state = kIdle;
                    W.shOcc[j] = 1u;
                }
            }
            else if (visit)
            {
                wide.template node_step<kCount>(wv, walk.ray, walk.maxDistance, ws, sstack, leafTop, &cnt, fetch);
                wideTop = ws.top;
            }
            if (state == kWide && wide.leafCount == 0u && !wide.has_nodes(ws)) { walk.leave_wide(S, src); state = kTop; }
            const uint32_t inside = __ballot_sync(0xFFFFFFFFu, state == kWide);
            if (inside == 0u) break;
            if (!exhausted && __popc(inside) < RTB_SW_REFILL) break;
        }
#else
        for (;;)
        {
            if (state == kWide)
            {
                WideStackAbove ws; ws.stack = stack; ws.sstack = sstack; ws.floor = walk.stackSize; ws.top = wideTop;
                const int r = wide.template step<kCount>(wv, walk.ray, walk.maxDistance, ws, &cnt, fetch);
                wideTop = ws.top;
                if (r == WideWalk::kOccluded)
                {
                    state = kIdle;
                    W.shOcc[j] = 1u;
                }
                else if (r == WideWalk::kMissed) { walk.leave_wide(S, src); state = kTop; }
            }
            const uint32_t inside = __ballot_sync(0xFFFFFFFFu, state == kWide);
            if (inside == 0u) break;
            if (!exhausted && __popc(inside) < RTB_SW_REFILL) break;
        }
#endif
    }
    for (int o = 16; o > 0; o >>= 1) visibleLocal += __shfl_down_sync(0xFFFFFFFFu, visibleLocal, o);
    if (lane == 0 && visibleLocal) atomicAdd(&W.tot->numShadowRaysHit, visibleLocal);
    if (kCount)
    {
        atomicAdd(&W.tot->numShadowBoxTests, (unsigned long long)cnt.boxTests);
        atomicAdd(&W.tot->numShadowTriTests, (unsigned long long)cnt.triTests);
    }
}

// ---- any-hit rays, split walk ------------------------------------------------------------------------------------------------
// "Occluded" is an OR over everything a ray meets, so the top level and the meshes need not be walked by the same thread
// in the same kernel.  k_shadow_top walks the exact top level for every ray (all lanes start together: converged), tests
// the analytic shapes on the spot and QUEUES every visit of a mesh that has an 8-wide tree as an item (the mesh-local
// ray).  k_shadow_mesh then does nothing but wide walks: persistent warps whose lanes fetch a new item the moment they
// finish one (an item is 48 bytes, there is no other per-ray state to set up), so the eight-child node visits run at
// nearly full warps.  A ray found occluded has its contribution zeroed and its flag set exactly once (the counter of
// visible rays is corrected through it).
#ifndef RTB_SM_THREADS
#define RTB_SM_THREADS 256
#endif
#ifndef RTB_SM_MINBLOCKS
#define RTB_SM_MINBLOCKS 4
#endif
#ifndef RTB_SM_REFILL
#define RTB_SM_REFILL 1      // idle lanes needed before the warp draws new items
#endif

template <bool kCount>
__global__ void __launch_bounds__(128, RTB_TRAV_MINBLOCKS) k_shadow_top(const __grid_constant__ SceneView S, const __grid_constant__ Wavefront W, uint32_t depth)
{
    const uint32_t count = W.ctl->tailTaken ? 0u : W.ctl->shadowRayCount[depth];
    const uint32_t lane = threadIdx.x & 31u;
    unsigned long long visibleLocal = 0;
    TraceCounters cnt; cnt.boxTests = 0; cnt.triTests = 0;
    extern __shared__ uint32_t smemStack[];
    SharedStack sstack = (SharedStack)__cvta_generic_to_shared(smemStack) + threadIdx.x * 4u;
    asm volatile("" : "+r"(sstack));
    uint32_t stack[1];
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < count; k += stride)
    {
        const uint32_t j = W.rayList[k];
        const RayInShadowQueue src = { W, j };
        ShadowWalk walk;
        bool occluded = false;
        if (walk.begin(S, stack, src, W.sr0[j].w, sstack))
        {
            for (;;)
            {
                const int r = walk.template step<kCount, RayInShadowQueue, true>(S, stack, src, &cnt, sstack);
                if (r == ShadowWalk::kContinue) continue;
                if (r == ShadowWalk::kEnterWide)
                {
                    const uint32_t slot = atomicAdd(&W.ctl->meshItemCount[depth], 1u);
                    if (slot < W.miCapacity)
                    {
                        W.mi0[slot] = make_float4(walk.ray.origin.x, walk.ray.origin.y, walk.ray.origin.z, walk.maxDistance);
                        W.mi1[slot] = make_float4(walk.ray.dir.x, walk.ray.dir.y, walk.ray.dir.z, __uint_as_float(j));
                        W.mi2[slot] = make_float4(walk.ray.invDir.x, walk.ray.invDir.y, walk.ray.invDir.z, __uint_as_float(walk.wideMesh));
                    }
                    else
                    {
                        // queue full (a ray visits more meshes than the queue was sized for): walk it here
                        const DMesh& mesh = S.meshes[walk.wideMesh];
                        WideView wv; wv.nodes = mesh.wideNodes; wv.tris = mesh.wideTris; wv.refNodes = mesh.nodes; wv.maxAbs = mesh.wideMaxAbs;
                        WideStackAbove ws; ws.stack = stack; ws.sstack = sstack; ws.floor = walk.stackSize; ws.top = walk.stackSize;
                        if (wide_any_hit<kCount>(wv, walk.ray, walk.maxDistance, ws, &cnt)) { occluded = true; break; }
                    }
                    walk.leave_wide(S, src);
                    continue;
                }
                occluded = r == ShadowWalk::kOccluded;
                break;
            }
        }
        if (occluded) { W.shOcc[j] = 1u; W.shFlag[j] = 1u; }      // SampleLight returns RayColor::Zero() (PathTracerMIS.cpp:92-96)
        else { W.shFlag[j] = 0u; visibleLocal++; }
    }
    for (int o = 16; o > 0; o >>= 1) visibleLocal += __shfl_down_sync(0xFFFFFFFFu, visibleLocal, o);
    if (lane == 0 && visibleLocal) atomicAdd(&W.tot->numShadowRaysHit, visibleLocal);
    if (kCount)
    {
        atomicAdd(&W.tot->numShadowBoxTests, (unsigned long long)cnt.boxTests);
        atomicAdd(&W.tot->numShadowTriTests, (unsigned long long)cnt.triTests);
    }
}

template <bool kCount>
__global__ void __launch_bounds__(RTB_SM_THREADS, RTB_SM_MINBLOCKS) k_shadow_mesh(const __grid_constant__ SceneView S, const __grid_constant__ Wavefront W, uint32_t depth, uint32_t panelBytes)
{
    extern __shared__ uint32_t smemStack[];
    const uint32_t numItems = W.ctl->tailTaken ? 0u : min(W.ctl->meshItemCount[depth], W.miCapacity);
    const uint32_t lane = threadIdx.x & 31u;
    SharedStack sstack = (SharedStack)__cvta_generic_to_shared(smemStack) + (threadIdx.x >> 7) * panelBytes + (threadIdx.x & 127u) * 4u;
    asm volatile("" : "+r"(sstack));
    uint32_t stack[1];
    unsigned long long newlyOccluded = 0;
    TraceCounters cnt; cnt.boxTests = 0; cnt.triTests = 0;
    bool active = false, exhausted = false;
    uint32_t j = 0, top = 0;
    Ray ray;
    float tMax = 0.0f;
    WideWalk wide;
    WideView wv; wv.nodes = nullptr; wv.tris = nullptr; wv.refNodes = nullptr; wv.maxAbs = 0.0f;
    const WideFetchGlobal fetch;
    for (;;)
    {
        const uint32_t idle = __ballot_sync(0xFFFFFFFFu, !active);
        if (!exhausted && __popc(idle) >= RTB_SM_REFILL)
        {
            const uint32_t want = __popc(idle);
            uint32_t k0 = 0;
            if (lane == 0) k0 = atomicAdd(&W.ctl->meshCursor[depth], want);
            k0 = __shfl_sync(0xFFFFFFFFu, k0, 0);
            exhausted = k0 + want >= numItems;
            const uint32_t k = k0 + __popc(idle & ((1u << lane) - 1u));
            if (!active && k < numItems)
            {
                const float4 a = W.mi0[k], b = W.mi1[k], c = W.mi2[k];
                j = __float_as_uint(b.w);
                ray.origin = v3(a.x, a.y, a.z); ray.dir = v3(b.x, b.y, b.z); ray.invDir = v3(c.x, c.y, c.z);
                ray.originDivDir = ray.origin * ray.invDir;                    // make_ray_unsafe (Ray.h:31-39): the same product
                tMax = a.w;
                const DMesh& mesh = S.meshes[__float_as_uint(c.w)];
                wv.nodes = mesh.wideNodes; wv.tris = mesh.wideTris; wv.refNodes = mesh.nodes;
                wide.begin(ray, mesh.wideMaxAbs);
                top = 0;
                active = true;
            }
        }
        if (__ballot_sync(0xFFFFFFFFu, active) == 0u)
        {
            if (exhausted) break;
            continue;
        }
        if (active)
        {
            WideStackAbove ws; ws.stack = stack; ws.sstack = sstack; ws.floor = 0; ws.top = top;
            const int r = wide.template step<kCount>(wv, ray, tMax, ws, &cnt, fetch);
            top = ws.top;
            if (r == WideWalk::kOccluded)
            {
                active = false;
                if (atomicExch(&W.shFlag[j], 1u) == 0u)
                {
                    newlyOccluded++;
                    W.shOcc[j] = 1u;
                }
            }
            else if (r == WideWalk::kMissed) active = false;
        }
    }
    for (int o = 16; o > 0; o >>= 1) newlyOccluded += __shfl_down_sync(0xFFFFFFFFu, newlyOccluded, o);
    if (lane == 0 && newlyOccluded) atomicAdd(&W.tot->numShadowRaysHit, 0ull - newlyOccluded);       // k_shadow_top counted them as visible
    if (kCount)
    {
        atomicAdd(&W.tot->numShadowBoxTests, (unsigned long long)cnt.boxTests);
        atomicAdd(&W.tot->numShadowTriTests, (unsigned long long)cnt.triTests);
    }
}

#endif  // !RTB_TU_TAIL

// ---- the tail of a wavefront ---------------------------------------------------------------------------------------------
// Once only a few ten thousand paths are left, every further bounce costs four launches whose run time is one warp's
// latency, not throughput (a 1/8 shard rendered pass by pass spends more than half of its 2.2 ms there).  k_tail is
// launched after every depth and does nothing while more than `tailThreshold` paths are alive; below that it runs every
// remaining path to its end in ONE launch: the reference's per-pixel loop (PathTracerMIS.cpp:270-396) on one thread per
// path, shadow rays traced on the spot in light order.  Same per-vertex functions, same order of the float additions
// (emission, then the NEE sum of the vertex, then the next vertex), so the film is bit-identical to the wavefront path.
#if RTB_TU_TAIL
struct TailSink
{
    const SceneView& S;
    SharedStack sstack;
    V3 result;
    V3 lightSum;
    uint32_t raysThisVertex, shadowRays, visible;
    __device__ TailSink(const SceneView& s, SharedStack st) : S(s), sstack(st) { result = v3s(0.0f); lightSum = v3s(0.0f); raysThisVertex = 0; shadowRays = 0; visible = 0; }
    __device__ void add(const V4& thr, const V4& c) { result = mul_add(xyz(thr), xyz(c), result); }
    __device__ void shadow(const Ray& ray, float tMax, const V4& c, uint32_t)
    {
        raysThisVertex++; shadowRays++;
        const bool occluded = scene_traverse_shadow<false, false>(S, ray, tMax, nullptr, sstack);
        if (!occluded) { visible++; lightSum = lightSum + xyz(c); }
        else lightSum = lightSum + v3s(0.0f);           // k_shadow zeroes the contribution, k_resolve still adds it
    }
    __device__ void end_lights(const V4& thr)
    {
        if (raysThisVertex) result = mul_add(xyz(thr), lightSum, result);       // k_resolve runs only for paths that own shadow rays
        lightSum = v3s(0.0f); raysThisVertex = 0;
    }
};

#ifndef RTB_TAIL_DISCARD
#define RTB_TAIL_DISCARD 0
#endif
#ifndef RTB_TAIL_PROBE
#define RTB_TAIL_PROBE 0
#endif
#ifndef RTB_TAIL_MINBLOCKS
#define RTB_TAIL_MINBLOCKS 4       // 2 compiles it free of spills (255 registers)
#endif
template <bool kTex>
__global__ void __launch_bounds__(128, RTB_TAIL_MINBLOCKS) k_tail(const __grid_constant__ SceneView S, const __grid_constant__ PassParams P, const __grid_constant__ Wavefront W, uint32_t depth)
{
    // tailTaken = 1 + the depth at which the tail took over.  CTAs of THIS launch may already see the value written by
    // its first CTA (a grid larger than what is resident starts late CTAs after early ones have finished): they must
    // carry on; launches for later depths see a different value and leave.
    const uint32_t taken = W.ctl->tailTaken;
    if (taken != 0u && taken != depth + 1u) return;
    const uint32_t count = W.ctl->pathCount[depth];
    if (count == 0u || count > W.tailThreshold) return;
    extern __shared__ uint32_t smemStack[];
    SharedStack sstack = (SharedStack)__cvta_generic_to_shared(smemStack) + threadIdx.x * 4u;
    asm volatile("" : "+r"(sstack));
    const PathBuf& in = W.buf[depth & 1];
    const uint32_t stride = gridDim.x * blockDim.x;
    const uint32_t lane = threadIdx.x & 31u;
    unsigned long long rays = 0, shadowRays = 0, visible = 0;
#if RTB_TAIL_PROBE      // (section 4c probe: order-independent checksums of what this kernel read and found, in the unused test counters)
    unsigned long long probePaths = 0, probeIn = 0, probeHits = 0, probeColor = 0;
#endif
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride)
    {
        PathState ps;
        load_path(in, i, W, P.width, P.numSeedDims, ps);
        const uint32_t colorIndex = ps.passSlot * W.numPixels + ps.pixel;
        const float4 pc = W.passColor[colorIndex];
#if RTB_TAIL_PROBE
        probePaths++;
        probeIn += (unsigned long long)__float_as_uint(ps.ray.origin.x) + __float_as_uint(ps.ray.dir.y) + __float_as_uint(ps.throughput.x) + ps.pixel + ps.sampler.salt + ps.sampler.dim;
        probeColor += (unsigned long long)__float_as_uint(pc.x) + __float_as_uint(pc.y) + __float_as_uint(pc.z);
#endif
        TailSink sink(S, sstack);
        sink.result = v3(pc.x, pc.y, pc.z);
        for (;;)
        {
            Hit hit;
            hit.objectId = RTB_INVALID_OBJECT; hit.subObjectId = 0; hit.distance = INFINITY; hit.u = 0.0f; hit.v = 0.0f;
            scene_traverse<false>(S, ps.ray, hit, nullptr, sstack);
#if RTB_TAIL_PROBE
            if (ps.depth == depth) probeHits += (unsigned long long)hit.objectId + hit.subObjectId + __float_as_uint(hit.distance);
#endif
            rays++;
            uint32_t numShadow = 0;
            bool alive;
            if (P.renderer == RTB_RENDERER_PATH_TRACER_MIS) alive = shade_vertex_mis<kTex>(S, P, ps, hit, sink, numShadow);
            else alive = shade_vertex_pt<kTex>(S, P, ps, hit, sink);
            if (!alive) break;
        }
#if !RTB_TAIL_DISCARD      // (section 4c probe: a k_tail that does all its work and writes nothing)
        W.passColor[colorIndex] = make_float4(sink.result.x, sink.result.y, sink.result.z, 0.0f);
#endif
        shadowRays += sink.shadowRays; visible += sink.visible;
    }
    for (int o = 16; o > 0; o >>= 1)
    {
        rays += __shfl_down_sync(0xFFFFFFFFu, rays, o);
        shadowRays += __shfl_down_sync(0xFFFFFFFFu, shadowRays, o);
        visible += __shfl_down_sync(0xFFFFFFFFu, visible, o);
    }
    if (lane == 0)
    {
        if (rays) atomicAdd(&W.tot->numRays, rays);
        if (shadowRays) atomicAdd(&W.tot->numShadowRays, shadowRays);
        if (visible) atomicAdd(&W.tot->numShadowRaysHit, visible);
    }
#if RTB_TAIL_PROBE
    atomicAdd(&W.tot->numBoxTests, probePaths); atomicAdd(&W.tot->numTriTests, probeIn);
    atomicAdd(&W.tot->numShadowBoxTests, probeHits); atomicAdd(&W.tot->numShadowTriTests, probeColor);
#endif
    if (blockIdx.x == 0 && threadIdx.x == 0) W.ctl->tailTaken = depth + 1u;
}


void rtb_launch_tail(bool textured, int blocks, size_t smemBytes, cudaStream_t stream, const SceneView& S, const PassParams& P, const Wavefront& W, uint32_t depth)
{
    if (textured) k_tail<true><<<blocks, 128, smemBytes, stream>>>(S, P, W, depth);
    else k_tail<false><<<blocks, 128, smemBytes, stream>>>(S, P, W, depth);
}

void rtb_configure_tail(int maxDynamicSmem, int carveoutPercent)
{
    cudaFuncSetAttribute(k_tail<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxDynamicSmem);
    cudaFuncSetAttribute(k_tail<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxDynamicSmem);
    if (carveoutPercent >= 0)
    {
        cudaFuncSetAttribute(k_tail<false>, cudaFuncAttributePreferredSharedMemoryCarveout, carveoutPercent);
        cudaFuncSetAttribute(k_tail<true>, cudaFuncAttributePreferredSharedMemoryCarveout, carveoutPercent);
    }
}
#endif  // RTB_TU_TAIL

#if !RTB_TU_TAIL
// PathTracerMIS::SampleLights: accumulatedColor += SampleLight(light) in light order (PathTracerMIS.cpp:142-148), then
// resultColor.MulAndAccumulate(throughput, accumulatedColor) (PathTracerMIS.cpp:320)
__global__ void __launch_bounds__(256) k_resolve(const __grid_constant__ Wavefront W, uint32_t depth)
{
    const uint32_t count = W.ctl->tailTaken ? 0u : W.ctl->shadowPathCount[depth];
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < count; j += stride)
    {
        const uint32_t i = W.shList[j];
        const uint32_t n = W.shCount[i];
        const uint32_t base = i * W.raysPerPath;
        V3 sum = v3s(0.0f);
        // SampleLight returns RayColor::Zero() for an occluded light (PathTracerMIS.cpp:92-96) and the zero is still added
        for (uint32_t r = 0; r < n; ++r) sum = sum + (W.shOcc[base + r] ? v3s(0.0f) : v3(W.sr1[base + r].w, W.sr2[base + r].w, W.sr3[base + r].w));
        const float4 t = W.shThr[i];
        const uint32_t pixel = __float_as_uint(t.w);
        const float4 pc = W.passColor[pixel];
        const V3 res = mul_add(v3(t.x, t.y, t.z), sum, v3(pc.x, pc.y, pc.z));
        W.passColor[pixel] = make_float4(res.x, res.y, res.z, 0.0f);
    }
}

// Film::AccumulateColor, Film.cpp:25-39: sum += colour; secondarySum += colour on even passes.  The passes of a batch
// are added in pass order, so the float sums are the ones the reference gets pass after pass.
// `snapshots` (optional): the film as of EVERY pass of the batch, snapshots[b * snapshotStride + 3 * slot], so that a
// caller who wants each pass' frame (progressive display, the end-to-end bench) does not have to render pass by pass.
__global__ void __launch_bounds__(256) k_film(const __grid_constant__ Wavefront W, float* __restrict__ sum, float* __restrict__ sum2,
                                              float* __restrict__ snapshots, size_t snapshotStride, uint32_t snapshotCount)
{
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < W.numPixels; i += stride)
    {
        float* s = sum + (size_t)i * 3;
        float* t = sum2 + (size_t)i * 3;
        float s0 = s[0], s1 = s[1], s2 = s[2], t0 = t[0], t1 = t[1], t2 = t[2];
        for (uint32_t b = 0; b < W.batch; ++b)
        {
            const float4 c = W.passColor[(size_t)b * W.numPixels + i];
            s0 = s0 + c.x; s1 = s1 + c.y; s2 = s2 + c.z;
            if (((W.firstPassIndex + b) & 1u) == 0u) { t0 = t0 + c.x; t1 = t1 + c.y; t2 = t2 + c.z; }
            if (b < snapshotCount)
            {
                float* q = snapshots + (size_t)b * snapshotStride + (size_t)i * 3;
                q[0] = s0; q[1] = s1; q[2] = s2;
            }
        }
        s[0] = s0; s[1] = s1; s[2] = s2;
        t[0] = t0; t[1] = t1; t[2] = t2;
    }
}

// raster (y * width + x) <- pixel slots; src may hold several shards back to back (all-gathered film)
__global__ void __launch_bounds__(256) k_unshard(const float* __restrict__ src, float* __restrict__ dstRaster, const uint32_t* __restrict__ pixelListAll, uint32_t total)
{
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride)
    {
        const uint32_t pix = pixelListAll[i];
        if (pix == 0xFFFFFFFFu) continue;     // padding slot
        dstRaster[(size_t)pix * 3 + 0] = src[(size_t)i * 3 + 0];
        dstRaster[(size_t)pix * 3 + 1] = src[(size_t)i * 3 + 1];
        dstRaster[(size_t)pix * 3 + 2] = src[(size_t)i * 3 + 2];
    }
}

// one shard's pixel slots -> their places in a raster frame the caller owns (device memory or mapped host memory: the 8 x 4
// pixel blocks of the slot order make every warp write four runs of 96 contiguous bytes)
__global__ void __launch_bounds__(256) k_scatter_shard(const float* __restrict__ src, float* dstRaster, const uint32_t* __restrict__ pixelList, uint32_t numPixels)
{
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < numPixels; i += stride)
    {
        const size_t o = (size_t)pixelList[i] * 3;
        dstRaster[o] = src[(size_t)i * 3]; dstRaster[o + 1] = src[(size_t)i * 3 + 1]; dstRaster[o + 2] = src[(size_t)i * 3 + 2];
    }
}

// The same for a destination in MAPPED HOST memory, where the shape of the writes decides the PCIe efficiency: 16-byte
// stores, consecutive lanes on consecutive addresses.  The shard's pixels are taken in raster order (rasterOrder[k] = slot
// of its k-th pixel, rasterPix[k] = that pixel's raster index); a run of the shard's pixels in one raster row is a tile
// row segment, a multiple of four pixels long when width and tile size are multiples of four, so the shard's floats in
// raster order split into float4 chunks none of which straddles two runs: chunk q starts in pixel k0 = 4q / 3 and lands
// at float rasterPix[k0] * 3 + (4q - 3 k0) of the frame.  A warp writes 512 contiguous bytes (less at the end of a run)
// where k_scatter_shard writes 96 four-byte words in three strided sweeps: the 8-GPU end-to-end figure of the bench
// was bound by exactly that (3.1 MB per rank and step at ~6 GB/s).
__global__ void __launch_bounds__(256) k_scatter_shard_rows(const float* __restrict__ src, float4* dstRaster, const uint32_t* __restrict__ rasterOrder,
                                                            const uint32_t* __restrict__ rasterPix, uint32_t numChunks)
{
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < numChunks; q += stride)
    {
        const uint32_t f0 = 4u * q, k0 = f0 / 3u, c0 = f0 - 3u * k0;
        float v[4];
#pragma unroll
        for (uint32_t j = 0; j < 4u; ++j)
        {
            const uint32_t f = c0 + j, k = k0 + f / 3u, ch = f % 3u;
            v[j] = src[(size_t)rasterOrder[k] * 3u + ch];
        }
        const size_t dstFloat = (size_t)rasterPix[k0] * 3u + c0;       // a multiple of 4 (see above)
        dstRaster[dstFloat >> 2] = make_float4(v[0], v[1], v[2], v[3]);
    }
}

// front buffer (Viewport::PostProcessTile): pixel slots -> tone-mapped 0x00RRGGBB in raster order; `blurred` (the
// RTB_BLOOM_LEVELS blurred raster images back to back) is only read when P.bloomFactor > 0; `frozenRaster` (optional):
// per raster pixel, the pass count at which adaptive rendering dropped it (0 = still sampled)
__global__ void __launch_bounds__(256) k_front(const float* __restrict__ src, uint32_t* __restrict__ dstRaster, const uint32_t* __restrict__ pixelListAll, uint32_t total,
                                               const float* __restrict__ blurred, size_t levelStride, const uint32_t* __restrict__ frozenRaster, const PostParams P)
{
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride)
    {
        const uint32_t pix = pixelListAll[i];
        if (pix == 0xFFFFFFFFu) continue;     // padding slot
        const V3 sum = v3(src[(size_t)i * 3 + 0], src[(size_t)i * 3 + 1], src[(size_t)i * 3 + 2]);
        V3 bloom = v3s(0.0f);
        if (P.bloomFactor > 0.0f) bloom = bloom_color(blurred, levelStride, pix);
        // a pixel adaptive rendering stopped sampling after n passes holds the sum of n passes (the reference keeps the
        // front-buffer value it had then: Viewport.cpp:474-485 post-processes the active tiles only)
        const uint32_t frozen = frozenRaster ? frozenRaster[pix] : 0u;
        dstRaster[pix] = postprocess_pixel(P, sum, bloom, pix, frozen ? fdiv(1.0f, (float)frozen) : P.pixelScaling);
    }
}

// Bitmap::GaussianBlur, one direction: thread = (line, colour channel), lines blurred element by element in the
// reference's order (rtb_postprocess.h).  Horizontal: line = row, elements 3 floats apart; vertical: line = column,
// elements one row apart (neighbouring threads then read neighbouring floats).
__global__ void __launch_bounds__(128) k_blur(const float* in, float* out, uint32_t width, uint32_t height, int vertical,
                                              float* scratchA, float* scratchB, const BlurSchedule schedule)
{
    const uint32_t numLines = vertical ? width : height;
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= numLines * 3u) return;
    const uint32_t line = t / 3u, channel = t % 3u;
    const size_t first = vertical ? (size_t)line * 3 + channel : (size_t)line * width * 3 + channel;
    const size_t elementStride = vertical ? (size_t)width * 3 : 3;
    blur_line(in + first, out + first, elementStride, scratchA + t, scratchB + t, (size_t)numLines * 3,
              schedule, vertical ? RTB_BLUR_KEPT_VERTICAL : RTB_BLUR_KEPT_HORIZONTAL, vertical ? height : width);
}

// Viewport::ComputeBlockError in three steps (rtb_postprocess.h): per-pixel estimates (parallel), row sums of every
// block (one thread per block row, left to right), block sums (one thread per block, top to bottom)
__global__ void __launch_bounds__(256) k_pixel_error(const float* __restrict__ sumRaster, const float* __restrict__ sum2Raster, uint32_t numPixels, float scale, float* __restrict__ out)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= numPixels) return;
    out[i] = pixel_error(v3(sumRaster[3 * (size_t)i], sumRaster[3 * (size_t)i + 1], sumRaster[3 * (size_t)i + 2]),
                         v3(sum2Raster[3 * (size_t)i], sum2Raster[3 * (size_t)i + 1], sum2Raster[3 * (size_t)i + 2]), scale);
}

// rowStart[b] = index of block b's first row in rowErrors (rowStart[numBlocks] = total)
__global__ void __launch_bounds__(128) k_row_error(const float* __restrict__ pixelErrors, uint32_t width, const RtbBlock* __restrict__ blocks, const uint32_t* __restrict__ rowStart,
                                                   uint32_t numBlocks, float* __restrict__ rowErrors)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= rowStart[numBlocks]) return;
    uint32_t lo = 0, hi = numBlocks;                 // last block whose first row is <= t
    while (hi - lo > 1u) { const uint32_t mid = (lo + hi) / 2u; if (rowStart[mid] <= t) lo = mid; else hi = mid; }
    const RtbBlock b = blocks[lo];
    const uint32_t y = b.minY + (t - rowStart[lo]);
    rowErrors[t] = row_error(pixelErrors + (size_t)y * width, b.minX, b.maxX);
}

__global__ void __launch_bounds__(128) k_block_error(const float* __restrict__ rowErrors, const RtbBlock* __restrict__ blocks, const uint32_t* __restrict__ rowStart,
                                                     uint32_t numBlocks, uint32_t totalArea, float* __restrict__ out)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= numBlocks) return;
    const RtbBlock b = blocks[i];
    out[i] = block_error(rowErrors + rowStart[i], b.maxY - b.minY, b.maxX - b.minX, totalArea);
}

// test hooks ------------------------------------------------------------------------------------------------------------

template <bool kCount>
__global__ void __launch_bounds__(128) k_trace_batch(const __grid_constant__ SceneView S, const RtbRay* __restrict__ rays, RtbHit* __restrict__ hits, unsigned long long n, uint32_t mode, Totals* tot)
{
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    TraceCounters cnt; cnt.boxTests = 0; cnt.triTests = 0;
#if RTB_SMEM_STACK > 0
    extern __shared__ uint32_t smemStack[];     // stackEntries x 128 words (dynamic, sized by the scene's BVH depth)
    SharedStack sstack = (SharedStack)__cvta_generic_to_shared(smemStack) + threadIdx.x * 4u;
    asm volatile("" : "+r"(sstack));             // keep it in a register (see stk_put)
#else
    const SharedStack sstack = 0;
#endif
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    {
        const float4* p = reinterpret_cast<const float4*>(rays + i);
        const float4 o = p[0], d = p[1], id = p[2], od = p[3];
        Ray ray;
        ray.origin = v3(o.x, o.y, o.z); ray.dir = v3(d.x, d.y, d.z); ray.invDir = v3(id.x, id.y, id.z); ray.originDivDir = v3(od.x, od.y, od.z);
        RtbHit out;
        if (mode == RTB_TRACE_SHADOW)
        {
            // hits[i].distance is the preset tMax (HitPoint::distance, PathTracerMIS.cpp:86)
            const float tMax = hits[i].distance;
            out.objectId = RTB_INVALID_OBJECT; out.subObjectId = 0; out.distance = tMax; out.u = 0.0f; out.v = 0.0f;
            out.occluded = scene_traverse_shadow<kCount>(S, ray, tMax, &cnt, sstack) ? 1u : 0u;
        }
        else
        {
            Hit hit;
            hit.objectId = RTB_INVALID_OBJECT; hit.subObjectId = 0; hit.distance = INFINITY; hit.u = 0.0f; hit.v = 0.0f;
            scene_traverse<kCount>(S, ray, hit, &cnt, sstack);
            out.objectId = hit.objectId; out.subObjectId = hit.subObjectId; out.distance = hit.distance; out.u = hit.u; out.v = hit.v; out.occluded = 0;
        }
        hits[i] = out;
    }
    if (kCount)
    {
        atomicAdd(mode == RTB_TRACE_SHADOW ? &tot->numShadowBoxTests : &tot->numBoxTests, (unsigned long long)cnt.boxTests);
        atomicAdd(mode == RTB_TRACE_SHADOW ? &tot->numShadowTriTests : &tot->numTriTests, (unsigned long long)cnt.triTests);
    }
}

__global__ void k_texture_batch(const __grid_constant__ SceneView S, uint32_t textureIndex, const float* __restrict__ uv, unsigned long long n, float4* __restrict__ out)
{
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const V4 c = texture_evaluate(S, textureIndex, uv[2 * i], uv[2 * i + 1]);
    out[i] = make_float4(c.x, c.y, c.z, c.w);
}

__global__ void k_sample_batch(const __grid_constant__ SceneView S, const __grid_constant__ PassParams P, const uint32_t* __restrict__ xy, unsigned long long n, uint32_t numDims, uint32_t* __restrict__ out)
{
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Sampler s;
    sampler_reset_pixel(s, xy[2 * i], xy[2 * i + 1]);
    for (uint32_t d = 0; d < numDims; ++d) out[i * numDims + d] = sampler_get_int(s, P, S.blueNoise);
}

__global__ void k_primary_rays(const __grid_constant__ SceneView S, const __grid_constant__ PassParams P, RtbRay* __restrict__ rays)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.width * P.height) return;
    PathState ps;
    init_path(P, S.blueNoise, i % P.width, i / P.width, 0, ps);
    RtbRay r;
    r.origin[0] = ps.ray.origin.x; r.origin[1] = ps.ray.origin.y; r.origin[2] = ps.ray.origin.z; r.origin[3] = 0.0f;
    r.dir[0] = ps.ray.dir.x; r.dir[1] = ps.ray.dir.y; r.dir[2] = ps.ray.dir.z; r.dir[3] = 0.0f;
    r.invDir[0] = ps.ray.invDir.x; r.invDir[1] = ps.ray.invDir.y; r.invDir[2] = ps.ray.invDir.z; r.invDir[3] = 0.0f;
    r.originDivDir[0] = ps.ray.originDivDir.x; r.originDivDir[1] = ps.ray.originDivDir.y; r.originDivDir[2] = ps.ray.originDivDir.z; r.originDivDir[3] = 0.0f;
    rays[i] = r;
}

// ---------------------------------------------------------------------------------------------------------------------
// host side: context + extern "C"
// ---------------------------------------------------------------------------------------------------------------------

static thread_local std::string gLastError;

static int fail(int code, const char* fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    gLastError = buf;
    return code;
}

#define RTB_CUDA(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(RTB_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); } while (0)

template <typename T> struct DevArray
{
    T* ptr = nullptr;
    size_t count = 0;
    cudaError_t alloc(size_t n)
    {
        release();
        count = n;
        if (n == 0) return cudaSuccess;
        return cudaMalloc(&ptr, n * sizeof(T));
    }
    cudaError_t upload(const std::vector<T>& v, cudaStream_t s)
    {
        cudaError_t e = alloc(v.size());
        if (e != cudaSuccess || v.empty()) return e;
        return cudaMemcpyAsync(ptr, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, s);
    }
    void release() { if (ptr) cudaFree(ptr); ptr = nullptr; count = 0; }
};

struct DeviceMesh
{
    DevArray<DNode> nodes;
    DevArray<DTri> tris;
    DevArray<DVertexIndices> indices;
    DevArray<DVertexShading> shading;
    DevArray<uint32_t> materialIds;
    DevArray<WNode> wideNodes;     // any-hit tree (rtb_wide.h), empty if the mesh keeps the binary walk
    DevArray<DTri> wideTris;
};

struct RtbContext
{
    int device = 0;
    uint32_t flags = 0;
    uint32_t numShards = 1, shardIndex = 0, tileSize = 32;
    cudaStream_t stream = nullptr;
    bool ownStream = false;
    int numSMs = 148;

    // scene
    bool hasScene = false;
    DevArray<DNode> tlas;
    DevArray<DObjectTrace> objTrace;
    DevArray<DObjectShade> objShade;
    DevArray<RtbShape> shapes;
    std::vector<DeviceMesh> meshes;
    DevArray<DMesh> meshViews;
    DevArray<RtbMaterial> materials;
    DevArray<DLight> lights;
    DevArray<uint32_t> globalLights;
    DevArray<uint16_t> blueNoise;
    DevArray<DTexture> textures;
    std::vector<DevArray<float>> texels;      // one per texture (empty for procedural ones)
    DevArray<RtbMaterialTextures> materialTextures;
    DevArray<float> bokehCdf;                 // texture-shaped bokeh (RtbSceneDesc::bokehCdf)
    uint32_t bokehWidth = 0, bokehHeight = 0;
    SceneView view;
    uint32_t numLights = 0;
    size_t stackBytes = 0;         // dynamic shared memory of the traversal kernels: the scene's worst-case stack x 128 threads x 4 B
    size_t widePanelBytes = 0;     // k_shadow_wide's column: top-level entries + the wide tree's groups + the queue of leaves to confirm
    uint32_t tailThreshold = 65536;   // k_tail finishes a wavefront once at most this many paths are alive (RTB_TAIL_PATHS, 0 = off)

    // film + wavefront
    uint32_t width = 0, height = 0;
    uint32_t numPixels = 0;        // pixels of this shard
    uint32_t shardSlots = 0;       // padded slot count (equal on every shard)
    uint32_t raysPerPath = 0;
    uint32_t maxBatch = 1;         // passes per wavefront (sized at resize so that a wavefront has a few million paths)
    uint32_t maxBatchFilm = 1;     // ... as the film size alone would have it (alloc_shadow lowers maxBatch for scenes with many lights)
    DevArray<uint32_t> pixelList;     // this shard
    DevArray<uint32_t> pixelListAll;  // all shards back to back, padded with 0xFFFFFFFF
    DevArray<uint32_t> rasterOrder, rasterPix;   // this shard's slots / raster indices in raster order (k_scatter_shard_rows); empty unless width and tile size are multiples of 4
    std::vector<uint32_t> hostPixelList;   // this shard's list, kept for rtb_set_active_tiles
    DevArray<uint32_t> activeList;    // adaptive rendering: the pixel slots that still get samples ...
    uint32_t numActive = 0;
    bool hasActiveList = false;       // ... when set (rtb_set_active_tiles); otherwise every pixel
    std::vector<uint32_t> hostFrozen; // per raster pixel: passesSinceReset when it left the active set (0 = active); the same on every shard
    DevArray<uint32_t> frozen;
    bool frozenDirty = false;
    DevArray<float> raster2, pixelError, rowError;   // rtb_block_errors scratch
    DevArray<uint32_t> blockRows;
    DevArray<float> sum, sum2;
    DevArray<float> raster;           // scratch for read-back
    DevArray<uint32_t> front;         // tone-mapped front buffer (rtb_read_front), allocated on first use
    DevArray<float> blurred, blurScratch;   // bloom: RTB_BLOOM_LEVELS raster images + two line-major scratch images
    // ring of raster snapshots for the asynchronous read-back: the render stream only waits for a copy when the ring
    // wraps, so a group of frames (rtb_film_snapshots) can be in flight on the copy stream while the next passes render
    DevArray<float> rasterAsync[RTB_SNAP_RING];
    cudaStream_t copyStream = nullptr;
    cudaEvent_t scatterReady = nullptr, scatterDone = nullptr;     // rtb_scatter_shard_async: source complete / last scatter complete
    bool scatterPending = false;                                   // a scatter on the copy stream may still read a film snapshot
    cudaEvent_t snapReady[RTB_SNAP_RING] = {}, snapCopied[RTB_SNAP_RING] = {};
    uint32_t snapIndex = 0;

    // Wavefront slots: consecutive wavefronts (one per rtb_render_pass call, or per batch of rtb_render_passes) go to
    // the slots round robin, each on its own stream with its own queues, so the sparse deep bounces of one wavefront
    // overlap the dense first bounces of the next (and a small shard's latency-bound launch chains run several at a
    // time).  Only k_film touches the film: the k_film launches are chained in pass order by events, so every pixel
    // still receives its passes in order and the sums are bit-identical to one-wavefront-at-a-time rendering.
    struct Slot
    {
        DevArray<float4> p[2][4];
        DevArray<uint4> p4[2];
        DevArray<uint4> hit0;
        DevArray<float> hit1;
        DevArray<float4> passColor, sr0, sr1, sr2, sr3, shThr;
        DevArray<uint32_t> shCount, shList, rayList, shFlag;
        DevArray<uint8_t> shOcc;
        DevArray<float4> mi0, mi1, mi2;
        DevArray<Control> ctl;
        DevArray<uint32_t> seed;       // maxBatch x numSeedDims seed integers, then maxBatch x 2 AA offsets
        std::vector<uint32_t> stagingSeeds;
        cudaStream_t stream = nullptr;
        cudaEvent_t filmDone = nullptr;   // recorded after this slot's last k_film
        Wavefront wf;
        // the launch chain of a wavefront (raygen + up to five kernels per depth) as a CUDA graph, re-used while its
        // parameters stay the same: one graph launch instead of ~47 dependent kernel launches
        cudaGraphExec_t graphExec = nullptr;
        std::vector<unsigned char> graphKey;
    };
    std::vector<Slot> slots;
    uint32_t nextSlot = 0;
    cudaEvent_t lastFilm = nullptr;       // filmDone of the most recent wavefront (null: none since the last join point)
    cudaEvent_t mainMark = nullptr;       // position of the main stream when a wavefront was enqueued
    bool slotsNeedMain = true;            // the main stream holds work the slots must wait for (reset / resize / upload)
    bool slotsOwnStreams = false;         // false: one slot on the main stream (default), true: RTB_SLOTS > 1
    bool serializeSlots = false, profileDepths = false;      // development aids, read from the environment at rtb_create
    bool useGraphs = true;                // RTB_GRAPH=0: plain launches
    bool exclusiveStream = false;         // the wavefront streams belong to this context alone (RtbCreateDesc::stream given, or RTB_SLOTS > 1)
    bool noWide = false;                  // RTB_NO_WIDE: any-hit rays walk the binary tree (A/B against rtb_wide.h)
    bool persistentTraverse = true;       // k_traverse_p (persistent warps with ray refill); RTB_TRAV_KERNEL=simple: k_traverse (one ray per thread, grid-stride)
    bool sortShade = true;                // RTB_SORT_SHADE=0: k_shade takes its tile of paths in queue order instead of binned by material
    int shadowKernel = 1;                 // RTB_SHADOW_KERNEL: 0 simple = k_shadow (binary tree), 1 wide = k_shadow_wide (the default), 2 split = k_shadow_top + k_shadow_mesh
    bool wideKernel = true;               // k_shadow_wide (8-wide tree, persistent warps, deferred leaf confirmations) for any-hit rays
                                          // (binary tree).  Same results; measured on the bench it does not beat the binary walk
    const WNode* hotNodes = nullptr;      // the scene's largest 8-wide tree ...
    uint32_t hotTotal = 0;                // ... and its node count
    uint32_t wideDepth = 0;               // deepest 8-wide tree of the scene (levels)
    int hotOverride = 0;                  // RTB_HOT_NODES: nodes of the tree top staged in shared memory (-1: as many as fit).  Measured on the bench
                                          // (k_shadow_wide per 8 passes): none 40.1 ms, 256 nodes 43.1, as many as fit (~1000) 43.2 - the copy costs L1
    DevArray<float> snapshots;            // rtb_film_snapshots: snapshotCapacity x shardSlots x 3 floats
    uint32_t snapshotCapacity = 0;
    DevArray<Totals> totals;
    DevArray<uint32_t> seed;              // single-pass helpers (rtb_sample_batch / rtb_primary_rays)

    RtbCounters counters;
    uint32_t passesSinceReset = 0;

    // multi-GPU frames (rtb_comm_init / rtb_finish_frame)
    ncclComm_t comm = nullptr;
    DevArray<float> gathered;             // numShards x shardSlots x 3: the all-gathered films

    // "VCM" renderer (rtb_vcm.cuh): buffers are allocated by the first VCM pass after a resize / parameter change
    struct Vcm
    {
        RtbVcmParams params = { 1u, 1u, 10u, 0.02f, 0.02f, 1.0f, 0x5EEDull };      // VertexConnectionAndMerging.cpp:47-62
        DevArray<float4> lightVertices, photons[2];
        DevArray<uint32_t> lightVertexCount, photonCount, cellStart, cellCursor, gridIndices, tileTotals;
        DevArray<float> splat, splat2;
        uint32_t hashSize = 0, photonCapacity = 0, vertexStride = 0;
        uint32_t current = 0;             // photons[current] holds the previous pass' photons
        uint32_t lastPassIndex = 0xFFFFFFFFu;   // passIndex of the pass that wrote them
        bool allocated = false;
        bool splatsPending = false;
    } vcm;

    // CTAs per SM of the grid-stride kernels.  8 are resident; launching 16x as many lets the hardware scheduler
    // even out the very uneven per-warp traversal times (measured: 8 -> 71.1, 32 -> 69.3, 128 -> 67.8, 1024 -> 70.0 ms / 8 passes)
    int simpleBlocksPerSM = 128;

    // profiling
    bool profiling = false;
    std::vector<cudaEvent_t> events;
    RtbPassTimings timings;
};

static void free_scene(RtbContext* c)
{
    c->tlas.release(); c->objTrace.release(); c->objShade.release(); c->shapes.release();
    for (DeviceMesh& m : c->meshes) { m.nodes.release(); m.tris.release(); m.indices.release(); m.shading.release(); m.materialIds.release(); m.wideNodes.release(); m.wideTris.release(); }
    c->meshes.clear();
    c->meshViews.release(); c->materials.release(); c->lights.release(); c->globalLights.release();
    c->textures.release(); c->materialTextures.release(); c->bokehCdf.release(); c->bokehWidth = c->bokehHeight = 0;
    for (DevArray<float>& t : c->texels) t.release();
    c->texels.clear();
    c->hasScene = false;
}

static void free_vcm(RtbContext* c)
{
    RtbContext::Vcm& v = c->vcm;
    v.lightVertices.release(); v.photons[0].release(); v.photons[1].release(); v.lightVertexCount.release(); v.photonCount.release();
    v.cellStart.release(); v.cellCursor.release(); v.gridIndices.release(); v.tileTotals.release(); v.splat.release(); v.splat2.release();
    v.allocated = false; v.splatsPending = false; v.lastPassIndex = 0xFFFFFFFFu;
}

static void free_film(RtbContext* c)
{
    free_vcm(c);
    c->pixelList.release(); c->pixelListAll.release(); c->rasterOrder.release(); c->rasterPix.release(); c->sum.release(); c->sum2.release(); c->raster.release(); c->activeList.release(); c->hasActiveList = false; c->frozen.release(); c->hostFrozen.clear(); c->raster2.release(); c->pixelError.release(); c->rowError.release(); c->blockRows.release(); c->front.release(); c->blurred.release(); c->blurScratch.release();
    for (int i = 0; i < RTB_SNAP_RING; ++i) c->rasterAsync[i].release();
    c->snapshots.release(); c->snapshotCapacity = 0;
    for (RtbContext::Slot& t : c->slots)
    {
        for (int b = 0; b < 2; ++b) { for (int k = 0; k < 4; ++k) t.p[b][k].release(); t.p4[b].release(); }
        t.hit0.release(); t.hit1.release(); t.passColor.release();
        t.sr0.release(); t.sr1.release(); t.sr2.release(); t.sr3.release(); t.shThr.release(); t.shCount.release(); t.shList.release(); t.rayList.release();
        t.shOcc.release(); t.shFlag.release(); t.mi0.release(); t.mi1.release(); t.mi2.release();
    }
}

static cudaError_t sync_all(RtbContext* c)
{
    cudaError_t e = cudaStreamSynchronize(c->stream);
    for (RtbContext::Slot& t : c->slots) if (e == cudaSuccess && t.stream) e = cudaStreamSynchronize(t.stream);
    if (e == cudaSuccess && c->copyStream) e = cudaStreamSynchronize(c->copyStream);
    c->lastFilm = nullptr;
    return e;
}

// the main stream waits for every wavefront in flight (the k_film launches are chained, so the last one is enough)
static cudaError_t join_slots(RtbContext* c)
{
    if (!c->lastFilm) return cudaSuccess;
    const cudaError_t e = cudaStreamWaitEvent(c->stream, c->lastFilm, 0);
    return e;
}

static int alloc_shadow(RtbContext* c)
{
    // shadow-ray storage depends on both the film size and the number of lights sampled per vertex
    const uint32_t rpp = c->numLights > 0 ? c->numLights : 1;
    if (c->numPixels == 0) return RTB_OK;
    // passes per wavefront: what the film size allows (rtb_resize), cut down when the scene's light count would make the
    // shadow-ray queues (68 B per path and light) outgrow half of the free device memory
    c->maxBatch = c->maxBatchFilm;
    {
        size_t freeB = 0, totalB = 0;
        if (cudaMemGetInfo(&freeB, &totalB) == cudaSuccess)
        {
            size_t held = 0;
            for (RtbContext::Slot& t : c->slots) held += (t.sr0.count * 4 * sizeof(float4)) + t.rayList.count * sizeof(uint32_t);
            const size_t budget = (freeB + held) / 2 / c->slots.size();
            const size_t perPass = (size_t)c->numPixels * rpp * (4 * sizeof(float4) + sizeof(uint32_t) + 1);
            const size_t cap = perPass ? budget / perPass : c->maxBatch;
            if (cap < c->maxBatch) c->maxBatch = cap < 1 ? 1u : (uint32_t)cap;
        }
    }
    const size_t n = (size_t)c->numPixels * c->maxBatch * rpp;
    if (rpp == c->raysPerPath && !c->slots.empty() && c->slots[0].sr0.count == n) return RTB_OK;
    c->raysPerPath = rpp;
    for (RtbContext::Slot& t : c->slots)
    {
        RTB_CUDA(t.sr0.alloc(n)); RTB_CUDA(t.sr1.alloc(n)); RTB_CUDA(t.sr2.alloc(n)); RTB_CUDA(t.sr3.alloc(n));
        RTB_CUDA(t.rayList.alloc(n)); RTB_CUDA(t.shOcc.alloc(n));
        if (c->shadowKernel == 2) { RTB_CUDA(t.shFlag.alloc(n)); RTB_CUDA(t.mi0.alloc(n)); RTB_CUDA(t.mi1.alloc(n)); RTB_CUDA(t.mi2.alloc(n)); }
    }
    return RTB_OK;
}

static void fill_wavefront(RtbContext* c)
{
    for (RtbContext::Slot& t : c->slots)
    {
        Wavefront& w = t.wf;
        for (int b = 0; b < 2; ++b)
        {
            w.buf[b].p0 = t.p[b][0].ptr; w.buf[b].p1 = t.p[b][1].ptr; w.buf[b].p2 = t.p[b][2].ptr; w.buf[b].p3 = t.p[b][3].ptr;
            w.buf[b].p4 = t.p4[b].ptr;
        }
        w.hit0 = t.hit0.ptr; w.hit1 = t.hit1.ptr; w.passColor = t.passColor.ptr;
        w.sr0 = t.sr0.ptr; w.sr1 = t.sr1.ptr; w.sr2 = t.sr2.ptr; w.sr3 = t.sr3.ptr;
        w.shThr = t.shThr.ptr; w.shCount = t.shCount.ptr; w.shList = t.shList.ptr; w.rayList = t.rayList.ptr;
        w.mi0 = t.mi0.ptr; w.mi1 = t.mi1.ptr; w.mi2 = t.mi2.ptr; w.shFlag = t.shFlag.ptr; w.shOcc = t.shOcc.ptr; w.miCapacity = (uint32_t)std::min<size_t>(t.mi0.count, 0xFFFFFFFFu);
        w.ctl = t.ctl.ptr;
        w.tot = c->totals.ptr;
        w.pixelList = c->pixelList.ptr;
        w.numPixels = c->numPixels;
        w.raysPerPath = c->raysPerPath;
        w.activeList = nullptr; w.numActive = RTB_ALL_ACTIVE;
    }
}

extern "C" {

const char* rtb_last_error(void) { return gLastError.c_str(); }
int rtb_api_version(void) { return RTB_API_VERSION; }

int rtb_create(const RtbCreateDesc* desc, RtbContext** outCtx)
{
    if (!desc || !outCtx) return fail(RTB_ERR_INVALID_ARG, "rtb_create: null argument");
    *outCtx = nullptr;
    int deviceCount = 0;
    cudaError_t e = cudaGetDeviceCount(&deviceCount);
    if (e != cudaSuccess || deviceCount == 0)
        return fail(RTB_ERR_NO_DEVICE, "rtb_create: no CUDA device (%s) - this library has no CPU fallback", e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    if (desc->deviceId < 0 || desc->deviceId >= deviceCount) return fail(RTB_ERR_INVALID_ARG, "rtb_create: device %d out of range (%d devices)", desc->deviceId, deviceCount);
    const uint32_t numShards = desc->numShards ? desc->numShards : 1;
    if (desc->shardIndex >= numShards) return fail(RTB_ERR_INVALID_ARG, "rtb_create: shardIndex %u >= numShards %u", desc->shardIndex, numShards);
    RTB_CUDA(cudaSetDevice(desc->deviceId));
    RtbContext* c = new (std::nothrow) RtbContext();
    if (!c) return fail(RTB_ERR_OUT_OF_MEMORY, "rtb_create: out of host memory");
    c->device = desc->deviceId;
    c->flags = desc->flags;
    c->numShards = numShards;
    c->shardIndex = desc->shardIndex;
    c->tileSize = desc->tileSize ? desc->tileSize : 32;
    if (desc->stream) { c->stream = (cudaStream_t)desc->stream; c->ownStream = false; c->exclusiveStream = true; }
    else
    {
        // contexts that do not bring their own stream share ONE stream per device: their kernels then never overlap each
        // other.  Overlapping kernels are not safe for this code as long as it has register spills;
        // a caller who passes streams of its own takes over that responsibility.
        static cudaStream_t sharedStreams[64] = {};
        static std::mutex sharedStreamsLock;
        std::lock_guard<std::mutex> guard(sharedStreamsLock);
        cudaStream_t& shared = sharedStreams[desc->deviceId & 63];
        if (!shared)
        {
            e = cudaStreamCreateWithFlags(&shared, cudaStreamNonBlocking);
            if (e != cudaSuccess) { shared = nullptr; delete c; return fail(RTB_ERR_CUDA, "cudaStreamCreate failed: %s", cudaGetErrorString(e)); }
        }
        c->stream = shared;
        c->ownStream = false;
    }
    cudaDeviceGetAttribute(&c->numSMs, cudaDevAttrMultiProcessorCount, c->device);
    if (const char* t = getenv("RTB_SIMPLE_BLOCKS")) c->simpleBlocksPerSM = atoi(t) > 0 ? atoi(t) : 128;
    // the traversal stacks live in dynamic shared memory (rtb_trace.h): allow up to the 227 KB a CTA may have
    {
        const int maxStack = (int)((size_t)RTB_MAX_SMEM_STACK_ENTRIES * RTB_STACK_THREADS * sizeof(uint32_t));
        cudaFuncSetAttribute(k_traverse<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxStack);
        cudaFuncSetAttribute(k_traverse<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxStack);
        cudaFuncSetAttribute(k_traverse_p<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxStack);
        cudaFuncSetAttribute(k_traverse_p<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxStack);
        cudaFuncSetAttribute(k_shadow<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxStack);
        cudaFuncSetAttribute(k_shadow<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxStack);
        cudaFuncSetAttribute(k_trace_batch<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxStack);
        cudaFuncSetAttribute(k_trace_batch<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxStack);
        rtb_configure_tail(maxStack, -1);
        cudaFuncSetAttribute(k_shadow_wide<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
        cudaFuncSetAttribute(k_shadow_wide<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
        cudaFuncSetAttribute(k_shadow_wide<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
        cudaFuncSetAttribute(k_shadow_wide<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
        cudaFuncSetAttribute(k_shadow_top<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxStack);
        cudaFuncSetAttribute(k_shadow_top<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxStack);
        if (const char* t = getenv("RTB_TAIL_PATHS")) c->tailThreshold = (uint32_t)strtoul(t, nullptr, 10);
    }
    // RTB_CARVEOUT=n (development aid): pin the L1 / shared-memory split of EVERY kernel of this library to n percent
    // shared memory, so that kernels of different streams never ask the SM for different splits
    if (const char* t = getenv("RTB_CARVEOUT"))
    {
        const int pct = atoi(t);
        const void* all[] = { (const void*)k_raygen, (const void*)k_traverse<false>, (const void*)k_traverse<true>, (const void*)k_traverse_p<false>, (const void*)k_traverse_p<true>,
                              (const void*)k_shade<false, false, RTB_RENDERER_PATH_TRACER_MIS>, (const void*)k_shade<true, false, RTB_RENDERER_PATH_TRACER_MIS>, (const void*)k_shade<false, true, RTB_RENDERER_PATH_TRACER_MIS>, (const void*)k_shade<true, true, RTB_RENDERER_PATH_TRACER_MIS>,
                              (const void*)k_shade<false, false, RTB_RENDERER_PATH_TRACER>, (const void*)k_shade<true, false, RTB_RENDERER_PATH_TRACER>, (const void*)k_shade<false, true, RTB_RENDERER_PATH_TRACER>, (const void*)k_shade<true, true, RTB_RENDERER_PATH_TRACER>,
                              (const void*)k_shade<false, false, RTB_RENDERER_DEBUG>, (const void*)k_shade<true, false, RTB_RENDERER_DEBUG>, (const void*)k_shade<false, true, RTB_RENDERER_DEBUG>, (const void*)k_shade<true, true, RTB_RENDERER_DEBUG>,
                              (const void*)k_shadow<false>, (const void*)k_shadow<true>, (const void*)k_shadow_wide<false, false>, (const void*)k_shadow_wide<true, false>, (const void*)k_shadow_wide<false, true>, (const void*)k_shadow_wide<true, true>,
                              (const void*)k_shadow_top<false>, (const void*)k_shadow_top<true>, (const void*)k_shadow_mesh<false>, (const void*)k_shadow_mesh<true>,
                              (const void*)k_resolve,
                              (const void*)k_film, (const void*)k_unshard, (const void*)k_front };
        for (const void* f : all) cudaFuncSetAttribute(f, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
        rtb_configure_tail((int)((size_t)RTB_MAX_SMEM_STACK_ENTRIES * RTB_STACK_THREADS * sizeof(uint32_t)), pct);
    }
    // wavefront slots (see RtbContext::Slot).  The default is ONE slot that runs on the main stream, i.e. strictly one
    // kernel at a time.  Several slots (RTB_SLOTS=2..4: +2 % device-resident on the bench) overlap the kernels of
    // consecutive wavefronts; round 1 saw different hits for the same ray in that mode, round 2 bisected it to k_tail
    // and to the optimisation level ptxas builds it at.
    int numSlots = 1;
    if (const char* t = getenv("RTB_SLOTS")) numSlots = atoi(t) >= 1 && atoi(t) <= 16 ? atoi(t) : numSlots;
    c->slotsOwnStreams = numSlots > 1;
    if (c->slotsOwnStreams) c->exclusiveStream = true;
    // (Round 2 bisected the overlap problem to k_tail and then to ptxas' -O2 / -O3 code for it: the kernel is built at ptxas -O1
    //  in a translation unit of its own, see "compiled TWICE" above and DESIGN.md section 4c, and contexts with several slots
    //  keep it.  RTB_NO_TAIL_WITH_SLOTS=1 restores the earlier workaround, slots without the tail kernel.)
    if (c->slotsOwnStreams && getenv("RTB_NO_TAIL_WITH_SLOTS")) c->tailThreshold = 0;
    c->serializeSlots = getenv("RTB_SERIALIZE_SLOTS") != nullptr;
    c->profileDepths = getenv("RTB_PROFILE_DEPTHS") != nullptr;
    c->noWide = getenv("RTB_NO_WIDE") != nullptr;
    if (const char* t = getenv("RTB_GRAPH")) c->useGraphs = atoi(t) != 0;
    // an instrumented context counts the REFERENCE's work per ray (box / triangle tests of its binary tree: the algorithmic
    // bytes of SURVEY 8(d)); RTB_COUNT_WIDE makes it count what the 8-wide walk does instead (8 "box tests" per node visited)
    if ((c->flags & RTB_FLAG_COUNT_TESTS) != 0 && !getenv("RTB_COUNT_WIDE")) c->noWide = true;
    if (const char* t = getenv("RTB_SORT_SHADE")) c->sortShade = atoi(t) != 0;
    if (const char* t = getenv("RTB_TRAV_KERNEL")) c->persistentTraverse = strcmp(t, "simple") != 0;
    if (const char* t = getenv("RTB_SHADOW_KERNEL")) { c->shadowKernel = !strcmp(t, "simple") ? 0 : !strcmp(t, "split") ? 2 : 1; c->wideKernel = c->shadowKernel != 0; }
    if (const char* t = getenv("RTB_HOT_NODES")) c->hotOverride = atoi(t);
    c->slots.resize(numSlots);
    e = c->totals.alloc(1);
    if (e == cudaSuccess) e = cudaMemsetAsync(c->totals.ptr, 0, sizeof(Totals), c->stream);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->mainMark, cudaEventDisableTiming);
    for (RtbContext::Slot& t : c->slots)
    {
        memset(&t.wf, 0, sizeof(t.wf));
        if (e == cudaSuccess) e = t.ctl.alloc(1);
        if (!c->slotsOwnStreams) t.stream = c->stream;
        else if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&t.stream, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&t.filmDone, cudaEventDisableTiming);
    }
    if (e != cudaSuccess) { rtb_destroy(c); return fail(RTB_ERR_CUDA, "control block allocation failed: %s", cudaGetErrorString(e)); }
    memset(&c->view, 0, sizeof(c->view));
    memset(&c->counters, 0, sizeof(c->counters));
    memset(&c->timings, 0, sizeof(c->timings));
    *outCtx = c;
    return RTB_OK;
}

static void release_comm(RtbContext* c);

void rtb_destroy(RtbContext* c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    for (RtbContext::Slot& t : c->slots) if (t.stream) cudaStreamSynchronize(t.stream);
    free_scene(c);
    free_film(c);
    c->blueNoise.release(); c->totals.release(); c->seed.release(); c->gathered.release();
    release_comm(c);
    for (RtbContext::Slot& t : c->slots)
    {
        t.ctl.release(); t.seed.release();
        if (t.graphExec) { cudaGraphExecDestroy(t.graphExec); t.graphExec = nullptr; }
        if (t.filmDone) cudaEventDestroy(t.filmDone);
        if (t.stream && c->slotsOwnStreams) cudaStreamDestroy(t.stream);
    }
    if (c->mainMark) cudaEventDestroy(c->mainMark);
    for (cudaEvent_t ev : c->events) cudaEventDestroy(ev);
    if (c->copyStream)
    {
        cudaStreamSynchronize(c->copyStream);
        for (int i = 0; i < RTB_SNAP_RING; ++i) { cudaEventDestroy(c->snapReady[i]); cudaEventDestroy(c->snapCopied[i]); }
        cudaEventDestroy(c->scatterReady); cudaEventDestroy(c->scatterDone);
        cudaStreamDestroy(c->copyStream);
    }
    if (c->ownStream) cudaStreamDestroy(c->stream);
    delete c;
}

int rtb_set_blue_noise(RtbContext* c, const uint16_t* data)
{
    if (!c) return fail(RTB_ERR_INVALID_ARG, "null context");
    RTB_CUDA(cudaSetDevice(c->device));
    if (!data) { c->blueNoise.release(); c->view.blueNoise = nullptr; return RTB_OK; }
    RTB_CUDA(c->blueNoise.alloc(128 * 128 * 4));
    RTB_CUDA(cudaMemcpyAsync(c->blueNoise.ptr, data, 128 * 128 * 4 * sizeof(uint16_t), cudaMemcpyHostToDevice, c->stream));
    RTB_CUDA(cudaStreamSynchronize(c->stream));
    c->view.blueNoise = c->blueNoise.ptr;
    return RTB_OK;
}

int rtb_upload_scene(RtbContext* c, const RtbSceneDesc* scene)
{
    if (!c || !scene) return fail(RTB_ERR_INVALID_ARG, "rtb_upload_scene: null argument");
    RTB_CUDA(cudaSetDevice(c->device));
    HostScene hs;
    std::string err;
    if (!build_host_scene(*scene, hs, err, RTB_MAX_SMEM_STACK_ENTRIES)) return fail(RTB_ERR_INVALID_ARG, "rtb_upload_scene: %s", err.c_str());
    if (!hs.materialTextures.empty() && hs.materials.size() > 4094u)     // 12 bits of the packed path state (PathState::lastMaterial)
        return fail(RTB_ERR_UNSUPPORTED, "rtb_upload_scene: %zu materials; textured scenes are limited to 4094", hs.materials.size());
    RTB_CUDA(sync_all(c));
    free_scene(c);
    c->slotsNeedMain = true;
    cudaStream_t s = c->stream;
    RTB_CUDA(c->tlas.upload(hs.tlas, s));
    RTB_CUDA(c->objTrace.upload(hs.objTrace, s));
    RTB_CUDA(c->objShade.upload(hs.objShade, s));
    RTB_CUDA(c->shapes.upload(hs.shapes, s));
    RTB_CUDA(c->materials.upload(hs.materials, s));
    RTB_CUDA(c->lights.upload(hs.lights, s));
    RTB_CUDA(c->globalLights.upload(hs.globalLights, s));
    c->meshes.resize(hs.meshes.size());
    std::vector<DMesh> views(hs.meshes.size());
    for (size_t i = 0; i < hs.meshes.size(); ++i)
    {
        DeviceMesh& dm = c->meshes[i];
        const HostMesh& hm = hs.meshes[i];
        RTB_CUDA(dm.nodes.upload(hm.nodes, s));
        RTB_CUDA(dm.tris.upload(hm.tris, s));
        RTB_CUDA(dm.indices.upload(hm.indices, s));
        RTB_CUDA(dm.shading.upload(hm.shading, s));
        RTB_CUDA(dm.materialIds.upload(hm.materialIds, s));
        const bool wide = !hm.wide.nodes.empty() && !c->noWide;
        if (wide) { RTB_CUDA(dm.wideNodes.upload(hm.wide.nodes, s)); RTB_CUDA(dm.wideTris.upload(hm.wide.tris, s)); }
        DMesh& v = views[i];
        v.wideNodes = wide ? dm.wideNodes.ptr : nullptr; v.wideTris = wide ? dm.wideTris.ptr : nullptr;
        v.wideMaxAbs = hm.wide.maxAbs; v.numWideNodes = wide ? (uint32_t)hm.wide.nodes.size() : 0u;
        v.nodes = dm.nodes.ptr; v.tris = dm.tris.ptr; v.indices = dm.indices.ptr; v.shading = dm.shading.ptr; v.materialIds = dm.materialIds.ptr;
        v.numNodes = (uint32_t)hm.nodes.size(); v.numTriangles = (uint32_t)hm.tris.size();
        v.numVertices = (uint32_t)hm.shading.size(); v.numMaterials = (uint32_t)hm.materialIds.size();
    }
    RTB_CUDA(c->meshViews.upload(views, s));
    c->hotNodes = nullptr; c->hotTotal = 0; c->wideDepth = 1;
    for (const HostMesh& hm : hs.meshes) c->wideDepth = std::max(c->wideDepth, hm.wide.depth);
    for (size_t i = 0; i < views.size(); ++i) if (views[i].numWideNodes > c->hotTotal) { c->hotTotal = views[i].numWideNodes; c->hotNodes = views[i].wideNodes; }
    // textures: bitmap texels stay float4 in HBM (rtb_texture.h); the descriptors get the device pointers
    c->texels.resize(hs.textures.size());
    for (size_t i = 0; i < hs.textures.size(); ++i)
    {
        if (!hs.texelSources[i]) continue;
        const size_t n = (size_t)hs.textures[i].width * hs.textures[i].height * 4;
        RTB_CUDA(c->texels[i].alloc(n));
        RTB_CUDA(cudaMemcpyAsync(c->texels[i].ptr, hs.texelSources[i], n * sizeof(float), cudaMemcpyHostToDevice, s));
        hs.textures[i].texels = c->texels[i].ptr;
    }
    RTB_CUDA(c->textures.upload(hs.textures, s));
    RTB_CUDA(c->materialTextures.upload(hs.materialTextures, s));
    RTB_CUDA(c->bokehCdf.upload(hs.bokehCdf, s));
    c->bokehWidth = hs.bokehWidth; c->bokehHeight = hs.bokehHeight;
    RTB_CUDA(cudaStreamSynchronize(s));   // host vectors go out of scope

    SceneView& v = c->view;
    v.tlas = c->tlas.ptr; v.objTrace = c->objTrace.ptr; v.objShade = c->objShade.ptr; v.shapes = c->shapes.ptr;
    v.meshes = c->meshViews.ptr; v.materials = c->materials.ptr; v.lights = c->lights.ptr; v.globalLights = c->globalLights.ptr;
    v.blueNoise = c->blueNoise.ptr;
    v.textures = c->textures.ptr; v.materialTextures = c->materialTextures.ptr;
    v.numTlasNodes = (uint32_t)hs.tlas.size(); v.numObjects = (uint32_t)hs.objTrace.size();
    v.numLights = (uint32_t)hs.lights.size(); v.numGlobalLights = (uint32_t)hs.globalLights.size();
    c->numLights = v.numLights;
    c->stackBytes = RTB_SMEM_STACK ? (size_t)hs.stackEntries * RTB_STACK_THREADS * sizeof(uint32_t) : 0;
    {
        const uint32_t topEntries = hs.maxTlasDepth + hs.maxTlasLeaf + 2u;
        const uint32_t wideEntries = std::max<uint32_t>(hs.stackEntries, topEntries + 2u * c->wideDepth + (RTB_SW_DEFER ? RTB_WIDE_LEAF_QUEUE : 0u));
        c->widePanelBytes = (size_t)wideEntries * RTB_STACK_THREADS * sizeof(uint32_t);
    }
    if (c->profileDepths) fprintf(stderr, "scene: top-level BVH depth %u (largest leaf %u), deepest mesh BVH %u -> %u stack entries, %zu B shared memory per CTA\n",
                                              hs.maxTlasDepth, hs.maxTlasLeaf, hs.maxBlasDepth, hs.stackEntries, c->stackBytes);
    c->hasScene = true;
    int rc = alloc_shadow(c);
    if (rc != RTB_OK) return rc;
    fill_wavefront(c);
    return RTB_OK;
}

int rtb_resize(RtbContext* c, uint32_t width, uint32_t height)
{
    if (!c) return fail(RTB_ERR_INVALID_ARG, "null context");
    if (width == 0 || height == 0 || width > 65536 || height > 65536) return fail(RTB_ERR_INVALID_ARG, "rtb_resize: invalid viewport size %ux%u", width, height);   // Viewport.cpp:52-58
    RTB_CUDA(cudaSetDevice(c->device));
    RTB_CUDA(sync_all(c));
    free_film(c);
    c->slotsNeedMain = true;
    c->width = width; c->height = height;

    // tile-interleaved sharding: tile t (row-major over the tile grid) belongs to shard t % numShards.  Every shard is padded to the same slot count so that one all-gather reassembles the film.
    const uint32_t ts = c->tileSize;
    const uint32_t tilesX = (width + ts - 1) / ts, tilesY = (height + ts - 1) / ts;
    std::vector<std::vector<uint32_t>> lists(c->numShards);
    for (uint32_t ty = 0; ty < tilesY; ++ty) for (uint32_t tx = 0; tx < tilesX; ++tx)
    {
        std::vector<uint32_t>& l = lists[(ty * tilesX + tx) % c->numShards];
        // inside a tile: 8x4 pixel blocks, row-major over the blocks and inside a block, so that a warp owns a compact
        // patch of the image (its 32 primary rays and their first hits stay close together)
        for (uint32_t by = ty * ts; by < (ty + 1) * ts && by < height; by += 4)
            for (uint32_t bx = tx * ts; bx < (tx + 1) * ts && bx < width; bx += 8)
                for (uint32_t y = by; y < by + 4 && y < (ty + 1) * ts && y < height; ++y)
                    for (uint32_t x = bx; x < bx + 8 && x < (tx + 1) * ts && x < width; ++x) l.push_back(y * width + x);
    }
    size_t maxSlots = 0;
    for (const auto& l : lists) if (l.size() > maxSlots) maxSlots = l.size();
    // the packed path state holds the pixel slot in 26 bits (store_path / load_path)
    if (maxSlots >= ((size_t)1 << 26)) { c->width = c->height = 0; c->numPixels = 0; }
    if (maxSlots >= ((size_t)1 << 26)) return fail(RTB_ERR_UNSUPPORTED, "rtb_resize: %zu pixels in one shard; the limit is 2^26 - 1 (use more shards)", maxSlots);
    c->shardSlots = (uint32_t)maxSlots;
    c->numPixels = (uint32_t)lists[c->shardIndex].size();
    std::vector<uint32_t> all((size_t)c->numShards * maxSlots, 0xFFFFFFFFu);
    for (uint32_t s = 0; s < c->numShards; ++s) memcpy(&all[(size_t)s * maxSlots], lists[s].data(), lists[s].size() * 4);
    c->hostPixelList = lists[c->shardIndex];
    RTB_CUDA(c->pixelList.upload(lists[c->shardIndex], c->stream));
    RTB_CUDA(c->pixelListAll.upload(all, c->stream));
    c->rasterOrder.release(); c->rasterPix.release();
    if (width % 4u == 0 && ts % 4u == 0 && c->numPixels != 0)
    {
        const std::vector<uint32_t>& mine = lists[c->shardIndex];
        std::vector<uint32_t> order(mine.size()), pix(mine.size());
        for (uint32_t i = 0; i < (uint32_t)order.size(); ++i) order[i] = i;
        std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return mine[a] < mine[b]; });
        for (size_t k = 0; k < order.size(); ++k) pix[k] = mine[order[k]];
        RTB_CUDA(c->rasterOrder.upload(order, c->stream));
        RTB_CUDA(c->rasterPix.upload(pix, c->stream));
    }
    RTB_CUDA(cudaStreamSynchronize(c->stream));

    // passes per wavefront: enough paths in flight to keep 148 SMs busy through the deep (sparse) bounces, bounded by memory
    {
        uint32_t target = 16u << 20;   // measured at 1080p: 4M paths 67.0, 8M 64.3, 16M 62.7, 32M 62.6 ms per 8 passes
        if (const char* t = getenv("RTB_BATCH_PATHS")) target = (uint32_t)strtoul(t, nullptr, 10);
        uint32_t b = c->numPixels ? (target + c->numPixels - 1) / c->numPixels : 1u;
        if (b < 1u) b = 1u;
        if (b > 32u) b = 32u;          // (a 1/8 shard of 1080p: 20 passes share one 5 M-path wavefront instead of two 2.6 M-path ones)
        if (const char* t = getenv("RTB_MAX_BATCH")) b = (uint32_t)atoi(t) > 0 ? (uint32_t)atoi(t) : b;
        if (b > 63u) b = 63u;
        c->maxBatch = b; c->maxBatchFilm = b;
    }
    const size_t n = (size_t)c->numPixels * c->maxBatch;
    RTB_CUDA(c->sum.alloc(maxSlots * 3)); RTB_CUDA(c->sum2.alloc(maxSlots * 3));
    RTB_CUDA(c->raster.alloc((size_t)width * height * 3));
    for (RtbContext::Slot& t : c->slots)
    {
        for (int b = 0; b < 2; ++b) { for (int k = 0; k < 4; ++k) RTB_CUDA(t.p[b][k].alloc(n)); RTB_CUDA(t.p4[b].alloc(n)); }
        RTB_CUDA(t.hit0.alloc(n)); RTB_CUDA(t.hit1.alloc(n)); RTB_CUDA(t.passColor.alloc(n));
        RTB_CUDA(t.shThr.alloc(n)); RTB_CUDA(t.shCount.alloc(n)); RTB_CUDA(t.shList.alloc(n));
    }
    c->raysPerPath = 0;
    int rc = alloc_shadow(c);
    if (rc != RTB_OK) return rc;
    fill_wavefront(c);
    return rtb_reset(c);
}

int rtb_reset(RtbContext* c)
{
    if (!c) return fail(RTB_ERR_INVALID_ARG, "null context");
    RTB_CUDA(cudaSetDevice(c->device));
    RTB_CUDA(join_slots(c));       // wavefronts still in flight belong to the frame being discarded: let them finish first
    c->lastFilm = nullptr;
    if (c->sum.ptr)
    {
        RTB_CUDA(cudaMemsetAsync(c->sum.ptr, 0, c->sum.count * sizeof(float), c->stream));
        RTB_CUDA(cudaMemsetAsync(c->sum2.ptr, 0, c->sum2.count * sizeof(float), c->stream));
    }
    RTB_CUDA(cudaMemsetAsync(c->totals.ptr, 0, sizeof(Totals), c->stream));
    c->slotsNeedMain = true;
    memset(&c->counters, 0, sizeof(c->counters));
    c->passesSinceReset = 0;
    c->hasActiveList = false; c->numActive = 0; c->hostFrozen.clear();      // Viewport::Reset -> BuildInitialBlocksList: everything active again
    c->vcm.lastPassIndex = 0xFFFFFFFFu;
    if (c->vcm.allocated)
    {
        RTB_CUDA(cudaMemsetAsync(c->vcm.splat.ptr, 0, c->vcm.splat.count * sizeof(float), c->stream));
        RTB_CUDA(cudaMemsetAsync(c->vcm.splat2.ptr, 0, c->vcm.splat2.count * sizeof(float), c->stream));
        c->vcm.splatsPending = false;
    }
    return RTB_OK;
}

static int fill_pass_params(RtbContext* c, const RtbPassDesc* pass, PassParams& P, bool uploadSeed)
{
    if (pass->renderer > RTB_RENDERER_LIGHT_TRACER) return fail(RTB_ERR_INVALID_ARG, "unknown renderer id %u", pass->renderer);   // CreateRenderer, Renderer.cpp:45-69
    if (pass->maxRayDepth >= 255u || pass->maxRayDepth >= RTB_MAX_DEPTH_SLOTS) return fail(RTB_ERR_INVALID_ARG, "maxRayDepth %u too large (limit %u)", pass->maxRayDepth, RTB_MAX_DEPTH_SLOTS - 1);
    if (pass->camera.barrelDistortionVariableFactor != 0.0f) return fail(RTB_ERR_UNSUPPORTED, "variable barrel distortion draws from a per-thread RNG in the reference and is not supported");
    if (pass->numSeedDims > 0 && !pass->seed) return fail(RTB_ERR_INVALID_ARG, "seed pointer is null");
    if (pass->numSeedDims > 4095u) return fail(RTB_ERR_INVALID_ARG, "numSeedDims %u too large (limit 4095)", pass->numSeedDims);   // 12 bits of the packed path state
    memset(&P, 0, sizeof(P));
    P.camLocalToWorld = load_m4(pass->camera.localToWorld);
    P.aspectRatio = pass->camera.aspectRatio; P.tanHalfFoV = pass->camera.tanHalfFoV;
    P.dofEnable = pass->camera.dofEnable; P.bokehShape = pass->camera.bokehShape;
    if (pass->camera.bokehShape > RTB_BOKEH_TEXTURE) return fail(RTB_ERR_INVALID_ARG, "unknown bokeh shape %u", pass->camera.bokehShape);
    if (pass->camera.dofEnable && pass->camera.bokehShape == RTB_BOKEH_NGON) return fail(RTB_ERR_UNSUPPORTED, "the n-gon bokeh is not implemented (neither in the reference: Camera.cpp:205-207)");
    if (pass->camera.dofEnable && pass->camera.bokehShape == RTB_BOKEH_TEXTURE && !c->bokehCdf.ptr)
        return fail(RTB_ERR_INVALID_ARG, "bokehShape = texture, but the scene carries no bokeh distribution (RtbSceneDesc::bokehCdf)");
    P.bokehCdf = c->bokehCdf.ptr; P.bokehWidth = c->bokehWidth; P.bokehHeight = c->bokehHeight;
    P.aperture = pass->camera.aperture; P.focalPlaneDistance = pass->camera.focalPlaneDistance;
    P.barrelConst = pass->camera.barrelDistortionConstFactor; P.barrelVar = 0.0f;
    P.width = c->width; P.height = c->height;
    P.sampleOffsetX = pass->sampleOffset[0]; P.sampleOffsetY = pass->sampleOffset[1];
    P.renderer = pass->renderer; P.debugMode = pass->debugMode;
    P.maxRayDepth = pass->maxRayDepth; P.minRussianRouletteDepth = pass->minRussianRouletteDepth;
    P.lightSamplingAll = pass->lightSamplingStrategy == RTB_LIGHT_SAMPLING_ALL ? 1u : 0u;
    P.useBlueNoise = (pass->useBlueNoise && c->blueNoise.ptr) ? 1u : 0u;
    P.passIndex = pass->passIndex; P.numSeedDims = pass->numSeedDims;
    if (uploadSeed)
    {
        // single-pass helpers (rtb_sample_batch / rtb_primary_rays)
        if (c->seed.count < pass->numSeedDims || !c->seed.ptr) { RTB_CUDA(cudaStreamSynchronize(c->stream)); RTB_CUDA(c->seed.alloc(pass->numSeedDims > 64 ? pass->numSeedDims : 64)); }
        if (pass->numSeedDims) RTB_CUDA(cudaMemcpyAsync(c->seed.ptr, pass->seed, pass->numSeedDims * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
    }
    P.seed = c->seed.ptr;
    return RTB_OK;
}

static cudaEvent_t next_event(RtbContext* c, size_t& cursor, cudaStream_t stream)
{
    if (cursor >= c->events.size()) { cudaEvent_t ev; cudaEventCreate(&ev); c->events.push_back(ev); }
    cudaEvent_t ev = c->events[cursor++];
    cudaEventRecord(ev, stream);
    return ev;
}

// one wavefront of `batch` consecutive passes that share everything but (seed, sampleOffset, passIndex)
static int render_batch(RtbContext* c, const RtbPassDesc* passes, uint32_t batch, uint32_t snapshotIndex)
{
    const RtbPassDesc* pass = &passes[0];
    PassParams P;
    int rc = fill_pass_params(c, pass, P, false);
    if (rc != RTB_OK) return rc;
    // this wavefront's slot: own stream, own queues.  Profiling measures one wavefront at a time on slot 0.
    const bool prof = c->profiling;
    RtbContext::Slot& slot = c->slots[prof ? 0u : (c->nextSlot++ % (uint32_t)c->slots.size())];
    cudaStream_t s = slot.stream;
    // position of the main stream now: everything the caller enqueued there before this call (film resets, reads of
    // the film such as snapshots or an all-gather) ...
    RTB_CUDA(cudaEventRecord(c->mainMark, c->stream));
    if (c->slotsNeedMain || prof)
    {
        // ... must be visible to the whole wavefront after a reset / resize / scene upload
        for (RtbContext::Slot& t : c->slots) RTB_CUDA(cudaStreamWaitEvent(t.stream, c->mainMark, 0));
        c->slotsNeedMain = false;
    }

    if (c->serializeSlots && c->lastFilm) RTB_CUDA(cudaStreamWaitEvent(s, c->lastFilm, 0));     // debugging aid (RTB_SERIALIZE_SLOTS): no overlap between wavefronts
    // per-pass seed vectors + AA offsets of the batch (stream ordered: the slot's previous wavefront is done with them)
    const uint32_t dims = pass->numSeedDims;
    // one upload: [batch x dims seed integers][batch x 2 AA offsets]
    const size_t words = (size_t)batch * (dims + 2u);
    if (slot.seed.count < (size_t)c->maxBatch * (dims + 2u) || !slot.seed.ptr)
    {
        RTB_CUDA(cudaStreamSynchronize(s));
        RTB_CUDA(slot.seed.alloc((size_t)c->maxBatch * ((dims > 64 ? dims : 64) + 2u)));
    }
    slot.stagingSeeds.resize(words);
    for (uint32_t b = 0; b < batch; ++b)
    {
        if (dims) memcpy(&slot.stagingSeeds[(size_t)b * dims], passes[b].seed, dims * sizeof(uint32_t));
        memcpy(&slot.stagingSeeds[(size_t)batch * dims + 2u * b], passes[b].sampleOffset, 2 * sizeof(float));
    }
    RTB_CUDA(cudaMemcpyAsync(slot.seed.ptr, slot.stagingSeeds.data(), words * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
    P.seed = slot.seed.ptr;
    P.sampleOffsets = reinterpret_cast<const float*>(slot.seed.ptr + (size_t)batch * dims);

    Wavefront W = slot.wf;
    W.batch = batch;
    W.firstPassIndex = pass->passIndex;
    const uint32_t perPass = c->hasActiveList ? c->numActive : c->numPixels;
    const uint32_t numPaths = perPass * batch;
    if (c->hasActiveList)
    {
        // the pixels outside the active tiles receive nothing from these passes (Viewport::RenderTile is not called for them)
        W.activeList = c->activeList.ptr; W.numActive = c->numActive;
        RTB_CUDA(cudaMemsetAsync(W.passColor, 0, (size_t)c->numPixels * batch * sizeof(float4), s));
    }

    const bool count = (c->flags & RTB_FLAG_COUNT_TESTS) != 0;
    // (k_raygen initialises the queue counters of this wavefront)

    // grids: oversubscribed grid-stride launches, but never more CTAs than the wavefront can have work for (a small
    // shard's single pass would otherwise spend its time retiring ~19 000 empty CTAs per launch)
    const int blocksCap = c->numSMs * c->simpleBlocksPerSM;
    const int blocks128 = (int)std::min<size_t>((size_t)blocksCap, std::max<size_t>(1, ((size_t)numPaths + 127) / 128));
    const int blocksShadow = (int)std::min<size_t>((size_t)blocksCap, std::max<size_t>(1, ((size_t)numPaths * std::max<uint32_t>(1u, W.raysPerPath) + 127) / 128));
    const int blocks256 = c->numSMs * 4;
    const int blocksShade = (int)std::min<size_t>((size_t)blocks128, std::max<size_t>(1, ((size_t)numPaths + RTB_SHADE_TILE - 1) / RTB_SHADE_TILE));
    // the instrumented (counting) context keeps the plain wavefront so that its per-ray test counters stay per kernel
    const uint32_t tailThreshold = (count || pass->renderer == RTB_RENDERER_DEBUG) ? 0u : c->tailThreshold;
    const int tailBlocks = tailThreshold ? (int)std::min<size_t>((size_t)blocks128, ((size_t)tailThreshold + 127) / 128) : 0;
    W.tailThreshold = tailThreshold;
    size_t ev = 0;
    std::vector<cudaEvent_t> marks;
    const uint32_t depthLimit = pass->renderer == RTB_RENDERER_DEBUG ? 0u : (numPaths ? pass->maxRayDepth : 0u);
    const bool shadows = pass->renderer == RTB_RENDERER_PATH_TRACER_MIS && c->numLights > 0;
    const bool textured = c->textures.count > 0;     // scenes without textures run the k_shade that has no texture calls compiled in
    auto enqueue = [&]()
    {
    if (prof) marks.push_back(next_event(c, ev, s));
    k_raygen<<<blocks256, 256, 0, s>>>(c->view, P, W);
    if (prof) marks.push_back(next_event(c, ev, s));
    for (uint32_t d = 0; d <= depthLimit; ++d)
    {
        if (c->persistentTraverse)
        {
            const int persistentBlocks = std::min(blocks128, c->numSMs * RTB_TRAV_MINBLOCKS);
            if (count) k_traverse_p<true><<<persistentBlocks, 128, c->stackBytes, s>>>(c->view, W, d);
            else k_traverse_p<false><<<persistentBlocks, 128, c->stackBytes, s>>>(c->view, W, d);
        }
        else if (count) k_traverse<true><<<blocks128, 128, c->stackBytes, s>>>(c->view, W, d);
        else k_traverse<false><<<blocks128, 128, c->stackBytes, s>>>(c->view, W, d);
        if (prof) marks.push_back(next_event(c, ev, s));
        // (a tile of RTB_SHADE_TILE paths per CTA and loop iteration)
        if (c->sortShade)
        {
            if (textured) launch_shade<true, true>(blocksShade, s, c->view, P, W, d);
            else launch_shade<false, true>(blocksShade, s, c->view, P, W, d);
        }
        else if (textured) launch_shade<true, false>(blocksShade, s, c->view, P, W, d);
        else launch_shade<false, false>(blocksShade, s, c->view, P, W, d);
        if (prof) marks.push_back(next_event(c, ev, s));
        if (shadows)
        {
            // resident warps only: they fetch rays dynamically until the list is empty
            const int shadowBlocks = RTB_SHADOW_REFILL ? c->numSMs * RTB_TRAV_MINBLOCKS : blocksShadow;
            if (c->shadowKernel == 2 && c->hotTotal > 0)
            {
                // top level for every ray, then the queued mesh visits on persistent warps (16 stack words per thread: 8 levels)
                const size_t meshPanel = (size_t)2 * c->wideDepth * RTB_STACK_THREADS * sizeof(uint32_t);
                const int meshBlocks = c->numSMs * RTB_SM_MINBLOCKS;
                if (count)
                {
                    k_shadow_top<true><<<shadowBlocks, 128, c->stackBytes, s>>>(c->view, W, d);
                    k_shadow_mesh<true><<<meshBlocks, RTB_SM_THREADS, (RTB_SM_THREADS / 128) * meshPanel, s>>>(c->view, W, d, (uint32_t)meshPanel);
                }
                else
                {
                    k_shadow_top<false><<<shadowBlocks, 128, c->stackBytes, s>>>(c->view, W, d);
                    k_shadow_mesh<false><<<meshBlocks, RTB_SM_THREADS, (RTB_SM_THREADS / 128) * meshPanel, s>>>(c->view, W, d, (uint32_t)meshPanel);
                }
                c->counters.numKernelLaunches++;
            }
            else if (c->wideKernel && c->hotTotal > 0)
            {
                // persistent CTAs; shared memory: barrier + staged tree top + one stack panel per 128 threads
                const size_t perCta = (233472u - 1024u * RTB_SW_MINBLOCKS) / RTB_SW_MINBLOCKS;
                const size_t stacks = (size_t)((RTB_SW_THREADS + 127) / 128) * c->widePanelBytes;
                size_t hot = perCta > stacks + 256u ? (perCta - stacks - 256u) / sizeof(WNode) : 0u;
                if (hot > c->hotTotal) hot = c->hotTotal;
                if (c->hotOverride >= 0 && (size_t)c->hotOverride < hot) hot = (size_t)c->hotOverride;      // (-1: all that fit)
                const size_t smem = 128u + ((hot * sizeof(WNode) + 127u) & ~(size_t)127u) + stacks;
                const int wideBlocks = (int)std::min<size_t>((size_t)c->numSMs * RTB_SW_MINBLOCKS, std::max<size_t>(1, ((size_t)numPaths * std::max<uint32_t>(1u, W.raysPerPath) + RTB_SW_THREADS - 1) / RTB_SW_THREADS));
                if (hot)
                {
                    if (count) k_shadow_wide<true, true><<<wideBlocks, RTB_SW_THREADS, smem, s>>>(c->view, W, d, c->hotNodes, (uint32_t)hot, (uint32_t)c->widePanelBytes);
                    else k_shadow_wide<false, true><<<wideBlocks, RTB_SW_THREADS, smem, s>>>(c->view, W, d, c->hotNodes, (uint32_t)hot, (uint32_t)c->widePanelBytes);
                }
                else if (count) k_shadow_wide<true, false><<<wideBlocks, RTB_SW_THREADS, smem, s>>>(c->view, W, d, c->hotNodes, 0u, (uint32_t)c->widePanelBytes);
                else k_shadow_wide<false, false><<<wideBlocks, RTB_SW_THREADS, smem, s>>>(c->view, W, d, c->hotNodes, 0u, (uint32_t)c->widePanelBytes);
            }
            else if (count) k_shadow<true><<<shadowBlocks, 128, c->stackBytes, s>>>(c->view, W, d);
            else k_shadow<false><<<shadowBlocks, 128, c->stackBytes, s>>>(c->view, W, d);
            k_resolve<<<blocks256, 256, 0, s>>>(W, d);
        }
        if (tailBlocks > 0 && d < depthLimit)
        {
            // does nothing unless at most tailThreshold paths are left, then finishes them all (see k_tail)
            rtb_launch_tail(textured, tailBlocks, c->stackBytes, s, c->view, P, W, d + 1);
        }
        if (prof) marks.push_back(next_event(c, ev, s));
    }
    };
    if (tailBlocks > 0) c->counters.numKernelLaunches += depthLimit;
    // The chain as a CUDA graph.  Everything a kernel of the chain reads from its arguments goes into the key; what changes
    // from wavefront to wavefront without the kernels looking at it is blanked first: the pass index (only the `Single`
    // light pick hashes it - such passes keep plain launches), pass slot 0's AA offset (the batched path reads the
    // offsets from the seed upload) and the film parity (k_film is launched outside the graph).
    // (only on a stream this context has to itself - the caller's, or a slot's: contexts without one share a stream per
    //  device, and another thread's launches on it would end up inside the capture)
    const bool graphable = c->useGraphs && c->exclusiveStream && !prof && !count && (P.lightSamplingAll || pass->renderer != RTB_RENDERER_PATH_TRACER_MIS);
    if (!graphable) enqueue();
    else
    {
        struct Key { PassParams P; Wavefront W; SceneView S; uint32_t numPaths, depthLimit, tailBlocks, flags; size_t stackBytes; const void* hot; int hotOverride, wide; };
        std::vector<unsigned char> key(sizeof(Key), 0);
        {
            Key k;
            memset(&k, 0, sizeof(k));
            memcpy(&k.P, &P, sizeof(P)); memcpy(&k.W, &W, sizeof(W)); memcpy(&k.S, &c->view, sizeof(SceneView));
            k.P.passIndex = 0; k.P.sampleOffsetX = 0.0f; k.P.sampleOffsetY = 0.0f; k.W.firstPassIndex = 0;
            k.numPaths = numPaths; k.depthLimit = depthLimit; k.tailBlocks = (uint32_t)tailBlocks; k.flags = (shadows ? 1u : 0u) | (textured ? 2u : 0u) | (c->sortShade ? 4u : 0u) | (c->persistentTraverse ? 8u : 0u);
            k.stackBytes = c->stackBytes; k.hot = c->hotNodes; k.hotOverride = c->hotOverride; k.wide = c->shadowKernel;
            memcpy(key.data(), &k, sizeof(k));
        }
        if (!slot.graphExec || slot.graphKey != key)
        {
            if (slot.graphExec) { cudaGraphExecDestroy(slot.graphExec); slot.graphExec = nullptr; }
            cudaGraph_t graph = nullptr;
            RTB_CUDA(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
            enqueue();
            const cudaError_t launchError = cudaGetLastError();
            RTB_CUDA(cudaStreamEndCapture(s, &graph));
            if (launchError != cudaSuccess) { if (graph) cudaGraphDestroy(graph); return fail(RTB_ERR_CUDA, "kernel launch failed during graph capture: %s", cudaGetErrorString(launchError)); }
            const cudaError_t ie = cudaGraphInstantiate(&slot.graphExec, graph, 0);
            cudaGraphDestroy(graph);
            if (ie != cudaSuccess) { slot.graphExec = nullptr; return fail(RTB_ERR_CUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(ie)); }
            slot.graphKey = key;
        }
        RTB_CUDA(cudaGraphLaunch(slot.graphExec, s));
    }
    // the film: after the previous wavefront's k_film (pass order) and after whatever the caller had enqueued on the main
    // stream when this wavefront was submitted (e.g. a snapshot or an all-gather still reading the film)
    if (c->lastFilm) RTB_CUDA(cudaStreamWaitEvent(s, c->lastFilm, 0));
    RTB_CUDA(cudaStreamWaitEvent(s, c->mainMark, 0));
    // film snapshots (rtb_film_snapshots): the passes of this call whose index is below the buffer's capacity
    float* snap = nullptr;
    uint32_t snapCount = 0;
    if (c->snapshots.ptr && snapshotIndex < c->snapshotCapacity)
    {
        snap = c->snapshots.ptr + (size_t)snapshotIndex * c->shardSlots * 3;
        snapCount = std::min(batch, c->snapshotCapacity - snapshotIndex);
    }
    if (snap && c->scatterPending) { RTB_CUDA(cudaStreamWaitEvent(s, c->scatterDone, 0)); c->scatterPending = false; }     // (rtb_scatter_shard_async of the previous call's snapshots)
    k_film<<<blocks256, 256, 0, s>>>(W, c->sum.ptr, c->sum2.ptr, snap, (size_t)c->shardSlots * 3, snapCount);
    RTB_CUDA(cudaEventRecord(slot.filmDone, s));
    c->lastFilm = slot.filmDone;
    if (prof) marks.push_back(next_event(c, ev, s));
    RTB_CUDA(cudaGetLastError());
    c->counters.numPrimaryRays += (uint64_t)perPass * batch;          // Viewport.cpp:429
    c->counters.numKernelLaunches += 2u + (depthLimit + 1u) * (shadows ? 4u : 2u);
    c->passesSinceReset += batch;

    if (prof)
    {
        RTB_CUDA(cudaStreamSynchronize(s));
        RtbPassTimings& t = c->timings;
        auto ms = [&](size_t a, size_t b) { float v = 0.0f; cudaEventElapsedTime(&v, marks[a], marks[b]); return v; };
        t.raygenMs += ms(0, 1);
        size_t k = 1;
        for (uint32_t d = 0; d <= depthLimit; ++d)
        {
            t.traverseMs += ms(k, k + 1); t.shadeMs += ms(k + 1, k + 2); t.shadowMs += ms(k + 2, k + 3);
            k += 3;
            t.traverseLaunches++; t.shadeLaunches++;
            if (shadows) t.shadowLaunches += 2;
        }
        t.totalMs += ms(0, marks.size() - 1);
        t.totalLaunches = 2 + t.traverseLaunches + t.shadeLaunches + t.shadowLaunches;
        Control h;
        RTB_CUDA(cudaMemcpy(&h, slot.ctl.ptr, sizeof(Control), cudaMemcpyDeviceToHost));
        for (uint32_t d = 0; d <= depthLimit; ++d) t.traverseRays += h.pathCount[d];
        for (uint32_t d = 0; d <= depthLimit; ++d) t.shadowRays += h.shadowRayCount[d];
        if (c->profileDepths)      // development aid (RTB_PROFILE_DEPTHS): per-depth breakdown of this wavefront
            for (uint32_t d = 0; d <= depthLimit; ++d)
                fprintf(stderr, "depth %u: paths %u traverse %.3f ms, shade %.3f ms, shadow rays %u shadow+resolve %.3f ms\n", d, h.pathCount[d],
                        ms(1 + 3 * d, 2 + 3 * d), ms(2 + 3 * d, 3 + 3 * d), h.shadowRayCount[d], ms(3 + 3 * d, 4 + 3 * d));
    }
    return RTB_OK;
}


// ---- "VCM" passes (rtb_vcm.cuh) --------------------------------------------------------------------------------------------
static int vcm_allocate(RtbContext* c)
{
    RtbContext::Vcm& v = c->vcm;
    const uint32_t stride = v.params.maxPathLength;
    if (v.allocated && v.vertexStride == stride && v.lightVertexCount.count == c->numPixels) return RTB_OK;
    RTB_CUDA(sync_all(c));
    free_vcm(c);
    v.vertexStride = stride;
    const size_t capacity = (size_t)c->numPixels * stride;
    if (capacity >= (1ull << 31)) return fail(RTB_ERR_UNSUPPORTED, "VCM: %zu light vertices per pass exceed 2^31", capacity);
    v.photonCapacity = (uint32_t)capacity;
    RTB_CUDA(v.lightVertices.alloc(capacity * (RTB_VCM_VERTEX_WORDS / 4u)));
    RTB_CUDA(v.lightVertexCount.alloc(c->numPixels));
    RTB_CUDA(v.photons[0].alloc(capacity * 3u)); RTB_CUDA(v.photons[1].alloc(capacity * 3u));
    RTB_CUDA(v.photonCount.alloc(2));
    uint32_t hashSize = 1u;                                   // HashGrid::Build: NextPowerOfTwo(particles.Size()) (HashGrid.h:40), here of the capacity
    while (hashSize < capacity) hashSize <<= 1;
    v.hashSize = hashSize;
    RTB_CUDA(v.cellStart.alloc((size_t)hashSize + 1u)); RTB_CUDA(v.cellCursor.alloc((size_t)hashSize + 1u));
    RTB_CUDA(v.gridIndices.alloc(capacity));
    RTB_CUDA(v.tileTotals.alloc(((size_t)hashSize + 1u + 1023u) / 1024u + 1u));
    const size_t raster = (size_t)c->width * c->height * 3u;
    RTB_CUDA(v.splat.alloc(raster)); RTB_CUDA(v.splat2.alloc(raster));
    RTB_CUDA(cudaMemsetAsync(v.splat.ptr, 0, raster * sizeof(float), c->stream));
    RTB_CUDA(cudaMemsetAsync(v.splat2.ptr, 0, raster * sizeof(float), c->stream));
    RTB_CUDA(cudaMemsetAsync(v.photonCount.ptr, 0, 2 * sizeof(uint32_t), c->stream));
    RTB_CUDA(cudaMemsetAsync(v.cellStart.ptr, 0, ((size_t)hashSize + 1u) * sizeof(uint32_t), c->stream));
    v.allocated = true; v.current = 0; v.lastPassIndex = 0xFFFFFFFFu; v.splatsPending = false;
    {
        const int maxStack = (int)((size_t)RTB_MAX_SMEM_STACK_ENTRIES * RTB_STACK_THREADS * sizeof(uint32_t));
        cudaFuncSetAttribute(k_vcm_light<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxStack);
        cudaFuncSetAttribute(k_vcm_light<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxStack);
        cudaFuncSetAttribute(k_vcm_camera<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxStack);
        cudaFuncSetAttribute(k_vcm_camera<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxStack);
    }
    return RTB_OK;
}

static int vcm_fold(RtbContext* c)
{
    RtbContext::Vcm& v = c->vcm;
    if (!v.allocated || !v.splatsPending) return RTB_OK;
    cudaStream_t s = c->stream;
    k_vcm_fold<<<c->numSMs * 4, 256, 0, s>>>(v.splat.ptr, c->sum.ptr, c->pixelList.ptr, c->numPixels);
    k_vcm_fold<<<c->numSMs * 4, 256, 0, s>>>(v.splat2.ptr, c->sum2.ptr, c->pixelList.ptr, c->numPixels);
    RTB_CUDA(cudaGetLastError());
    RTB_CUDA(cudaMemsetAsync(v.splat.ptr, 0, v.splat.count * sizeof(float), s));
    RTB_CUDA(cudaMemsetAsync(v.splat2.ptr, 0, v.splat2.count * sizeof(float), s));
    c->counters.numKernelLaunches += 2;
    v.splatsPending = false;
    return RTB_OK;
}

static int render_vcm_pass(RtbContext* c, const RtbPassDesc* pass)
{
    RtbContext::Vcm& v = c->vcm;
    // RTB_RENDERER_LIGHT_TRACER shares the splat machinery (raster splat buffers, fold, all-reduce when sharded) and nothing else
    const bool lightTracer = pass->renderer == RTB_RENDERER_LIGHT_TRACER;
    if (c->hasActiveList) return fail(RTB_ERR_UNSUPPORTED, "VCM / Light Tracer: adaptive rendering (rtb_set_active_tiles) is not supported");
    if (!lightTracer && c->numShards > 1 && v.params.useVertexMerging) return fail(RTB_ERR_UNSUPPORTED, "VCM: vertex merging needs the photons of every shard; sharded contexts support vertex connection only");
    PassParams P;
    int rc = fill_pass_params(c, pass, P, false);
    if (rc != RTB_OK) return rc;
    rc = vcm_allocate(c);
    if (rc != RTB_OK) return rc;
    RTB_CUDA(join_slots(c));
    RtbContext::Slot& slot = c->slots[0];
    cudaStream_t s = c->stream;
    if (c->slotsOwnStreams) RTB_CUDA(cudaStreamSynchronize(slot.stream));
    // seed vector + AA offset of this pass
    const uint32_t dims = pass->numSeedDims;
    if (slot.seed.count < (size_t)c->maxBatch * (dims + 2u) || !slot.seed.ptr)
    {
        RTB_CUDA(cudaStreamSynchronize(s));
        RTB_CUDA(slot.seed.alloc((size_t)c->maxBatch * ((dims > 64 ? dims : 64) + 2u)));
    }
    slot.stagingSeeds.resize(dims + 2u);
    if (dims) memcpy(slot.stagingSeeds.data(), pass->seed, dims * sizeof(uint32_t));
    memcpy(&slot.stagingSeeds[dims], pass->sampleOffset, 2 * sizeof(float));
    RTB_CUDA(cudaMemcpyAsync(slot.seed.ptr, slot.stagingSeeds.data(), (dims + 2u) * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
    P.seed = slot.seed.ptr;
    P.sampleOffsets = reinterpret_cast<const float*>(slot.seed.ptr + dims);

    VcmDevice V;
    memset(&V, 0, sizeof(V));
    // PreRender (VertexConnectionAndMerging.cpp:80-125): the radii follow from the pass number alone
    V.pass.useVertexConnection = v.params.useVertexConnection ? 1u : 0u; V.pass.useVertexMerging = v.params.useVertexMerging ? 1u : 0u;
    V.pass.maxPathLength = v.params.maxPathLength;
    for (uint32_t p = 0; p <= pass->passIndex; ++p)
        bidir_begin_pass(V.pass, p, (float)(c->width * c->height), v.params.initialMergingRadius, v.params.minMergingRadius, v.params.mergingRadiusMultiplier);
    // photons of the previous pass of this frame, if that is what the buffer holds
    const bool havePhotons = !lightTracer && v.params.useVertexMerging && pass->passIndex > 0 && v.lastPassIndex + 1u == pass->passIndex;
    if (!lightTracer && v.params.useVertexMerging && pass->passIndex > 0 && !havePhotons)
        return fail(RTB_ERR_INVALID_ARG, "VCM: pass %u merges with the photons of pass %u, which was not the previous pass rendered", pass->passIndex, pass->passIndex - 1u);
    V.camera.localToWorld = P.camLocalToWorld; V.camera.tanHalfFoV = P.tanHalfFoV; V.camera.aspectRatio = P.aspectRatio;
    V.camera.worldToScreen = camera_world_to_screen(P.camLocalToWorld, P.aspectRatio, P.tanHalfFoV);
    const uint32_t prev = v.current, next = v.current ^ 1u;
    V.lightVertices = v.lightVertices.ptr; V.lightVertexCount = v.lightVertexCount.ptr;
    V.newPhotons = v.photons[next].ptr; V.newPhotonCount = v.photonCount.ptr + next;
    V.photons = v.photons[prev].ptr; V.photonCount = v.photonCount.ptr + prev;
    V.cellStart = v.cellStart.ptr; V.gridIndices = v.gridIndices.ptr;
    V.hashMask = v.hashSize - 1u; V.photonCapacity = v.photonCapacity;
    V.splat = v.splat.ptr; V.splat2 = (pass->passIndex & 1u) == 0u ? v.splat2.ptr : nullptr;      // Film: the secondary sum takes the even passes (Viewport.cpp:283)
    V.vertexStride = v.vertexStride;
    V.seedLo = (uint32_t)v.params.lightPathSeed; V.seedHi = (uint32_t)(v.params.lightPathSeed >> 32);

    const int blocks256 = c->numSMs * 4;
    uint32_t launches = 0;
    if (lightTracer)
    {
        // LightTracer::RenderPixel returns black for the pixel itself (LightTracer.cpp:174): the film only receives the splats
        const int blocks = (int)std::min<size_t>((size_t)c->numSMs * 16, std::max<size_t>(1, ((size_t)c->numPixels + 127) / 128));
        if (c->textures.count > 0) k_light_tracer<true><<<blocks, 128, c->stackBytes, s>>>(c->view, P, V, c->pixelList.ptr, c->numPixels, c->totals.ptr);
        else k_light_tracer<false><<<blocks, 128, c->stackBytes, s>>>(c->view, P, V, c->pixelList.ptr, c->numPixels, c->totals.ptr);
        RTB_CUDA(cudaGetLastError());
        v.splatsPending = true;
        v.lastPassIndex = 0xFFFFFFFEu;                 // no photons of this pass for a later VCM pass to merge with
        c->counters.numPrimaryRays += c->numPixels;
        c->counters.numKernelLaunches += 1;
        c->passesSinceReset += 1;
        if (c->numShards == 1) return vcm_fold(c);
        return RTB_OK;
    }
    if (!havePhotons) RTB_CUDA(cudaMemsetAsync(v.photonCount.ptr + prev, 0, sizeof(uint32_t), s));
    if (havePhotons)
    {
        // HashGrid::Build (HashGrid.h:17-82): count per cell, exclusive prefix sum, fill
        const uint32_t cells = v.hashSize + 1u, tiles = (cells + 1023u) / 1024u;
        const float invCell = 1.0f / (V.pass.radiusVM * 2.0f);
        RTB_CUDA(cudaMemsetAsync(v.cellStart.ptr, 0, (size_t)cells * sizeof(uint32_t), s));
        k_vcm_grid_count<<<blocks256, 256, 0, s>>>(V.photons, V.photonCount, v.photonCapacity, invCell, V.hashMask, v.cellStart.ptr);
        k_vcm_scan_tiles<<<tiles, 256, 0, s>>>(v.cellStart.ptr, cells, v.tileTotals.ptr);
        k_vcm_scan_totals<<<1, 1024, 0, s>>>(v.tileTotals.ptr, tiles, v.tileTotals.ptr + tiles);
        k_vcm_scan_offsets<<<(cells + 255u) / 256u, 256, 0, s>>>(v.cellStart.ptr, cells, v.tileTotals.ptr);
        RTB_CUDA(cudaMemcpyAsync(v.cellCursor.ptr, v.cellStart.ptr, (size_t)cells * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
        k_vcm_grid_fill<<<blocks256, 256, 0, s>>>(V.photons, V.photonCount, v.photonCapacity, invCell, V.hashMask, v.cellCursor.ptr, v.gridIndices.ptr);
        launches += 5;
    }
    RTB_CUDA(cudaMemsetAsync(v.photonCount.ptr + next, 0, sizeof(uint32_t), s));
    const bool textured = c->textures.count > 0;
    const int blocks128 = (int)std::min<size_t>((size_t)c->numSMs * 16, std::max<size_t>(1, ((size_t)c->numPixels + 127) / 128));
    Wavefront W = slot.wf;
    W.batch = 1; W.firstPassIndex = pass->passIndex;
    if (textured) k_vcm_light<true><<<blocks128, 128, c->stackBytes, s>>>(c->view, P, V, c->pixelList.ptr, c->numPixels, c->totals.ptr);
    else k_vcm_light<false><<<blocks128, 128, c->stackBytes, s>>>(c->view, P, V, c->pixelList.ptr, c->numPixels, c->totals.ptr);
    if (textured) k_vcm_camera<true><<<blocks128, 128, c->stackBytes, s>>>(c->view, P, V, c->pixelList.ptr, c->numPixels, W.passColor, c->totals.ptr);
    else k_vcm_camera<false><<<blocks128, 128, c->stackBytes, s>>>(c->view, P, V, c->pixelList.ptr, c->numPixels, W.passColor, c->totals.ptr);
    k_film<<<blocks256, 256, 0, s>>>(W, c->sum.ptr, c->sum2.ptr, nullptr, 0, 0);
    RTB_CUDA(cudaGetLastError());
    launches += 3;
    v.splatsPending = true;
    v.current = next; v.lastPassIndex = pass->passIndex;
    c->counters.numPrimaryRays += c->numPixels;
    c->counters.numKernelLaunches += launches;
    c->passesSinceReset += 1;
    if (c->numShards == 1) return vcm_fold(c);
    return RTB_OK;
}

extern "C" int rtb_render_passes(RtbContext* c, const RtbPassDesc* passes, uint32_t count)
{
    if (!c || (count && !passes)) return fail(RTB_ERR_INVALID_ARG, "rtb_render_passes: null argument");
    if (!c->hasScene) return fail(RTB_ERR_NO_SCENE, "rtb_render_passes: no scene uploaded");       // Viewport.cpp:194-198 ("Missing renderer")
    if (c->width == 0 || c->height == 0) return fail(RTB_ERR_INVALID_ARG, "rtb_render_passes: viewport has no size");
    RTB_CUDA(cudaSetDevice(c->device));

    for (uint32_t i = 0; i < count; ++i)
    {
        const RtbPassDesc* pass = &passes[i];
        // the reference silently switches to a per-thread RNG beyond SamplingParams::dimensions (GenericSampler.cpp:106-109);
        // that is irreproducible, so refuse it up front.  Worst case per path: SURVEY appendix B.
        if (pass->renderer == RTB_RENDERER_VCM)
        {
            // camera path: per vertex three numbers per light (SampleLights) and three for the BSDF sample (AdvancePath)
            const uint32_t needed = (pass->camera.dofEnable ? 2u : 0u) + (3u * c->numLights + 3u) * c->vcm.params.maxPathLength;
            if (needed > pass->numSeedDims)
                return fail(RTB_ERR_INVALID_ARG, "rtb_render_pass: VCM needs %u sample dimensions (lights=%u, maxPathLength=%u) but only %u are seeded", needed, c->numLights, c->vcm.params.maxPathLength, pass->numSeedDims);
        }
        else if (pass->renderer != RTB_RENDERER_DEBUG && pass->renderer != RTB_RENDERER_LIGHT_TRACER)       // (the Light Tracer draws nothing from the pixel sampler)
        {
            const uint32_t L = pass->renderer == RTB_RENDERER_PATH_TRACER_MIS ? (pass->lightSamplingStrategy == RTB_LIGHT_SAMPLING_ALL ? c->numLights : (c->numLights ? 1u : 0u)) : 0u;
            const uint32_t D = pass->maxRayDepth;
            const uint32_t needed = (pass->camera.dofEnable ? 2u : 0u) + 3u * L * (D + 1u) + 4u * D;
            if (needed > pass->numSeedDims)
                return fail(RTB_ERR_INVALID_ARG, "rtb_render_pass: %u sample dimensions needed (lights=%u, maxRayDepth=%u) but only %u seeded", needed, L, D, pass->numSeedDims);
        }
        PassParams P;
        int rc = fill_pass_params(c, pass, P, false);     // validates renderer id, depth, camera, seed pointer
        if (rc != RTB_OK) return rc;
    }
    if (c->profiling) memset(&c->timings, 0, sizeof(c->timings));
    if (c->numPixels == 0) { c->passesSinceReset += count; return RTB_OK; }

    // consecutive passes that differ only in (seed, sampleOffset, passIndex = previous + 1) share a wavefront
    auto same = [](const RtbPassDesc& a, const RtbPassDesc& b)
    {
        return a.renderer == b.renderer && a.debugMode == b.debugMode && memcmp(&a.camera, &b.camera, sizeof(RtbCamera)) == 0 &&
               a.maxRayDepth == b.maxRayDepth && a.minRussianRouletteDepth == b.minRussianRouletteDepth &&
               a.lightSamplingStrategy == b.lightSamplingStrategy && a.useBlueNoise == b.useBlueNoise && a.numSeedDims == b.numSeedDims;
    };
    uint32_t i = 0;
    while (i < count)
    {
        if (passes[i].renderer == RTB_RENDERER_VCM || passes[i].renderer == RTB_RENDERER_LIGHT_TRACER)
        {
            int rc = render_vcm_pass(c, passes + i);
            if (rc != RTB_OK) return rc;
            i++;
            continue;
        }
        // run of compatible passes, then cut into equal wavefronts (64 passes with maxBatch 9 -> 8 x 8, not 7 x 9 + 1)
        uint32_t run = 1;
        while (i + run < count && same(passes[i], passes[i + run]) && passes[i + run].passIndex == passes[i].passIndex + run) run++;
        const uint32_t groups = (run + c->maxBatch - 1) / c->maxBatch;
        for (uint32_t g = 0; g < groups; ++g)
        {
            const uint32_t n = run / groups + (g < run % groups ? 1u : 0u);
            int rc = render_batch(c, passes + i, n, i);
            if (rc != RTB_OK) return rc;
            i += n;
        }
    }
    return RTB_OK;
}

int rtb_render_pass(RtbContext* c, const RtbPassDesc* pass)
{
    if (!c || !pass) return fail(RTB_ERR_INVALID_ARG, "rtb_render_pass: null argument");
    return rtb_render_passes(c, pass, 1);
}

int rtb_synchronize(RtbContext* c)
{
    if (!c) return fail(RTB_ERR_INVALID_ARG, "null context");
    RTB_CUDA(cudaSetDevice(c->device));
    RTB_CUDA(sync_all(c));
    return RTB_OK;
}

int rtb_join(RtbContext* c)
{
    if (!c) return fail(RTB_ERR_INVALID_ARG, "null context");
    RTB_CUDA(cudaSetDevice(c->device));
    RTB_CUDA(join_slots(c));
    return RTB_OK;
}

int rtb_film_device_buffer(RtbContext* c, int which, void** outDevicePtr, uint64_t* outNumFloats)
{
    if (!c || !outDevicePtr || !outNumFloats) return fail(RTB_ERR_INVALID_ARG, "null argument");
    if (!c->sum.ptr) return fail(RTB_ERR_INVALID_ARG, "viewport has no size");
    RTB_CUDA(cudaSetDevice(c->device));
    RTB_CUDA(join_slots(c));      // what the caller enqueues on the main stream next sees every pass rendered so far
    *outDevicePtr = which == 0 ? c->sum.ptr : c->sum2.ptr;
    *outNumFloats = (uint64_t)c->shardSlots * 3;
    return RTB_OK;
}

int rtb_film_snapshots(RtbContext* c, uint32_t maxPasses, void** outDevicePtr, uint64_t* outStrideFloats)
{
    if (!c) return fail(RTB_ERR_INVALID_ARG, "null context");
    if (!c->sum.ptr) return fail(RTB_ERR_INVALID_ARG, "viewport has no size");
    RTB_CUDA(cudaSetDevice(c->device));
    RTB_CUDA(sync_all(c));
    c->snapshots.release(); c->snapshotCapacity = 0;
    if (maxPasses)
    {
        RTB_CUDA(c->snapshots.alloc((size_t)maxPasses * c->shardSlots * 3));
        RTB_CUDA(cudaMemsetAsync(c->snapshots.ptr, 0, c->snapshots.count * sizeof(float), c->stream));
        RTB_CUDA(cudaStreamSynchronize(c->stream));
        c->snapshotCapacity = maxPasses;
    }
    if (outDevicePtr) *outDevicePtr = c->snapshots.ptr;
    if (outStrideFloats) *outStrideFloats = (uint64_t)c->shardSlots * 3;
    return RTB_OK;
}

int rtb_unshard_film(RtbContext* c, const void* gatheredDevice, float* hostRgb)
{
    if (!c || !gatheredDevice || !hostRgb) return fail(RTB_ERR_INVALID_ARG, "null argument");
    if (!c->sum.ptr) return fail(RTB_ERR_INVALID_ARG, "viewport has no size");
    RTB_CUDA(cudaSetDevice(c->device));
    RTB_CUDA(join_slots(c));
    const uint32_t total = c->numShards * c->shardSlots;
    k_unshard<<<c->numSMs * 4, 256, 0, c->stream>>>((const float*)gatheredDevice, c->raster.ptr, c->pixelListAll.ptr, total);
    RTB_CUDA(cudaGetLastError());
    RTB_CUDA(cudaMemcpyAsync(hostRgb, c->raster.ptr, (size_t)c->width * c->height * 3 * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    RTB_CUDA(cudaStreamSynchronize(c->stream));
    return RTB_OK;
}

// Asynchronous read-back: the film is snapshotted to one of two raster buffers on the render stream, and copied to the
// host on a second stream, so the copy of pass k overlaps the kernels of pass k+1.  `gatheredDevice` null: this
// context's own film (rtb_read_sum semantics), else the all-gathered shard films (rtb_unshard_film semantics).
// the second stream of a context (read-backs, rtb_scatter_shard_async) and its events
static int ensure_copy_stream(RtbContext* c)
{
    if (c->copyStream) return RTB_OK;
    RTB_CUDA(cudaStreamCreateWithFlags(&c->copyStream, cudaStreamNonBlocking));
    for (int i = 0; i < RTB_SNAP_RING; ++i)
    {
        RTB_CUDA(cudaEventCreateWithFlags(&c->snapReady[i], cudaEventDisableTiming));
        RTB_CUDA(cudaEventCreateWithFlags(&c->snapCopied[i], cudaEventDisableTiming));
        RTB_CUDA(cudaEventRecord(c->snapCopied[i], c->copyStream));
    }
    RTB_CUDA(cudaEventCreateWithFlags(&c->scatterReady, cudaEventDisableTiming));
    RTB_CUDA(cudaEventCreateWithFlags(&c->scatterDone, cudaEventDisableTiming));
    return RTB_OK;
}

static int read_back_async(RtbContext* c, const void* gatheredDevice, float* hostRgb, int which)
{
    if (!c || !hostRgb) return fail(RTB_ERR_INVALID_ARG, "null argument");
    if (!c->sum.ptr) return fail(RTB_ERR_INVALID_ARG, "viewport has no size");
    RTB_CUDA(cudaSetDevice(c->device));
    { const int rc = ensure_copy_stream(c); if (rc != RTB_OK) return rc; }
    RTB_CUDA(join_slots(c));
    const uint32_t b = c->snapIndex++ % RTB_SNAP_RING;
    const size_t n = (size_t)c->width * c->height * 3;
    if (c->rasterAsync[b].count != n) RTB_CUDA(c->rasterAsync[b].alloc(n));
    float* dst = c->rasterAsync[b].ptr;
    RTB_CUDA(cudaStreamWaitEvent(c->stream, c->snapCopied[b], 0));       // the copy that last used this buffer is done
    if (gatheredDevice)
    {
        k_unshard<<<c->numSMs * 4, 256, 0, c->stream>>>((const float*)gatheredDevice, dst, c->pixelListAll.ptr, c->numShards * c->shardSlots);
    }
    else if (c->numShards > 1)
    {
        RTB_CUDA(cudaMemsetAsync(dst, 0, n * sizeof(float), c->stream));
        k_unshard<<<c->numSMs * 4, 256, 0, c->stream>>>(which == 0 ? c->sum.ptr : c->sum2.ptr, dst, c->pixelList.ptr, c->numPixels);
    }
    else
    {
        k_unshard<<<c->numSMs * 4, 256, 0, c->stream>>>(which == 0 ? c->sum.ptr : c->sum2.ptr, dst, c->pixelListAll.ptr, c->shardSlots);
    }
    RTB_CUDA(cudaGetLastError());
    c->counters.numKernelLaunches++;
    RTB_CUDA(cudaEventRecord(c->snapReady[b], c->stream));
    RTB_CUDA(cudaStreamWaitEvent(c->copyStream, c->snapReady[b], 0));
    RTB_CUDA(cudaMemcpyAsync(hostRgb, dst, n * sizeof(float), cudaMemcpyDeviceToHost, c->copyStream));
    RTB_CUDA(cudaEventRecord(c->snapCopied[b], c->copyStream));
    return RTB_OK;
}

int rtb_read_sum_async(RtbContext* c, float* rgb, int which) { return read_back_async(c, nullptr, rgb, which); }

int rtb_unshard_film_async(RtbContext* c, const void* gatheredDevice, float* hostRgb)
{
    if (!gatheredDevice) return fail(RTB_ERR_INVALID_ARG, "null argument");
    return read_back_async(c, gatheredDevice, hostRgb, 0);
}

void* rtb_host_alloc(uint64_t bytes)
{
    void* p = nullptr;
    const cudaError_t e = cudaHostAlloc(&p, (size_t)bytes, cudaHostAllocPortable | cudaHostAllocMapped);
    if (e != cudaSuccess) { fail(RTB_ERR_CUDA, "rtb_host_alloc(%llu): %s", (unsigned long long)bytes, cudaGetErrorString(e)); return nullptr; }
    return p;
}

void rtb_host_free(void* p)
{
    if (p) cudaFreeHost(p);
}

int rtb_read_sum(RtbContext* c, float* rgb, int which)
{
    if (!c || !rgb) return fail(RTB_ERR_INVALID_ARG, "null argument");
    if (!c->sum.ptr) return fail(RTB_ERR_INVALID_ARG, "viewport has no size");
    RTB_CUDA(cudaSetDevice(c->device));
    RTB_CUDA(join_slots(c));
    if (c->numShards > 1)
    {
        // only this shard's pixels are known here: the rest of the raster stays zero
        RTB_CUDA(cudaMemsetAsync(c->raster.ptr, 0, c->raster.count * sizeof(float), c->stream));
        k_unshard<<<c->numSMs * 4, 256, 0, c->stream>>>(which == 0 ? c->sum.ptr : c->sum2.ptr, c->raster.ptr, c->pixelList.ptr, c->numPixels);
    }
    else
    {
        k_unshard<<<c->numSMs * 4, 256, 0, c->stream>>>(which == 0 ? c->sum.ptr : c->sum2.ptr, c->raster.ptr, c->pixelListAll.ptr, c->shardSlots);
    }
    RTB_CUDA(cudaGetLastError());
    RTB_CUDA(cudaMemcpyAsync(rgb, c->raster.ptr, (size_t)c->width * c->height * 3 * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    RTB_CUDA(cudaStreamSynchronize(c->stream));
    return RTB_OK;
}

int rtb_read_front(RtbContext* c, const RtbPostprocessDesc* post, const void* gatheredDevice, uint32_t* bgra)
{
    if (!c || !post || !bgra) return fail(RTB_ERR_INVALID_ARG, "null argument");
    if (!c->sum.ptr) return fail(RTB_ERR_INVALID_ARG, "viewport has no size");
    if (post->tonemapper > RTB_TONEMAP_ACES) return fail(RTB_ERR_INVALID_ARG, "unknown tonemapper %u", post->tonemapper);
    if (post->numPasses == 0) return fail(RTB_ERR_INVALID_ARG, "numPasses is 0: nothing has been accumulated");
    const bool bloom = post->bloomFactor > 0.0f;
    BlurSchedule schedules[RTB_BLOOM_LEVELS];
    if (bloom)
    {
        if (c->numShards > 1 && !gatheredDevice) return fail(RTB_ERR_UNSUPPORTED, "bloom blurs the whole image: pass the all-gathered film");
        // where the reference's blur leaves its image (Bitmap.cpp:981-1009 walks columns four at a time; BoxBlur_Internal
        // reads 2 * radius + 1 elements of a line) the result would be whatever lay next to it in memory
        if (c->width % 4u != 0u || c->width > 4096u || c->height > 4096u) return fail(RTB_ERR_UNSUPPORTED, "bloom needs a width that is a multiple of 4 and at most 4096 x 4096 pixels (Bitmap.cpp:926-931, 981)");
        float sigma = 2.0f;                                        // Viewport.cpp:438-444
        for (uint32_t l = 0; l < RTB_BLOOM_LEVELS; ++l, sigma *= 2.5f)
        {
            schedules[l] = blur_schedule(sigma);
            for (uint32_t i = 0; i < RTB_BLUR_ITERATIONS; ++i)
                if (2u * schedules[l].radius[i] >= (c->width < c->height ? c->width : c->height))
                    return fail(RTB_ERR_UNSUPPORTED, "bloom needs an image larger than its widest box blur (%u pixels)", 2u * schedules[l].radius[i] + 1u);
        }
    }
    RTB_CUDA(cudaSetDevice(c->device));
    RTB_CUDA(join_slots(c));
    const size_t n = (size_t)c->width * c->height;
    if (c->front.count != n) RTB_CUDA(c->front.alloc(n));

    PostParams P;
    const float exposureScale = powf(2.0f, post->exposure);       // Viewport.cpp:453
    P.colorScale = v3(post->colorFilter[0] * exposureScale, post->colorFilter[1] * exposureScale, post->colorFilter[2] * exposureScale);
    P.pixelScaling = 1.0f / (float)post->numPasses;
    P.contrast = post->contrast; P.saturation = post->saturation;
    P.ditheringStrength = post->ditheringStrength; P.bloomFactor = post->bloomFactor;
    P.tonemapper = post->tonemapper; P.ditherSeed = post->ditherSeed;

    const float* src = gatheredDevice ? (const float*)gatheredDevice : c->sum.ptr;
    const size_t levelStride = n * 3;
    const uint32_t* frozen = nullptr;
    if (c->hasActiveList)
    {
        if (c->frozenDirty)
        {
            RTB_CUDA(c->frozen.upload(c->hostFrozen, c->stream));
            c->frozenDirty = false;
        }
        frozen = c->frozen.ptr;
    }
    if (bloom)
    {
        // Viewport::PerformPostProcess (Viewport.cpp:434-445): level 0 = blur(sum), level i = blur(level i - 1)
        if (c->blurred.count != levelStride * RTB_BLOOM_LEVELS) RTB_CUDA(c->blurred.alloc(levelStride * RTB_BLOOM_LEVELS));
        if (c->blurScratch.count != levelStride * 2) RTB_CUDA(c->blurScratch.alloc(levelStride * 2));
        k_unshard<<<c->numSMs * 4, 256, 0, c->stream>>>(src, c->raster.ptr, c->pixelListAll.ptr, c->numShards * c->shardSlots);
        for (uint32_t l = 0; l < RTB_BLOOM_LEVELS; ++l)
        {
            const float* from = l == 0 ? c->raster.ptr : c->blurred.ptr + (l - 1) * levelStride;
            float* level = c->blurred.ptr + l * levelStride;
            k_blur<<<(c->height * 3 + 127) / 128, 128, 0, c->stream>>>(from, level, c->width, c->height, 0, c->blurScratch.ptr, c->blurScratch.ptr + levelStride, schedules[l]);
            k_blur<<<(c->width * 3 + 127) / 128, 128, 0, c->stream>>>(level, level, c->width, c->height, 1, c->blurScratch.ptr, c->blurScratch.ptr + levelStride, schedules[l]);
        }
        c->counters.numKernelLaunches += 1 + 2 * RTB_BLOOM_LEVELS;
    }
    if (gatheredDevice)
    {
        k_front<<<c->numSMs * 4, 256, 0, c->stream>>>(src, c->front.ptr, c->pixelListAll.ptr, c->numShards * c->shardSlots, c->blurred.ptr, levelStride, frozen, P);
    }
    else if (c->numShards > 1)
    {
        RTB_CUDA(cudaMemsetAsync(c->front.ptr, 0, n * sizeof(uint32_t), c->stream));
        k_front<<<c->numSMs * 4, 256, 0, c->stream>>>(src, c->front.ptr, c->pixelList.ptr, c->numPixels, c->blurred.ptr, levelStride, frozen, P);
    }
    else
    {
        k_front<<<c->numSMs * 4, 256, 0, c->stream>>>(src, c->front.ptr, c->pixelListAll.ptr, c->shardSlots, c->blurred.ptr, levelStride, frozen, P);
    }
    RTB_CUDA(cudaGetLastError());
    c->counters.numKernelLaunches++;
    RTB_CUDA(cudaMemcpyAsync(bgra, c->front.ptr, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    RTB_CUDA(cudaStreamSynchronize(c->stream));
    return RTB_OK;
}

int rtb_block_errors(RtbContext* c, const void* gatheredSum, const void* gatheredSum2, uint32_t numPasses, const RtbBlock* blocks, uint32_t numBlocks, float* errors)
{
    if (!c || (numBlocks && (!blocks || !errors))) return fail(RTB_ERR_INVALID_ARG, "null argument");
    if (!c->sum.ptr) return fail(RTB_ERR_INVALID_ARG, "viewport has no size");
    if ((gatheredSum == nullptr) != (gatheredSum2 == nullptr)) return fail(RTB_ERR_INVALID_ARG, "pass both gathered films or neither");
    if (c->numShards > 1 && !gatheredSum) return fail(RTB_ERR_UNSUPPORTED, "a block error sums whole rows of the image: pass the all-gathered films");
    if (numPasses == 0 || (numPasses & 1u)) return fail(RTB_ERR_INVALID_ARG, "block errors compare all passes with the even ones: numPasses must be even and > 0 (Viewport.cpp:559)");
    if (numBlocks == 0) return RTB_OK;
    std::vector<uint32_t> rowStart(numBlocks + 1, 0u);
    for (uint32_t i = 0; i < numBlocks; ++i)
    {
        const RtbBlock& b = blocks[i];
        if (b.minX >= b.maxX || b.minY >= b.maxY || b.maxX > c->width || b.maxY > c->height)
            return fail(RTB_ERR_INVALID_ARG, "block %u [%u, %u) x [%u, %u) is empty or outside the %u x %u image", i, b.minX, b.maxX, b.minY, b.maxY, c->width, c->height);
        const uint64_t next = (uint64_t)rowStart[i] + (b.maxY - b.minY);
        if (next > 0xFFFFFFFFull) return fail(RTB_ERR_INVALID_ARG, "too many block rows");
        rowStart[i + 1] = (uint32_t)next;
    }
    RTB_CUDA(cudaSetDevice(c->device));
    RTB_CUDA(join_slots(c));
    const size_t n = (size_t)c->width * c->height;
    if (c->raster2.count != n * 3) RTB_CUDA(c->raster2.alloc(n * 3));
    if (c->pixelError.count != n) RTB_CUDA(c->pixelError.alloc(n));
    if (c->rowError.count < rowStart[numBlocks] + (size_t)numBlocks) RTB_CUDA(c->rowError.alloc(rowStart[numBlocks] + (size_t)numBlocks));
    // blocks and row offsets in one upload: [numBlocks x RtbBlock][numBlocks + 1 offsets]
    std::vector<uint32_t> staging((size_t)numBlocks * 4 + numBlocks + 1);
    memcpy(staging.data(), blocks, (size_t)numBlocks * sizeof(RtbBlock));
    memcpy(staging.data() + (size_t)numBlocks * 4, rowStart.data(), rowStart.size() * sizeof(uint32_t));
    if (c->blockRows.count < staging.size()) RTB_CUDA(c->blockRows.alloc(staging.size()));
    RTB_CUDA(cudaMemcpyAsync(c->blockRows.ptr, staging.data(), staging.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
    const RtbBlock* dBlocks = reinterpret_cast<const RtbBlock*>(c->blockRows.ptr);
    const uint32_t* dRowStart = c->blockRows.ptr + (size_t)numBlocks * 4;

    const uint32_t slots = c->numShards * c->shardSlots;
    k_unshard<<<c->numSMs * 4, 256, 0, c->stream>>>(gatheredSum ? (const float*)gatheredSum : c->sum.ptr, c->raster.ptr, c->pixelListAll.ptr, slots);
    k_unshard<<<c->numSMs * 4, 256, 0, c->stream>>>(gatheredSum2 ? (const float*)gatheredSum2 : c->sum2.ptr, c->raster2.ptr, c->pixelListAll.ptr, slots);
    k_pixel_error<<<(uint32_t)((n + 255) / 256), 256, 0, c->stream>>>(c->raster.ptr, c->raster2.ptr, (uint32_t)n, 1.0f / (float)numPasses, c->pixelError.ptr);
    k_row_error<<<(rowStart[numBlocks] + 127) / 128, 128, 0, c->stream>>>(c->pixelError.ptr, c->width, dBlocks, dRowStart, numBlocks, c->rowError.ptr);
    float* dErrors = c->rowError.ptr + rowStart[numBlocks];
    k_block_error<<<(numBlocks + 127) / 128, 128, 0, c->stream>>>(c->rowError.ptr, dBlocks, dRowStart, numBlocks, c->width * c->height, dErrors);
    RTB_CUDA(cudaGetLastError());
    c->counters.numKernelLaunches += 5;
    RTB_CUDA(cudaMemcpyAsync(errors, dErrors, (size_t)numBlocks * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    RTB_CUDA(cudaStreamSynchronize(c->stream));
    return RTB_OK;
}

int rtb_set_active_tiles(RtbContext* c, const RtbBlock* tiles, uint32_t numTiles)
{
    if (!c) return fail(RTB_ERR_INVALID_ARG, "null context");
    if (!c->sum.ptr) return fail(RTB_ERR_INVALID_ARG, "viewport has no size");
    RTB_CUDA(cudaSetDevice(c->device));
    if (!tiles)
    {
        c->hasActiveList = false; c->numActive = 0; c->hostFrozen.clear();
        return RTB_OK;
    }
    std::vector<uint8_t> mask((size_t)c->width * c->height, 0);
    for (uint32_t i = 0; i < numTiles; ++i)
    {
        const RtbBlock& b = tiles[i];
        if (b.minX >= b.maxX || b.minY >= b.maxY || b.maxX > c->width || b.maxY > c->height)
            return fail(RTB_ERR_INVALID_ARG, "tile %u [%u, %u) x [%u, %u) is empty or outside the %u x %u image", i, b.minX, b.maxX, b.minY, b.maxY, c->width, c->height);
        for (uint32_t y = b.minY; y < b.maxY; ++y)
            for (uint32_t x = b.minX; x < b.maxX; ++x)
            {
                // a pixel in two tiles would get two samples per pass from the reference; its block lists never overlap
                if (mask[(size_t)y * c->width + x]) return fail(RTB_ERR_INVALID_ARG, "tile %u overlaps an earlier tile at pixel (%u, %u)", i, x, y);
                mask[(size_t)y * c->width + x] = 1;
            }
    }
    // pixels that leave the active set now hold the sum of the passes rendered so far (every shard records the same)
    if (c->hostFrozen.size() != mask.size()) c->hostFrozen.assign(mask.size(), 0u);
    for (size_t p = 0; p < mask.size(); ++p)
    {
        if (mask[p]) c->hostFrozen[p] = 0u;
        else if (c->hostFrozen[p] == 0u) c->hostFrozen[p] = c->passesSinceReset;
    }
    c->frozenDirty = true;
    // this shard's active pixel slots, in slot order (neighbouring paths stay neighbouring pixels)
    std::vector<uint32_t> active;
    active.reserve(c->hostPixelList.size());
    for (uint32_t slot = 0; slot < (uint32_t)c->hostPixelList.size(); ++slot) if (mask[c->hostPixelList[slot]]) active.push_back(slot);
    RTB_CUDA(sync_all(c));        // wavefronts in flight still read the previous list
    if (!active.empty()) RTB_CUDA(c->activeList.upload(active, c->stream));
    RTB_CUDA(cudaStreamSynchronize(c->stream));
    c->numActive = (uint32_t)active.size();
    c->hasActiveList = true;
    c->slotsNeedMain = true;
    return RTB_OK;
}

int rtb_get_counters(RtbContext* c, RtbCounters* out)
{
    if (!c || !out) return fail(RTB_ERR_INVALID_ARG, "null argument");
    RTB_CUDA(cudaSetDevice(c->device));
    RTB_CUDA(sync_all(c));
    Totals h;
    RTB_CUDA(cudaMemcpy(&h, c->totals.ptr, sizeof(Totals), cudaMemcpyDeviceToHost));
    *out = c->counters;
    out->numRays = h.numRays;
    out->numShadowRays = h.numShadowRays;
    out->numShadowRaysHit = h.numShadowRaysHit;
    out->numRayBoxTests = h.numBoxTests;
    out->numRayTriangleTests = h.numTriTests;
    out->numShadowBoxTests = h.numShadowBoxTests;
    out->numShadowTriangleTests = h.numShadowTriTests;
    return RTB_OK;
}

int rtb_trace_batch_device(RtbContext* c, const void* raysDevice, uint64_t n, void* hitsDevice, uint32_t mode, float* outMs)
{
    if (!c || (n && (!raysDevice || !hitsDevice))) return fail(RTB_ERR_INVALID_ARG, "null argument");
    if (!c->hasScene) return fail(RTB_ERR_NO_SCENE, "rtb_trace_batch: no scene uploaded");
    if (mode > RTB_TRACE_SHADOW) return fail(RTB_ERR_INVALID_ARG, "unknown trace mode %u", mode);
    RTB_CUDA(cudaSetDevice(c->device));
    if (n == 0) { if (outMs) *outMs = 0.0f; return RTB_OK; }
    const bool count = (c->flags & RTB_FLAG_COUNT_TESTS) != 0;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (outMs) { cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventRecord(e0, c->stream); }
    const int blocks = c->numSMs * 8;
    if (count) k_trace_batch<true><<<blocks, 128, c->stackBytes, c->stream>>>(c->view, (const RtbRay*)raysDevice, (RtbHit*)hitsDevice, n, mode, c->totals.ptr);
    else k_trace_batch<false><<<blocks, 128, c->stackBytes, c->stream>>>(c->view, (const RtbRay*)raysDevice, (RtbHit*)hitsDevice, n, mode, c->totals.ptr);
    RTB_CUDA(cudaGetLastError());
    if (outMs)
    {
        cudaEventRecord(e1, c->stream);
        RTB_CUDA(cudaEventSynchronize(e1));
        cudaEventElapsedTime(outMs, e0, e1);
        cudaEventDestroy(e0); cudaEventDestroy(e1);
    }
    return RTB_OK;
}

int rtb_trace_batch(RtbContext* c, const RtbRay* rays, uint64_t n, RtbHit* hits, uint32_t mode)
{
    if (!c || (n && (!rays || !hits))) return fail(RTB_ERR_INVALID_ARG, "null argument");
    if (!c->hasScene) return fail(RTB_ERR_NO_SCENE, "rtb_trace_batch: no scene uploaded");
    RTB_CUDA(cudaSetDevice(c->device));
    if (n == 0) return RTB_OK;
    RtbRay* dRays = nullptr; RtbHit* dHits = nullptr;
    RTB_CUDA(cudaMalloc(&dRays, n * sizeof(RtbRay)));
    cudaError_t e = cudaMalloc(&dHits, n * sizeof(RtbHit));
    if (e != cudaSuccess) { cudaFree(dRays); return fail(RTB_ERR_CUDA, "cudaMalloc failed: %s", cudaGetErrorString(e)); }
    cudaMemcpyAsync(dRays, rays, n * sizeof(RtbRay), cudaMemcpyHostToDevice, c->stream);
    cudaMemcpyAsync(dHits, hits, n * sizeof(RtbHit), cudaMemcpyHostToDevice, c->stream);   // shadow mode: carries tMax
    int rc = rtb_trace_batch_device(c, dRays, n, dHits, mode, nullptr);
    if (rc == RTB_OK)
    {
        e = cudaMemcpyAsync(hits, dHits, n * sizeof(RtbHit), cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) rc = fail(RTB_ERR_CUDA, "trace batch failed: %s", cudaGetErrorString(e));
    }
    cudaFree(dRays); cudaFree(dHits);
    return rc;
}

int rtb_texture_batch(RtbContext* c, uint32_t textureIndex, const float* uv, uint64_t n, float* rgba)
{
    if (!c || (n && (!uv || !rgba))) return fail(RTB_ERR_INVALID_ARG, "null argument");
    if (!c->hasScene) return fail(RTB_ERR_NO_SCENE, "rtb_texture_batch: no scene uploaded");
    if (textureIndex >= c->textures.count) return fail(RTB_ERR_INVALID_ARG, "texture index %u out of range (%zu textures)", textureIndex, c->textures.count);
    RTB_CUDA(cudaSetDevice(c->device));
    if (n == 0) return RTB_OK;
    float* dUv = nullptr; float4* dOut = nullptr;
    RTB_CUDA(cudaMalloc(&dUv, n * 2 * sizeof(float)));
    cudaError_t e = cudaMalloc(&dOut, n * sizeof(float4));
    if (e != cudaSuccess) { cudaFree(dUv); return fail(RTB_ERR_CUDA, "cudaMalloc failed: %s", cudaGetErrorString(e)); }
    cudaMemcpyAsync(dUv, uv, n * 2 * sizeof(float), cudaMemcpyHostToDevice, c->stream);
    k_texture_batch<<<(unsigned)((n + 127) / 128), 128, 0, c->stream>>>(c->view, textureIndex, dUv, n, dOut);
    e = cudaMemcpyAsync(rgba, dOut, n * sizeof(float4), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(dUv); cudaFree(dOut);
    if (e != cudaSuccess) return fail(RTB_ERR_CUDA, "texture batch failed: %s", cudaGetErrorString(e));
    return RTB_OK;
}

int rtb_sample_batch(RtbContext* c, const RtbPassDesc* pass, const uint32_t* xy, uint64_t n, uint32_t numDims, uint32_t* out)
{
    if (!c || !pass || (n && (!xy || !out))) return fail(RTB_ERR_INVALID_ARG, "null argument");
    RTB_CUDA(cudaSetDevice(c->device));
    if (n == 0 || numDims == 0) return RTB_OK;
    PassParams P;
    int rc = fill_pass_params(c, pass, P, true);
    if (rc != RTB_OK) return rc;
    uint32_t* dXy = nullptr; uint32_t* dOut = nullptr;
    RTB_CUDA(cudaMalloc(&dXy, n * 2 * sizeof(uint32_t)));
    cudaError_t e = cudaMalloc(&dOut, n * numDims * sizeof(uint32_t));
    if (e != cudaSuccess) { cudaFree(dXy); return fail(RTB_ERR_CUDA, "cudaMalloc failed: %s", cudaGetErrorString(e)); }
    cudaMemcpyAsync(dXy, xy, n * 2 * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream);
    SceneView v = c->view;
    k_sample_batch<<<(unsigned)((n + 127) / 128), 128, 0, c->stream>>>(v, P, dXy, n, numDims, dOut);
    e = cudaMemcpyAsync(out, dOut, n * numDims * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(dXy); cudaFree(dOut);
    if (e != cudaSuccess) return fail(RTB_ERR_CUDA, "sample batch failed: %s", cudaGetErrorString(e));
    return RTB_OK;
}

int rtb_primary_rays(RtbContext* c, const RtbPassDesc* pass, RtbRay* rays)
{
    if (!c || !pass || !rays) return fail(RTB_ERR_INVALID_ARG, "null argument");
    if (c->width == 0) return fail(RTB_ERR_INVALID_ARG, "viewport has no size");
    RTB_CUDA(cudaSetDevice(c->device));
    PassParams P;
    int rc = fill_pass_params(c, pass, P, true);
    if (rc != RTB_OK) return rc;
    const size_t n = (size_t)c->width * c->height;
    RtbRay* d = nullptr;
    RTB_CUDA(cudaMalloc(&d, n * sizeof(RtbRay)));
    k_primary_rays<<<(unsigned)((n + 127) / 128), 128, 0, c->stream>>>(c->view, P, d);
    cudaError_t e = cudaMemcpyAsync(rays, d, n * sizeof(RtbRay), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(d);
    if (e != cudaSuccess) return fail(RTB_ERR_CUDA, "primary rays failed: %s", cudaGetErrorString(e));
    return RTB_OK;
}

// ---- NCCL, loaded on demand (a single-GPU user of this library never needs it) ------------------------------------------------
struct NcclApi
{
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    std::string error;
};

static NcclApi* nccl_api()
{
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, []()
    {
        // a process that already has an NCCL (e.g. the one PyTorch ships) gets that one back
        api.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!api.lib) api.lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!api.lib) { api.error = std::string("cannot load libnccl.so.2: ") + dlerror(); return; }
        bool ok = true;
        auto sym = [&](const char* name) { void* p = dlsym(api.lib, name); if (!p) { ok = false; api.error = std::string("libnccl has no ") + name; } return p; };
        api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
        api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
        api.CommInitAll = (decltype(api.CommInitAll))sym("ncclCommInitAll");
        api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
        api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
        api.AllReduce = (decltype(api.AllReduce))sym("ncclAllReduce");
        api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
        api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
        api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
        if (!ok) { dlclose(api.lib); api.lib = nullptr; }
    });
    return api.lib ? &api : nullptr;
}

static void release_comm(RtbContext* c)
{
    if (!c->comm) return;
    if (NcclApi* n = nccl_api()) n->CommDestroy(c->comm);
    c->comm = nullptr;
}

#define RTB_NCCL(api, call) do { ncclResult_t r_ = (call); if (r_ != ncclSuccess) return fail(RTB_ERR_CUDA, "%s failed: %s", #call, (api)->GetErrorString(r_)); } while (0)

int rtb_comm_unique_id(void* outId)
{
    if (!outId) return fail(RTB_ERR_INVALID_ARG, "null argument");
    NcclApi* n = nccl_api();
    if (!n) return fail(RTB_ERR_UNSUPPORTED, "NCCL is not available: %s", nccl_api() ? "" : "libnccl.so.2 could not be loaded");
    static_assert(sizeof(ncclUniqueId) == RTB_COMM_ID_BYTES, "ncclUniqueId size");
    ncclUniqueId id;
    RTB_NCCL(n, n->GetUniqueId(&id));
    memcpy(outId, &id, sizeof(id));
    return RTB_OK;
}

int rtb_comm_init(RtbContext* c, const void* idBytes)
{
    if (!c || !idBytes) return fail(RTB_ERR_INVALID_ARG, "null argument");
    NcclApi* n = nccl_api();
    if (!n) return fail(RTB_ERR_UNSUPPORTED, "NCCL is not available (libnccl.so.2 could not be loaded)");
    if (c->comm) return fail(RTB_ERR_INVALID_ARG, "the context already has a communicator");
    RTB_CUDA(cudaSetDevice(c->device));
    ncclUniqueId id;
    memcpy(&id, idBytes, sizeof(id));
    RTB_NCCL(n, n->CommInitRank(&c->comm, (int)c->numShards, id, (int)c->shardIndex));
    return RTB_OK;
}

int rtb_comm_init_all(RtbContext** ctxs, uint32_t count)
{
    if (!ctxs || count == 0) return fail(RTB_ERR_INVALID_ARG, "null argument");
    NcclApi* n = nccl_api();
    if (!n) return fail(RTB_ERR_UNSUPPORTED, "NCCL is not available (libnccl.so.2 could not be loaded)");
    std::vector<int> devices(count);
    for (uint32_t i = 0; i < count; ++i)
    {
        if (!ctxs[i] || ctxs[i]->numShards != count || ctxs[i]->shardIndex != i) return fail(RTB_ERR_INVALID_ARG, "rtb_comm_init_all: context %u must be shard %u of %u", i, i, count);
        if (ctxs[i]->comm) return fail(RTB_ERR_INVALID_ARG, "context %u already has a communicator", i);
        devices[i] = ctxs[i]->device;
        for (uint32_t k = 0; k < i; ++k) if (devices[k] == devices[i]) return fail(RTB_ERR_INVALID_ARG, "rtb_comm_init_all: contexts %u and %u share device %d", k, i, devices[i]);
    }
    std::vector<ncclComm_t> comms(count);
    RTB_NCCL(n, n->CommInitAll(comms.data(), (int)count, devices.data()));
    for (uint32_t i = 0; i < count; ++i) ctxs[i]->comm = comms[i];
    return RTB_OK;
}

int rtb_finish_frame(RtbContext* c, int which, void** outGatheredDevice)
{
    if (!c || !outGatheredDevice) return fail(RTB_ERR_INVALID_ARG, "null argument");
    if (!c->sum.ptr) return fail(RTB_ERR_INVALID_ARG, "viewport has no size");
    RTB_CUDA(cudaSetDevice(c->device));
    RTB_CUDA(join_slots(c));
    const size_t shardFloats = (size_t)c->shardSlots * 3;
    if (c->numShards == 1)
    {
        int rc = vcm_fold(c);
        if (rc != RTB_OK) return rc;
        *outGatheredDevice = which == 0 ? c->sum.ptr : c->sum2.ptr;
        return RTB_OK;
    }
    if (!c->comm) return fail(RTB_ERR_INVALID_ARG, "rtb_finish_frame: the context has no communicator (rtb_comm_init)");
    NcclApi* n = nccl_api();
    if (c->vcm.allocated && c->vcm.splatsPending)
    {
        // the camera connections of every shard's light paths, summed once per frame, then this shard's pixels into its film
        RTB_NCCL(n, n->GroupStart());
        RTB_NCCL(n, n->AllReduce(c->vcm.splat.ptr, c->vcm.splat.ptr, c->vcm.splat.count, ncclFloat, ncclSum, c->comm, c->stream));
        RTB_NCCL(n, n->AllReduce(c->vcm.splat2.ptr, c->vcm.splat2.ptr, c->vcm.splat2.count, ncclFloat, ncclSum, c->comm, c->stream));
        RTB_NCCL(n, n->GroupEnd());
        int rc = vcm_fold(c);
        if (rc != RTB_OK) return rc;
    }
    if (c->gathered.count != shardFloats * c->numShards) RTB_CUDA(c->gathered.alloc(shardFloats * c->numShards));
    RTB_NCCL(n, n->AllGather(which == 0 ? c->sum.ptr : c->sum2.ptr, c->gathered.ptr, shardFloats, ncclFloat, c->comm, c->stream));
    *outGatheredDevice = c->gathered.ptr;
    return RTB_OK;
}

// single process, one thread driving every device: the collectives of all contexts inside ONE NCCL group
int rtb_finish_frame_all(RtbContext** ctxs, uint32_t count, int which, void** outGatheredDevice)
{
    if (!ctxs || !outGatheredDevice || count == 0) return fail(RTB_ERR_INVALID_ARG, "null argument");
    if (count == 1) return rtb_finish_frame(ctxs[0], which, outGatheredDevice);
    NcclApi* n = nccl_api();
    if (!n) return fail(RTB_ERR_UNSUPPORTED, "NCCL is not available (libnccl.so.2 could not be loaded)");
    bool splats = false;
    for (uint32_t i = 0; i < count; ++i)
    {
        RtbContext* c = ctxs[i];
        if (!c || !c->comm || c->numShards != count || !c->sum.ptr) return fail(RTB_ERR_INVALID_ARG, "rtb_finish_frame_all: context %u is not a sized shard with a communicator", i);
        RTB_CUDA(cudaSetDevice(c->device));
        RTB_CUDA(join_slots(c));
        splats = splats || (c->vcm.allocated && c->vcm.splatsPending);
        const size_t total = (size_t)c->shardSlots * 3 * c->numShards;
        if (c->gathered.count != total) RTB_CUDA(c->gathered.alloc(total));
    }
    if (splats)
    {
        RTB_NCCL(n, n->GroupStart());
        for (uint32_t i = 0; i < count; ++i)
        {
            RtbContext* c = ctxs[i];
            if (!c->vcm.allocated) { n->GroupEnd(); return fail(RTB_ERR_INVALID_ARG, "rtb_finish_frame_all: context %u has rendered no VCM pass while others have", i); }
            RTB_NCCL(n, n->AllReduce(c->vcm.splat.ptr, c->vcm.splat.ptr, c->vcm.splat.count, ncclFloat, ncclSum, c->comm, c->stream));
            RTB_NCCL(n, n->AllReduce(c->vcm.splat2.ptr, c->vcm.splat2.ptr, c->vcm.splat2.count, ncclFloat, ncclSum, c->comm, c->stream));
        }
        RTB_NCCL(n, n->GroupEnd());
        for (uint32_t i = 0; i < count; ++i)
        {
            RTB_CUDA(cudaSetDevice(ctxs[i]->device));
            ctxs[i]->vcm.splatsPending = true;
            int rc = vcm_fold(ctxs[i]);
            if (rc != RTB_OK) return rc;
        }
    }
    RTB_NCCL(n, n->GroupStart());
    for (uint32_t i = 0; i < count; ++i)
    {
        RtbContext* c = ctxs[i];
        RTB_NCCL(n, n->AllGather(which == 0 ? c->sum.ptr : c->sum2.ptr, c->gathered.ptr, (size_t)c->shardSlots * 3, ncclFloat, c->comm, c->stream));
        outGatheredDevice[i] = c->gathered.ptr;
    }
    RTB_NCCL(n, n->GroupEnd());
    return RTB_OK;
}

int rtb_scatter_shard(RtbContext* c, const void* shardFilmDevice, int which, void* rasterDst)
{
    if (!c || !rasterDst) return fail(RTB_ERR_INVALID_ARG, "null argument");
    if (!c->sum.ptr) return fail(RTB_ERR_INVALID_ARG, "viewport has no size");
    RTB_CUDA(cudaSetDevice(c->device));
    RTB_CUDA(join_slots(c));
    if (!shardFilmDevice && c->numShards == 1) { int rc = vcm_fold(c); if (rc != RTB_OK) return rc; }
    const float* src = shardFilmDevice ? (const float*)shardFilmDevice : (which == 0 ? c->sum.ptr : c->sum2.ptr);
    if (c->rasterOrder.ptr && ((uintptr_t)rasterDst & 15u) == 0)
        k_scatter_shard_rows<<<c->numSMs * 4, 256, 0, c->stream>>>(src, (float4*)rasterDst, c->rasterOrder.ptr, c->rasterPix.ptr, (uint32_t)((size_t)c->numPixels * 3u / 4u));
    else
        k_scatter_shard<<<c->numSMs * 4, 256, 0, c->stream>>>(src, (float*)rasterDst, c->pixelList.ptr, c->numPixels);
    RTB_CUDA(cudaGetLastError());
    c->counters.numKernelLaunches++;
    return RTB_OK;
}

int rtb_scatter_shard_async(RtbContext* c, const void* shardFilmDevice, int which, void* rasterDst)
{
    if (!c || !rasterDst || !shardFilmDevice) return fail(RTB_ERR_INVALID_ARG, "null argument (the asynchronous scatter takes a film snapshot, not the live film)");
    if (!c->sum.ptr) return fail(RTB_ERR_INVALID_ARG, "viewport has no size");
    (void)which;
    RTB_CUDA(cudaSetDevice(c->device));
    { const int rc = ensure_copy_stream(c); if (rc != RTB_OK) return rc; }
    RTB_CUDA(join_slots(c));
    // the stores go over PCIe while the next wavefront renders: the kernel runs on the copy stream, after everything enqueued
    // so far (the wavefront whose k_film wrote the snapshot), with a small grid - it waits on the bus, not on the SMs
    RTB_CUDA(cudaEventRecord(c->scatterReady, c->stream));
    RTB_CUDA(cudaStreamWaitEvent(c->copyStream, c->scatterReady, 0));
    const float* src = (const float*)shardFilmDevice;
    if (c->rasterOrder.ptr && ((uintptr_t)rasterDst & 15u) == 0)
        k_scatter_shard_rows<<<c->numSMs, 256, 0, c->copyStream>>>(src, (float4*)rasterDst, c->rasterOrder.ptr, c->rasterPix.ptr, (uint32_t)((size_t)c->numPixels * 3u / 4u));
    else
        k_scatter_shard<<<c->numSMs, 256, 0, c->copyStream>>>(src, (float*)rasterDst, c->pixelList.ptr, c->numPixels);
    RTB_CUDA(cudaGetLastError());
    RTB_CUDA(cudaEventRecord(c->scatterDone, c->copyStream));
    c->scatterPending = true;
    c->counters.numKernelLaunches++;
    return RTB_OK;
}

int rtb_host_register(void* p, uint64_t bytes)
{
    if (!p || !bytes) return fail(RTB_ERR_INVALID_ARG, "null argument");
    RTB_CUDA(cudaHostRegister(p, (size_t)bytes, cudaHostRegisterPortable | cudaHostRegisterMapped));
    return RTB_OK;
}

int rtb_host_unregister(void* p)
{
    if (!p) return fail(RTB_ERR_INVALID_ARG, "null argument");
    RTB_CUDA(cudaHostUnregister(p));
    return RTB_OK;
}

void* rtb_host_device_pointer(void* p)
{
    void* d = nullptr;
    const cudaError_t e = cudaHostGetDevicePointer(&d, p, 0);
    if (e != cudaSuccess) { fail(RTB_ERR_CUDA, "cudaHostGetDevicePointer: %s", cudaGetErrorString(e)); return nullptr; }
    return d;
}

int rtb_set_vcm_params(RtbContext* c, const RtbVcmParams* params)
{
    if (!c || !params) return fail(RTB_ERR_INVALID_ARG, "null argument");
    // the asserts of VertexConnectionAndMerging::PreRender, VertexConnectionAndMerging.cpp:82-85
    if (params->maxPathLength == 0 || params->maxPathLength > RTB_VCM_MAX_PATH_LENGTH) return fail(RTB_ERR_INVALID_ARG, "maxPathLength %u out of range (1 .. %u)", params->maxPathLength, RTB_VCM_MAX_PATH_LENGTH);
    if (!(params->initialMergingRadius >= params->minMergingRadius) || !(params->minMergingRadius > 0.0f)) return fail(RTB_ERR_INVALID_ARG, "merging radii: initial >= min > 0 required");
    if (!(params->mergingRadiusMultiplier > 0.0f) || !(params->mergingRadiusMultiplier <= 1.0f)) return fail(RTB_ERR_INVALID_ARG, "mergingRadiusMultiplier must be in (0, 1]");
    c->vcm.params = *params;
    c->vcm.lastPassIndex = 0xFFFFFFFFu;
    return RTB_OK;
}

int rtb_vcm_splat_buffers(RtbContext* c, void** outSplat, void** outSplat2, uint64_t* outNumFloats)
{
    if (!c || !outSplat || !outSplat2 || !outNumFloats) return fail(RTB_ERR_INVALID_ARG, "null argument");
    if (!c->sum.ptr) return fail(RTB_ERR_INVALID_ARG, "viewport has no size");
    RTB_CUDA(cudaSetDevice(c->device));
    int rc = vcm_allocate(c);
    if (rc != RTB_OK) return rc;
    RTB_CUDA(join_slots(c));
    *outSplat = c->vcm.splat.ptr; *outSplat2 = c->vcm.splat2.ptr; *outNumFloats = (uint64_t)c->width * c->height * 3u;
    return RTB_OK;
}

int rtb_vcm_fold_splats(RtbContext* c)
{
    if (!c) return fail(RTB_ERR_INVALID_ARG, "null context");
    RTB_CUDA(cudaSetDevice(c->device));
    RTB_CUDA(join_slots(c));
    return vcm_fold(c);
}

int rtb_set_profiling(RtbContext* c, int enable)
{
    if (!c) return fail(RTB_ERR_INVALID_ARG, "null context");
    c->profiling = enable != 0;
    return RTB_OK;
}

int rtb_get_pass_timings(RtbContext* c, RtbPassTimings* out)
{
    if (!c || !out) return fail(RTB_ERR_INVALID_ARG, "null argument");
    *out = c->timings;
    return RTB_OK;
}

} // extern "C"
#endif  // !RTB_TU_TAIL
