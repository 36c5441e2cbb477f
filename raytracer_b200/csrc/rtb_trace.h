// rtb_trace.h — two-level BVH ray/scene intersection: closest hit (Scene::Traverse) and any hit
// (Scene::Traverse_Shadow).  Per-ray device functions; the wavefront kernels in rtb_kernels.cu call them with one
// ray per thread.
//
// Mirrors, operation for operation:
//   Intersect_BoxRay / Intersect_BoxRay_TwoSided / Intersect_TriangleRay   Core/Math/Geometry.h:57-189 (SSE branch)
//   GenericTraverse / GenericTraverse_Shadow                               Core/Traversal/Traversal_Single.h:15-179
//   Scene::Traverse / Traverse_Shadow / Traverse_Object(_Shadow) / Traverse_Leaf(_Shadow)  Core/Scene/Scene.cpp:128-261
//   IShape::Traverse / Traverse_Shadow                                     Core/Shapes/Shape.cpp:19-57
//   SphereShape / BoxShape / RectShape ::Intersect                         SphereShape.cpp:29-46, BoxShape.cpp:118-125, RectShape.cpp:32-49
//   MeshShape::Traverse_Leaf / Traverse_Leaf_Shadow                        Core/Shapes/MeshShape.cpp:134-207
//   LightSceneObject::Traverse / Traverse_Shadow, AreaLight::TestRayHit    SceneObject_Light.cpp:27-53, AreaLight.cpp:30-41
#pragma once

#include <type_traits>

#include "rtb_scene.h"

namespace rtb {

// HitPoint (Core/Traversal/HitPoint.h:14-51)
struct Hit
{
    uint32_t objectId;
    uint32_t subObjectId;
    float    distance;
    float    u, v;
};

struct TraceCounters { uint32_t boxTests, triTests; };

#define RTB_STACK_SIZE 64
#ifndef RTB_LDG256
#define RTB_LDG256 1
#endif
// RTB_SPLIT_FINITE = 1 compiles the node loops twice (finite-slab rays / generic test) instead of choosing the box test
// per iteration.  Measured on the bench (ms per 8 passes): k_traverse 20.4 -> 21.6, k_shadow 32.6 -> 36.2.  Kept off.
#ifndef RTB_SPLIT_FINITE
#define RTB_SPLIT_FINITE 0
#endif
#ifndef RTB_TRAV_VOTE
#define RTB_TRAV_VOTE 6        // > 0: the closest-hit node loop is left once fewer than this many lanes of the warp are still in it
#endif
#ifndef RTB_SHADOW_VOTE
#define RTB_SHADOW_VOTE 6      // the same for the binary any-hit walk
#endif
#ifndef RTB_SHADOW_NEAR_FIRST
#define RTB_SHADOW_NEAR_FIRST 1
#endif      // BVH::MaxDepth is 128 (BVH.h:20); upload verifies the real depth of every tree fits

// Intersect_BoxRay, Geometry.h:57-86
RTB_HD bool box_ray(const Ray& ray, float minx, float miny, float minz, float maxx, float maxy, float maxz, float& outDistance)
{
    const float t1x = fma_(minx, ray.invDir.x, -ray.originDivDir.x);
    const float t1y = fma_(miny, ray.invDir.y, -ray.originDivDir.y);
    const float t1z = fma_(minz, ray.invDir.z, -ray.originDivDir.z);
    const float t2x = fma_(maxx, ray.invDir.x, -ray.originDivDir.x);
    const float t2y = fma_(maxy, ray.invDir.y, -ray.originDivDir.y);
    const float t2z = fma_(maxz, ray.invDir.z, -ray.originDivDir.z);
    const float lminx = sse_min(t1x, t2x), lminy = sse_min(t1y, t2y), lminz = sse_min(t1z, t2z);
    const float lmaxx = sse_max(t1x, t2x), lmaxy = sse_max(t1y, t2y), lmaxz = sse_max(t1z, t2z);
    const float tNear = sse_max(lminx, sse_max(lminy, lminz));
    const float tFar = sse_min(lmaxx, sse_min(lmaxy, lmaxz));
    outDistance = tNear;
    return (tFar >= tNear) && (tFar >= 0.0f);
}

// Same test for a ray whose invDir and originDivDir are all finite.  Then no slab distance can be NaN (an fma of
// finite operands is finite or +-inf), and on NaN-free inputs _mm_min_ps / _mm_max_ps differ from fminf / fmaxf
// only in the sign of a zero result, which none of the consumers (>=, <) can see.  fminf / fmaxf are single FMNMX
// instructions; the select form above costs a compare + select each.
RTB_HD bool box_ray_finite(const Ray& ray, float minx, float miny, float minz, float maxx, float maxy, float maxz, float& outDistance)
{
    const float t1x = fma_(minx, ray.invDir.x, -ray.originDivDir.x);
    const float t1y = fma_(miny, ray.invDir.y, -ray.originDivDir.y);
    const float t1z = fma_(minz, ray.invDir.z, -ray.originDivDir.z);
    const float t2x = fma_(maxx, ray.invDir.x, -ray.originDivDir.x);
    const float t2y = fma_(maxy, ray.invDir.y, -ray.originDivDir.y);
    const float t2z = fma_(maxz, ray.invDir.z, -ray.originDivDir.z);
    const float tNear = fmaxf(fminf(t1x, t2x), fmaxf(fminf(t1y, t2y), fminf(t1z, t2z)));
    const float tFar = fminf(fmaxf(t1x, t2x), fminf(fmaxf(t1y, t2y), fmaxf(t1z, t2z)));
    outDistance = tNear;
    return (tFar >= tNear) && (tFar >= 0.0f);
}

RTB_HD bool is_finite_bits(float x) { return (f2bits(x) & 0x7F800000u) != 0x7F800000u; }
RTB_HD bool ray_is_finite(const Ray& r)
{
    return is_finite_bits(r.invDir.x) && is_finite_bits(r.invDir.y) && is_finite_bits(r.invDir.z) &&
           is_finite_bits(r.originDivDir.x) && is_finite_bits(r.originDivDir.y) && is_finite_bits(r.originDivDir.z);
}

// Intersect_BoxRay_TwoSided, Geometry.h:100-130
RTB_HD bool box_ray_two_sided(const Ray& ray, const float* size, float& outNear, float& outFar)
{
    // Box(-mSize, mSize): operator-() is 0 - v
    const float minx = 0.0f - size[0], miny = 0.0f - size[1], minz = 0.0f - size[2];
    const float t1x = fma_(minx, ray.invDir.x, -ray.originDivDir.x);
    const float t1y = fma_(miny, ray.invDir.y, -ray.originDivDir.y);
    const float t1z = fma_(minz, ray.invDir.z, -ray.originDivDir.z);
    const float t2x = fma_(size[0], ray.invDir.x, -ray.originDivDir.x);
    const float t2y = fma_(size[1], ray.invDir.y, -ray.originDivDir.y);
    const float t2z = fma_(size[2], ray.invDir.z, -ray.originDivDir.z);
    const float lminx = sse_min(t1x, t2x), lminy = sse_min(t1y, t2y), lminz = sse_min(t1z, t2z);
    const float lmaxx = sse_max(t1x, t2x), lmaxy = sse_max(t1y, t2y), lmaxz = sse_max(t1z, t2z);
    outNear = sse_max(lminx, sse_max(lminy, lminz));
    outFar = sse_min(lmaxx, sse_min(lmaxy, lmaxz));
    return outNear < outFar;
}

// Intersect_TriangleRay (SSE branch), Geometry.h:132-167
RTB_HD bool triangle_ray(const Ray& ray, const V3& v0, const V3& e1, const V3& e2, float& outU, float& outV, float& outT)
{
    const V3 tvec = ray.origin - v0;
    const V3 pvec = cross3(ray.dir, e2);
    const V3 qvec = cross3(tvec, e1);
    const float det = dot3(e1, pvec);
    const float u = dot3(tvec, pvec);
    const float v = dot3(ray.dir, qvec);
    const float t = dot3(e2, qvec);
    const float upv = (u + v) / det;
    outU = u / det;
    outV = v / det;
    outT = t / det;
    // mask 0xE: (u+v > 1) false, v > 0, t > 0, u > 0
    return !(upv > 1.0f) && (outV > 0.0f) && (outT > 0.0f) && (outU > 0.0f);
}

// The same test as "triangle_ray(...) && outT < tLimit" with the four IEEE divisions (a subroutine each on the device)
// taken only by candidates: the numerators are the reference's, and a quotient can only be positive if numerator and
// determinant agree in sign, can only stay <= 1 if |u + v| <= |det| up to its rounding, can only stay below tLimit if
// |t| < tLimit * |det| up to its rounding.  Everything that survives this division-free screen takes the reference's
// divisions and comparisons, so the result is the reference's bit for bit; what the screen rejects the reference rejects.
RTB_HD bool triangle_ray_below(const Ray& ray, const V3& v0, const V3& e1, const V3& e2, float tLimit, float& outU, float& outV, float& outT)
{
    const V3 tvec = ray.origin - v0;
    const V3 pvec = cross3(ray.dir, e2);
    const V3 qvec = cross3(tvec, e1);
    const float det = dot3(e1, pvec);
    const float u = dot3(tvec, pvec);
    const float v = dot3(ray.dir, qvec);
    const float t = dot3(e2, qvec);
    const uint32_t sign = f2bits(det) & 0x80000000u;
    const float ad = bits2f(f2bits(det) ^ sign);                          // |det|
    const float su = bits2f(f2bits(u) ^ sign), sv = bits2f(f2bits(v) ^ sign), st = bits2f(f2bits(t) ^ sign);
    const float slack = 1.00000095367431640625f;                          // 1 + 2^-20
    if (!(su > 0.0f && sv > 0.0f && st > 0.0f && bit_abs(u + v) <= ad * slack && st < (tLimit * ad) * slack)) return false;
    const float upv = (u + v) / det;
    outU = u / det;
    outV = v / det;
    outT = t / det;
    return !(upv > 1.0f) && (outV > 0.0f) && (outT > 0.0f) && (outU > 0.0f) && (outT < tLimit);
}

// "triangle_ray(...) && outT < tLimit" as a boolean only (any-hit rays).  On top of the screen above: when every one of the
// reference's comparisons is decided with room to spare the quotients need not be formed at all -
//   u / det > 0 (FTZ division): numerator and determinant agree in sign and the quotient cannot underflow (>= 2^-100);
//   !((u + v) / det > 1):       |fl(u + v)| <= |det| * (1 - 2^-20): the quotient is below 1 before and after rounding;
//   t / det < tLimit:           |t| <= fl(fl(tLimit * |det|) * (1 - 2^-20)) with a finite product: the real quotient is
//                               below tLimit * (1 - 2^-21), so its rounding is below tLimit.
// Candidates inside those margins (hits within 2^-20 of an edge or of tLimit) take the reference's divisions.
RTB_HD bool triangle_ray_occludes(const Ray& ray, const V3& v0, const V3& e1, const V3& e2, float tLimit)
{
    const V3 tvec = ray.origin - v0;
    const V3 pvec = cross3(ray.dir, e2);
    const V3 qvec = cross3(tvec, e1);
    const float det = dot3(e1, pvec);
    const float u = dot3(tvec, pvec);
    const float v = dot3(ray.dir, qvec);
    const float t = dot3(e2, qvec);
    const uint32_t sign = f2bits(det) & 0x80000000u;
    const float ad = bits2f(f2bits(det) ^ sign);                          // |det|
    const float su = bits2f(f2bits(u) ^ sign), sv = bits2f(f2bits(v) ^ sign), st = bits2f(f2bits(t) ^ sign);
    const float slack = 1.00000095367431640625f;                          // 1 + 2^-20
    const float sum = bit_abs(u + v), lim = tLimit * ad;
    if (!(su > 0.0f && sv > 0.0f && st > 0.0f && sum <= ad * slack && st < lim * slack)) return false;
    const float tight = 0.99999904632568359375f;                          // 1 - 2^-20
    const float tiny = ad * 7.8886090522101181e-31f;                      // 2^-100 (flushes to 0 for |det| < 2^-26: then any normal numerator is far enough)
    if (su >= tiny && sv >= tiny && st >= tiny && sum <= ad * tight && is_finite_bits(lim) && st <= lim * tight) return true;
    const float upv = (u + v) / det;
    const float outU = u / det, outV = v / det, outT = t / det;
    return !(upv > 1.0f) && (outV > 0.0f) && (outT > 0.0f) && (outU > 0.0f) && (outT < tLimit);
}

struct ShapeIntersection { float nearDist, farDist; uint32_t subObjectId; };

// SphereShape::Intersect (double precision quadratic), SphereShape.cpp:29-46
RTB_HD bool sphere_intersect(const Ray& ray, float radius, ShapeIntersection& out)
{
    const double radiusD = (double)radius;
    const double v = (double)dot3(ray.dir, neg(ray.origin));
    const double det = radiusD * radiusD - (double)sqr_length3(ray.origin) + v * v;
    if (det <= 0.0) return false;
    const double sqrtDet = sqrt(det);
    out.nearDist = (float)(v - sqrtDet);
    out.farDist = (float)(v + sqrtDet);
    out.subObjectId = 0;
    return out.farDist > out.nearDist;
}

// RectShape::Intersect, RectShape.cpp:32-49
RTB_HD bool rect_intersect(const Ray& ray, const float* size, ShapeIntersection& out)
{
    const float t = -ray.origin.z * ray.invDir.z;
    if (t > FLT_EPSILON)
    {
        const float px = fma_(ray.dir.x, t, ray.origin.x);
        const float py = fma_(ray.dir.y, t, ray.origin.y);
        if (bit_abs(px) < size[0] && bit_abs(py) < size[1])
        {
            out.nearDist = t;
            out.farDist = t;
            return true;
        }
    }
    return false;
}

RTB_HD bool shape_intersect(uint32_t shapeType, const float* size, const Ray& ray, ShapeIntersection& out)
{
    out.subObjectId = 0xFFFFFFFFu;   // ShapeIntersection default, Shape.h:25-30
    if (shapeType == RTB_SHAPE_SPHERE) return sphere_intersect(ray, size[0], out);
    if (shapeType == RTB_SHAPE_BOX) { out.subObjectId = 0; return box_ray_two_sided(ray, size, out.nearDist, out.farDist); }
    if (shapeType == RTB_SHAPE_RECT) return rect_intersect(ray, size, out);
    return false;
}

#ifndef RTB_PLAIN_LOADS
#define RTB_PLAIN_LOADS 0     // 1: ld.global instead of ld.global.nc for nodes and triangles (experiment)
#endif
#if RTB_PLAIN_LOADS
#define RTB_LDG(p) (*(p))
#define RTB_LD_NC "ld.global"
#else
#define RTB_LDG(p) __ldg(p)
#define RTB_LD_NC "ld.global.nc"
#endif

RTB_HD void load_tri(const DTri* tris, uint32_t index, V3& v0, V3& e1, V3& e2)
{
#if defined(__CUDA_ARCH__)
    const float4* p = reinterpret_cast<const float4*>(tris + index);
    const float4 a = RTB_LDG(p), b = RTB_LDG(p + 1), c = RTB_LDG(p + 2);
    v0 = v3(a.x, a.y, a.z); e1 = v3(b.x, b.y, b.z); e2 = v3(c.x, c.y, c.z);
#else
    const DTri& t = tris[index];
    v0 = v3(t.v0x, t.v0y, t.v0z); e1 = v3(t.e1x, t.e1y, t.e1z); e2 = v3(t.e2x, t.e2y, t.e2z);
#endif
}

RTB_HD DNode load_node(const DNode* nodes, uint32_t index)
{
#if defined(__CUDA_ARCH__) && RTB_LDG256
    // one 256-bit load per node (LDG.E.ENL2.256, sm_100+): the traversal kernels are bound by L1 data-pipe wavefronts
    // (70-80 % of peak in ncu), and a divergent lane costs one wavefront per load instruction whatever its width
    DNode n;
    asm(RTB_LD_NC ".v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=f"(n.minx), "=f"(n.miny), "=f"(n.minz), "=r"(n.childIndex), "=f"(n.maxx), "=f"(n.maxy), "=f"(n.maxz), "=r"(n.leafInfo)
        : "l"(nodes + index));
    return n;
#elif defined(__CUDA_ARCH__)
    const float4* p = reinterpret_cast<const float4*>(nodes + index);
    const float4 a = RTB_LDG(p), b = RTB_LDG(p + 1);
    DNode n;
    n.minx = a.x; n.miny = a.y; n.minz = a.z; n.childIndex = __float_as_uint(a.w);
    n.maxx = b.x; n.maxy = b.y; n.maxz = b.z; n.leafInfo = __float_as_uint(b.w);
    return n;
#else
    return nodes[index];
#endif
}

// both children of an inner node: they sit in one aligned 64-byte block, so ONE address + an immediate offset
RTB_HD void load_node_pair(const DNode* nodes, uint32_t index, DNode& a, DNode& b)
{
#if defined(__CUDA_ARCH__) && RTB_LDG256
    asm(RTB_LD_NC ".v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%16];\n\t"
        RTB_LD_NC ".v8.f32 {%8,%9,%10,%11,%12,%13,%14,%15}, [%16+32];"
        : "=f"(a.minx), "=f"(a.miny), "=f"(a.minz), "=r"(a.childIndex), "=f"(a.maxx), "=f"(a.maxy), "=f"(a.maxz), "=r"(a.leafInfo),
          "=f"(b.minx), "=f"(b.miny), "=f"(b.minz), "=r"(b.childIndex), "=f"(b.maxx), "=f"(b.maxy), "=f"(b.maxz), "=r"(b.leafInfo)
        : "l"(nodes + index));
#else
    a = load_node(nodes, index);
    b = load_node(nodes, index + 1);
#endif
}


// IShape::Traverse (Shape.cpp:19-45) / LightSceneObject::Traverse (SceneObject_Light.cpp:27-38) for analytic shapes
RTB_HD void analytic_closest(const DObjectTrace& obj, const Ray& ray, uint32_t objectID, Hit& hit)
{
    if (obj.shapeType == RTB_NO_SHAPE) return;          // point / spot lights: TestRayHit is always false
    ShapeIntersection isect;
    if (!shape_intersect(obj.shapeType, obj.size, ray, isect)) return;
    if (obj.kind == RTB_OBJECT_SHAPE)
    {
        if (isect.nearDist > 0.0f && isect.nearDist < hit.distance)
        {
            hit.distance = isect.nearDist; hit.objectId = objectID; hit.subObjectId = isect.subObjectId;
            return;
        }
        if (isect.farDist > 0.0f && isect.farDist < hit.distance)
        {
            hit.distance = isect.farDist; hit.objectId = objectID; hit.subObjectId = isect.subObjectId;
        }
    }
    else
    {
        const float lightDistance = isect.nearDist;     // AreaLight::TestRayHit, AreaLight.cpp:30-41
        if (lightDistance > 0.0f && lightDistance < hit.distance)
        {
            hit.distance = lightDistance; hit.objectId = objectID; hit.subObjectId = RTB_LIGHT_OBJECT;
        }
    }
}

// IShape::Traverse_Shadow (Shape.cpp:47-57) / LightSceneObject::Traverse_Shadow (SceneObject_Light.cpp:40-53)
RTB_HD bool analytic_shadow(const DObjectTrace& obj, const Ray& ray, float maxDistance)
{
    if (obj.shapeType == RTB_NO_SHAPE) return false;
    ShapeIntersection isect;
    if (!shape_intersect(obj.shapeType, obj.size, ray, isect)) return false;
    if (obj.kind == RTB_OBJECT_SHAPE) return isect.farDist > 0.0f && isect.nearDist < maxDistance;
    return isect.nearDist < maxDistance;
}

// One 32-bit word says what a node visit needs: index (26 bits) | count (6 bits).  count 0 = inner node (index of its
// first child), 1..62 = leaf with that many items starting at index, 63 = "traverse object `index`" (stack only).
// rtb_upload_scene refuses trees that do not fit (>= 2^26 nodes / items, leaves with more than 62 items) and stores that
// word in the `leafInfo` slot of every node of the DEVICE copy (build_host_scene, rtb_upload.h: the reference's
// numLeaves : 30 | splitAxis : 2 word is not needed by any walk), so the walks carry ONE register per pending node, a
// push is a plain store of the far child's word and a pop a plain load - no pack / unpack arithmetic on the way (push and
// pop were 13 % of k_traverse_p's warp instructions at 3 active lanes).  A node's box is only needed when its parent is
// visited (both children sit in one aligned 64-byte block), so neither a popped node nor the nearer child we descend
// into is ever fetched again: memory traffic is exactly one 64-byte block per visited inner node.
//
// On the device the whole stack lives in SHARED memory (RTB_SMEM_STACK = 1, the default): entry i of thread t at
// sstack[i * RTB_STACK_THREADS + t], so whatever depth each lane of a warp is at, the 32 lanes touch 32 different
// banks (one conflict-free wavefront per push / pop; a local-memory stack costs one L1 wavefront per distinct depth
// in the warp), and the kernels are launched with `stackEntries * 512` bytes of dynamic shared memory, stackEntries =
// the scene's worst-case stack occupancy computed at upload.  Measured on the bench against the local-memory stack:
// k_shadow 2.36 -> 2.31 ms per launch, k_traverse unchanged.  RTB_SMEM_STACK = 0 keeps the local-memory stack (the
// host build of these headers always uses a local array).
#ifndef RTB_SMEM_STACK
#define RTB_SMEM_STACK 1
#endif
#define RTB_STACK_THREADS 128
#if defined(__CUDA_ARCH__) && RTB_SMEM_STACK > 0
#define RTB_SSTACK 1
#define RTB_LOCAL_STACK_SIZE 1
#else
#define RTB_SSTACK 0
#define RTB_LOCAL_STACK_SIZE RTB_STACK_SIZE
#endif
#define RTB_PACK_INDEX_BITS 26u
#define RTB_PACK_MAX_INDEX ((1u << RTB_PACK_INDEX_BITS) - 1u)
#define RTB_PACK_OBJECT 63u
#define RTB_PACK_MAX_LEAF 62u
#define RTB_NODE_COUNT(p) ((p) >> RTB_PACK_INDEX_BITS)
#define RTB_NODE_INDEX(p) ((p) & RTB_PACK_MAX_INDEX)
#define RTB_OBJECT_ENTRY(i) ((i) | (RTB_PACK_OBJECT << RTB_PACK_INDEX_BITS))
RTB_HD uint32_t pack_node_word(uint32_t childIndex, uint32_t referenceLeafInfo) { return childIndex | ((referenceLeafInfo & 0x3FFFFFFFu) << RTB_PACK_INDEX_BITS); }

// the kernels hand the stack over as a 32-bit shared-space byte address (word 0 of this thread's column) behind an
// optimisation barrier: with a generic pointer the compiler re-derived the address (thread id, CTA shared window:
// 5 instructions incl. S2R / S2UR) in every iteration of the node loop instead of keeping it in a register
typedef uint32_t SharedStack;
RTB_HD void stk_put(uint32_t* stack, SharedStack sstack, uint32_t i, uint32_t packed)
{
#if RTB_SSTACK
    (void)stack;
    asm volatile("st.shared.u32 [%0], %1;" :: "r"(sstack + i * (RTB_STACK_THREADS * 4u)), "r"(packed));
#else
    (void)sstack;
    stack[i] = packed;
#endif
}

RTB_HD uint32_t stk_get(const uint32_t* stack, SharedStack sstack, uint32_t i)
{
#if RTB_SSTACK
    (void)stack;
    uint32_t packed;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(packed) : "r"(sstack + i * (RTB_STACK_THREADS * 4u)));
    return packed;
#else
    (void)sstack;
    return stack[i];
#endif
}

// the packed word of node `index` itself (the root, when a walk starts)
RTB_HD uint32_t node_header(const DNode* nodes, uint32_t index) { return nodes[index].leafInfo; }

// Scene::Traverse (Scene.cpp:219-243) = GenericTraverse<Scene> -> Scene::Traverse_Leaf -> Traverse_Object ->
// GenericTraverse<MeshShape> -> MeshShape::Traverse_Leaf, flattened into ONE loop over ONE stack so that all lanes of
// a warp execute the same instruction stream whatever level of the two-level BVH they are in (the nested form ran
// at 2-8 active lanes per warp instruction).  The visiting order is exactly the reference's: near child first,
// objects of a top-level leaf one after the other, each mesh traversed to completion before the next object.
//
// `hit` must be preset (objectId = RTB_INVALID_OBJECT, distance = +inf).  The top-level box tests use
// ray.originDivDir exactly as stored: the integrator offsets ray.origin AFTER constructing the ray without
// refreshing originDivDir (PathTracerMIS.cpp:392-393) and that stale value is part of the reference's behaviour;
// Traverse_Object rebuilds the ray in object space.
// Where the world-space ray comes from.  The walk needs it three times: at the start, for every object it enters
// (Traverse_Object transforms the WORLD ray) and when it returns from a mesh to the top level.  Holding it in
// registers across a whole mesh traversal makes ptxas spill it to local memory at 64 registers per thread; the
// kernels therefore hand over a source that simply loads it again from the ray queue (three L1-resident 16-byte
// loads per object, against a dozen local stores and loads), and the traversal kernels end up free of long-lived spills.
struct RayInRegisters
{
    const Ray& r;
    RTB_HD Ray load() const { return r; }
};

template <bool kCount, typename WorldRay>
RTB_HD void scene_traverse_from(const SceneView& S, const WorldRay& worldRaySource, Hit& hit, TraceCounters* cnt, SharedStack sstack = 0);

template <bool kCount>
RTB_HD void scene_traverse(const SceneView& S, const Ray& worldRay, Hit& hit, TraceCounters* cnt, SharedStack sstack = 0)
{
    const RayInRegisters src = { worldRay };
    scene_traverse_from<kCount>(S, src, hit, cnt, sstack);
}

template <bool kCount, typename WorldRay>
RTB_HD void scene_traverse_from(const SceneView& S, const WorldRay& worldRaySource, Hit& hit, TraceCounters* cnt, SharedStack sstack)
{
    if (S.numObjects == 0) return;
    uint32_t stack[RTB_LOCAL_STACK_SIZE];
    uint32_t stackSize = 0;
    uint32_t blasBase = 0;
    uint32_t popFloor = 0xFFFFFFFFu;     // entries above it are plain nodes of the mesh being traversed
    bool inBlas = false;
    const DNode* nodes = S.tlas;
    const DTri* tris = nullptr;
    Ray ray = worldRaySource.load();
    const bool worldFinite = ray_is_finite(ray);
    bool finite = worldFinite;
    uint32_t objectID = 0;
    bool have = false;                 // `cur` (packed word) describes a node still to be visited
    uint32_t cur = 0;

    if (S.numObjects == 1) { stk_put(stack, sstack, 0, RTB_OBJECT_ENTRY(0u)); stackSize = 1; }      // BVH bypass, Scene.cpp:231-235
    else if (S.numTlasNodes != 0) { cur = node_header(S.tlas, 0); have = true; }

    for (;;)
    {
        // inner nodes: GenericTraverse, Traversal_Single.h:36-91 (kMode as in ShadowWalk::node_phase)
        auto innerNodes = [&](auto modeTag)
        {
            constexpr int kMode = decltype(modeTag)::value;
            while (have && RTB_NODE_COUNT(cur) == 0)
            {
                DNode a, b;
                load_node_pair(nodes, cur, a, b);       // (an inner node's word is the index of its first child)
                float distanceA, distanceB;
                bool hitA, hitB;
                if (kMode == 1 || (kMode == 0 && finite))
                {
                    hitA = box_ray_finite(ray, a.minx, a.miny, a.minz, a.maxx, a.maxy, a.maxz, distanceA);
                    hitB = box_ray_finite(ray, b.minx, b.miny, b.minz, b.maxx, b.maxy, b.maxz, distanceB);
                }
                else
                {
                    hitA = box_ray(ray, a.minx, a.miny, a.minz, a.maxx, a.maxy, a.maxz, distanceA);
                    hitB = box_ray(ray, b.minx, b.miny, b.minz, b.maxx, b.maxy, b.maxz, distanceB);
                }
                hitA &= (distanceA < hit.distance);
                hitB &= (distanceB < hit.distance);
                if (kCount) cnt->boxTests += 2;
                if (hitA && hitB)
                {
                    // visit the nearer child first, push the other (Traversal_Single.h:65-75)
                    const bool bFirst = distanceB < distanceA;
                    stk_put(stack, sstack, stackSize++, bFirst ? a.leafInfo : b.leafInfo);
                    cur = bFirst ? b.leafInfo : a.leafInfo;
                }
                else if (hitA) cur = a.leafInfo;
                else if (hitB) cur = b.leafInfo;
                else if (stackSize > popFloor)
                {
                    // both children missed: take the next pending node of this mesh right here, so that the lane stays in
                    // the inner loop with the rest of its warp instead of idling until every lane has left it
                    cur = stk_get(stack, sstack, --stackSize);
                }
                else have = false;
#if defined(__CUDA_ARCH__) && RTB_TRAV_VOTE > 0
                // the lanes that have reached a leaf wait for this loop to end: leave it once only a few are still
                // walking nodes (each lane's own sequence of visits is unchanged, only the interleaving between lanes)
                if (__popc(__activemask()) < RTB_TRAV_VOTE) break;
#endif
            }
        };
#if RTB_SPLIT_FINITE
        if (finite) innerNodes(std::integral_constant<int, 1>()); else innerNodes(std::integral_constant<int, 2>());
#else
        innerNodes(std::integral_constant<int, 0>());
#endif

#if RTB_TRAV_VOTE > 0
        if (have && RTB_NODE_COUNT(cur) == 0) continue;       // left the node loop by vote: carry on with the same node
#endif
        if (have)
        {
            const uint32_t numLeaves = RTB_NODE_COUNT(cur), firstItem = RTB_NODE_INDEX(cur);
            if (inBlas)
            {
                // MeshShape::Traverse_Leaf, MeshShape.cpp:134-168
                if (kCount) cnt->triTests += numLeaves;
                for (uint32_t i = 0; i < numLeaves; ++i)
                {
                    const uint32_t triangleIndex = firstItem + i;
                    V3 v0, e1, e2;
                    load_tri(tris, triangleIndex, v0, e1, e2);
                    float u, v, distance;
                    if (triangle_ray_below(ray, v0, e1, e2, hit.distance, u, v, distance))       // Intersect_TriangleRay && distance < hitPoint.distance
                    {
                        hit.distance = distance;
                        hit.subObjectId = triangleIndex;
                        hit.objectId = objectID;
                        hit.u = u;
                        hit.v = v;
                    }
                }
            }
            else
            {
                // Scene::Traverse_Leaf, Scene.cpp:167-178: objects in order -> push in reverse
                for (uint32_t i = numLeaves; i-- > 0;) { stk_put(stack, sstack, stackSize, RTB_OBJECT_ENTRY(firstItem + i)); stackSize++; }
            }
            have = false;
        }

        // pop
        if (inBlas && stackSize == blasBase)
        {
            // mesh finished: back to the top level with the world-space ray
            inBlas = false; popFloor = 0xFFFFFFFFu; nodes = S.tlas; ray = worldRaySource.load(); finite = worldFinite;
        }
        if (stackSize == 0) break;
        const uint32_t e = stk_get(stack, sstack, --stackSize);
        if (RTB_NODE_COUNT(e) == RTB_PACK_OBJECT)
        {
            // Scene::Traverse_Object, Scene.cpp:128-145
            objectID = RTB_NODE_INDEX(e);
            const DObjectTrace& obj = S.objTrace[objectID];
            const Ray localRay = transform_ray_unsafe(obj.invTransform, worldRaySource.load());
            if (obj.kind == RTB_OBJECT_SHAPE && obj.shapeType == RTB_SHAPE_MESH)
            {
                const DMesh& mesh = S.meshes[obj.meshIndex];
                if (mesh.numNodes != 0)
                {
                    inBlas = true; blasBase = stackSize; popFloor = stackSize; nodes = mesh.nodes; tris = mesh.tris; ray = localRay; finite = ray_is_finite(localRay);
                    cur = node_header(nodes, 0);
                    have = true;
                }
            }
            else
            {
                analytic_closest(obj, localRay, objectID, hit);
            }
        }
        else
        {
            cur = e; have = true;
        }
    }
}

// scene_traverse_from as a resumable walk (begin() + step() until it returns false) for the persistent closest-hit kernel
// (k_traverse_p), whose warps give a finished lane a new ray instead of letting it idle until the slowest ray of the
// warp is done.  One step() = what one iteration of scene_traverse_from's outer loop does, in the same order: node
// phase, leaf (triangle tests / top-level leaf), pop.  A lane's own sequence of box and triangle tests is exactly the
// one above - hit records are bit-identical - only the interleaving between lanes differs.
struct TraverseWalk
{
    uint32_t stackSize, blasBase, popFloor, objectID, cur;
    bool inBlas, finite, worldFinite, have;
    const DNode* nodes;
    const DTri* tris;
    Ray ray;
    Hit hit;

    // false: nothing to traverse
    template <typename WorldRay>
    RTB_HD bool begin(const SceneView& S, uint32_t* stack, const WorldRay& worldRaySource, SharedStack sstack = 0)
    {
        hit.objectId = RTB_INVALID_OBJECT; hit.subObjectId = 0; hit.distance = INFINITY; hit.u = 0.0f; hit.v = 0.0f;
        if (S.numObjects == 0) return false;
        stackSize = 0; blasBase = 0; popFloor = 0xFFFFFFFFu; inBlas = false; nodes = S.tlas; tris = nullptr;
        ray = worldRaySource.load();
        worldFinite = ray_is_finite(ray); finite = worldFinite;
        objectID = 0; have = false; cur = 0;
        if (S.numObjects == 1) { stk_put(stack, sstack, 0, RTB_OBJECT_ENTRY(0u)); stackSize = 1; }
        else if (S.numTlasNodes != 0) { cur = node_header(S.tlas, 0); have = true; }
        else return false;
        return true;
    }

    // false: the walk is finished (hit holds the result)
    template <bool kCount, typename WorldRay>
    RTB_HD bool step(const SceneView& S, uint32_t* stack, const WorldRay& worldRaySource, TraceCounters* cnt, SharedStack sstack = 0)
    {
        while (have && RTB_NODE_COUNT(cur) == 0)
        {
            DNode a, b;
            load_node_pair(nodes, cur, a, b);       // (an inner node's word is the index of its first child)
            float distanceA, distanceB;
            bool hitA, hitB;
            if (finite)
            {
                hitA = box_ray_finite(ray, a.minx, a.miny, a.minz, a.maxx, a.maxy, a.maxz, distanceA);
                hitB = box_ray_finite(ray, b.minx, b.miny, b.minz, b.maxx, b.maxy, b.maxz, distanceB);
            }
            else
            {
                hitA = box_ray(ray, a.minx, a.miny, a.minz, a.maxx, a.maxy, a.maxz, distanceA);
                hitB = box_ray(ray, b.minx, b.miny, b.minz, b.maxx, b.maxy, b.maxz, distanceB);
            }
            hitA &= (distanceA < hit.distance);
            hitB &= (distanceB < hit.distance);
            if (kCount) cnt->boxTests += 2;
            if (hitA && hitB)
            {
                const bool bFirst = distanceB < distanceA;
                stk_put(stack, sstack, stackSize++, bFirst ? a.leafInfo : b.leafInfo);
                cur = bFirst ? b.leafInfo : a.leafInfo;
            }
            else if (hitA) cur = a.leafInfo;
            else if (hitB) cur = b.leafInfo;
            else if (stackSize > popFloor)
            {
                cur = stk_get(stack, sstack, --stackSize);
            }
            else have = false;
#if defined(__CUDA_ARCH__) && RTB_TRAV_VOTE > 0
            if (__popc(__activemask()) < RTB_TRAV_VOTE) break;
#endif
        }
        if (have && RTB_NODE_COUNT(cur) == 0) return true;         // left the node loop by vote
        if (have)
        {
            const uint32_t numLeaves = RTB_NODE_COUNT(cur), firstItem = RTB_NODE_INDEX(cur);
            if (inBlas)
            {
                if (kCount) cnt->triTests += numLeaves;
                for (uint32_t i = 0; i < numLeaves; ++i)
                {
                    const uint32_t triangleIndex = firstItem + i;
                    V3 v0, e1, e2;
                    load_tri(tris, triangleIndex, v0, e1, e2);
                    float u, v, distance;
                    if (triangle_ray_below(ray, v0, e1, e2, hit.distance, u, v, distance))
                    {
                        hit.distance = distance; hit.subObjectId = triangleIndex; hit.objectId = objectID; hit.u = u; hit.v = v;
                    }
                }
            }
            else
            {
                for (uint32_t i = numLeaves; i-- > 0;) { stk_put(stack, sstack, stackSize, RTB_OBJECT_ENTRY(firstItem + i)); stackSize++; }
            }
            have = false;
        }
        if (inBlas && stackSize == blasBase)
        {
            inBlas = false; popFloor = 0xFFFFFFFFu; nodes = S.tlas; ray = worldRaySource.load(); finite = worldFinite;
        }
        if (stackSize == 0) return false;
        const uint32_t e = stk_get(stack, sstack, --stackSize);
        if (RTB_NODE_COUNT(e) == RTB_PACK_OBJECT)
        {
            objectID = RTB_NODE_INDEX(e);
            const DObjectTrace& obj = S.objTrace[objectID];
            const Ray localRay = transform_ray_unsafe(obj.invTransform, worldRaySource.load());
            if (obj.kind == RTB_OBJECT_SHAPE && obj.shapeType == RTB_SHAPE_MESH)
            {
                const DMesh& mesh = S.meshes[obj.meshIndex];
                if (mesh.numNodes != 0)
                {
                    inBlas = true; blasBase = stackSize; popFloor = stackSize; nodes = mesh.nodes; tris = mesh.tris; ray = localRay; finite = ray_is_finite(localRay);
                    cur = node_header(nodes, 0);
                    have = true;
                }
            }
            else
            {
                analytic_closest(obj, localRay, objectID, hit);
            }
        }
        else
        {
            cur = e; have = true;
        }
        return true;
    }
};

// stack of the wide any-hit walk (rtb_wide.h) on top of the entries the two-level walk already holds: a group is two
// consecutive 4-byte entries
struct WideStackAbove
{
    uint32_t* stack; SharedStack sstack; uint32_t floor, top;
    RTB_HD bool empty() const { return top == floor; }
    RTB_HD void push(uint32_t a, uint32_t b)
    {
#if RTB_SSTACK
        asm volatile("st.shared.u32 [%0], %1;" :: "r"(sstack + top * (RTB_STACK_THREADS * 4u)), "r"(a));
        asm volatile("st.shared.u32 [%0], %1;" :: "r"(sstack + (top + 1u) * (RTB_STACK_THREADS * 4u)), "r"(b));
#else
        stack[top] = a; stack[top + 1u] = b;
#endif
        top += 2u;
    }
    RTB_HD void pop(uint32_t& a, uint32_t& b)
    {
        top -= 2u;
#if RTB_SSTACK
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(a) : "r"(sstack + top * (RTB_STACK_THREADS * 4u)));
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(b) : "r"(sstack + (top + 1u) * (RTB_STACK_THREADS * 4u)));
#else
        a = stack[top]; b = stack[top + 1u];
#endif
    }
};

} // namespace rtb
#include "rtb_wide.h"      // (after the stack definitions it builds on)
namespace rtb {

#ifndef RTB_WIDE_SHADOW
#define RTB_WIDE_SHADOW 1      // 0: any-hit rays walk the reference's binary tree like closest-hit rays do
#endif

// Scene::Traverse_Shadow (Scene.cpp:245-261), same flattening, as a resumable walk: begin() + step() until done.
// The result is a boolean "any triangle / shape closer than maxDistance among the leaves whose ancestors pass the box
// test".  It does not depend on the visiting order, so this any-hit walk is free to (a) descend into the NEARER child
// first (the reference goes A then B, Traversal_Single.h:151-156) and (b) postpone the triangle tests of one leaf
// and keep walking nodes, which gives fewer and fuller leaf phases per warp.  One step() = one node phase (until the
// lane holds up to two leaves or runs out of nodes) + the leaf tests + one pop; k_shadow refills finished lanes
// between steps.
struct ShadowWalk
{
    enum { kContinue = 0, kOccluded = 1, kVisible = 2, kEnterWide = 3 };

    // (the stack itself is a separate local array handed to begin()/step(): with the array inside, the whole struct
    //  would live in local memory instead of registers)
    uint32_t stackSize, blasBase;
    uint32_t popFloor;                     // entries above it are plain nodes of the mesh being traversed
    bool inBlas, finite, worldFinite, have;
    const DNode* nodes;
    const DTri* tris;
    Ray ray;
    float maxDistance;
    uint32_t cur;
    uint32_t pend;                         // a mesh leaf whose triangle tests are postponed (its packed word; 0 = none)
    uint32_t wideMesh;                     // kEnterWide: the mesh whose 8-wide tree the caller walks with `ray` (mesh-local), then leave_wide()

    // false: nothing to traverse (the ray is unoccluded)
    template <typename WorldRay>
    RTB_HD bool begin(const SceneView& S, uint32_t* stack, const WorldRay& worldRaySource, float tMax, SharedStack sstack = 0)
    {
        if (S.numObjects == 0) return false;
        stackSize = 0; blasBase = 0; popFloor = 0xFFFFFFFFu; inBlas = false;
        nodes = S.tlas; tris = nullptr;
        ray = worldRaySource.load(); maxDistance = tMax;
        worldFinite = ray_is_finite(ray); finite = worldFinite;
        have = false; cur = 0; pend = 0;
        if (S.numObjects == 1) { stk_put(stack, sstack, 0, RTB_OBJECT_ENTRY(0u)); stackSize = 1; }
        else if (S.numTlasNodes != 0) { cur = node_header(S.tlas, 0); have = true; }
        else return false;
        return true;
    }

    template <bool kCount>
    RTB_HD bool leaf_occludes(uint32_t leaf, TraceCounters* cnt) const
    {
        // MeshShape::Traverse_Leaf_Shadow, MeshShape.cpp:175-207
        const uint32_t numLeaves = RTB_NODE_COUNT(leaf), first = RTB_NODE_INDEX(leaf);
        if (kCount) cnt->triTests += numLeaves;
        bool occluded = false;
        for (uint32_t i = 0; i < numLeaves; ++i)
        {
            V3 v0, e1, e2;
            load_tri(tris, first + i, v0, e1, e2);
            if (triangle_ray_occludes(ray, v0, e1, e2, maxDistance)) occluded = true;
        }
        return occluded;
    }

    // the node phase of step().  kMode 0: the finite / generic box test is chosen per iteration; 1 / 2: the phase is
    // compiled once for rays with finite slabs and once for the generic (SSE NaN semantics) test and step() picks one
    // when it enters the phase - `finite` only changes when a mesh is entered or left, never inside the phase.
    template <bool kCount, int kMode>
    RTB_HD void node_phase(uint32_t* stack, TraceCounters* cnt, SharedStack sstack)
    {
        while (have)
        {
            if (RTB_NODE_COUNT(cur) != 0)
            {
                if (inBlas && pend == 0 && stackSize > popFloor)
                {
                    pend = cur;
                    cur = stk_get(stack, sstack, --stackSize);
                    continue;
                }
                break;
            }
            DNode a, b;
            load_node_pair(nodes, cur, a, b);       // (an inner node's word is the index of its first child)
            float distanceA, distanceB;
            bool hitA, hitB;
            if (kMode == 1 || (kMode == 0 && finite))
            {
                hitA = box_ray_finite(ray, a.minx, a.miny, a.minz, a.maxx, a.maxy, a.maxz, distanceA);
                hitB = box_ray_finite(ray, b.minx, b.miny, b.minz, b.maxx, b.maxy, b.maxz, distanceB);
            }
            else
            {
                hitA = box_ray(ray, a.minx, a.miny, a.minz, a.maxx, a.maxy, a.maxz, distanceA);
                hitB = box_ray(ray, b.minx, b.miny, b.minz, b.maxx, b.maxy, b.maxz, distanceB);
            }
            hitA &= (distanceA < maxDistance);
            hitB &= (distanceB < maxDistance);
            if (kCount) cnt->boxTests += 2;
            if (hitA && hitB)
            {
                const bool swap = distanceB < distanceA;
                stk_put(stack, sstack, stackSize++, swap ? a.leafInfo : b.leafInfo);
                cur = swap ? b.leafInfo : a.leafInfo;
            }
            else if (hitA) cur = a.leafInfo;
            else if (hitB) cur = b.leafInfo;
            else if (stackSize > popFloor)
            {
                // both children missed: take the next pending node of this mesh right here, so that the lane stays in
                // the node loop with the rest of its warp instead of idling until every lane has left it
                cur = stk_get(stack, sstack, --stackSize);
            }
            else have = false;
#if defined(__CUDA_ARCH__) && RTB_SHADOW_VOTE > 0
            if (__popc(__activemask()) < RTB_SHADOW_VOTE) break;      // (see scene_traverse_from)
#endif
        }

    }

    // after the caller has walked the wide tree of `wideMesh` without finding an occluder
    template <typename WorldRay>
    RTB_HD void leave_wide(const SceneView& S, const WorldRay& worldRaySource)
    {
        nodes = S.tlas; ray = worldRaySource.load(); finite = worldFinite;
        // (what step() left behind when it returned kEnterWide, restated so that none of it has to live through the wide walk)
        have = false; inBlas = false; pend = 0; popFloor = 0xFFFFFFFFu; blasBase = 0; cur = 0; tris = nullptr;
    }

    // kSuspendWide: a mesh that has an 8-wide tree is not walked here; step() returns kEnterWide with the mesh-local ray
    // in `ray` and the caller walks it (k_shadow_wide does that for all lanes of a warp together)
    // kAllowWide = false: meshes are walked on the reference's binary tree whether they have a wide one or not (k_shadow and
    // k_tail: the round-1 kernels, whose register budget has no room for both walks)
    template <bool kCount, typename WorldRay, bool kSuspendWide = false, bool kAllowWide = true>
    RTB_HD int step(const SceneView& S, uint32_t* stack, const WorldRay& worldRaySource, TraceCounters* cnt, SharedStack sstack = 0)
    {
        // (the form of the node loop matters: `while (have) { if (leaf) break; ... }` WITHOUT the postponing measured 35 %
        //  slower than `while (have && !leaf)`: ptxas places the reconvergence points differently)
#if RTB_SPLIT_FINITE
        if (finite) node_phase<kCount, 1>(stack, cnt, sstack);
        else node_phase<kCount, 2>(stack, cnt, sstack);
#else
        node_phase<kCount, 0>(stack, cnt, sstack);
#endif

        if (pend != 0)
        {
            if (leaf_occludes<kCount>(pend, cnt)) return kOccluded;
            pend = 0;
        }
#if RTB_SHADOW_VOTE > 0
        if (have && RTB_NODE_COUNT(cur) == 0) return kContinue;       // left the node loop by vote: carry on with the same node
#endif
        if (have)
        {
            if (inBlas)
            {
                if (leaf_occludes<kCount>(cur, cnt)) return kOccluded;
            }
            else
            {
                // Scene::Traverse_Leaf_Shadow, Scene.cpp:180-194: one stack entry per object of the top-level leaf
                for (uint32_t i = RTB_NODE_COUNT(cur); i-- > 0;) { stk_put(stack, sstack, stackSize, RTB_OBJECT_ENTRY(RTB_NODE_INDEX(cur) + i)); stackSize++; }
            }
            have = false;
        }

        if (inBlas && stackSize == blasBase)
        {
            inBlas = false; popFloor = 0xFFFFFFFFu; nodes = S.tlas; ray = worldRaySource.load(); finite = worldFinite;
        }
        if (stackSize == 0) return kVisible;
        const uint32_t e = stk_get(stack, sstack, --stackSize);
        if (RTB_NODE_COUNT(e) == RTB_PACK_OBJECT)
        {
            // Scene::Traverse_Object_Shadow, Scene.cpp:147-165
            const DObjectTrace& obj = S.objTrace[RTB_NODE_INDEX(e)];
            const Ray localRay = transform_ray_unsafe(obj.invTransform, worldRaySource.load());
            if (obj.kind == RTB_OBJECT_SHAPE && obj.shapeType == RTB_SHAPE_MESH)
            {
                const DMesh& mesh = S.meshes[obj.meshIndex];
#if RTB_WIDE_SHADOW
                if (kAllowWide && mesh.numNodes != 0 && mesh.wideNodes != nullptr && wide_ray_ok(localRay))
                {
                    if (kSuspendWide) { wideMesh = obj.meshIndex; ray = localRay; return kEnterWide; }
                    // the whole mesh in one go on its 8-wide tree: same boolean as the binary walk (rtb_wide.h)
                    WideView wv; wv.nodes = mesh.wideNodes; wv.tris = mesh.wideTris; wv.refNodes = mesh.nodes; wv.maxAbs = mesh.wideMaxAbs;
                    WideStackAbove ws; ws.stack = stack; ws.sstack = sstack; ws.floor = stackSize; ws.top = stackSize;
                    if (wide_any_hit<kCount>(wv, localRay, maxDistance, ws, cnt)) return kOccluded;
                }
                else
#endif
                if (mesh.numNodes != 0)
                {
                    inBlas = true; blasBase = stackSize; popFloor = stackSize; nodes = mesh.nodes; tris = mesh.tris; ray = localRay; finite = ray_is_finite(localRay);
                    cur = node_header(nodes, 0);
                    have = true;
                }
            }
            else if (analytic_shadow(obj, localRay, maxDistance))
            {
                return kOccluded;
            }
        }
        else
        {
            cur = e; have = true;
        }
        return kContinue;
    }
};

template <bool kCount, typename WorldRay, bool kAllowWide = true>
RTB_HD bool scene_traverse_shadow_from(const SceneView& S, const WorldRay& worldRaySource, float maxDistance, TraceCounters* cnt, SharedStack sstack = 0)
{
    ShadowWalk w;
    uint32_t stack[RTB_LOCAL_STACK_SIZE];
    if (!w.begin(S, stack, worldRaySource, maxDistance, sstack)) return false;
    for (;;)
    {
        const int r = w.template step<kCount, WorldRay, false, kAllowWide>(S, stack, worldRaySource, cnt, sstack);
        if (r != ShadowWalk::kContinue) return r == ShadowWalk::kOccluded;
    }
}

template <bool kCount, bool kAllowWide = true>
RTB_HD bool scene_traverse_shadow(const SceneView& S, const Ray& worldRay, float maxDistance, TraceCounters* cnt, SharedStack sstack = 0)
{
    const RayInRegisters src = { worldRay };
    return scene_traverse_shadow_from<kCount, RayInRegisters, kAllowWide>(S, src, maxDistance, cnt, sstack);
}

} // namespace rtb
