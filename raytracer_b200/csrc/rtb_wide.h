// rtb_wide.h — 8-wide BVH with 8-bit quantised child boxes for ANY-HIT (shadow) rays inside one mesh.
//
// Why this is allowed next to a bit-exact bar.  The reference's shadow walk (GenericTraverse_Shadow,
// Core/Traversal/Traversal_Single.h:100-179 -> MeshShape::Traverse_Leaf_Shadow, MeshShape.cpp:175-207) returns
//     "some triangle T passes Intersect_TriangleRay with distance < maxDistance  AND  every node on the way from the root
//      (exclusive) to T's leaf passes Intersect_BoxRay with distance < maxDistance".
// The builder makes every node's box the union of its leaves' boxes (BVHBuilder.cpp:124-172), so a child box lies inside
// its parent's box component by component, and for a ray whose invDir / originDivDir are finite the slab test
// fma(box, invDir, -originDivDir) is monotonic in the box coordinate: if a LEAF's box passes, all its ancestors' boxes
// pass too (t_near can only shrink, t_far only grow).  Hence for such rays
//     occluded  <=>  exists a leaf L (not the root) : box test of L passes  AND  a triangle of L passes the triangle test
// and the result does not depend on the tree above the leaves at all.  Any CONSERVATIVE acceleration structure over the
// same leaves therefore gives the same boolean, provided every candidate hit is confirmed with the reference's exact
// triangle test and the exact box test of the triangle's own leaf.  build_wide_bvh() verifies the containment
// precondition node by node and refuses the mesh otherwise (the exact binary walk of rtb_trace.h is used then, as it is
// for rays with a non-finite slab such as axis-aligned ones).
//
// Structure (after Ylitie, Karras, Laine: "Efficient Incoherent Ray Traversal on GPUs Through Compressed Wide BVHs",
// HPG 2017, re-derived for this use): the reference's binary tree is collapsed into nodes of up to 8 children, the
// child with the largest surface area being opened first; a child is either an inner node or a leaf of up to
// RTB_WIDE_LEAF_TRIS triangles (small binary subtrees are folded into one leaf child).  A node is 96 bytes, 32-byte
// aligned: 2 x LDG.256 + 1 x LDG.128.  Child boxes are quantised to 8 bits per coordinate on a per-node grid
// (origin p, power-of-two step 2^e per axis), rounded outwards.  Nodes are stored in BREADTH-FIRST order, so that the
// first K nodes are the top of the tree (the part the persistent shadow kernel stages in shared memory).
#pragma once

// (included from rtb_trace.h, after the box / triangle tests and the node loads it builds on)

namespace rtb {

#ifndef RTB_WIDE_LEAF_TRIS
#define RTB_WIDE_LEAF_TRIS 3u       // at most 4 (2 bits of the child meta byte)
#endif
#define RTB_WIDE_NO_LEAF 0xFFFFFFFFu   // triangle whose reference leaf is the root (no box test on the way to it)

struct __attribute__((aligned(32))) WNode
{
    float    px, py, pz;       // origin of the quantisation grid (min corner of the union of the child boxes)
    uint32_t eimask;           // ex | ey << 8 | ez << 16 | imask << 24: e* = IEEE exponent byte of the grid step, bit i of imask: child i is an inner node
    uint32_t childBase;        // node index of the first inner child (inner children are consecutive, in slot order)
    uint32_t triBaseValid;     // first triangle of the leaf children in the mesh's wide triangle array (24 bits) | mask of the non-empty children << 24
    uint32_t meta[2];          // one byte per child.  leaf: 0x80 | (count - 1) << 5 | offset from the node's first triangle; inner / empty: 0
    uint32_t qlo[3][2];        // [axis][children 0-3, 4-7]: one byte per child
    uint32_t qhi[3][2];
    uint32_t pad[4];
};
static_assert(sizeof(WNode) == 96, "wide node layout");

// A ray may use the wide tree if the reference's box test is monotonic for it (finite slabs) and nothing here can
// overflow into inf - inf.
RTB_HD bool wide_ray_ok(const Ray& r)
{
    const uint32_t kMaxExp = (127u + 60u) << 23;      // |x| < 2^60
    const uint32_t ix = f2bits(r.invDir.x) & 0x7FFFFFFFu, iy = f2bits(r.invDir.y) & 0x7FFFFFFFu, iz = f2bits(r.invDir.z) & 0x7FFFFFFFu;
    const uint32_t ox = f2bits(r.origin.x) & 0x7FFFFFFFu, oy = f2bits(r.origin.y) & 0x7FFFFFFFu, oz = f2bits(r.origin.z) & 0x7FFFFFFFu;
    const uint32_t kMinExp = (127u - 60u) << 23;      // |invDir| > 2^-60 (a direction component cannot be that large anyway)
    return ix < kMaxExp && iy < kMaxExp && iz < kMaxExp && ix > kMinExp && iy > kMinExp && iz > kMinExp &&
           ox < kMaxExp && oy < kMaxExp && oz < kMaxExp;
}

// slack per slab distance.  Both tests evaluate (x - o) * invDir for a box plane x: the reference as
// fma(x, invDir, -(o * invDir)), this one as fma(f, step * invDir, (p - o) * invDir - step' * invDir).  Each rounding is
// relative (2^-24) to a term bounded by (|o| + maxAbs) * |invDir|, there are fewer than eight of them, so 2^-19 of that
// bound covers the difference with a factor of four to spare.  (The grid-step term is added per node, see wide_node_test.)
// The bound holds axis by axis, and the slack must be per axis as well: one shared slack (the largest of the three) makes
// a ray with one tiny direction component - invDir 10^5 on that axis, a slack of several scene lengths - pass EVERY box
// on the other two axes too and walk the whole tree.  A few thousand such rays per launch kept a handful of warps busy
// for 3 M cycles after all others had finished (ncu: sm__cycles_active 12.8 M min / 16.1 M max, profiles/r2j_*).
RTB_HD V3 wide_ray_eps(const Ray& r, float maxAbs)
{
    const float ex = (bit_abs(r.origin.x) + maxAbs) * bit_abs(r.invDir.x);
    const float ey = (bit_abs(r.origin.y) + maxAbs) * bit_abs(r.invDir.y);
    const float ez = (bit_abs(r.origin.z) + maxAbs) * bit_abs(r.invDir.z);
    return v3(ex * 1.9073486328125e-06f, ey * 1.9073486328125e-06f, ez * 1.9073486328125e-06f);     // 2^-19
}

// byte k (0..3) of w as the float 1 + byte * 2^-15 (exact): one PRMT on the device
#if defined(__CUDACC__)
static __constant__ uint32_t g_wideOne = 0x3F800000u;      // bits of 1.0f, see wide_byte_float (static: the library is two translation units)
#endif
// (`one` = 0x3F800000 held in a register the compiler cannot see through: PRMT takes one immediate, and with both the
//  constant and the selector known ptxas kept the constant as the immediate and rebuilt the SELECTOR in a register
//  before every one of the 48 PRMTs of a node visit - 47 extra IMAD.U32 of 297 instructions)
template <int k>
RTB_HD float wide_byte_float(uint32_t w, uint32_t one)
{
#if defined(__CUDA_ARCH__)
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(w), "r"(one), "n"(0x7604 | (k << 4)));
    return __uint_as_float(r);
#else
    return bits2f(one | (((w >> (8 * k)) & 0xFFu) << 8));
#endif
}

// The 8 child tests of one node.  Returns the mask of children whose (outward-rounded, slack-widened) box the ray
// segment [0, tMax] touches.  n1 = words 8..15 of the node (qlo x, y, z, qhi x), n2 = words 16..19 (qhi y, z).
RTB_HD uint32_t wide_node_test(float px, float py, float pz, uint32_t eimask, const uint32_t* q /* 12 words: qlo[3][2], qhi[3][2] */,
                               const Ray& ray, float tMax, const V3& eps)
{
    // grid step 2^e per axis, pre-multiplied by 2^15 (the child coordinate arrives as 1 + q * 2^-15)
    const float sx = bits2f(((eimask & 0xFFu) + 15u) << 23);
    const float sy = bits2f((((eimask >> 8) & 0xFFu) + 15u) << 23);
    const float sz = bits2f((((eimask >> 16) & 0xFFu) + 15u) << 23);
    const float ax = sx * ray.invDir.x, ay = sy * ray.invDir.y, az = sz * ray.invDir.z;          // exact scalings of invDir
    // t(q) = (p + q * 2^e - o) * invDir = f * a + (b - a),  f = 1 + q * 2^-15,  b = (p - o) * invDir
    const float bx = (px - ray.origin.x) * ray.invDir.x, by = (py - ray.origin.y) * ray.invDir.y, bz = (pz - ray.origin.z) * ray.invDir.z;
    // slack: the ray's (see wide_ray_eps) + 2^-22 of |a| for the rounding of (b - a) (= 1/128 of a grid step)
    const float kx = fma_(bit_abs(ax), 2.384185791015625e-07f, eps.x), ky = fma_(bit_abs(ay), 2.384185791015625e-07f, eps.y), kz = fma_(bit_abs(az), 2.384185791015625e-07f, eps.z);
    const float cx = bx - ax, cy = by - ay, cz = bz - az;
    const float clx = cx - kx, chx = cx + kx, cly = cy - ky, chy = cy + ky, clz = cz - kz, chz = cz + kz;
    // entry planes: qlo for a positive direction, qhi for a negative one
    const bool nx = ray.invDir.x < 0.0f, ny = ray.invDir.y < 0.0f, nz = ray.invDir.z < 0.0f;
    const uint32_t nearX0 = nx ? q[6] : q[0], nearX1 = nx ? q[7] : q[1], farX0 = nx ? q[0] : q[6], farX1 = nx ? q[1] : q[7];
    const uint32_t nearY0 = ny ? q[8] : q[2], nearY1 = ny ? q[9] : q[3], farY0 = ny ? q[2] : q[8], farY1 = ny ? q[3] : q[9];
    const uint32_t nearZ0 = nz ? q[10] : q[4], nearZ1 = nz ? q[11] : q[5], farZ0 = nz ? q[4] : q[10], farZ1 = nz ? q[5] : q[11];
    uint32_t mask = 0;
#if defined(__CUDA_ARCH__)
    const uint32_t one = g_wideOne;          // (from constant memory: a value ptxas cannot fold back into an immediate)
#else
    const uint32_t one = 0x3F800000u;
#endif
#define RTB_WIDE_CHILD(i, k, NX, NY, NZ, FX, FY, FZ)                                                          \
    {                                                                                                           \
        const float tlx = fma_(wide_byte_float<k>(NX, one), ax, clx), tly = fma_(wide_byte_float<k>(NY, one), ay, cly), tlz = fma_(wide_byte_float<k>(NZ, one), az, clz); \
        const float thx = fma_(wide_byte_float<k>(FX, one), ax, chx), thy = fma_(wide_byte_float<k>(FY, one), ay, chy), thz = fma_(wide_byte_float<k>(FZ, one), az, chz); \
        const float tmin = fmaxf(fmaxf(tlx, tly), fmaxf(tlz, 0.0f));                                            \
        const float tmax = fminf(fminf(thx, thy), fminf(thz, tMax));                                            \
        if (tmin <= tmax) mask |= 1u << (i);                                                                    \
    }
    RTB_WIDE_CHILD(0, 0, nearX0, nearY0, nearZ0, farX0, farY0, farZ0)
    RTB_WIDE_CHILD(1, 1, nearX0, nearY0, nearZ0, farX0, farY0, farZ0)
    RTB_WIDE_CHILD(2, 2, nearX0, nearY0, nearZ0, farX0, farY0, farZ0)
    RTB_WIDE_CHILD(3, 3, nearX0, nearY0, nearZ0, farX0, farY0, farZ0)
    RTB_WIDE_CHILD(4, 0, nearX1, nearY1, nearZ1, farX1, farY1, farZ1)
    RTB_WIDE_CHILD(5, 1, nearX1, nearY1, nearZ1, farX1, farY1, farZ1)
    RTB_WIDE_CHILD(6, 2, nearX1, nearY1, nearZ1, farX1, farY1, farZ1)
    RTB_WIDE_CHILD(7, 3, nearX1, nearY1, nearZ1, farX1, farY1, farZ1)
#undef RTB_WIDE_CHILD
    return mask;
}

// the mesh-level view the traversal needs
struct WideView
{
    const WNode* nodes;
    const DTri*  tris;
    const DNode* refNodes;     // the reference's binary nodes: the exact box of a triangle's leaf
    float        maxAbs;
};

// exact confirmation of a candidate: the reference's triangle test, then the reference's box test of the triangle's leaf
// (measured, round 2: loading the leaf's box as soon as the triangle's first 16 bytes are there, so that its latency runs under
//  the triangle test - 18.36 -> 18.32 ms per 8 passes, noise: the kernel is bound by issue slots and L1 wavefronts, not by
//  the latencies ncu's per-line view shows on these loads)
RTB_HD bool wide_confirm(const WideView& W, const Ray& ray, float tMax, uint32_t triIndex)
{
#if defined(__CUDA_ARCH__)
    const float4* p = reinterpret_cast<const float4*>(W.tris + triIndex);
    const float4 a = RTB_LDG(p), b = RTB_LDG(p + 1), c = RTB_LDG(p + 2);
    const V3 v0 = v3(a.x, a.y, a.z), e1 = v3(b.x, b.y, b.z), e2 = v3(c.x, c.y, c.z);
    const uint32_t leaf = __float_as_uint(a.w);
#else
    const DTri& tr = W.tris[triIndex];
    const V3 v0 = v3(tr.v0x, tr.v0y, tr.v0z), e1 = v3(tr.e1x, tr.e1y, tr.e1z), e2 = v3(tr.e2x, tr.e2y, tr.e2z);
    const uint32_t leaf = f2bits(tr.pad0);
#endif
    if (!triangle_ray_occludes(ray, v0, e1, e2, tMax)) return false;
    if (leaf == RTB_WIDE_NO_LEAF) return true;
    const DNode n = load_node(W.refNodes, leaf);
    // (a mesh-local ray comes from make_ray_unsafe: originDivDir IS origin * invDir - forming the product here instead of
    //  reading the member lets the kernel drop three registers from the state it carries through the walk)
    Ray r = ray;
    r.originDivDir = ray.origin * ray.invDir;
    float d;
    return box_ray_finite(r, n.minx, n.miny, n.minz, n.maxx, n.maxy, n.maxz, d) && d < tMax;
}

// Any-hit walk of one mesh for a ray that passed wide_ray_ok().  Stack: groups of pending children, two words each
// (first inner child | pending mask + inner mask); at most `depth` of them.  PUSH(a, b) / POP(a, b) are supplied by the caller.
// where the 20 words of a node come from
struct WideFetchGlobal
{
    RTB_HD void operator()(const WNode* nodes, uint32_t index, uint32_t* n) const
    {
        const uint32_t* w = reinterpret_cast<const uint32_t*>(nodes + index);
#if defined(__CUDA_ARCH__)
        asm(RTB_LD_NC ".v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(n[0]), "=r"(n[1]), "=r"(n[2]), "=r"(n[3]), "=r"(n[4]), "=r"(n[5]), "=r"(n[6]), "=r"(n[7]) : "l"(w));
        asm(RTB_LD_NC ".v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8+32];" : "=r"(n[8]), "=r"(n[9]), "=r"(n[10]), "=r"(n[11]), "=r"(n[12]), "=r"(n[13]), "=r"(n[14]), "=r"(n[15]) : "l"(w));
        asm(RTB_LD_NC ".v4.u32 {%0,%1,%2,%3}, [%4+64];" : "=r"(n[16]), "=r"(n[17]), "=r"(n[18]), "=r"(n[19]) : "l"(w));
#else
        for (int i = 0; i < 20; ++i) n[i] = w[i];
#endif
    }
};

// The walk, one node visit per step() so that a kernel can interleave it with other work (k_shadow_wide refills idle
// lanes between visits).  State: the current group of pending children (first inner child | pending mask + inner mask
// << 8); groups that still have pending children wait on the caller's stack, at most `depth` of them.
struct WideWalk
{
    enum { kContinue = 0, kOccluded = 1, kMissed = 2 };
    uint32_t base, bits;
    V3 eps;

    RTB_HD void begin(const Ray& ray, float maxAbs)
    {
        eps = wide_ray_eps(ray, maxAbs);
        base = 0u; bits = 0x0101u;               // pseudo group: the root is "inner child 0" of nothing
    }

    template <bool kCount, typename Stack, typename Fetch>
    RTB_HD int step(const WideView& W, const Ray& ray, float tMax, Stack& stk, TraceCounters* cnt, const Fetch& fetch)
    {
        if ((bits & 0xFFu) == 0u)
        {
            if (stk.empty()) return kMissed;
            stk.pop(base, bits);
        }
        const uint32_t pending = bits & 0xFFu;
#if defined(__CUDA_ARCH__)
        const uint32_t slot = 31u - (uint32_t)__clz(pending & (0u - pending));
#else
        const uint32_t slot = 31u - (uint32_t)__builtin_clz(pending & (0u - pending));
#endif
        bits &= ~(1u << slot);
        const uint32_t below = (bits >> 8) & ((1u << slot) - 1u);
#if defined(__CUDA_ARCH__)
        const uint32_t index = base + __popc(below);
#else
        const uint32_t index = base + (uint32_t)__builtin_popcount(below);
#endif
        if (bits & 0xFFu) stk.push(base, bits);

        uint32_t n[20];
        fetch(W.nodes, index, n);
        const uint32_t imask = n[3] >> 24;
        const uint32_t hit = wide_node_test(bits2f(n[0]), bits2f(n[1]), bits2f(n[2]), n[3], n + 8, ray, tMax, eps) & (n[5] >> 24);
        if (kCount) cnt->boxTests += 8;
        uint32_t leaves = hit & ~imask;
        while (leaves)
        {
#if defined(__CUDA_ARCH__)
            const uint32_t s = 31u - (uint32_t)__clz(leaves & (0u - leaves));
#else
            const uint32_t s = 31u - (uint32_t)__builtin_clz(leaves & (0u - leaves));
#endif
            leaves &= leaves - 1u;
            const uint32_t m = (((s & 4u) ? n[7] : n[6]) >> (8u * (s & 3u))) & 0xFFu;      // (no dynamic index: n[] stays in registers)
            const uint32_t firstTri = (n[5] & 0x00FFFFFFu) + (m & 31u), numTris = ((m >> 5) & 3u) + 1u;
            if (kCount) cnt->triTests += numTris;
            for (uint32_t i = 0; i < numTris; ++i) if (wide_confirm(W, ray, tMax, firstTri + i)) return kOccluded;
        }
        base = n[4];
        bits = (hit & imask) | (imask << 8);
        return kContinue;
    }
};

// The same walk with the leaf confirmations taken OUT of the node visit (k_shadow_wide, RTB_SW_DEFER).  ncu on WideWalk
// showed why the 8-wide tree bought so little: the eight child tests ran at 26 of 32 lanes, but the triangle / leaf-box
// confirmations inside step() ran at 2-4 lanes (each lane meets its leaves at a different visit) and took 37 % of the
// kernel's warp instructions.  Here a node visit only QUEUES the leaves it touches (one word per visit: node index | mask
// of the touched leaf children << 24) and leaf_step() confirms one queued leaf; the kernel runs either as a warp-wide
// phase, chosen by vote, so that both run at many lanes.  Any-hit: the order in which leaves are confirmed cannot
// change the boolean.  The queue grows downwards from entry `leafTop` of the lane's shared-memory column (the node
// groups grow upwards from the floor); the caller keeps RTB_WIDE_LEAF_QUEUE entries free for it and lets a lane visit
// a node only while can_visit() holds (one more entry fits).
#define RTB_WIDE_LEAF_QUEUE 8u
struct WideWalkDeferred
{
    uint32_t base, bits, leafCount;
    V3 eps;

    RTB_HD void begin(const Ray& ray, float maxAbs)
    {
        eps = wide_ray_eps(ray, maxAbs);
        base = 0u; bits = 0x0101u; leafCount = 0u;
    }
    template <typename Stack> RTB_HD bool has_nodes(const Stack& stk) const { return (bits & 0xFFu) != 0u || !stk.empty(); }
    RTB_HD bool can_visit() const { return leafCount < RTB_WIDE_LEAF_QUEUE; }

    // one node visit; the lane has_nodes() and can_visit()
    template <bool kCount, typename Stack, typename Fetch>
    RTB_HD void node_step(const WideView& W, const Ray& ray, float tMax, Stack& stk, SharedStack sstack, uint32_t leafTop, TraceCounters* cnt, const Fetch& fetch)
    {
        if ((bits & 0xFFu) == 0u) stk.pop(base, bits);
        const uint32_t pending = bits & 0xFFu;
#if defined(__CUDA_ARCH__)
        const uint32_t slot = 31u - (uint32_t)__clz(pending & (0u - pending));
#else
        const uint32_t slot = 31u - (uint32_t)__builtin_clz(pending & (0u - pending));
#endif
        bits &= ~(1u << slot);
        const uint32_t below = (bits >> 8) & ((1u << slot) - 1u);
#if defined(__CUDA_ARCH__)
        const uint32_t index = base + __popc(below);
#else
        const uint32_t index = base + (uint32_t)__builtin_popcount(below);
#endif
        if (bits & 0xFFu) stk.push(base, bits);

        uint32_t n[20];
        fetch(W.nodes, index, n);
        const uint32_t imask = n[3] >> 24;
        const uint32_t hit = wide_node_test(bits2f(n[0]), bits2f(n[1]), bits2f(n[2]), n[3], n + 8, ray, tMax, eps) & (n[5] >> 24);
        if (kCount) cnt->boxTests += 8;
        // ONE queue entry per visit whatever the number of leaves touched (node index | mask of the touched leaf children << 24);
        // leaf_step() decodes a child from the node's meta words when its turn comes.  (A loop that pushed one entry per
        // leaf ran at 3.8 lanes and took 5 % of the kernel's warp instructions.)
        const uint32_t leaves = hit & ~imask;
        if (leaves)
        {
            leafCount++;
            const uint32_t entry = index | (leaves << 24);
#if RTB_SSTACK
            asm volatile("st.shared.u32 [%0], %1;" :: "r"(sstack + (leafTop - leafCount) * (RTB_STACK_THREADS * 4u)), "r"(entry));
#else
            (void)sstack; stk.stack[leafTop - leafCount] = entry;
#endif
        }
        base = n[4];
        bits = (hit & imask) | (imask << 8);
    }

    // confirm one queued leaf (leafCount != 0); true: the ray is occluded
    template <bool kCount, typename Stack>
    RTB_HD bool leaf_step(const WideView& W, const Ray& ray, float tMax, const Stack& stk, SharedStack sstack, uint32_t leafTop, TraceCounters* cnt)
    {
        uint32_t entry;
#if RTB_SSTACK
        (void)stk;
        const uint32_t slotAddr = sstack + (leafTop - leafCount) * (RTB_STACK_THREADS * 4u);
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(entry) : "r"(slotAddr));
#else
        (void)sstack; entry = stk.stack[leafTop - leafCount];
#endif
        const uint32_t index = entry & 0x00FFFFFFu;
        uint32_t mask = entry >> 24;
#if defined(__CUDA_ARCH__)
        const uint32_t s = 31u - (uint32_t)__clz(mask & (0u - mask));
        const uint4 w = RTB_LDG(reinterpret_cast<const uint4*>(W.nodes + index) + 1);       // childBase, triBaseValid, meta[0], meta[1]
#else
        const uint32_t s = 31u - (uint32_t)__builtin_clz(mask & (0u - mask));
        const WNode& nd = W.nodes[index];
        struct { uint32_t x, y, z, w; } w = { nd.childBase, nd.triBaseValid, nd.meta[0], nd.meta[1] };
#endif
        mask &= mask - 1u;
        if (mask)
        {
            // more touched leaves in this node: the entry stays, without the child taken now
#if RTB_SSTACK
            asm volatile("st.shared.u32 [%0], %1;" :: "r"(slotAddr), "r"(index | (mask << 24)));
#else
            stk.stack[leafTop - leafCount] = index | (mask << 24);
#endif
        }
        else leafCount--;
        const uint32_t m = (((s & 4u) ? w.w : w.z) >> (8u * (s & 3u))) & 0xFFu;
        const uint32_t firstTri = (w.y & 0x00FFFFFFu) + (m & 31u), numTris = ((m >> 5) & 3u) + 1u;
        if (kCount) cnt->triTests += numTris;
        bool occluded = false;
        for (uint32_t i = 0; i < numTris; ++i) if (wide_confirm(W, ray, tMax, firstTri + i)) occluded = true;
        return occluded;
    }
};

template <bool kCount, typename Stack, typename Fetch = WideFetchGlobal>
RTB_HD bool wide_any_hit(const WideView& W, const Ray& ray, float tMax, Stack& stk, TraceCounters* cnt, const Fetch& fetch = Fetch())
{
    WideWalk walk;
    walk.begin(ray, W.maxAbs);
    for (;;)
    {
        const int r = walk.template step<kCount>(W, ray, tMax, stk, cnt, fetch);
        if (r != WideWalk::kContinue) return r == WideWalk::kOccluded;
    }
}

} // namespace rtb
