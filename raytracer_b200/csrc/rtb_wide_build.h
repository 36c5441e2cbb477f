// rtb_wide_build.h — host side of rtb_wide.h: collapse of the reference's binary BVH (Core/BVH/BVHBuilder.cpp:37-277
// builds it; host/rt_bvh.h restates that builder) into the 8-wide quantised tree the any-hit walk uses.  Pure host C++.
#pragma once

#include "rtb_trace.h"

#include <vector>
#include <string>
#include <algorithm>
#include <cmath>
#include <cstring>

namespace rtb {

struct HostWide
{
    std::vector<WNode> nodes;      // breadth-first
    std::vector<DTri>  tris;       // leaf triangles in node order; pad0 = bits of the reference leaf's node index (RTB_WIDE_NO_LEAF: root leaf), pad1 = bits of the reference triangle index
    uint32_t depth = 0;            // levels of wide nodes (= worst-case number of pending groups on the traversal stack)
    float    maxAbs = 0.0f;        // largest |coordinate| of any box (wide_ray_eps)
    uint32_t maxLeafTris = 0;
};

// false: the mesh keeps the exact binary walk (empty tree, a box that sticks out of its parent's, leaves that are not
// contiguous triangle ranges, more triangles than 24 bits address, a root leaf that does not fit one node)
inline bool build_wide_bvh(const std::vector<DNode>& n2, const std::vector<DTri>& tris2, HostWide& out, uint32_t leafTris = RTB_WIDE_LEAF_TRIS)
{
    out = HostWide();
    const uint32_t numNodes = (uint32_t)n2.size();
    if (numNodes == 0 || tris2.empty() || tris2.size() >= (1u << 24)) return false;
    // leafTris = 0: every leaf child is ONE triangle with its own (inflated, quantised) box - a leaf of the reference's tree
    // with two triangles takes two child slots.  Triangle tests run at a few lanes per warp, child tests at nearly full
    // warps, so trading the former for the latter pays (k_shadow_mesh).
    const bool perTriangle = leafTris == 0u;
    if (perTriangle) leafTris = 1u;
    if (leafTris > 4u) leafTris = 4u;
    auto leafCount = [&](uint32_t i) { return n2[i].leafInfo & 0x3FFFFFFFu; };

    // per binary node: triangle range of its subtree (post-order), containment of the child boxes
    std::vector<uint32_t> first(numNodes, 0u), count(numNodes, 0u), parentLeafDepth(numNodes, 0u);
    {
        struct Item { uint32_t node; bool expanded; };
        std::vector<Item> stack;
        stack.push_back({ 0u, false });
        std::vector<uint8_t> seen(numNodes, 0);
        while (!stack.empty())
        {
            Item it = stack.back();
            stack.pop_back();
            const DNode& n = n2[it.node];
            if (leafCount(it.node) != 0)
            {
                first[it.node] = n.childIndex; count[it.node] = leafCount(it.node);
                if ((uint64_t)n.childIndex + count[it.node] > tris2.size()) return false;
                continue;
            }
            const uint32_t a = n.childIndex, b = a + 1;
            if (b >= numNodes) return false;
            if (!it.expanded)
            {
                if (seen[it.node]) return false;
                seen[it.node] = 1;
                stack.push_back({ it.node, true });
                stack.push_back({ a, false });
                stack.push_back({ b, false });
                // the monotonicity argument needs child boxes inside the parent's (the root's own box is never tested)
                for (uint32_t c = a; c <= b; ++c)
                {
                    const DNode& ch = n2[c];
                    if (!(ch.minx <= ch.maxx && ch.miny <= ch.maxy && ch.minz <= ch.maxz)) return false;      // also rejects NaN
                    if (it.node != 0u && !(ch.minx >= n.minx && ch.miny >= n.miny && ch.minz >= n.minz && ch.maxx <= n.maxx && ch.maxy <= n.maxy && ch.maxz <= n.maxz)) return false;
                }
            }
            else
            {
                if (first[b] != first[a] + count[a]) return false;      // subtrees are consecutive triangle ranges (BVHBuilder::GenerateLeaf)
                first[it.node] = first[a]; count[it.node] = count[a] + count[b];
            }
        }
    }
    for (const DNode& n : n2)
    {
        const float v[6] = { n.minx, n.miny, n.minz, n.maxx, n.maxy, n.maxz };
        for (float x : v) { if (!(std::fabs(x) < 1.0e18f)) return false; out.maxAbs = std::max(out.maxAbs, std::fabs(x)); }
    }

    auto area = [&](uint32_t i)
    {
        const DNode& n = n2[i];
        const double dx = (double)n.maxx - n.minx, dy = (double)n.maxy - n.miny, dz = (double)n.maxz - n.minz;
        return dx * dy + dy * dz + dz * dx;
    };
    // copies the triangles of binary subtree `node` to the wide array, tagging each with its reference leaf
    auto emitTriangles = [&](uint32_t node)
    {
        std::vector<uint32_t> st(1, node);
        std::vector<std::pair<uint32_t, uint32_t>> leaves;      // (first triangle, leaf node)
        while (!st.empty())
        {
            const uint32_t i = st.back(); st.pop_back();
            if (leafCount(i) != 0) leaves.push_back({ n2[i].childIndex, i });
            else { st.push_back(n2[i].childIndex); st.push_back(n2[i].childIndex + 1); }
        }
        std::sort(leaves.begin(), leaves.end());
        for (const auto& l : leaves)
            for (uint32_t t = 0; t < leafCount(l.second); ++t)
            {
                DTri tri = tris2[l.first + t];
                const uint32_t tag = l.second == 0u ? RTB_WIDE_NO_LEAF : l.second, index = l.first + t;
                memcpy(&tri.pad0, &tag, 4); memcpy(&tri.pad1, &index, 4);
                out.tris.push_back(tri);
            }
    };

    // Optimal collapse (Ylitie et al. 2017, section 3.1) by dynamic programming over the binary tree, bottom up:
    //   cost[n][i] = cheapest way to stand for subtree n with at most i + 1 children of one wide node (i = 0 .. 6)
    //   one child:  a leaf (area * triangles * kTriCost, if it has few enough triangles) or an inner wide node
    //               (area * kNodeCost + the cheapest split of ITS 8 slots over n's two children)
    //   more:       the cheapest split of the slots over n's two children, or the same as with one slot fewer
    // Costs are in surface-area units (SAH): an inner node costs its 8 child tests once, a leaf one exact test per triangle.
    const double kNodeCost = 1.0, kTriCost = 0.3;
    struct Dp { float cost[7]; uint8_t split[7]; uint8_t innerSplit; uint8_t leaf; };     // split[i]: slots given to the left child (0: same as i - 1)
    std::vector<Dp> dp(numNodes);
    {
        // post-order without recursion: children have larger indices than their parent (BVHBuilder allocates them after it),
        // but do not rely on that
        std::vector<uint32_t> order; order.reserve(numNodes);
        std::vector<uint32_t> st(1, 0u);
        while (!st.empty()) { const uint32_t i = st.back(); st.pop_back(); order.push_back(i); if (leafCount(i) == 0) { st.push_back(n2[i].childIndex); st.push_back(n2[i].childIndex + 1); } }
        for (size_t k = order.size(); k-- > 0;)
        {
            const uint32_t n = order[k];
            Dp& d = dp[n];
            const double a = area(n);
            const double leafCost = count[n] <= leafTris ? a * count[n] * kTriCost : 1e300;
            if (leafCount(n) != 0)
            {
                // per-triangle mode: a leaf of c triangles needs c slots (its cost: c boxes tested where the parent is visited)
                const uint32_t need = perTriangle ? count[n] : 1u;
                const double cst = perTriangle ? a * count[n] * kTriCost : (count[n] <= leafTris ? leafCost : a * count[n] * kTriCost);
                if (need > 7u) return false;
                for (int i = 0; i < 7; ++i) { d.cost[i] = (uint32_t)(i + 1) >= need ? (float)cst : INFINITY; d.split[i] = 0; }
                d.leaf = 1; d.innerSplit = 0;
                continue;
            }
            const Dp& l = dp[n2[n].childIndex]; const Dp& r = dp[n2[n].childIndex + 1];
            auto distribute = [&](int slots, uint8_t& bestK)        // slots >= 2: left gets k, right slots - k (each 1 .. 7)
            {
                double best = 1e300; bestK = 1;
                for (int kk = 1; kk < slots; ++kk)
                {
                    if (kk > 7 || slots - kk > 7) continue;
                    const double c = (double)l.cost[kk - 1] + (double)r.cost[slots - kk - 1];
                    if (c < best) { best = c; bestK = (uint8_t)kk; }
                }
                return best;
            };
            uint8_t k8;
            const double innerCost = distribute(8, k8) + a * kNodeCost;
            d.innerSplit = k8;
            d.leaf = leafCost <= innerCost ? 1 : 0;
            d.cost[0] = (float)std::min(leafCost, innerCost); d.split[0] = 0;
            for (int i = 1; i < 7; ++i)
            {
                uint8_t kk;
                const double c = distribute(i + 1, kk);
                if (c < d.cost[i - 1]) { d.cost[i] = (float)c; d.split[i] = kk; }
                else { d.cost[i] = d.cost[i - 1]; d.split[i] = 0; }
            }
        }
    }
    // the children that stand for subtree n when it may use `slots` slots
    struct Kid { uint32_t node; int tri; double box[6]; };      // tri >= 0: triangle `tri` (reference order) of leaf `node`, alone
    const double margin = 1.0e-4 * (double)out.maxAbs;        // a float triangle hit lies inside the triangle's box inflated by this (MT's error is a few ulps of maxAbs)
    auto makeKid = [&](uint32_t node, int tri)
    {
        Kid k; k.node = node; k.tri = tri;
        if (tri < 0)
        {
            const DNode& c = n2[node];
            k.box[0] = c.minx; k.box[1] = c.miny; k.box[2] = c.minz; k.box[3] = c.maxx; k.box[4] = c.maxy; k.box[5] = c.maxz;
        }
        else
        {
            const DTri& t = tris2[(size_t)tri];
            const double v[3][3] = { { t.v0x, t.v0y, t.v0z }, { (double)t.v0x + t.e1x, (double)t.v0y + t.e1y, (double)t.v0z + t.e1z }, { (double)t.v0x + t.e2x, (double)t.v0y + t.e2y, (double)t.v0z + t.e2z } };
            for (int a = 0; a < 3; ++a)
            {
                k.box[a] = std::min(v[0][a], std::min(v[1][a], v[2][a])) - margin;
                k.box[3 + a] = std::max(v[0][a], std::max(v[1][a], v[2][a])) + margin;
            }
        }
        return k;
    };
    auto pushLeaf = [&](uint32_t node, std::vector<Kid>& kids)
    {
        if (!perTriangle) { kids.push_back(makeKid(node, -1)); return; }
        for (uint32_t t = 0; t < count[node]; ++t) kids.push_back(makeKid(node, (int)(first[node] + t)));
    };
    auto gather = [&](uint32_t root, int rootSlots, std::vector<Kid>& kids)
    {
        struct G { uint32_t node; int slots; };
        std::vector<G> st(1, { root, rootSlots });
        while (!st.empty())
        {
            G g = st.back(); st.pop_back();
            const Dp& d = dp[g.node];
            if (leafCount(g.node) != 0) { pushLeaf(g.node, kids); continue; }
            while (g.slots > 1 && d.split[g.slots - 1] == 0) g.slots--;
            if (g.slots <= 1) { kids.push_back(makeKid(g.node, -1)); continue; }
            const int kk = d.split[g.slots - 1];
            st.push_back({ n2[g.node].childIndex + 1, g.slots - kk });
            st.push_back({ n2[g.node].childIndex, kk });
        }
    };

    struct Work { uint32_t node; uint32_t wide; uint32_t level; };      // binary subtree root -> wide node index
    std::vector<Work> queue;
    out.nodes.resize(1);
    queue.push_back({ 0u, 0u, 1u });
    for (size_t head = 0; head < queue.size(); ++head)
    {
        const Work w = queue[head];
        out.depth = std::max(out.depth, w.level);
        std::vector<Kid> kids;
        if (leafCount(w.node) != 0)
        {
            // (only the root can arrive here as a leaf) one node whose children are that leaf / its triangles
            if (!perTriangle && count[w.node] > leafTris) return false;
            pushLeaf(w.node, kids);
        }
        else
        {
            const int kk = dp[w.node].innerSplit;
            gather(n2[w.node].childIndex, kk, kids);
            gather(n2[w.node].childIndex + 1, 8 - kk, kids);
        }
        if (kids.empty() || kids.size() > 8) return false;
        // slot order: largest surface first (an any-hit ray is most likely to be stopped there)
        auto kidArea = [](const Kid& k) { const double dx = k.box[3] - k.box[0], dy = k.box[4] - k.box[1], dz = k.box[5] - k.box[2]; return dx * dy + dy * dz + dz * dx; };
        std::stable_sort(kids.begin(), kids.end(), [&](const Kid& a, const Kid& b) { return kidArea(a) > kidArea(b); });

        WNode node;
        memset(&node, 0, sizeof(node));
        double lo[3] = { 1e300, 1e300, 1e300 }, hi[3] = { -1e300, -1e300, -1e300 };
        for (const Kid& k : kids)
            for (int a = 0; a < 3; ++a) { lo[a] = std::min(lo[a], k.box[a]); hi[a] = std::max(hi[a], k.box[3 + a]); }
        // the grid origin must be a float at or below the smallest coordinate (inflated triangle boxes are doubles)
        for (int a = 0; a < 3; ++a) { float f = (float)lo[a]; if ((double)f > lo[a]) f = std::nextafterf(f, -INFINITY); lo[a] = (double)f; }
        node.px = (float)lo[0]; node.py = (float)lo[1]; node.pz = (float)lo[2];
        int e[3];
        for (int a = 0; a < 3; ++a)
        {
            const double extent = hi[a] - lo[a];
            int ex = -100;
            if (extent > 0.0) { int fe; std::frexp(extent / 255.0, &fe); ex = fe; }      // 2^fe > extent / 255
            while (std::ldexp(255.0, ex) < extent) ++ex;
            ex = std::max(ex, -100);
            if (ex + 127 + 15 > 254) return false;
            e[a] = ex;
        }
        uint32_t imask = 0, valid = 0, numInner = 0, triOffset = 0;
        const uint32_t triBase = (uint32_t)out.tris.size();
        uint8_t meta[8] = { 0 }, qlo[3][8], qhi[3][8];
        memset(qlo, 255, sizeof(qlo)); memset(qhi, 0, sizeof(qhi));               // empty children: inverted boxes
        for (size_t s = 0; s < kids.size(); ++s)
        {
            const Kid& kid = kids[s];
            const uint32_t k = kid.node;
            for (int a = 0; a < 3; ++a)
            {
                const double step = std::ldexp(1.0, e[a]);
                double l = std::floor((kid.box[a] - lo[a]) / step), h = std::ceil((kid.box[3 + a] - lo[a]) / step);
                l = std::min(std::max(l, 0.0), 255.0); h = std::min(std::max(h, 0.0), 255.0);
                if (lo[a] + l * step > kid.box[a] || lo[a] + h * step < kid.box[3 + a]) return false;     // (cannot happen)
                qlo[a][s] = (uint8_t)l; qhi[a][s] = (uint8_t)h;
            }
            valid |= 1u << s;
            if (kid.tri >= 0)
            {
                // one triangle of a reference leaf
                if (triOffset + 1u > 32u) return false;
                meta[s] = (uint8_t)(0x80u | triOffset);
                triOffset += 1u;
                out.maxLeafTris = std::max(out.maxLeafTris, 1u);
                DTri tri = tris2[(size_t)kid.tri];
                const uint32_t tag = k == 0u ? RTB_WIDE_NO_LEAF : k, index = (uint32_t)kid.tri;
                memcpy(&tri.pad0, &tag, 4); memcpy(&tri.pad1, &index, 4);
                out.tris.push_back(tri);
            }
            else if (dp[k].leaf || leafCount(k) != 0)
            {
                if (count[k] > 4u || triOffset + count[k] > 32u) return false;
                meta[s] = (uint8_t)(0x80u | ((count[k] - 1u) << 5) | triOffset);
                triOffset += count[k];
                out.maxLeafTris = std::max(out.maxLeafTris, count[k]);
                emitTriangles(k);
            }
            else { imask |= 1u << s; numInner++; }
        }
        node.eimask = (uint32_t)(e[0] + 127) | ((uint32_t)(e[1] + 127) << 8) | ((uint32_t)(e[2] + 127) << 16) | (imask << 24);
        node.childBase = (uint32_t)out.nodes.size();
        if (out.tris.size() >= (1u << 24)) return false;
        node.triBaseValid = triBase | (valid << 24);
        memcpy(node.meta, meta, 8);
        for (int a = 0; a < 3; ++a) { memcpy(node.qlo[a], qlo[a], 8); memcpy(node.qhi[a], qhi[a], 8); }
        out.nodes.resize(out.nodes.size() + numInner);
        uint32_t rank = 0;
        for (size_t s = 0; s < kids.size(); ++s) if (imask & (1u << s)) queue.push_back({ kids[s].node, node.childBase + rank++, w.level + 1u });
        out.nodes[w.wide] = node;
    }
    return true;
}

} // namespace rtb
