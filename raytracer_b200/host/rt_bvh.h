// rt_bvh.h — host BVH (the reference's 32-byte node format) and its full-sweep SAH builder.
//
// The device traverses exactly the tree the reference would build, so closest-hit tie cases resolve identically.
// The builder therefore follows BVHBuilder (Core/BVH/BVHBuilder.cpp:37-277) decision for decision:
//   * root occupies two node slots (children pairs start at even indices, slot 1 stays unused)          :86
//   * leaves hold <= maxLeafNodeSize (2) items                                                           :124-128, BVHBuilder.h:16
//   * per node: sort the items along each axis by (box.min + box.max)[axis] with std::sort, the axis the
//     parent split on keeps its inherited order                                                           :242-277
//   * sweep all split positions of all three axes, cost = area(L) * nL + area(R) * nR with the half-area
//     of Box::SurfaceArea; first strictly smaller cost wins                                              :139-201
//   * children are built depth first, left then right; leaf items are appended to the output order      :203-239
#pragma once

#include <algorithm>
#include <cstring>
#include <vector>

#include "rt_math.h"
#include "../../include/rtb200.h"

namespace rt {

struct BvhBuildingParams
{
    enum class Heuristics { SurfaceArea, Volume };
    uint32_t maxLeafNodeSize = 2;
    Heuristics heuristics = Heuristics::SurfaceArea;
};

class BVH
{
public:
    static constexpr uint32_t MaxDepth = 128;      // BVH.h:20
    std::vector<RtbBvhNode> nodes;
    uint32_t GetNumNodes() const { return (uint32_t)nodes.size(); }
};

class BVHBuilder
{
public:
    using Indices = std::vector<uint32_t>;

    explicit BVHBuilder(BVH& target) : mTarget(target) {}

    bool Build(const math::Box* data, uint32_t numLeaves, const BvhBuildingParams& params, Indices& outLeavesOrder)
    {
        mLeafBoxes = data;
        mNumLeaves = numLeaves;
        mParams = params;
        mTarget.nodes.assign((size_t)2 * numLeaves, RtbBvhNode());
        for (RtbBvhNode& n : mTarget.nodes) memset(&n, 0, sizeof(n));
        mNumGeneratedNodes = 0;
        mNumGeneratedLeaves = 0;
        mLeavesOrder.clear();
        mLeavesOrder.reserve(numLeaves);
        if (numLeaves == 0) { outLeavesOrder.clear(); return true; }

        math::Box overall = math::Box::Empty();
        for (uint32_t i = 0; i < numLeaves; ++i) overall = math::Box::Merge(overall, data[i]);

        WorkSet root;
        root.box = overall;
        root.leafIndices.resize(numLeaves);
        for (uint32_t i = 0; i < numLeaves; ++i) root.leafIndices[i] = i;

        mLeft.resize(numLeaves);
        mRight.resize(numLeaves);
        mNumGeneratedNodes += 2;
        BuildNode(root, 0);

        mTarget.nodes.resize(mNumGeneratedNodes);
        outLeavesOrder = mLeavesOrder;
        return true;
    }

private:
    struct WorkSet
    {
        math::Box box;
        Indices leafIndices;
        uint32_t sortedBy = 0xFFFFFFFFu;
        uint32_t depth = 0;
    };

    void BuildNode(const WorkSet& ws, uint32_t nodeIndex)
    {
        const uint32_t n = (uint32_t)ws.leafIndices.size();
        {
            RtbBvhNode& node = mTarget.nodes[nodeIndex];
            node.min[0] = ws.box.min.x; node.min[1] = ws.box.min.y; node.min[2] = ws.box.min.z;
            node.max[0] = ws.box.max.x; node.max[1] = ws.box.max.y; node.max[2] = ws.box.max.z;
        }
        if (n <= mParams.maxLeafNodeSize)
        {
            RtbBvhNode& node = mTarget.nodes[nodeIndex];
            node.leafInfo = (node.leafInfo & 0xC0000000u) | n;
            node.childIndex = mNumGeneratedLeaves;
            for (uint32_t i = 0; i < n; ++i) mLeavesOrder.push_back(ws.leafIndices[i]);
            mNumGeneratedLeaves += n;
            return;
        }

        // sort along every axis except the one the items are already sorted by
        Indices sorted[3];
        for (uint32_t axis = 0; axis < 3; ++axis)
        {
            if (ws.sortedBy == axis) { sorted[axis] = ws.leafIndices; continue; }
            sorted[axis] = ws.leafIndices;
            const math::Box* boxes = mLeafBoxes;
            std::sort(sorted[axis].begin(), sorted[axis].end(), [boxes, axis](const uint32_t a, const uint32_t b)
            {
                const math::Vector4 ca = boxes[a].max + boxes[a].min;
                const math::Vector4 cb = boxes[b].max + boxes[b].min;
                return ca[(int)axis] < cb[(int)axis];
            });
        }

        uint32_t bestAxis = 0, bestSplitPos = 0;
        float bestCost = FLT_MAX;
        math::Box bestLeft = math::Box::Empty(), bestRight = math::Box::Empty();
        for (uint32_t axis = 0; axis < 3; ++axis)
        {
            const Indices& idx = sorted[axis];
            math::Box acc = math::Box::Empty();
            for (uint32_t i = 0; i < n; ++i) { acc = math::Box::Merge(acc, mLeafBoxes[idx[i]]); mLeft[i] = acc; }
            acc = math::Box::Empty();
            for (uint32_t i = n; i-- > 0;) { acc = math::Box::Merge(acc, mLeafBoxes[idx[i]]); mRight[i] = acc; }
            for (uint32_t splitPos = 0; splitPos < n - 1; ++splitPos)
            {
                const math::Box& l = mLeft[splitPos];
                const math::Box& r = mRight[splitPos + 1];
                float leftCost, rightCost;
                if (mParams.heuristics == BvhBuildingParams::Heuristics::SurfaceArea) { leftCost = l.SurfaceArea(); rightCost = r.SurfaceArea(); }
                else { leftCost = l.Volume(); rightCost = r.Volume(); }
                const uint32_t leftCount = splitPos + 1, rightCount = n - leftCount;
                const float totalCost = leftCost * static_cast<float>(leftCount) + rightCost * static_cast<float>(rightCount);
                if (totalCost < bestCost)
                {
                    bestCost = totalCost; bestAxis = axis; bestSplitPos = splitPos; bestLeft = l; bestRight = r;
                }
            }
        }

        const uint32_t leftCount = bestSplitPos + 1, rightCount = n - leftCount;
        const uint32_t leftNodeIndex = mNumGeneratedNodes;
        mNumGeneratedNodes += 2;
        {
            RtbBvhNode& node = mTarget.nodes[nodeIndex];
            node.childIndex = leftNodeIndex;
            node.leafInfo = bestAxis << 30;     // numLeaves = 0, splitAxis
        }

        WorkSet child;
        child.sortedBy = bestAxis;
        child.depth = ws.depth + 1;
        const Indices& s = sorted[bestAxis];
        Indices rightIndices(s.begin() + leftCount, s.end());
        child.box = bestLeft;
        child.leafIndices.assign(s.begin(), s.begin() + leftCount);
        sorted[0].clear(); sorted[0].shrink_to_fit();
        sorted[1].clear(); sorted[1].shrink_to_fit();
        sorted[2].clear(); sorted[2].shrink_to_fit();
        BuildNode(child, leftNodeIndex);
        child.box = bestRight;
        child.leafIndices = std::move(rightIndices);
        BuildNode(child, leftNodeIndex + 1);
        (void)rightCount;
    }

    BvhBuildingParams mParams;
    const math::Box* mLeafBoxes = nullptr;
    uint32_t mNumLeaves = 0;
    uint32_t mNumGeneratedNodes = 0;
    uint32_t mNumGeneratedLeaves = 0;
    Indices mLeavesOrder;
    std::vector<math::Box> mLeft, mRight;
    BVH& mTarget;
};

} // namespace rt
