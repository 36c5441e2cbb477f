// rtb_upload.h — host-side conversion of the flat RtbSceneDesc (reference layouts) into the device layouts of
// rtb_scene.h.  Pure host C++ (no CUDA calls): rtb_api.cu copies the resulting arrays to HBM.
#pragma once

#include <vector>
#include <string>
#include <cstring>
#include <algorithm>
#include <cstdlib>

#include "rtb_trace.h"
#include "rtb_wide_build.h"

namespace rtb {

struct HostMesh
{
    std::vector<DNode> nodes;
    std::vector<DTri> tris;
    std::vector<DVertexIndices> indices;
    std::vector<DVertexShading> shading;
    std::vector<uint32_t> materialIds;
    HostWide wide;                 // any-hit tree (rtb_wide.h); wide.nodes empty: not built, the binary walk is used
};

struct HostScene
{
    std::vector<DNode> tlas;
    std::vector<DObjectTrace> objTrace;
    std::vector<DObjectShade> objShade;
    std::vector<RtbShape> shapes;
    std::vector<HostMesh> meshes;
    std::vector<RtbMaterial> materials;
    std::vector<DLight> lights;
    std::vector<uint32_t> globalLights;
    std::vector<DTexture> textures;                      // texels pointers are patched after the device upload
    std::vector<const float*> texelSources;              // host texels of every bitmap texture (caller-owned), else null
    std::vector<RtbMaterialTextures> materialTextures;   // empty = no material has a texture
    std::vector<float> bokehCdf;                         // texture-shaped bokeh: bokehWidth * bokehHeight + 1 entries, or empty
    uint32_t bokehWidth = 0, bokehHeight = 0;
    bool hasLightTextures = false;
    uint32_t maxTlasDepth = 0, maxBlasDepth = 0, maxTlasLeaf = 0;
    uint32_t stackEntries = 0;      // worst-case occupancy of the traversal stack (rtb_trace.h)
};

inline M4 to_m4(const RtbMat4& m) { return load_m4(m); }

// depth of a reference-format BVH (BVH::CalculateStats, BVH.cpp:112-143); returns 0 on a malformed tree
inline uint32_t bvh_depth(const RtbBvhNode* nodes, uint32_t numNodes, uint32_t numLeafItems, std::string& err, uint32_t* outMaxLeaf = nullptr)
{
    if (numNodes == 0) return 0;
    if (numNodes > RTB_PACK_MAX_INDEX || numLeafItems > RTB_PACK_MAX_INDEX) { err = "BVH too large for the packed traversal stack (2^26 nodes / items)"; return 0; }
    struct Item { uint32_t node, depth; };
    std::vector<Item> stack;
    stack.push_back({ 0u, 1u });
    uint32_t maxDepth = 0, visited = 0;
    while (!stack.empty())
    {
        const Item it = stack.back();
        stack.pop_back();
        if (it.node >= numNodes) { err = "BVH child index out of range"; return 0; }
        if (++visited > numNodes) { err = "BVH has a cycle"; return 0; }
        const RtbBvhNode& n = nodes[it.node];
        if (it.depth > maxDepth) maxDepth = it.depth;
        const uint32_t numLeaves = n.leafInfo & 0x3FFFFFFFu;
        if (numLeaves != 0)
        {
            if ((uint64_t)n.childIndex + numLeaves > numLeafItems) { err = "BVH leaf range out of bounds"; return 0; }
            if (numLeaves > RTB_PACK_MAX_LEAF) { err = "BVH leaf with more than 62 items"; return 0; }
            if (outMaxLeaf && numLeaves > *outMaxLeaf) *outMaxLeaf = numLeaves;
        }
        else
        {
            stack.push_back({ n.childIndex, it.depth + 1 });
            stack.push_back({ n.childIndex + 1, it.depth + 1 });
        }
    }
    return maxDepth;
}

inline bool build_host_scene(const RtbSceneDesc& d, HostScene& out, std::string& err, uint32_t maxStackEntries = RTB_STACK_SIZE)
{
    out = HostScene();
    static_assert(sizeof(DNode) == sizeof(RtbBvhNode), "node layout");
    static_assert(sizeof(DTri) == 48, "triangle layout");

    out.tlas.resize(d.numTlasNodes);
    if (d.numTlasNodes) memcpy(out.tlas.data(), d.tlasNodes, sizeof(DNode) * d.numTlasNodes);
    out.shapes.assign(d.shapes, d.shapes + d.numShapes);
    out.materials.assign(d.materials, d.materials + d.numMaterials);
    out.globalLights.assign(d.globalLights, d.globalLights + d.numGlobalLights);

    for (uint32_t i = 0; i < d.numMaterials; ++i)
    {
        if (d.materials[i].bsdf > RTB_BSDF_ROUGH_PLASTIC) { err = "unknown BSDF id"; return false; }
        if ((d.materials[i].flags & 1u) && !d.materialTextures) { err = "material is flagged as textured but the scene carries no materialTextures"; return false; }
    }

    // textures: a mix texture may only reference EARLIER textures (so there are no cycles) and nests at most RTB_MAX_MIX_DEPTH deep
    if (d.numTextures && !d.textures) { err = "numTextures > 0 but textures is null"; return false; }
    out.textures.resize(d.numTextures);
    out.texelSources.assign(d.numTextures, nullptr);
    std::vector<uint32_t> mixDepth(d.numTextures, 0u);
    for (uint32_t i = 0; i < d.numTextures; ++i)
    {
        const RtbTexture& s = d.textures[i];
        DTexture& t = out.textures[i];
        memset(&t, 0, sizeof(t));
        t.type = s.type; t.width = s.width; t.height = s.height; t.numOctaves = s.numOctaves;
        t.texA = s.texA; t.texB = s.texB; t.texWeight = s.texWeight;
        t.colorA = v4(s.colorA[0], s.colorA[1], s.colorA[2], s.colorA[3]);
        t.colorB = v4(s.colorB[0], s.colorB[1], s.colorB[2], s.colorB[3]);
        switch (s.type)
        {
        case RTB_TEXTURE_BITMAP:
            if (s.width == 0 || s.height == 0 || s.width > 65536u || s.height > 65536u || !s.texels) { err = "bitmap texture without texels or with an invalid size"; return false; }
            out.texelSources[i] = s.texels;
            break;
        case RTB_TEXTURE_CHECKERBOARD:
            break;
        case RTB_TEXTURE_NOISE:
            if (s.numOctaves == 0 || s.numOctaves > 20) { err = "noise texture: numOctaves must be in 1..20"; return false; }     // NoiseTexture.cpp:51
            break;
        case RTB_TEXTURE_MIX:
            if (s.texA >= i || s.texB >= i || s.texWeight >= i) { err = "mix texture must reference earlier textures"; return false; }   // MixTexture.cpp:15-17 asserts non-null
            mixDepth[i] = 1u + std::max(mixDepth[s.texA], std::max(mixDepth[s.texB], mixDepth[s.texWeight]));
            if (mixDepth[i] > RTB_MAX_MIX_DEPTH) { err = "mix textures are nested too deep"; return false; }
            break;
        default:
            err = "unknown texture type"; return false;
        }
    }
    if (d.materialTextures)
    {
        bool any = false;
        for (uint32_t i = 0; i < d.numMaterials; ++i)
        {
            const RtbMaterialTextures& mt = d.materialTextures[i];
            const uint32_t ids[6] = { mt.baseColor, mt.emission, mt.roughness, mt.metalness, mt.normalMap, mt.maskMap };
            for (uint32_t id : ids)
            {
                if (id == RTB_NO_TEXTURE) continue;
                if (id >= d.numTextures) { err = "material texture index out of bounds"; return false; }
                any = true;
            }
        }
        if (any) out.materialTextures.assign(d.materialTextures, d.materialTextures + d.numMaterials);
    }

    if (d.bokehCdf)
    {
        if (d.bokehWidth == 0 || d.bokehHeight == 0 || d.bokehWidth > 65536u || d.bokehHeight > 65536u) { err = "bokeh distribution with an invalid size"; return false; }
        const size_t n = (size_t)d.bokehWidth * d.bokehHeight;
        out.bokehCdf.assign(d.bokehCdf, d.bokehCdf + n + 1);
        for (size_t i = 0; i < n; ++i) if (!(out.bokehCdf[i] <= out.bokehCdf[i + 1])) { err = "bokeh distribution is not a cumulative distribution"; return false; }
        out.bokehWidth = d.bokehWidth; out.bokehHeight = d.bokehHeight;
    }

    out.meshes.resize(d.numMeshes);
    for (uint32_t m = 0; m < d.numMeshes; ++m)
    {
        const RtbMesh& src = d.meshes[m];
        HostMesh& dst = out.meshes[m];
        dst.nodes.resize(src.numNodes);
        if (src.numNodes) memcpy(dst.nodes.data(), src.nodes, sizeof(DNode) * src.numNodes);
        dst.tris.resize(src.numTriangles);
        for (uint32_t t = 0; t < src.numTriangles; ++t)
        {
            const float* p = src.triangles + (size_t)t * 9;
            DTri& o = dst.tris[t];
            o.v0x = p[0]; o.v0y = p[1]; o.v0z = p[2]; o.pad0 = 0.0f;
            o.e1x = p[3]; o.e1y = p[4]; o.e1z = p[5]; o.pad1 = 0.0f;
            o.e2x = p[6]; o.e2y = p[7]; o.e2z = p[8]; o.pad2 = 0.0f;
        }
        dst.indices.resize(src.numTriangles);
        if (src.numTriangles) memcpy(dst.indices.data(), src.indices, sizeof(DVertexIndices) * src.numTriangles);
        dst.shading.resize(src.numVertices);
        if (src.numVertices) memcpy(dst.shading.data(), src.shading, sizeof(DVertexShading) * src.numVertices);
        dst.materialIds.assign(src.materialIds, src.materialIds + src.numMaterials);
        for (uint32_t t = 0; t < src.numTriangles; ++t)
        {
            const DVertexIndices& ix = dst.indices[t];
            if (ix.i0 >= src.numVertices || ix.i1 >= src.numVertices || ix.i2 >= src.numVertices) { err = "vertex index out of bounds"; return false; }
            if (ix.materialIndex != 0xFFFFFFFFu && ix.materialIndex >= src.numMaterials) { err = "mesh material index out of bounds"; return false; }
        }
        for (uint32_t id : dst.materialIds) if (id >= d.numMaterials) { err = "material id out of bounds"; return false; }
        const uint32_t depth = bvh_depth(src.nodes, src.numNodes, src.numTriangles, err);
        if (src.numNodes && depth == 0) return false;
        if (depth > out.maxBlasDepth) out.maxBlasDepth = depth;
        // any-hit tree; a mesh it cannot be built for keeps the binary walk.  Its stack: two entries per level
#if RTB_WIDE_SHADOW
        const char* leafEnv = getenv("RTB_WIDE_LEAF_TRIS");      // development knob: triangles per leaf child (1..4)
        if (!build_wide_bvh(dst.nodes, dst.tris, dst.wide, leafEnv ? (uint32_t)atoi(leafEnv) : RTB_WIDE_LEAF_TRIS)) dst.wide = HostWide();
        if (2u * dst.wide.depth > out.maxBlasDepth) out.maxBlasDepth = 2u * dst.wide.depth;
#endif
    }

    out.lights.resize(d.numLights);
    for (uint32_t i = 0; i < d.numLights; ++i)
    {
        const RtbLight& s = d.lights[i];
        DLight& l = out.lights[i];
        memset(&l, 0, sizeof(l));
        l.type = s.type; l.flags = s.flags; l.shapeIndex = s.shapeIndex; l.isDelta = s.isDelta;
        l.color = v4(s.color[0], s.color[1], s.color[2], s.color[3]);
        l.cosAngle = s.cosAngle;
        l.textureIndex = RTB_NO_TEXTURE;
        if (d.lightTextures && d.lightTextures[i] != RTB_NO_TEXTURE)
        {
            if (d.lightTextures[i] >= d.numTextures) { err = "light texture index out of bounds"; return false; }
            // only BackgroundLight evaluates its texture (BackgroundLight.cpp:45-61); AreaLight's is commented out (AreaLight.cpp:49-58, 136-146)
            if (s.type == RTB_LIGHT_BACKGROUND) { l.textureIndex = d.lightTextures[i]; out.hasLightTextures = true; }
        }
        l.transform = to_m4(s.transform);
        l.invTransform = to_m4(s.invTransform);
        if (s.type > RTB_LIGHT_SPOT) { err = "unknown light type"; return false; }
        if (s.type == RTB_LIGHT_AREA)
        {
            if (s.shapeIndex >= d.numShapes) { err = "area light shape out of bounds"; return false; }
            if (d.shapes[s.shapeIndex].type == RTB_SHAPE_MESH) { err = "mesh area lights are not supported (MeshShape::Sample is unimplemented in the reference)"; return false; }
        }
    }
    for (uint32_t g : out.globalLights) if (g >= d.numLights) { err = "global light index out of bounds"; return false; }

    out.objTrace.resize(d.numObjects);
    out.objShade.resize(d.numObjects);
    for (uint32_t i = 0; i < d.numObjects; ++i)
    {
        const RtbObject& s = d.objects[i];
        DObjectTrace& t = out.objTrace[i];
        DObjectShade& h = out.objShade[i];
        memset(&t, 0, sizeof(t));
        memset(&h, 0, sizeof(h));
        t.invTransform = to_m4(s.invTransform);
        t.kind = s.kind;
        t.shapeType = RTB_NO_SHAPE;
        t.meshIndex = 0xFFFFFFFFu;
        h.transform = to_m4(s.transform);
        h.kind = s.kind;
        h.shapeIndex = s.shapeIndex; h.materialIndex = s.materialIndex; h.lightIndex = s.lightIndex;
        if (s.kind == RTB_OBJECT_SHAPE)
        {
            if (s.shapeIndex >= d.numShapes) { err = "object shape out of bounds"; return false; }
            if (s.materialIndex >= d.numMaterials) { err = "object material out of bounds"; return false; }
        }
        else if (s.kind == RTB_OBJECT_LIGHT)
        {
            if (s.lightIndex >= d.numLights) { err = "object light out of bounds"; return false; }
        }
        else { err = "unknown object kind"; return false; }
        if (s.shapeIndex != 0xFFFFFFFFu)
        {
            if (s.shapeIndex >= d.numShapes) { err = "object shape out of bounds"; return false; }
            const RtbShape& sh = d.shapes[s.shapeIndex];
            t.shapeType = sh.type;
            t.meshIndex = sh.meshIndex;
            memcpy(t.size, sh.size, sizeof(t.size));
            if (sh.type == RTB_SHAPE_MESH && sh.meshIndex >= d.numMeshes) { err = "mesh index out of bounds"; return false; }
            if (sh.type > RTB_SHAPE_MESH) { err = "unknown shape type"; return false; }
        }
    }

    if (d.numObjects > 1)
    {
        const uint32_t depth = bvh_depth(d.tlasNodes, d.numTlasNodes, d.numObjects, err, &out.maxTlasLeaf);
        if (d.numTlasNodes && depth == 0) return false;
        out.maxTlasDepth = depth;
    }
    if (d.numObjects > RTB_PACK_MAX_INDEX) { err = "too many objects for the packed traversal stack"; return false; }
    // one pending sibling per level of either tree + the object entries of one top-level leaf (+ the 1-object bypass entry)
    out.stackEntries = out.maxTlasDepth + out.maxBlasDepth + out.maxTlasLeaf + 2;
    if (out.stackEntries > maxStackEntries)
    {
        err = "BVH too deep for the traversal stack";
        return false;
    }
    // the walks read ONE packed word per node (rtb_trace.h): index | count << 26, kept in the leafInfo slot of OUR copies
    // of the trees (the caller's arrays stay in the reference's format; everything above, the any-hit tree's collapse
    // included, has read that format by now)
    for (DNode& n : out.tlas) n.leafInfo = pack_node_word(n.childIndex, n.leafInfo);
    for (HostMesh& m : out.meshes) for (DNode& n : m.nodes) n.leafInfo = pack_node_word(n.childIndex, n.leafInfo);
    return true;
}

} // namespace rtb
