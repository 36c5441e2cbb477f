// rtb_scene.h — device-resident scene layout (what rtb_upload_scene builds from RtbSceneDesc).
//
// Layout for HBM / L2 (see DESIGN.md "Data layout"):
//   * BVH nodes stay in the reference's 32-byte format (BVH.h:22-30) and are read as two 16-byte words; children
//     are allocated in adjacent pairs at even indices (BVHBuilder.cpp:203-206), so one inner-node step is ONE
//     aligned 64-byte fetch (4 x LDG.128) of both child boxes.
//   * triangles are padded from the reference's unaligned 36 bytes (Triangle.h:25-41) to 48 bytes = 3 x float4
//     {v0, edge1, edge2}; same values, aligned 128-bit loads.
//   * per-object data is split into what traversal touches (DObjectTrace: inverse transform + shape header,
//     96 B) and what shading touches (DObjectShade).
#pragma once

#include "rtb_math.h"
#include "../../include/rtb200.h"

namespace rtb {

struct __attribute__((aligned(16))) DNode   // == RtbBvhNode
{
    float minx, miny, minz; uint32_t childIndex;
    float maxx, maxy, maxz; uint32_t leafInfo;    // numLeaves : 30, splitAxis : 2
};

struct __attribute__((aligned(16))) DTri { float v0x, v0y, v0z, pad0, e1x, e1y, e1z, pad1, e2x, e2y, e2z, pad2; };

#define RTB_NO_SHAPE 0xFFFFFFFFu
#define RTB_NO_MATERIAL 0xFFFFFFFFu

// traversal-time view of one traceable object (objectId == index)
struct __attribute__((aligned(16))) DObjectTrace
{
    M4       invTransform;     // world -> local
    uint32_t kind;             // RTB_OBJECT_SHAPE / RTB_OBJECT_LIGHT
    uint32_t shapeType;        // RtbShapeType, RTB_NO_SHAPE for lights that cannot be hit (point / spot)
    uint32_t meshIndex;
    uint32_t pad;
    float    size[4];          // sphere: radius; box: half extents; rect: half extents.xy
};

struct __attribute__((aligned(16))) DObjectShade
{
    M4       transform;        // local -> world
    uint32_t kind;
    uint32_t shapeIndex;
    uint32_t materialIndex;
    uint32_t lightIndex;
};

struct __attribute__((aligned(16))) DVertexIndices { uint32_t i0, i1, i2, materialIndex; };      // VertexBuffer.h:18-24
struct __attribute__((aligned(16))) DVertexShading { float nx, ny, nz, tx, ty, tz, u, v; };      // VertexBuffer.h:26-31

struct WNode;      // rtb_wide.h

struct DMesh
{
    const DNode*          nodes;
    const DTri*           tris;
    // any-hit acceleration structure (rtb_wide.h); wideNodes == nullptr: the mesh only has the reference's binary tree
    const WNode*          wideNodes;
    const DTri*           wideTris;
    float                 wideMaxAbs;
    uint32_t              numWideNodes;
    const DVertexIndices* indices;
    const DVertexShading* shading;
    const uint32_t*       materialIds;
    uint32_t              numNodes;
    uint32_t              numTriangles;
    uint32_t              numVertices;
    uint32_t              numMaterials;
};

struct __attribute__((aligned(16))) DLight
{
    uint32_t type, flags, shapeIndex, isDelta;
    V4       color;
    float    cosAngle;
    uint32_t textureIndex;   // background light: environment map (RTB_NO_TEXTURE = none)
    float    pad1, pad2;
    M4       transform;      // light -> world
    M4       invTransform;   // world -> light
};

// one ITexture (rtb_texture.h); bitmap texels = width*height RGBA float quadruples in HBM, row-major: one bilinear
// lookup is 4 aligned 16-byte loads
struct __attribute__((aligned(16))) DTexture
{
    uint32_t type, width, height, numOctaves;
    uint32_t texA, texB, texWeight, pad0;
    V4       colorA, colorB;
    const float* texels;
    uint64_t pad1;
};

struct SceneView
{
    const DNode*        tlas;
    const DObjectTrace* objTrace;
    const DObjectShade* objShade;
    const RtbShape*     shapes;
    const DMesh*        meshes;
    const RtbMaterial*  materials;
    const DLight*       lights;
    const uint32_t*     globalLights;
    const uint16_t*     blueNoise;       // 128*128*4, may be null
    const DTexture*     textures;        // null without textures
    const RtbMaterialTextures* materialTextures;   // per material, null without textured materials
    uint32_t            numTlasNodes;
    uint32_t            numObjects;
    uint32_t            numLights;
    uint32_t            numGlobalLights;
};

RTB_HD M4 load_m4(const RtbMat4& m)
{
    M4 r;
    r.r0 = v4(m.r[0][0], m.r[0][1], m.r[0][2], m.r[0][3]);
    r.r1 = v4(m.r[1][0], m.r[1][1], m.r[1][2], m.r[1][3]);
    r.r2 = v4(m.r[2][0], m.r[2][1], m.r[2][2], m.r[2][3]);
    r.r3 = v4(m.r[3][0], m.r[3][1], m.r[3][2], m.r[3][3]);
    return r;
}

} // namespace rtb
