/*
 * rtb200.h — C ABI of the B200-native path-tracing hot path.
 *
 * Drop-in boundary for the reference's (Witek902/Raytracer) per-pixel integrator loop and the
 * two-level BVH ray/scene intersection it drives.  Every entry point below replaces one piece of the
 * reference's C++ interface (file:line relative to the reference checkout):
 *
 *   rtb_create / rtb_destroy      Viewport::Viewport / ~Viewport          Core/Rendering/Viewport.cpp:21-28
 *   rtb_upload_scene              Scene::BuildBVH result + CreateRenderer Core/Scene/Scene.cpp:36-126, Core/Rendering/Renderer.cpp:45-69
 *   rtb_resize                    Viewport::Resize                        Core/Rendering/Viewport.cpp:52-115
 *   rtb_reset                     Viewport::Reset                         Core/Rendering/Viewport.cpp:117-138
 *   rtb_render_pass(es)           Viewport::Render (one sample per pixel) Core/Rendering/Viewport.cpp:185-289
 *   rtb_read_sum                  Viewport::GetSumBuffer                  Core/Rendering/Viewport.h:50
 *   rtb_get_counters              Viewport::GetCounters                   Core/Rendering/Viewport.h:58, Core/Rendering/Counters.h:36-93
 *   rtb_trace_batch               Scene::Traverse / Scene::Traverse_Shadow Core/Scene/Scene.cpp:219-261
 *   rtb_sample_batch              GenericSampler::ResetPixel/GetInt       Core/Sampling/GenericSampler.cpp:75-112
 *   rtb_primary_rays              Camera::GenerateRay via RenderTile      Core/Scene/Camera.cpp:81-118, Core/Rendering/Viewport.cpp:320-330
 *
 * All structs are POD, all pointers are caller-owned HOST pointers that are copied during the call
 * (unless a name says "device").  Every function returns 0 on success or a negative RtbStatus;
 * rtb_last_error() returns a human readable message for the calling thread's last failure.
 * There is no CPU fallback: without a CUDA device rtb_create fails with RTB_ERR_NO_DEVICE.
 * A context is not thread-safe; calls are synchronous unless stated otherwise.
 */
#ifndef RTB200_H
#define RTB200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RTB_API_VERSION 5   /* 2: textures (RtbTexture, RtbMaterialTextures, RtbSceneDesc::textures ...)  3: RTB_RENDERER_VCM, rtb_set_vcm_params, rtb_vcm_*  4: RTB_RENDERER_LIGHT_TRACER  5: rtb_scatter_shard_async */

typedef enum RtbStatus
{
    RTB_OK                = 0,
    RTB_ERR_INVALID_ARG   = -1,
    RTB_ERR_NO_DEVICE     = -2,
    RTB_ERR_CUDA          = -3,
    RTB_ERR_NO_SCENE      = -4,
    RTB_ERR_UNSUPPORTED   = -5,
    RTB_ERR_OUT_OF_MEMORY = -6
} RtbStatus;

/* sentinel ids, HitPoint.h:8-9 */
#define RTB_INVALID_OBJECT 0xFFFFFFFFu
#define RTB_LIGHT_OBJECT   0xFFFFFFFEu

/* ---- scene description (flattened rt::Scene after BuildBVH) ------------------------------------ */

/* BVH::Node, Core/BVH/BVH.h:22-30, verbatim 32-byte layout.  leafInfo = numLeaves | splitAxis << 30 */
typedef struct RtbBvhNode
{
    float    min[3];
    uint32_t childIndex;
    float    max[3];
    uint32_t leafInfo;
} RtbBvhNode;

/* math::Matrix4, four rows of four floats (row-vector convention, Core/Math/Matrix4.h:110-139) */
typedef struct RtbMat4 { float r[4][4]; } RtbMat4;

typedef enum RtbShapeType
{
    RTB_SHAPE_SPHERE = 0,   /* Core/Shapes/SphereShape.cpp */
    RTB_SHAPE_BOX    = 1,   /* Core/Shapes/BoxShape.cpp    */
    RTB_SHAPE_RECT   = 2,   /* Core/Shapes/RectShape.cpp   */
    RTB_SHAPE_MESH   = 3    /* Core/Shapes/MeshShape.cpp   */
} RtbShapeType;

/*
 * sphere: size[0] = radius, invSize[0] = 1/radius
 * box   : size = half extents (w = 0), invSize = 1/size (w = 0), aux.xyz = face cdf (BoxShape.cpp:97-111)
 * rect  : size.xy = half extents, aux.xy = texture scale
 * mesh  : meshIndex selects RtbSceneDesc::meshes[]
 */
typedef struct RtbShape
{
    uint32_t type;
    uint32_t meshIndex;
    uint32_t pad0, pad1;
    float    size[4];
    float    invSize[4];
    float    aux[4];
} RtbShape;

typedef enum RtbBsdfType
{
    RTB_BSDF_NULL = 0, RTB_BSDF_DIFFUSE, RTB_BSDF_ROUGH_DIFFUSE, RTB_BSDF_DIELECTRIC, RTB_BSDF_ROUGH_DIELECTRIC,
    RTB_BSDF_METAL, RTB_BSDF_ROUGH_METAL, RTB_BSDF_PLASTIC, RTB_BSDF_ROUGH_PLASTIC
} RtbBsdfType;

/* rt::Material constant parameters (Core/Material/Material.h:44-75); the textures that modulate them are bound by the
 * parallel RtbSceneDesc::materialTextures array */
typedef struct RtbMaterial
{
    uint32_t bsdf;
    uint32_t flags;
    float    roughness;
    float    metalness;
    float    baseColor[4];
    float    emission[4];
    float    IoR;
    float    K;
    float    pad0, pad1;
} RtbMaterial;

/* ---- textures (ITexture, Core/Textures/Texture.h:14-34) ---------------------------------------------- */

typedef enum RtbTextureType
{
    RTB_TEXTURE_BITMAP       = 0,  /* BitmapTexture,       Core/Textures/BitmapTexture.cpp:32-93 (wrap + smooth-step bilinear) */
    RTB_TEXTURE_CHECKERBOARD = 1,  /* CheckerboardTexture, Core/Textures/CheckerboardTexture.cpp:32-41 */
    RTB_TEXTURE_NOISE        = 2,  /* NoiseTexture,        Core/Textures/NoiseTexture.cpp:59-162 (2-D simplex noise octaves) */
    RTB_TEXTURE_MIX          = 3   /* MixTexture,          Core/Textures/MixTexture.cpp:24-31 */
} RtbTextureType;

#define RTB_NO_TEXTURE 0xFFFFFFFFu
#define RTB_MAX_MIX_DEPTH 4u       /* a mix texture may nest other mix textures this deep */

/*
 * bitmap      : texels = width*height RGBA float quadruples, row-major, texel (x, y) at texels[4*(y*width + x)], holding
 *               EXACTLY what Bitmap::GetPixel(x, y) returns (Core/Utils/Bitmap.cpp:335-518): the storage format is decoded
 *               and the sRGB -> linear conversion of gamma-space bitmaps is already applied by the host
 * checkerboard: colorA, colorB
 * noise       : colorA, colorB, numOctaves
 * mix         : texA, texB, texWeight = indices of EARLIER entries of RtbSceneDesc::textures
 */
typedef struct RtbTexture
{
    uint32_t type;           /* RtbTextureType */
    uint32_t width, height;
    uint32_t numOctaves;
    uint32_t texA, texB, texWeight;
    uint32_t pad0;
    float    colorA[4];
    float    colorB[4];
    const float* texels;
} RtbTexture;

/* the optional textures of one rt::Material (Material.h:63-77, MaterialParameter.h:12-33); RTB_NO_TEXTURE = none */
typedef struct RtbMaterialTextures
{
    uint32_t baseColor;          /* value = baseValue * texture (4 lanes) */
    uint32_t emission;
    uint32_t roughness;          /* value = baseValue * texture.x */
    uint32_t metalness;
    uint32_t normalMap;          /* Material::GetNormalVector, Material.cpp:120-138; Scene.cpp:328-337 */
    uint32_t maskMap;            /* carried for completeness: Material::GetMaskValue has no caller in the reference */
    float    normalMapStrength;
    uint32_t pad0;
} RtbMaterialTextures;

typedef enum RtbLightType
{
    RTB_LIGHT_AREA = 0, RTB_LIGHT_BACKGROUND, RTB_LIGHT_DIRECTIONAL, RTB_LIGHT_POINT, RTB_LIGHT_SPOT
} RtbLightType;

#define RTB_LIGHT_FLAG_FINITE 1u   /* ILight::Flag_IsFinite, Light.h:41 */
#define RTB_LIGHT_FLAG_DELTA  2u   /* ILight::Flag_IsDelta,  Light.h:42 */

/* one LightSceneObject (+ its ILight), in Scene::GetLights() order (Scene.cpp:43-49) */
typedef struct RtbLight
{
    uint32_t type;
    uint32_t flags;
    uint32_t shapeIndex;     /* area lights: RtbSceneDesc::shapes[] */
    uint32_t isDelta;        /* directional / spot */
    float    color[4];
    float    cosAngle;       /* directional / spot: cosf(angle) computed by the host */
    float    pad0, pad1, pad2;
    RtbMat4  transform;      /* light -> world */
    RtbMat4  invTransform;   /* world -> light */
} RtbLight;

#define RTB_OBJECT_SHAPE 0u
#define RTB_OBJECT_LIGHT 1u

/* one traceable object, in top-level BVH leaf order (Scene.cpp:87-94): objectId == index here */
typedef struct RtbObject
{
    uint32_t kind;           /* RTB_OBJECT_SHAPE | RTB_OBJECT_LIGHT */
    uint32_t shapeIndex;     /* shapes[] (for lights: the area light's shape, or UINT32_MAX) */
    uint32_t materialIndex;  /* default material of a shape object */
    uint32_t lightIndex;     /* lights[] for RTB_OBJECT_LIGHT */
    RtbMat4  transform;
    RtbMat4  invTransform;
} RtbObject;

/* one MeshShape (+VertexBuffer), triangles already in BVH leaf order (MeshShape.cpp:82-95) */
typedef struct RtbMesh
{
    uint32_t numNodes;
    uint32_t numTriangles;
    uint32_t numVertices;
    uint32_t numMaterials;
    const RtbBvhNode* nodes;        /* numNodes                                               */
    const float*      triangles;    /* numTriangles * 9: v0, edge1, edge2 (Triangle.h:25-41)   */
    const uint32_t*   indices;      /* numTriangles * 4: i0 i1 i2 materialIndex (VertexBuffer.h:18-24) */
    const float*      shading;      /* numVertices * 8: normal3 tangent3 uv2 (VertexBuffer.h:26-31)    */
    const uint32_t*   materialIds;  /* numMaterials: index into RtbSceneDesc::materials        */
} RtbMesh;

typedef struct RtbSceneDesc
{
    uint32_t numTlasNodes;   const RtbBvhNode*  tlasNodes;
    uint32_t numObjects;     const RtbObject*   objects;
    uint32_t numShapes;      const RtbShape*    shapes;
    uint32_t numMeshes;      const RtbMesh*     meshes;
    uint32_t numMaterials;   const RtbMaterial* materials;
    uint32_t numLights;      const RtbLight*    lights;
    uint32_t numGlobalLights; const uint32_t*   globalLights;  /* indices into lights[], Scene::GetGlobalLights() order */
    /* textures: all optional (numTextures = 0, null pointers = a scene without textures) */
    uint32_t numTextures;    const RtbTexture*  textures;
    const RtbMaterialTextures* materialTextures;   /* numMaterials entries, or NULL */
    const uint32_t*            lightTextures;      /* numLights entries (BackgroundLight::mTexture, BackgroundLight.cpp:45-61;
                                                      RTB_NO_TEXTURE = none), or NULL */
    /* texture-shaped bokeh (Camera::mDOF.bokehTexture, Camera.cpp:208-213): the importance map of the bokeh bitmap as
     * BitmapTexture::MakeSamplable builds it (BitmapTexture.cpp:121-152) - the normalised cumulative distribution of
     * Distribution::Initialize (Math/Distribution.cpp:27-84), bokehWidth * bokehHeight + 1 floats; NULL = none */
    uint32_t bokehWidth, bokehHeight;
    const float* bokehCdf;
} RtbSceneDesc;

/* ---- per-pass description ----------------------------------------------------------------------- */

typedef enum RtbRenderer
{
    RTB_RENDERER_PATH_TRACER     = 0,  /* "Path Tracer"     Core/Rendering/PathTracer.cpp    */
    RTB_RENDERER_PATH_TRACER_MIS = 1,  /* "Path Tracer MIS" Core/Rendering/PathTracerMIS.cpp */
    RTB_RENDERER_DEBUG           = 2,  /* "Debug"           Core/Rendering/DebugRenderer.cpp */
    RTB_RENDERER_VCM             = 3,  /* "VCM"             Core/Rendering/VertexConnectionAndMerging.cpp (rtb_set_vcm_params) */
    RTB_RENDERER_LIGHT_TRACER    = 4   /* "Light Tracer"    Core/Rendering/LightTracer.cpp: one light sub-path per pixel and pass, every vertex connected
                                          to the camera and splatted (Film::AccumulateColor(pos, ...)); draws from a generator seeded by
                                          (pixel, pass, RtbVcmParams::lightPathSeed) where the reference uses per-thread generators:
                                          reproducible here, statistical against the reference; rendered one pass at a time like VCM */
} RtbRenderer;

/* DebugRenderingMode, Core/Rendering/DebugRenderer.h:7-33 (same numeric values) */
typedef enum RtbDebugMode
{
    RTB_DEBUG_CAMERA_LIGHT = 0, RTB_DEBUG_TRIANGLE_ID, RTB_DEBUG_DEPTH, RTB_DEBUG_POSITION, RTB_DEBUG_NORMALS,
    RTB_DEBUG_TANGENTS, RTB_DEBUG_BITANGENTS, RTB_DEBUG_TEXCOORDS, RTB_DEBUG_BASE_COLOR, RTB_DEBUG_EMISSION,
    RTB_DEBUG_ROUGHNESS, RTB_DEBUG_METALNESS, RTB_DEBUG_IOR
} RtbDebugMode;

/* BokehShape, Core/Scene/Camera.h:21-28 (same numeric values).  RTB_BOKEH_TEXTURE samples the lens point from
 * RtbSceneDesc::bokehCdf; RTB_BOKEH_NGON is refused (not implemented in the reference either, Camera.cpp:205-207) */
typedef enum RtbBokehShape { RTB_BOKEH_CIRCLE = 0, RTB_BOKEH_HEXAGON = 1, RTB_BOKEH_SQUARE = 2, RTB_BOKEH_NGON = 3, RTB_BOKEH_TEXTURE = 4 } RtbBokehShape;

/* rt::Camera state consumed by Camera::GenerateRay (Camera.cpp:81-118) */
typedef struct RtbCamera
{
    RtbMat4  localToWorld;
    float    aspectRatio;
    float    tanHalfFoV;        /* tanf(fov/2) computed by the host, Camera.cpp:37 */
    uint32_t dofEnable;
    uint32_t bokehShape;
    float    aperture;
    float    focalPlaneDistance;
    float    barrelDistortionConstFactor;
    float    barrelDistortionVariableFactor;   /* must be 0: the reference draws it from a per-thread RNG */
} RtbCamera;

#define RTB_LIGHT_SAMPLING_SINGLE 0u
#define RTB_LIGHT_SAMPLING_ALL    1u

typedef struct RtbPassDesc
{
    uint32_t        renderer;               /* RtbRenderer */
    uint32_t        debugMode;              /* RtbDebugMode, RTB_RENDERER_DEBUG only */
    RtbCamera       camera;
    uint32_t        maxRayDepth;            /* RenderingParams::maxRayDepth            Context.h:69 */
    uint32_t        minRussianRouletteDepth;/* RenderingParams::minRussianRouletteDepth Context.h:72 */
    uint32_t        lightSamplingStrategy;  /* RTB_LIGHT_SAMPLING_*                    Context.h:82 */
    uint32_t        useBlueNoise;           /* SamplingParams::useBlueNoiseDithering   Context.h:52 */
    uint32_t        passIndex;              /* Viewport progress.passesFinished: even passes also feed secondarySum */
    uint32_t        numSeedDims;            /* SamplingParams::dimensions              Context.h:49 */
    const uint32_t* seed;                   /* numSeedDims Halton integers of this pass, Viewport.cpp:200-205 */
    float           sampleOffset[2];        /* AA jitter of this pass, Viewport.cpp:235-242 */
} RtbPassDesc;

/* RayTracingCounters, Core/Rendering/Counters.h:36-41 (+ traversal work counters, Counters.h:8-33) */
typedef struct RtbCounters
{
    uint64_t numRays;
    uint64_t numShadowRays;
    uint64_t numShadowRaysHit;
    uint64_t numPrimaryRays;
    uint64_t numRayBoxTests;        /* closest-hit rays; only filled when the context was created with RTB_FLAG_COUNT_TESTS */
    uint64_t numRayTriangleTests;
    uint64_t numShadowBoxTests;     /* shadow rays, same flag */
    uint64_t numShadowTriangleTests;
    uint64_t numKernelLaunches;     /* CUDA kernels launched by rtb_render_pass(es) since the last reset */
} RtbCounters;

/* ---- batch test hooks ---------------------------------------------------------------------------- */

/* math::Ray as Scene::Traverse sees it (Ray.h:13-28): 4 x float4 */
typedef struct RtbRay
{
    float origin[4];
    float dir[4];
    float invDir[4];
    float originDivDir[4];
} RtbRay;

/* HitPoint, Core/Traversal/HitPoint.h:14-51 (distance doubles as tMax for shadow queries) */
typedef struct RtbHit
{
    uint32_t objectId;
    uint32_t subObjectId;
    float    distance;
    float    u;
    float    v;
    uint32_t occluded;    /* shadow mode: 1 if Scene::Traverse_Shadow returned true */
} RtbHit;

#define RTB_TRACE_CLOSEST 0u
#define RTB_TRACE_SHADOW  1u

/* ---- context ------------------------------------------------------------------------------------ */

#define RTB_FLAG_COUNT_TESTS 1u   /* per-ray box/triangle test counters (roofline algorithmic bytes) */

typedef struct RtbCreateDesc
{
    int32_t  deviceId;        /* CUDA device ordinal */
    uint32_t flags;
    /* multi-GPU tile sharding: this context renders tiles t with t % numShards == shardIndex (1/0 = whole frame) */
    uint32_t numShards;
    uint32_t shardIndex;
    uint32_t tileSize;        /* 0 -> 32 (RenderingParams::tileSize, Context.h:76) */
    void*    stream;          /* cudaStream_t to launch on; NULL -> the library's stream for that device, shared by every
                                 context created this way, so that their kernels never overlap each other */
} RtbCreateDesc;

typedef struct RtbContext RtbContext;

const char* rtb_last_error(void);
int  rtb_api_version(void);

int  rtb_create(const RtbCreateDesc* desc, RtbContext** outCtx);
void rtb_destroy(RtbContext* ctx);

int  rtb_upload_scene(RtbContext* ctx, const RtbSceneDesc* scene);
int  rtb_set_blue_noise(RtbContext* ctx, const uint16_t* data /* 128*128*4 u16, Data/BlueNoise128_RGBA16.dat */);

int  rtb_resize(RtbContext* ctx, uint32_t width, uint32_t height);
int  rtb_reset(RtbContext* ctx);

/* enqueue one pass (asynchronous on the context's stream); rtb_synchronize / rtb_read_* wait for it */
int  rtb_render_pass(RtbContext* ctx, const RtbPassDesc* pass);
/* `count` passes at once: consecutive passes that differ only in (seed, sampleOffset, passIndex + 1) are traced in ONE
 * wavefront (more rays per launch, fewer launches); the film receives them in pass order, so the result is bit-identical
 * to `count` calls of rtb_render_pass */
int  rtb_render_passes(RtbContext* ctx, const RtbPassDesc* passes, uint32_t count);
int  rtb_synchronize(RtbContext* ctx);
/* Passes are enqueued on the context's wavefront stream(s).  What ships is ONE wavefront slot: a context runs one kernel
 * at a time, on the stream of RtbCreateDesc (or one shared stream per device if none was given); with RTB_SLOTS = 2..4 in
 * the environment several wavefronts are in flight on streams of their own (the film still receives them in pass order;
 * bit-identical, tested with 2 and 4 slots; measured gain: 2 % device resident).  Everything in this API that looks at the result (rtb_read_*, rtb_unshard_*, rtb_get_counters,
 * rtb_film_device_buffer, rtb_reset, rtb_synchronize) orders itself after them.  rtb_join does only that: it makes the
 * main stream (RtbCreateDesc::stream) wait, on the device, for every pass enqueued so far - for callers that go on to
 * enqueue their own work on that stream.  It does not block the host. */
int  rtb_join(RtbContext* ctx);

/* which: 0 = sum, 1 = secondarySum.  rgb = width*height*3 floats, row 0 = top (Film.cpp:31-39) */
int  rtb_read_sum(RtbContext* ctx, float* rgb, int which);
/* device pointer + element count of this shard's compact film (for an NCCL all-gather by the caller) */
int  rtb_film_device_buffer(RtbContext* ctx, int which, void** outDevicePtr, uint64_t* outNumFloats);
/* Streaming frames: with a snapshot buffer of `maxPasses` entries (owned by the context, freed by maxPasses = 0 or
 * rtb_resize), pass i of every following rtb_render_passes call (i < maxPasses) also leaves the film (sum) AS OF THAT
 * PASS at outDevicePtr + i * strideFloats, in the layout of rtb_film_device_buffer - what Viewport::GetSumBuffer would
 * return after that pass (Viewport.h:50) - without giving up the shared wavefronts of a batched call.  Each entry can
 * be all-gathered / handed to rtb_unshard_film(_async) like the film itself. */
int  rtb_film_snapshots(RtbContext* ctx, uint32_t maxPasses, void** outDevicePtr, uint64_t* outStrideFloats);
/* scatter gathered shard buffers (numShards * shardFloats, device) back to raster order into hostRgb */
int  rtb_unshard_film(RtbContext* ctx, const void* gatheredDevice, float* hostRgb);
/* asynchronous variants: the film is snapshotted in stream order (so later passes do not leak into it) and copied to
 * the host on a second stream while the next passes render; the data is complete after rtb_synchronize.  At most twelve
 * read-backs are in flight (the thirteenth waits for the first).  The host buffer should be page-locked (rtb_host_alloc). */
int  rtb_read_sum_async(RtbContext* ctx, float* rgb, int which);
int  rtb_unshard_film_async(RtbContext* ctx, const void* gatheredDevice, float* hostRgb);
/* ---- multi-GPU frames (the reference renders its tiles on a thread pool, Viewport.cpp:216-268) -------
 * Every shard context owns tile t % numShards of the frame.  Two ways to put a frame together:
 *
 * (1) host-destined frames: rtb_scatter_shard writes THIS shard's pixels of the film (or of a film snapshot / any buffer in
 *     the layout of rtb_film_device_buffer) into a raster frame `rasterDst` - device memory, or page-locked host memory that
 *     is mapped into this device (rtb_host_alloc / rtb_host_register; pass the pointer rtb_host_device_pointer returns).
 *     All shards write into the same frame, each over its own PCIe link, no collective and no funnel through one GPU.
 *     Stream ordered; complete after rtb_synchronize.
 * (2) device-resident frames: rtb_comm_init gives the context an NCCL communicator (the library loads libnccl.so.2 at that
 *     point; one rank per shard, rank = shardIndex), rtb_finish_frame all-gathers the shard films once per frame on the
 *     context's stream and returns the gathered buffer (numShards * shardFloats) that rtb_unshard_film / rtb_read_front /
 *     rtb_block_errors accept.  For VCM it first all-reduces the splat buffers and folds them into the film.
 *     rtb_comm_init_all is the single-process flavour (ncclCommInitAll over the contexts' devices). */
int  rtb_scatter_shard(RtbContext* ctx, const void* shardFilmDevice /* NULL: own film */, int which, void* rasterDst);
/* the same for a film SNAPSHOT (rtb_film_snapshots), on the context's second stream: the stores cross PCIe while the passes
 * enqueued after this call render (the renderer's next k_film waits for the scatter before it overwrites the snapshots).
 * Complete after rtb_synchronize.  What a progressive multi-GPU display calls once per frame of a group of passes. */
int  rtb_scatter_shard_async(RtbContext* ctx, const void* shardFilmSnapshotDevice, int which, void* rasterDst);
#define RTB_COMM_ID_BYTES 128
int  rtb_comm_unique_id(void* outId /* RTB_COMM_ID_BYTES */);                                   /* ncclGetUniqueId: rank 0 calls it and hands the bytes to the others */
int  rtb_comm_init(RtbContext* ctx, const void* id /* RTB_COMM_ID_BYTES */);                  /* ncclCommInitRank(numShards, id, shardIndex) */
int  rtb_comm_init_all(RtbContext** ctxs, uint32_t count);                                    /* one process, `count` = numShards contexts on different devices */
int  rtb_finish_frame(RtbContext* ctx, int which, void** outGatheredDevice);
int  rtb_finish_frame_all(RtbContext** ctxs, uint32_t count, int which, void** outGatheredDevice);   /* after rtb_comm_init_all: all contexts' collectives in one NCCL group */
/* host memory the devices can write into directly: rtb_host_alloc'ed memory is mapped already; rtb_host_register pins and
 * maps memory the caller owns (e.g. a shared-memory segment several rank processes attach to) */
int  rtb_host_register(void* p, uint64_t bytes);
int  rtb_host_unregister(void* p);
void* rtb_host_device_pointer(void* p);

/* page-locked host memory for the rgb / hostRgb arguments above (the reference's Viewport owns its sum Bitmap,
 * Viewport.h:121; an adapter that wants full-speed read-back allocates that bitmap here).  Any host pointer works,
 * pinned ones copy at PCIe rate.  NULL on failure (rtb_last_error). */
void* rtb_host_alloc(uint64_t bytes);
void  rtb_host_free(void* p);

/* Viewport::SetPostprocessParams / GetFrontBuffer (Viewport.h:43,49; PostprocessParams, Rendering/PostProcess.h:11-29):
 * the tone-mapped 8-bit image of the film.  Tonemapper values are the reference's (Color/ColorHelpers.h:77-83). */
typedef enum RtbTonemapper
{
    RTB_TONEMAP_CLAMPED  = 0,
    RTB_TONEMAP_REINHARD = 1,
    RTB_TONEMAP_JIM_HEJL = 2,
    RTB_TONEMAP_ACES     = 3
} RtbTonemapper;

typedef struct RtbPostprocessDesc
{
    float    colorFilter[4];     /* defaults of PostprocessParams::PostprocessParams (PostProcess.cpp:6-14): 1,1,1,1 */
    float    exposure;           /* log2 scale, 0 */
    float    contrast;           /* 0.8 */
    float    saturation;         /* 0.98 */
    float    ditheringStrength;  /* 0.005; 0 = none.  The reference dithers from per-thread generators (Viewport.cpp:494,545:
                                    not reproducible across thread counts); here the noise is a hash of (pixel, ditherSeed) */
    float    bloomFactor;        /* 0 = off (the reference's default) */
    uint32_t tonemapper;         /* RtbTonemapper; ACES */
    uint32_t numPasses;          /* passes accumulated in the film: the image is sum / numPasses (Viewport.cpp:499) */
    uint32_t ditherSeed;
} RtbPostprocessDesc;

/* Viewport::PerformPostProcess + PostProcessTile (Viewport.cpp:432-552): film (sum) -> bgra[y*width + x] = 0x00RRGGBB,
 * row 0 = top.  `gatheredDevice` NULL: this context's own film (pixels of other shards stay 0), else the all-gathered
 * shard films as for rtb_unshard_film.  Synchronous. */
int  rtb_read_front(RtbContext* ctx, const RtbPostprocessDesc* post, const void* gatheredDevice, uint32_t* bgra);

/* Adaptive rendering (AdaptiveRenderingSettings, Rendering/Context.h:35-43; Viewport.cpp:552-732).  The block list and its
 * subdivision policy stay with the caller (host/rt_viewport.h mirrors Viewport::UpdateBlocksList); the device side is the
 * error estimate of a list of blocks and the restriction of the following passes to a list of tiles. */
typedef struct RtbBlock { uint32_t minX, maxX, minY, maxY; } RtbBlock;     /* math::Rectangle<uint32>, Viewport.h:64 */

/* errors[i] = Viewport::ComputeBlockError(blocks[i]) (Viewport.cpp:552-582) from sum and secondarySum after `numPasses`
 * (even) passes; blocks may overlap (ComputeError passes the whole image, Viewport.cpp:179-183).  `gatheredSum` /
 * `gatheredSum2` NULL: this context's film (numShards must be 1: an estimate sums whole rows of a block), else the
 * all-gathered films as for rtb_unshard_film.  Synchronous. */
int  rtb_block_errors(RtbContext* ctx, const void* gatheredSum, const void* gatheredSum2, uint32_t numPasses,
                      const RtbBlock* blocks, uint32_t numBlocks, float* errors);
/* The following passes sample only the pixels inside `tiles` (Viewport::GenerateRenderingTiles, Viewport.cpp:584-616 -
 * the reference renders exactly its tile list, Viewport.cpp:261); the other pixels keep their sums.  tiles == NULL: the
 * whole image again (also after rtb_resize).  numTiles == 0 with tiles != NULL: nothing is left to render. */
int  rtb_set_active_tiles(RtbContext* ctx, const RtbBlock* tiles, uint32_t numTiles);

int  rtb_get_counters(RtbContext* ctx, RtbCounters* out);   /* counters of the passes since the last reset */

int  rtb_trace_batch(RtbContext* ctx, const RtbRay* rays, uint64_t n, RtbHit* hits, uint32_t mode);
/* device-resident variant used by bench.py for the kernel-only number; returns kernel milliseconds */
int  rtb_trace_batch_device(RtbContext* ctx, const void* raysDevice, uint64_t n, void* hitsDevice, uint32_t mode, float* outMs);

/* rgba[4i..] = ITexture::Evaluate of textures[textureIndex] at (uv[2i], uv[2i+1])   (test hook, Core/Textures/Texture.h:24) */
int  rtb_texture_batch(RtbContext* ctx, uint32_t textureIndex, const float* uv, uint64_t n, float* rgba);
/* out[i*numDims + d] = d-th GenericSampler::GetInt() of pixel (xy[2i], xy[2i+1]) */
int  rtb_sample_batch(RtbContext* ctx, const RtbPassDesc* pass, const uint32_t* xy, uint64_t n, uint32_t numDims, uint32_t* out);
/* rays[y*width + x] = primary ray of film pixel (x, y) for this pass */
int  rtb_primary_rays(RtbContext* ctx, const RtbPassDesc* pass, RtbRay* rays);

/* per-kernel device time of the last rendered pass (CUDA events), for bench.py's roofline */
typedef struct RtbPassTimings
{
    float raygenMs, traverseMs, shadeMs, shadowMs, totalMs;
    uint32_t traverseLaunches, shadeLaunches, shadowLaunches, totalLaunches;
    uint64_t traverseRays, shadowRays;
} RtbPassTimings;
int  rtb_set_profiling(RtbContext* ctx, int enable);

/* ---- "VCM": vertex connection and merging (Core/Rendering/VertexConnectionAndMerging.h:31-58) -------------------------
 * The public members of the reference's renderer object; defaults as its constructor sets them
 * (VertexConnectionAndMerging.cpp:47-62).  Passes with renderer = RTB_RENDERER_VCM are rendered one at a time: a pass traces
 * one light sub-path and one camera path per pixel, merges with the photons of the PREVIOUS pass (passIndex - 1 of the
 * same frame) and splats the light paths' camera connections onto the film with float atomics.  The camera path draws
 * from the pass' Halton seed like the path tracers; the light path draws from a generator seeded with lightPathSeed, the
 * pixel and the pass (the reference uses its per-thread generators there): the image is reproducible here, comparable
 * with the reference statistically.  maxRayDepth / minRussianRouletteDepth / lightSamplingStrategy of the pass are not
 * used by this renderer (nor by the reference's). */
typedef struct RtbVcmParams
{
    uint32_t useVertexConnection;      /* mUseVertexConnection */
    uint32_t useVertexMerging;         /* mUseVertexMerging */
    uint32_t maxPathLength;            /* mMaxPathLength (1 .. 32) */
    float    initialMergingRadius;     /* mInitialMergingRadius */
    float    minMergingRadius;         /* mMinMergingRadius */
    float    mergingRadiusMultiplier;  /* mMergingRadiusMultiplier (0 < m <= 1) */
    uint64_t lightPathSeed;
} RtbVcmParams;
int  rtb_set_vcm_params(RtbContext* ctx, const RtbVcmParams* params);
/* Sharded contexts (numShards > 1): a light path's camera connection can land on any pixel, so every shard collects the
 * splats of ITS light paths in two raster buffers (width * height * 3 floats each: sum, secondary sum) until the frame is
 * read.  The caller sums them over the shards (one all-reduce per frame) in place and calls rtb_vcm_fold_splats, which adds
 * this shard's pixels to its film and clears the buffers.  Single-shard contexts fold after every pass by themselves.
 * Vertex merging needs every shard's photons: not built - a sharded context refuses useVertexMerging. */
int  rtb_vcm_splat_buffers(RtbContext* ctx, void** outSplat, void** outSplat2, uint64_t* outNumFloats);
int  rtb_vcm_fold_splats(RtbContext* ctx);
int  rtb_get_pass_timings(RtbContext* ctx, RtbPassTimings* out);

#ifdef __cplusplus
}
#endif
#endif /* RTB200_H */
