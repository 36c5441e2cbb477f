// rt_capi.cpp — C entry points of the host object model (librtb200_host.so) for Python tests / bench.py and other FFI
// users.  Thin: every function forwards to the C++ classes in rt_scene.h / rt_loader.h / rt_viewport.h, which in turn
// call the extern "C" rtb_* boundary of librtb200.so.
#include <cstring>
#include <exception>
#include <string>

#include "rt_loader.h"
#include "rt_viewport.h"

namespace {

struct SceneHandle
{
    rt::Scene scene;
    rt::Camera camera;
};

void SetErr(char* buf, size_t len, const std::string& msg)
{
    if (!buf || len == 0) return;
    strncpy(buf, msg.c_str(), len - 1);
    buf[len - 1] = 0;
}

} // namespace

extern "C" {

// helpers::LoadScene + Scene::BuildBVH (Demo/Demo.cpp:156-198 SwitchScene)
void* rtbh_scene_load(const char* jsonPath, const char* dataPath, char* err, size_t errLen)
{
    // no C++ exception may cross the C boundary (std::bad_alloc / std::length_error from a hostile file)
    SceneHandle* h = nullptr;
    try
    {
        h = new SceneHandle();
        std::string msg;
        if (!rt::helpers::LoadScene(jsonPath, dataPath ? dataPath : "", h->scene, h->camera, msg)) { SetErr(err, errLen, msg); delete h; return nullptr; }
        if (!h->scene.BuildBVH()) { SetErr(err, errLen, "Scene::BuildBVH failed"); delete h; return nullptr; }
        return h;
    }
    catch (const std::exception& e) { SetErr(err, errLen, std::string("scene load failed: ") + e.what()); delete h; return nullptr; }
    catch (...) { SetErr(err, errLen, "scene load failed: unknown exception"); delete h; return nullptr; }
}

void rtbh_scene_free(void* scene) { delete static_cast<SceneHandle*>(scene); }

// flat RTBS file of the scene + camera at the given aspect (what rtb_upload_scene receives)
int rtbh_scene_save_flat(void* scene, const char* path, uint32_t width, uint32_t height)
{
    try
    {
        SceneHandle* h = static_cast<SceneHandle*>(scene);
        rtb::SceneFile flat;
        h->scene.Flatten(flat);
        rt::Camera cam = h->camera;
        cam.SetPerspective((float)width / (float)height, cam.mFieldOfView);
        cam.Describe(flat.camera);
        cam.DescribeBokeh(flat);
        flat.fieldOfView = cam.mFieldOfView;
        return flat.Save(path) ? 0 : -1;
    }
    catch (...) { return -1; }
}

void rtbh_scene_stats(void* scene, uint32_t* numObjects, uint32_t* numLights, uint32_t* numTlasNodes)
{
    SceneHandle* h = static_cast<SceneHandle*>(scene);
    if (numObjects) *numObjects = (uint32_t)h->scene.GetNumTraceableObjects();
    if (numLights) *numLights = (uint32_t)h->scene.GetLights().size();
    if (numTlasNodes) *numTlasNodes = h->scene.GetBVH().GetNumNodes();
}

void* rtbh_viewport_create(void* scene, int deviceId, uint32_t numShards, uint32_t shardIndex, uint32_t flags, const char* blueNoisePath, void* cudaStream, char* err, size_t errLen)
{
    rt::ViewportB200* vp = nullptr;
    try
    {
        vp = new rt::ViewportB200();
        if (!vp->Create(deviceId, numShards, shardIndex, flags, blueNoisePath ? blueNoisePath : "", cudaStream)) { SetErr(err, errLen, vp->GetLastError()); delete vp; return nullptr; }
        if (scene && !vp->SetScene(static_cast<SceneHandle*>(scene)->scene, &static_cast<SceneHandle*>(scene)->camera)) { SetErr(err, errLen, vp->GetLastError()); delete vp; return nullptr; }
        return vp;
    }
    catch (const std::exception& e) { SetErr(err, errLen, std::string("viewport creation failed: ") + e.what()); delete vp; return nullptr; }
    catch (...) { SetErr(err, errLen, "viewport creation failed: unknown exception"); delete vp; return nullptr; }
}

void rtbh_viewport_free(void* vp) { delete static_cast<rt::ViewportB200*>(vp); }
const char* rtbh_viewport_last_error(void* vp) { return static_cast<rt::ViewportB200*>(vp)->GetLastError().c_str(); }
void* rtbh_viewport_context(void* vp) { return static_cast<rt::ViewportB200*>(vp)->GetContext(); }

int rtbh_viewport_set_renderer(void* vp, const char* name, uint32_t debugMode)
{
    rt::ViewportB200* v = static_cast<rt::ViewportB200*>(vp);
    v->SetDebugMode(debugMode);
    return v->SetRenderer(name) ? 0 : -1;
}

int rtbh_viewport_set_params(void* vp, uint32_t dimensions, uint32_t useBlueNoise, float aaSpread, uint32_t maxRayDepth, uint32_t minRRDepth, uint32_t strategy, uint64_t seed)
{
    rt::ViewportB200* v = static_cast<rt::ViewportB200*>(vp);
    rt::RenderingParams p = v->GetRenderingParams();
    p.dimensions = dimensions; p.useBlueNoiseDithering = useBlueNoise != 0; p.antiAliasingSpread = aaSpread;
    p.maxRayDepth = maxRayDepth; p.minRussianRouletteDepth = minRRDepth; p.lightSamplingStrategy = strategy;
    v->SetRenderingParams(p);
    v->SetSeed(seed);
    return 0;
}

// AdaptiveRenderingSettings (Context.h:35-43); takes effect at the next Reset / Resize like the reference's
int rtbh_viewport_set_adaptive(void* vp, uint32_t enable, uint32_t numInitialPasses, uint32_t minBlockSize, uint32_t maxBlockSize, float subdivisionTreshold, float convergenceTreshold)
{
    rt::ViewportB200* v = static_cast<rt::ViewportB200*>(vp);
    rt::RenderingParams p = v->GetRenderingParams();
    p.adaptiveSettings.enable = enable != 0; p.adaptiveSettings.numInitialPasses = numInitialPasses;
    p.adaptiveSettings.minBlockSize = minBlockSize; p.adaptiveSettings.maxBlockSize = maxBlockSize;
    p.adaptiveSettings.subdivisionTreshold = subdivisionTreshold; p.adaptiveSettings.convergenceTreshold = convergenceTreshold;
    v->SetRenderingParams(p);
    return 0;
}

// Viewport::GetProgress; out5 = passesFinished, activePixels, activeBlocks, converged (float bits), averageError (float bits)
int rtbh_viewport_progress(void* vp, uint32_t* out5)
{
    rt::RenderingProgress pr;
    if (!static_cast<rt::ViewportB200*>(vp)->GetProgress(pr)) return -1;
    out5[0] = pr.passesFinished; out5[1] = pr.activePixels; out5[2] = pr.activeBlocks;
    memcpy(&out5[3], &pr.converged, 4); memcpy(&out5[4], &pr.averageError, 4);
    return 0;
}

// the block list (Viewport::mBlocks): returns the number of blocks, copies at most `capacity`
uint32_t rtbh_viewport_blocks(void* vp, RtbBlock* out, uint32_t capacity)
{
    const std::vector<RtbBlock>& b = static_cast<rt::ViewportB200*>(vp)->GetBlocks();
    for (uint32_t i = 0; i < b.size() && i < capacity; ++i) out[i] = b[i];
    return (uint32_t)b.size();
}

// ViewportB200::VisualizeBlocks on caller data (Viewport::VisualizeActiveBlocks; host only)
void rtbh_visualize_blocks(uint32_t* bgra, uint32_t width, const RtbBlock* blocks, uint32_t numBlocks)
{
    rt::ViewportB200::VisualizeBlocks(bgra, width, blocks, numBlocks);
}

// ViewportB200::PassScript: seeds[numPasses * dimensions], sampleOffsets[numPasses * 2] (host prologue only, no device)
void rtbh_pass_script(uint64_t seed, uint32_t dimensions, float antiAliasingSpread, uint32_t width, uint32_t height, uint32_t extraResets, uint32_t numPasses, uint32_t* seeds, float* sampleOffsets)
{
    rt::ViewportB200::PassScript(seed, dimensions, antiAliasingSpread, width, height, extraResets, numPasses, seeds, sampleOffsets);
}

// ViewportB200::ReviseBlocks on caller data (no device): blocks[0..n) + errors[0..n) in, the revised list out (at most
// `capacity` entries are written; a revision at most doubles the list); returns its length
uint32_t rtbh_revise_blocks(RtbBlock* blocks, uint32_t n, uint32_t capacity, const float* errors, uint32_t minBlockSize, float subdivisionTreshold, float convergenceTreshold)
{
    rt::AdaptiveRenderingSettings settings;
    settings.minBlockSize = minBlockSize; settings.subdivisionTreshold = subdivisionTreshold; settings.convergenceTreshold = convergenceTreshold;
    std::vector<RtbBlock> list(blocks, blocks + n);
    rt::ViewportB200::ReviseBlocks(list, std::vector<float>(errors, errors + n), settings);
    for (uint32_t i = 0; i < list.size() && i < capacity; ++i) blocks[i] = list[i];
    return (uint32_t)list.size();
}

int rtbh_viewport_set_vcm(void* vp, uint32_t connection, uint32_t merging, uint32_t maxPathLength, float initialRadius, float minRadius, float radiusMultiplier, uint64_t seed)
{
    RtbVcmParams p;
    p.useVertexConnection = connection; p.useVertexMerging = merging; p.maxPathLength = maxPathLength;
    p.initialMergingRadius = initialRadius; p.minMergingRadius = minRadius; p.mergingRadiusMultiplier = radiusMultiplier; p.lightPathSeed = seed;
    return static_cast<rt::ViewportB200*>(vp)->SetVcmParams(p) ? 0 : -1;
}

int rtbh_viewport_resize(void* vp, uint32_t w, uint32_t h) { return static_cast<rt::ViewportB200*>(vp)->Resize(w, h) ? 0 : -1; }
void rtbh_viewport_reset(void* vp) { static_cast<rt::ViewportB200*>(vp)->Reset(); }

// Viewport::Render x numPasses with the scene's camera at the viewport's aspect ratio
int rtbh_viewport_render(void* vp, void* scene, uint32_t numPasses)
{
    rt::ViewportB200* v = static_cast<rt::ViewportB200*>(vp);
    rt::Camera cam = static_cast<SceneHandle*>(scene)->camera;
    cam.SetPerspective((float)v->GetWidth() / (float)v->GetHeight(), cam.mFieldOfView);
    return v->RenderPasses(cam, numPasses) ? 0 : -1;
}

int rtbh_viewport_synchronize(void* vp) { return static_cast<rt::ViewportB200*>(vp)->Synchronize() ? 0 : -1; }
int rtbh_viewport_read_sum(void* vp, float* rgb, int which) { return static_cast<rt::ViewportB200*>(vp)->GetSumBuffer(rgb, which) ? 0 : -1; }
// Viewport::SetPostprocessParams / GetFrontBuffer
int rtbh_viewport_set_postprocess(void* vp, const float* colorFilter3, float exposure, float contrast, float saturation, float ditheringStrength, float bloomFactor, uint32_t tonemapper)
{
    rt::PostprocessParams p;
    for (int i = 0; i < 3; ++i) p.colorFilter[i] = colorFilter3[i];
    p.exposure = exposure; p.contrast = contrast; p.saturation = saturation; p.ditheringStrength = ditheringStrength;
    p.bloomFactor = bloomFactor; p.tonemapper = tonemapper;
    return static_cast<rt::ViewportB200*>(vp)->SetPostprocessParams(p) ? 0 : -1;
}
int rtbh_viewport_read_front(void* vp, uint32_t* bgra) { return static_cast<rt::ViewportB200*>(vp)->GetFrontBuffer(bgra) ? 0 : -1; }
int rtbh_viewport_counters(void* vp, RtbCounters* out) { return static_cast<rt::ViewportB200*>(vp)->GetCounters(*out) ? 0 : -1; }

} // extern "C"
