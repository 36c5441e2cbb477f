// rtadapter.cpp - the reference-side adapter of INTEGRATION.md section 2, COMPILED against the reference:
//
//     helpers::LoadScene -> Scene::BuildBVH -> rt::ViewportB200 (this file) -> rtb_* (librtb200.so)
//
// next to the reference's own rt::Viewport in the SAME process, on the SAME Scene and Camera objects, and compares the
// two sum buffers bit for bit.  rt::ViewportB200 mirrors rt::Viewport's interface (Core/Rendering/Viewport.h:37-58) and
// consumes the reference's own HaltonSequence and Random exactly like Viewport::Render does (Viewport.cpp:200-205, 235),
// so a maintainer swaps the class the application instantiates (Demo/Demo.cpp:156-198) and nothing else.
//
// TEST INFRASTRUCTURE (oracle/): built by oracle/build_ref.sh into oracle/_ref/exact/rtadapter against the patched copy of
// the reference (-fno-access-control), linked with libCore.so and librtb200.so.  Run by tests/test_gpu_adapter.py.
//
//   rtadapter --scene S.json --data DIR --width W --height H --passes N [--renderer "Path Tracer MIS"] [--dims 128] [--depth 8]
#include "PCH.h"
#include "Demo.h"
#include "SceneLoader.h"
#include "../Core/Rendering/Renderer.h"
#include "../Core/Rendering/DebugRenderer.h"
#include "../Core/Rendering/VertexConnectionAndMerging.h"
#include "../Core/Rendering/Viewport.h"
#include "../Core/Rendering/Context.h"
#include "../Core/Scene/Scene.h"
#include "../Core/Scene/Camera.h"
#include "../Core/Scene/Object/SceneObject_Shape.h"
#include "../Core/Scene/Object/SceneObject_Light.h"
#include "../Core/Scene/Light/AreaLight.h"
#include "../Core/Scene/Light/BackgroundLight.h"
#include "../Core/Scene/Light/DirectionalLight.h"
#include "../Core/Scene/Light/PointLight.h"
#include "../Core/Scene/Light/SpotLight.h"
#include "../Core/Shapes/SphereShape.h"
#include "../Core/Shapes/BoxShape.h"
#include "../Core/Shapes/RectShape.h"
#include "../Core/Shapes/MeshShape.h"
#include "../Core/Material/Material.h"
#include "../Core/Textures/BitmapTexture.h"
#include "../Core/Textures/CheckerboardTexture.h"
#include "../Core/Textures/NoiseTexture.h"
#include "../Core/Textures/MixTexture.h"
#include "../Core/Utils/Bitmap.h"
#include "../Core/Math/Distribution.h"
#include "../Core/Math/SamplingHelpers.h"
#include "../Core/Sampling/GenericSampler.h"
#include "../Core/Sampling/HaltonSampler.h"
#include "../Core/Utils/Timer.h"

#include "rtb200.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

Options gOptions;

using namespace rt;
using namespace rt::math;

#include "ref_flatten.h"

namespace rt {

// RtbSceneDesc over a Flat (the pointers live as long as this object)
struct FlatView
{
    std::vector<RtbMesh> meshes;
    std::vector<RtbTexture> textures;
    RtbSceneDesc desc;
    explicit FlatView(const Flat& f)
    {
        memset(&desc, 0, sizeof(desc));
        meshes.resize(f.meshes.size());
        for (size_t i = 0; i < f.meshes.size(); ++i)
        {
            const Flat::Mesh& m = f.meshes[i];
            RtbMesh& v = meshes[i];
            v.numNodes = (uint32_t)m.nodes.size(); v.numTriangles = (uint32_t)(m.triangles.size() / 9);
            v.numVertices = m.numVertices; v.numMaterials = (uint32_t)m.materialIds.size();
            v.nodes = m.nodes.data(); v.triangles = m.triangles.data(); v.indices = m.indices.data(); v.shading = m.shading.data(); v.materialIds = m.materialIds.data();
        }
        desc.numTlasNodes = (uint32_t)f.tlas.size(); desc.tlasNodes = f.tlas.data();
        desc.numObjects = (uint32_t)f.objects.size(); desc.objects = f.objects.data();
        desc.numShapes = (uint32_t)f.shapes.size(); desc.shapes = f.shapes.data();
        desc.numMeshes = (uint32_t)meshes.size(); desc.meshes = meshes.data();
        desc.numMaterials = (uint32_t)f.materials.size(); desc.materials = f.materials.data();
        desc.numLights = (uint32_t)f.lights.size(); desc.lights = f.lights.data();
        desc.numGlobalLights = (uint32_t)f.globalLights.size(); desc.globalLights = f.globalLights.data();
        if (f.anyTexture)
        {
            textures.resize(f.textures.size());
            for (size_t i = 0; i < f.textures.size(); ++i)
            {
                textures[i] = f.textures[i].desc;
                textures[i].texels = f.textures[i].texels.empty() ? nullptr : f.textures[i].texels.data();
            }
            desc.numTextures = (uint32_t)textures.size(); desc.textures = textures.empty() ? nullptr : textures.data();
            desc.materialTextures = f.materialTextures.size() == f.materials.size() && !f.materials.empty() ? f.materialTextures.data() : nullptr;
            desc.lightTextures = f.lightTextures.size() == f.lights.size() && !f.lights.empty() ? f.lightTextures.data() : nullptr;
        }
    }
};

// rt::Viewport's interface (Viewport.h:37-58) on the B200 path
class ViewportB200
{
public:
    ViewportB200()
    {
        RtbCreateDesc d;
        memset(&d, 0, sizeof(d));
        if (rtb_create(&d, &mCtx) != RTB_OK) Die(rtb_last_error());
        // the blue noise table the reference loads itself (GenericSampler.cpp:13-49): same 128 x 128 x 4 u16 layout
        FILE* f = fopen("../Data/BlueNoise128_RGBA16.dat", "rb");
        if (f)
        {
            std::vector<uint16_t> table(128 * 128 * 4);
            if (fread(table.data(), 2, table.size(), f) == table.size()) rtb_set_blue_noise(mCtx, table.data());
            fclose(f);
        }
    }
    ~ViewportB200() { rtb_destroy(mCtx); }

    bool Resize(uint32 width, uint32 height)
    {
        mWidth = width; mHeight = height;
        // Viewport::Resize draws one Vector4 per pixel from its generator for a salt table (Viewport.cpp:99-102)
        for (uint32 i = 0; i < width * height; ++i) mRandomGenerator.GetVector4();
        if (rtb_resize(mCtx, width, height) != RTB_OK) return false;
        Reset();                                                                // Viewport::Resize ends with Reset() (Viewport.cpp:104)
        return true;
    }
    bool SetRenderingParams(const RenderingParams& params)
    {
        mParams = params;          // (the Halton sequence is initialised by Reset, as in Viewport::Reset, Viewport.cpp:126)
        return true;
    }
    bool SetRenderer(const std::string& name, const Scene& scene)
    {
        if (name == "Path Tracer") mRendererId = RTB_RENDERER_PATH_TRACER;
        else if (name == "Path Tracer MIS") mRendererId = RTB_RENDERER_PATH_TRACER_MIS;
        else if (name == "Debug") mRendererId = RTB_RENDERER_DEBUG;
        else if (name == "VCM") mRendererId = RTB_RENDERER_VCM;
        else return false;
        Flat flat;
        flat.Build(scene);
        FlatView view(flat);
        return rtb_upload_scene(mCtx, &view.desc) == RTB_OK;
    }
    void Reset()
    {
        rtb_reset(mCtx);
        mPassesFinished = 0;
        mHaltonSequence.Initialize(mParams.samplingParams.dimensions);
    }
    bool Render(const Camera& camera)
    {
        mHaltonSequence.NextSample();                                                         // Viewport.cpp:200-205
        std::vector<uint32_t> seed(mHaltonSequence.GetNumDimensions());
        for (uint32 i = 0; i < (uint32)seed.size(); ++i) seed[i] = mHaltonSequence.GetInt(i);
        const Vector4 u = SamplingHelpers::GetFloatNormal2(mRandomGenerator.GetFloat2());   // Viewport.cpp:235
        const Vector4 offset = u * mParams.antiAliasingSpread;
        RtbPassDesc p;
        memset(&p, 0, sizeof(p));
        p.renderer = mRendererId; p.debugMode = RTB_DEBUG_TRIANGLE_ID;
        p.camera.localToWorld = ToMat(camera.mLocalToWorld);
        p.camera.aspectRatio = camera.mAspectRatio; p.camera.tanHalfFoV = camera.mTanHalfFoV;
        p.camera.dofEnable = camera.mDOF.enable; p.camera.bokehShape = (uint32_t)camera.mDOF.bokehShape;
        p.camera.aperture = camera.mDOF.aperture; p.camera.focalPlaneDistance = camera.mDOF.focalPlaneDistance;
        p.camera.barrelDistortionConstFactor = camera.barrelDistortionConstFactor;
        p.camera.barrelDistortionVariableFactor = camera.barrelDistortionVariableFactor;
        p.maxRayDepth = mParams.maxRayDepth; p.minRussianRouletteDepth = mParams.minRussianRouletteDepth;
        p.lightSamplingStrategy = mParams.lightSamplingStrategy == LightSamplingStrategy::All ? RTB_LIGHT_SAMPLING_ALL : RTB_LIGHT_SAMPLING_SINGLE;
        p.useBlueNoise = mParams.samplingParams.useBlueNoiseDithering ? 1u : 0u;
        p.passIndex = mPassesFinished++; p.numSeedDims = (uint32_t)seed.size(); p.seed = seed.data();
        p.sampleOffset[0] = offset.x; p.sampleOffset[1] = offset.y;
        return rtb_render_pass(mCtx, &p) == RTB_OK;
    }
    bool GetSumBuffer(float* rgb) { return rtb_read_sum(mCtx, rgb, 0) == RTB_OK; }
    bool GetCounters(RtbCounters& c) { return rtb_get_counters(mCtx, &c) == RTB_OK; }

    // Both classes seed their generators from the process-wide entropy stream when they are constructed, so two objects in
    // one process start from different seeds.  For a comparison the adapter takes over the seeds of a freshly constructed
    // rt::Viewport (the viewport's generator and the generator inside its Halton sequence, which Initialize() draws the
    // permutations from); from there on each object runs its own code on its own state.
    void SeedLike(const Viewport& vp) { mRandomGenerator = vp.mRandomGenerator; mHaltonSequence.mRandom = vp.mHaltonSequence.mRandom; }

private:
    RtbContext* mCtx = nullptr;
    uint32_t mRendererId = RTB_RENDERER_PATH_TRACER_MIS, mPassesFinished = 0;
    uint32 mWidth = 0, mHeight = 0;
    RenderingParams mParams;
    Random mRandomGenerator;
    HaltonSequence mHaltonSequence;
};

} // namespace rt

struct Args
{
    std::map<std::string, std::string> kv;
    std::string Str(const char* k, const char* d = "") const { auto it = kv.find(k); return it == kv.end() ? d : it->second; }
    long Int(const char* k, long d) const { auto it = kv.find(k); return it == kv.end() ? d : atol(it->second.c_str()); }
};

int main(int argc, char** argv)
{
    Args a;
    for (int i = 1; i + 1 < argc; i += 2) if (!strncmp(argv[i], "--", 2)) a.kv[argv[i] + 2] = argv[i + 1];
    const uint32 w = (uint32)a.Int("width", 64), h = (uint32)a.Int("height", 64), passes = (uint32)a.Int("passes", 2);
    const std::string rendererName = a.Str("renderer", "Path Tracer MIS");
    Scene scene; Camera camera;
    gOptions.dataPath = a.Str("data", "");
    if (!helpers::LoadScene(a.Str("scene"), scene, camera)) Die("scene load failed");
    if (!scene.BuildBVH()) Die("BuildBVH failed");
    camera.SetPerspective((float)w / (float)h, camera.mFieldOfView);

    RenderingParams params;
    params.maxRayDepth = (uint32)a.Int("depth", 8);
    params.samplingParams.dimensions = (uint32)a.Int("dims", 128);
    params.lightSamplingStrategy = LightSamplingStrategy::All;

    // ---- the reference's own viewport (CPU) and the adapter, constructed side by side, same seeds ---------------------------
    Viewport vp;
    ViewportB200 vpb;
    vpb.SeedLike(vp);

    vp.SetRenderingParams(params);
    vp.Resize(w, h);
    RendererPtr renderer = CreateRenderer(rendererName, scene);
    if (!renderer) Die("unknown renderer");
    vp.SetRenderer(renderer);
    vp.Reset();
    Timer cpuTimer;
    RayTracingCounters rc;      // Viewport::GetCounters holds the LAST pass' counters (Viewport.cpp:423-430): add them up like the adapter's
    rc.numRays = rc.numShadowRays = rc.numShadowRaysHit = rc.numPrimaryRays = 0;
    for (uint32 i = 0; i < passes; ++i)
    {
        vp.Render(camera);
        const RayTracingCounters& c = vp.GetCounters();
        rc.numRays += c.numRays; rc.numShadowRays += c.numShadowRays; rc.numShadowRaysHit += c.numShadowRaysHit; rc.numPrimaryRays += c.numPrimaryRays;
    }
    const double cpuSeconds = cpuTimer.Stop();
    const Bitmap& sum = vp.GetSumBuffer();

    // ---- the same calls on the adapter ---------------------------------------------------------------------------------------
    if (!vpb.SetRenderingParams(params)) Die("SetRenderingParams");
    if (!vpb.Resize(w, h)) Die(rtb_last_error());
    if (!vpb.SetRenderer(rendererName, scene)) Die(rtb_last_error());
    vpb.Reset();
    Timer gpuTimer;
    for (uint32 i = 0; i < passes; ++i) if (!vpb.Render(camera)) Die(rtb_last_error());
    std::vector<float> mine((size_t)w * h * 3);
    if (!vpb.GetSumBuffer(mine.data())) Die(rtb_last_error());
    const double gpuSeconds = gpuTimer.Stop();
    RtbCounters c;
    vpb.GetCounters(c);

    // ---- bit-for-bit ------------------------------------------------------------------------------------------------------------
    uint64_t different = 0;
    for (uint32 y = 0; y < h; ++y)
        for (uint32 x = 0; x < w; ++x)
        {
            const Float3& r = sum.GetPixelRef<Float3>(x, y);
            const float* m = &mine[((size_t)y * w + x) * 3];
            if (memcmp(&r.x, &m[0], 4) || memcmp(&r.y, &m[1], 4) || memcmp(&r.z, &m[2], 4)) different++;
        }
    printf("{\"width\": %u, \"height\": %u, \"passes\": %u, \"renderer\": \"%s\", \"differentPixels\": %llu, "
           "\"refRays\": %llu, \"refShadowRays\": %llu, \"refShadowRaysHit\": %llu, \"rays\": %llu, \"shadowRays\": %llu, \"shadowRaysHit\": %llu, "
           "\"cpuSeconds\": %.4f, \"adapterSeconds\": %.4f}\n", w, h, passes, rendererName.c_str(), (unsigned long long)different,
           (unsigned long long)rc.numRays, (unsigned long long)rc.numShadowRays, (unsigned long long)rc.numShadowRaysHit,
           (unsigned long long)c.numRays, (unsigned long long)c.numShadowRays, (unsigned long long)c.numShadowRaysHit, cpuSeconds, gpuSeconds);
    return different == 0 && rc.numRays == c.numRays && rc.numShadowRays == c.numShadowRays && rc.numShadowRaysHit == c.numShadowRaysHit ? 0 : 3;
}
