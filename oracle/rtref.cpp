// rtref — headless harness around the UNMODIFIED algorithms of the reference (Witek902/Raytracer).
//
// TEST INFRASTRUCTURE ONLY (parity oracle + CPU timing baseline).  It links the patched copy of the reference
// built by oracle/build_ref.sh and is compiled with -fno-access-control so it can read private state of
// rt::Scene / rt::Viewport / shapes (Scene.h:83-92, SceneObject_Shape.h:29-31, MeshShape.h:53-64 ...) without
// editing reference sources.  Nothing in the product links or executes this file.
//
// Sub-commands (all take --scene X.json [--data DIR]):
//   render   Viewport::Render loop -> sum buffer, secondary sum, per-pass seed/AA script, counters
//   flat     flatten the reference's Scene (after BuildBVH) + Camera into the RtbSceneDesc file format ("RTBS")
//   trace    Scene::Traverse / Scene::Traverse_Shadow over a file of RtbRay -> file of RtbHit
//   sampler  GenericSampler streams for given pixels of given passes
//   primary  Camera::GenerateRay for every pixel exactly as Viewport::RenderTile calls it
//   texeval  ITexture::Evaluate of every texture of the scene at a list of (u, v)
//   emit     ILight::Emit of every light of the scene for a list of (position sample x3, direction sample x2)
//   illuminate ILight::Illuminate without solid-angle sampling (as VCM calls it) for a list of shading points
//   radiance ILight::GetRadiance without solid-angle sampling, with both pdf outputs, for a list of light-space rays
//   bsdfpdf  BSDF::Pdf (forward and reverse) of every material of the scene for a list of direction pairs
//   camera   Camera::WorldToFilm and Camera::PdfW for a list of world-space points
//   writeexr a seeded random HDR image written by the reference's own EXR writer (tinyexr) - test input for EXR readers
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
#include "../Core/Traversal/TraversalContext.h"
#include "../Core/Math/SamplingHelpers.h"
#include "../Core/Sampling/GenericSampler.h"
#include "../Core/Utils/Timer.h"
#include "../Core/Utils/Memory.h"
#include "../External/tinyexr/tinyexr.h"

#include "rtb200.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>
#include <thread>
#include <execinfo.h>
#include <signal.h>
#include <unistd.h>

Options gOptions;

using namespace rt;
using namespace rt::math;

namespace {

struct Args
{
    std::map<std::string, std::string> kv;
    std::string cmd;

    bool Has(const char* k) const { return kv.count(k) > 0; }
    std::string Str(const char* k, const char* def = "") const { auto it = kv.find(k); return it == kv.end() ? def : it->second; }
    long Int(const char* k, long def) const { auto it = kv.find(k); return it == kv.end() ? def : atol(it->second.c_str()); }
    double Dbl(const char* k, double def) const { auto it = kv.find(k); return it == kv.end() ? def : atof(it->second.c_str()); }
};

Args ParseArgs(int argc, char** argv)
{
    Args a;
    if (argc > 1) a.cmd = argv[1];
    for (int i = 2; i < argc; ++i)
    {
        std::string k = argv[i];
        if (k.rfind("--", 0) == 0)
        {
            k = k.substr(2);
            if (i + 1 < argc && strncmp(argv[i + 1], "--", 2) != 0) { a.kv[k] = argv[++i]; }
            else { a.kv[k] = "1"; }
        }
    }
    return a;
}

#include "ref_flatten.h"

bool LoadSceneFile(const Args& a, Scene& scene, Camera& camera, uint32 w, uint32 h)
{
    gOptions.dataPath = a.Str("data", "");
    if (!helpers::LoadScene(a.Str("scene"), scene, camera)) { fprintf(stderr, "rtref: scene load failed\n"); return false; }
    if (!scene.BuildBVH()) return false;
    camera.SetPerspective((float)w / (float)h, camera.mFieldOfView);
    // camera state the JSON schema cannot express (Camera.h:43, 99): set on the reference's Camera object directly
    if (a.Has("bokeh")) camera.mDOF.bokehShape = (BokehShape)a.Int("bokeh", 0);
    if (a.Has("barrel")) camera.barrelDistortionConstFactor = (float)a.Dbl("barrel", 0.01);
    return true;
}

RenderingParams MakeParams(const Args& a)
{
    RenderingParams p;
    const long threads = a.Int("threads", 0);
    p.numThreads = threads > 0 ? (uint32)threads : std::thread::hardware_concurrency();
    p.maxRayDepth = (uint32)a.Int("depth", 8);
    p.minRussianRouletteDepth = (uint32)a.Int("rr-depth", 1);
    p.samplingParams.dimensions = (uint32)a.Int("dims", 128);
    p.samplingParams.useBlueNoiseDithering = a.Int("blue-noise", 1) != 0;
    p.lightSamplingStrategy = a.Str("strategy", "all") == "single" ? LightSamplingStrategy::Single : LightSamplingStrategy::All;
    p.antiAliasingSpread = (float)a.Dbl("aa-spread", 0.5);
    p.adaptiveSettings.enable = a.Int("adaptive", 0) != 0;
    p.adaptiveSettings.numInitialPasses = (uint32)a.Int("adaptive-initial", p.adaptiveSettings.numInitialPasses);
    p.adaptiveSettings.minBlockSize = (uint32)a.Int("adaptive-min-block", p.adaptiveSettings.minBlockSize);
    p.adaptiveSettings.maxBlockSize = (uint32)a.Int("adaptive-max-block", p.adaptiveSettings.maxBlockSize);
    p.adaptiveSettings.subdivisionTreshold = (float)a.Dbl("adaptive-subdivision", p.adaptiveSettings.subdivisionTreshold);
    p.adaptiveSettings.convergenceTreshold = (float)a.Dbl("adaptive-convergence", p.adaptiveSettings.convergenceTreshold);
    return p;
}

int CmdRender(const Args& a)
{
    const uint32 w = (uint32)a.Int("width", 256), h = (uint32)a.Int("height", 256), passes = (uint32)a.Int("passes", 1);
    const std::string rendererName = a.Str("renderer", "Path Tracer MIS");
    Scene scene; Camera camera;
    if (!LoadSceneFile(a, scene, camera, w, h)) return 1;

    Viewport vp;
    const RenderingParams params = MakeParams(a);
    vp.SetRenderingParams(params);
    vp.Resize(w, h);
    RendererPtr renderer = CreateRenderer(rendererName, scene);
    if (!renderer) Die("unknown renderer");
    if (rendererName == "Debug") static_cast<DebugRenderer*>(renderer.get())->mRenderingMode = (DebugRenderingMode)a.Int("debug-mode", 1);
    if (rendererName == "VCM")
    {
        // --vcm-connection 0|1 --vcm-merging 0|1: the renderer's own switches (VertexConnectionAndMerging.h)
        VertexConnectionAndMerging* vcm = static_cast<VertexConnectionAndMerging*>(renderer.get());
        vcm->mUseVertexConnection = a.Int("vcm-connection", 1) != 0;
        vcm->mUseVertexMerging = a.Int("vcm-merging", 1) != 0;
    }
    vp.SetRenderer(renderer);
    if (a.Has("dump-front"))
    {
        // dithering draws from the per-thread generators (Viewport.cpp:494): off unless asked for
        PostprocessParams pp;
        pp.ditheringStrength = (float)a.Dbl("pp-dither", 0.0);
        pp.exposure = (float)a.Dbl("pp-exposure", pp.exposure);
        pp.contrast = (float)a.Dbl("pp-contrast", pp.contrast);
        pp.saturation = (float)a.Dbl("pp-saturation", pp.saturation);
        pp.bloomFactor = (float)a.Dbl("pp-bloom", pp.bloomFactor);
        pp.tonemapper = (Tonemapper)a.Int("pp-tonemapper", (long)pp.tonemapper);
        if (a.Has("pp-filter"))
        {
            float r = 1.0f, g = 1.0f, b = 1.0f;
            if (sscanf(a.Str("pp-filter").c_str(), "%f,%f,%f", &r, &g, &b) != 3) Die("--pp-filter r,g,b");
            pp.colorFilter = Vector4(r, g, b, 1.0f);
        }
        vp.SetPostprocessParams(pp);
    }
    vp.Reset();

    // --dump-progress: after every Render(): RenderingProgress (Viewport.h:25-32) + the block list (Viewport.h:126)
    std::vector<math::Rectangle<uint32>> previousBlocks(vp.mBlocks.begin(), vp.mBlocks.end());
    FILE* progress = a.Has("dump-progress") ? fopen(a.Str("dump-progress").c_str(), "wb") : nullptr;
    FILE* script = a.Has("dump-passes") ? fopen(a.Str("dump-passes").c_str(), "wb") : nullptr;
    if (script) { WriteU32(script, passes); WriteU32(script, params.samplingParams.dimensions); }

    const uint32 warmup = (uint32)a.Int("warmup", 0);
    uint64 rays = 0, shadow = 0, shadowHit = 0, primary = 0;
    double seconds = 0.0;
    for (uint32 i = 0; i < passes + warmup; ++i)
    {
        // same draw Viewport::Render makes from its private generator (Viewport.cpp:235) - taken from a COPY so the
        // real stream is untouched
        Random rngCopy = vp.mRandomGenerator;
        const Vector4 u = SamplingHelpers::GetFloatNormal2(rngCopy.GetFloat2());
        const Vector4 sampleOffset = u * params.antiAliasingSpread;

        Timer t;
        vp.Render(camera);
        const double dt = t.Stop();

        if (script && i >= warmup)
        {
            for (uint32 d = 0; d < params.samplingParams.dimensions; ++d) WriteU32(script, vp.mHaltonSequence.GetInt(d));
            WritePod(script, sampleOffset.x);
            WritePod(script, sampleOffset.y);
        }
        if (progress)
        {
            const RenderingProgress& pr = vp.GetProgress();
            WriteU32(progress, pr.passesFinished); WriteU32(progress, pr.activePixels); WriteU32(progress, pr.activeBlocks);
            WritePod(progress, pr.converged); WritePod(progress, pr.averageError);
            WriteU32(progress, vp.mBlocks.Size());
            for (const auto& b : vp.mBlocks) { WriteU32(progress, b.minX); WriteU32(progress, b.maxX); WriteU32(progress, b.minY); WriteU32(progress, b.maxY); }
            // what Viewport::UpdateBlocksList saw inside this Render(): the film is as it was then, the blocks are those
            // of the previous pass -> their ComputeBlockError values (only defined after an even number of passes)
            const bool even = pr.passesFinished > 0 && pr.passesFinished % 2 == 0;
            WriteU32(progress, even ? (uint32)previousBlocks.size() : 0u);
            if (even) for (const auto& b : previousBlocks) WritePod(progress, vp.ComputeBlockError(b));
            previousBlocks.assign(vp.mBlocks.begin(), vp.mBlocks.end());
        }
        if (i >= warmup)
        {
            seconds += dt;
            const RayTracingCounters& c = vp.GetCounters();
            rays += c.numRays; shadow += c.numShadowRays; shadowHit += c.numShadowRaysHit; primary += c.numPrimaryRays;
        }
    }
    if (script) fclose(script);
    if (progress) fclose(progress);

    if (a.Has("dump-sum"))
    {
        FILE* f = fopen(a.Str("dump-sum").c_str(), "wb");
        WriteBytes(f, vp.GetSumBuffer().GetData(), (size_t)w * h * 12);
        fclose(f);
    }
    if (a.Has("block-errors"))
    {
        // --block-errors blocks.bin (u32 minX, maxX, minY, maxY each) --dump-block-errors out.f32: Viewport::ComputeBlockError
        FILE* f = fopen(a.Str("block-errors").c_str(), "rb");
        if (!f) Die("cannot open --block-errors");
        uint32 r[4];
        std::vector<float> errors;
        while (fread(r, 4, 4, f) == 4) errors.push_back(vp.ComputeBlockError(math::Rectangle<uint32>(r[0], r[1], r[2], r[3])));
        fclose(f);
        f = fopen(a.Str("dump-block-errors").c_str(), "wb");
        if (!f) Die("cannot open --dump-block-errors");
        WriteBytes(f, errors.data(), errors.size() * 4);
        fclose(f);
    }
    if (a.Has("dump-overlay"))
    {
        // Viewport::VisualizeActiveBlocks on a copy of the front buffer
        Bitmap overlay;
        Bitmap::InitData init;
        init.width = w; init.height = h; init.format = Bitmap::Format::B8G8R8A8_UNorm; init.linearSpace = false;
        if (!overlay.Init(init)) Die("overlay bitmap");
        memcpy(overlay.GetData(), vp.GetFrontBuffer().GetData(), (size_t)w * h * 4);
        vp.VisualizeActiveBlocks(overlay);
        FILE* f = fopen(a.Str("dump-overlay").c_str(), "wb");
        WriteBytes(f, overlay.GetData(), (size_t)w * h * 4);
        fclose(f);
    }
    if (a.Has("dump-front"))
    {
        FILE* f = fopen(a.Str("dump-front").c_str(), "wb");
        WriteBytes(f, vp.GetFrontBuffer().GetData(), (size_t)w * h * 4);
        fclose(f);
    }
    if (a.Has("dump-sum2"))
    {
        FILE* f = fopen(a.Str("dump-sum2").c_str(), "wb");
        WriteBytes(f, vp.mSecondarySum.GetData(), (size_t)w * h * 12);
        fclose(f);
    }
    printf("{\"scene\": \"%s\", \"width\": %u, \"height\": %u, \"passes\": %u, \"renderer\": \"%s\", \"threads\": %u, "
           "\"seconds\": %.6f, \"numPrimaryRays\": %llu, \"numRays\": %llu, \"numShadowRays\": %llu, \"numShadowRaysHit\": %llu, "
           "\"mrays_per_s\": %.4f, \"msamples_per_s\": %.4f}\n",
           a.Str("scene").c_str(), w, h, passes, rendererName.c_str(), params.numThreads, seconds,
           (unsigned long long)primary, (unsigned long long)rays, (unsigned long long)shadow, (unsigned long long)shadowHit,
           seconds > 0 ? (rays + shadow) / seconds * 1e-6 : 0.0, seconds > 0 ? primary / seconds * 1e-6 : 0.0);
    return 0;
}

int CmdFlat(const Args& a)
{
    const uint32 w = (uint32)a.Int("width", 256), h = (uint32)a.Int("height", 256);
    Scene scene; Camera camera;
    if (!LoadSceneFile(a, scene, camera, w, h)) return 1;
    Flat flat;
    flat.Build(scene);
    flat.Write(a.Str("out").c_str(), camera);
    printf("{\"objects\": %zu, \"shapes\": %zu, \"meshes\": %zu, \"materials\": %zu, \"lights\": %zu, \"tlasNodes\": %zu}\n",
           flat.objects.size(), flat.shapes.size(), flat.meshes.size(), flat.materials.size(), flat.lights.size(), flat.tlas.size());
    return 0;
}

Ray ToRay(const RtbRay& r)
{
    Ray ray;
    ray.origin = Vector4(r.origin);
    ray.dir = Vector4(r.dir);
    ray.invDir = Vector4(r.invDir);
    ray.originDivDir = Vector4(r.originDivDir);
    return ray;
}

void FromRay(const Ray& ray, RtbRay& r)
{
    ToF4(ray.origin, r.origin); ToF4(ray.dir, r.dir); ToF4(ray.invDir, r.invDir); ToF4(ray.originDivDir, r.originDivDir);
}

int CmdTrace(const Args& a)
{
    Scene scene; Camera camera;
    if (!LoadSceneFile(a, scene, camera, 256, 256)) return 1;
    FILE* f = fopen(a.Str("rays").c_str(), "rb");
    if (!f) Die("cannot open rays");
    fseek(f, 0, SEEK_END); const size_t bytes = ftell(f); fseek(f, 0, SEEK_SET);
    const size_t n = bytes / sizeof(RtbRay);
    std::vector<RtbRay> rays(n);
    if (n && fread(rays.data(), sizeof(RtbRay), n, f) != n) Die("read failed");
    fclose(f);
    // shadow mode: per-ray tMax comes from a parallel float file (HitPoint::distance preset, PathTracerMIS.cpp:86)
    const bool shadow = a.Int("shadow", 0) != 0;
    std::vector<float> tmax(n, std::numeric_limits<float>::infinity());
    if (a.Has("tmax"))
    {
        FILE* tf = fopen(a.Str("tmax").c_str(), "rb");
        if (!tf || fread(tmax.data(), 4, n, tf) != n) Die("cannot read tmax");
        fclose(tf);
    }
    std::vector<RtbHit> hits(n);
    RenderingContext ctx;
    Timer t;
    for (size_t i = 0; i < n; ++i)
    {
        const Ray ray = ToRay(rays[i]);
        HitPoint hp;
        hp.u = 0.0f; hp.v = 0.0f; hp.subObjectId = 0;
        RtbHit& o = hits[i];
        memset(&o, 0, sizeof(o));
        if (shadow)
        {
            hp.distance = tmax[i];
            o.occluded = scene.Traverse_Shadow({ ray, hp, ctx }) ? 1u : 0u;
        }
        else
        {
            scene.Traverse({ ray, hp, ctx });
        }
        o.objectId = hp.objectId; o.subObjectId = hp.subObjectId; o.distance = hp.distance; o.u = hp.u; o.v = hp.v;
    }
    const double s = t.Stop();
    FILE* of = fopen(a.Str("out").c_str(), "wb");
    WriteBytes(of, hits.data(), n * sizeof(RtbHit));
    fclose(of);
    printf("{\"rays\": %zu, \"seconds\": %.6f, \"mrays_per_s\": %.4f}\n", n, s, s > 0 ? n / s * 1e-6 : 0.0);
    return 0;
}

// pass script reader (written by CmdRender --dump-passes)
struct PassScript
{
    uint32 numPasses = 0, numDims = 0;
    std::vector<uint32> seeds;
    std::vector<float> offsets;
    void Load(const std::string& path)
    {
        FILE* f = fopen(path.c_str(), "rb");
        if (!f) Die("cannot open pass script");
        if (fread(&numPasses, 4, 1, f) != 1 || fread(&numDims, 4, 1, f) != 1) Die("bad pass script");
        seeds.resize((size_t)numPasses * numDims); offsets.resize((size_t)numPasses * 2);
        for (uint32 p = 0; p < numPasses; ++p)
        {
            if (fread(&seeds[(size_t)p * numDims], 4, numDims, f) != numDims) Die("bad pass script");
            if (fread(&offsets[(size_t)p * 2], 4, 2, f) != 2) Die("bad pass script");
        }
        fclose(f);
    }
};

int CmdTexEval(const Args& a)
{
    // ITexture::Evaluate for every texture of the flattened scene (in RtbSceneDesc::textures order) at a list of (u, v)
    Scene scene; Camera camera;
    if (!LoadSceneFile(a, scene, camera, 64, 64)) return 1;
    Flat flat; flat.Build(scene);
    std::vector<const ITexture*> byIndex(flat.textures.size(), nullptr);
    for (const auto& it : flat.textureIds) byIndex[it.second] = it.first;
    std::vector<float> uv;
    {
        FILE* f = fopen(a.Str("uv").c_str(), "rb");
        if (!f) Die("cannot open uv file");
        float pair[2];
        while (fread(pair, 4, 2, f) == 2) { uv.push_back(pair[0]); uv.push_back(pair[1]); }
        fclose(f);
    }
    FILE* of = fopen(a.Str("out").c_str(), "wb");
    if (!of) Die("cannot open output");
    for (const ITexture* t : byIndex)
        for (size_t i = 0; i + 1 < uv.size(); i += 2)
        {
            float c[4];
            ToF4(t->Evaluate(Vector4(uv[i], uv[i + 1], 0.0f, 0.0f)), c);
            WriteBytes(of, c, 16);
        }
    fclose(of);
    printf("{\"textures\": %zu, \"points\": %zu}\n", byIndex.size(), uv.size() / 2);
    return 0;
}

int CmdSampler(const Args& a)
{
    // streams for a grid of pixels: out[pass][pixel][dim]
    PassScript ps; ps.Load(a.Str("passes"));
    const uint32 w = (uint32)a.Int("width", 16), h = (uint32)a.Int("height", 16), nd = (uint32)a.Int("num-dims", 16);
    const bool blue = a.Int("blue-noise", 1) != 0;
    FILE* of = fopen(a.Str("out").c_str(), "wb");
    GenericSampler sampler;
    Random fallback;
    sampler.fallbackGenerator = &fallback;
    for (uint32 p = 0; p < ps.numPasses; ++p)
    {
        DynArray<uint32> seed(ps.numDims);
        for (uint32 d = 0; d < ps.numDims; ++d) seed[d] = ps.seeds[(size_t)p * ps.numDims + d];
        sampler.ResetFrame(seed, blue);
        for (uint32 y = 0; y < h; ++y) for (uint32 x = 0; x < w; ++x)
        {
            sampler.ResetPixel(x, y);
            for (uint32 d = 0; d < nd; ++d) WriteU32(of, sampler.GetInt());
        }
    }
    fclose(of);
    return 0;
}

int CmdPrimary(const Args& a)
{
    const uint32 w = (uint32)a.Int("width", 64), h = (uint32)a.Int("height", 64);
    Scene scene; Camera camera;
    if (!LoadSceneFile(a, scene, camera, w, h)) return 1;
    PassScript ps; ps.Load(a.Str("passes"));
    const uint32 pass = (uint32)a.Int("pass", 0);
    RenderingContext ctx;
    DynArray<uint32> seed(ps.numDims);
    for (uint32 d = 0; d < ps.numDims; ++d) seed[d] = ps.seeds[(size_t)pass * ps.numDims + d];
    ctx.sampler.ResetFrame(seed, a.Int("blue-noise", 1) != 0);
    ctx.sampler.fallbackGenerator = &ctx.randomGenerator;
    const Vector4 sampleOffset(ps.offsets[pass * 2], ps.offsets[pass * 2 + 1], 0.0f, 0.0f);
    // mirrors Viewport::RenderTile (Viewport.cpp:300-330)
    const Vector4 filmSize = Vector4::FromIntegers(w, h, 1, 1);
    const Vector4 invSize = VECTOR_ONE2 / filmSize;
    std::vector<RtbRay> rays((size_t)w * h);
    for (uint32 y = 0; y < h; ++y)
    {
        const uint32 realY = h - 1u - y;
        for (uint32 x = 0; x < w; ++x)
        {
            const Vector4 coords = (Vector4::FromIntegers(x, realY, 0, 0) + sampleOffset) * invSize;
            ctx.sampler.ResetPixel(x, y);
            ctx.time = 0.0f;
            const Ray ray = camera.GenerateRay(coords, ctx);
            FromRay(ray, rays[(size_t)y * w + x]);
        }
    }
    FILE* of = fopen(a.Str("out").c_str(), "wb");
    WriteBytes(of, rays.data(), rays.size() * sizeof(RtbRay));
    fclose(of);
    return 0;
}

} // namespace

static void CrashHandler(int sig)
{
    void* frames[64];
    const int n = backtrace(frames, 64);
    fprintf(stderr, "rtref: fatal signal %d\n", sig);
    backtrace_symbols_fd(frames, n, 2);
    _exit(128 + sig);
}

// emit --scene X.json --samples F (5 floats each) --out F: per light (Scene::GetLights order), per sample, 12 floats:
// position xyz, direction xyz, directPdfA, emissionPdfW, cosAtLight, radiance rgb
int CmdEmit(const Args& a)
{
    Scene scene; Camera camera;
    if (!LoadSceneFile(a, scene, camera, 64, 64)) return 1;
    FILE* f = fopen(a.Str("samples").c_str(), "rb");
    if (!f) Die("cannot open --samples");
    std::vector<float> samples;
    float v[5];
    while (fread(v, 4, 5, f) == 5) samples.insert(samples.end(), v, v + 5);
    fclose(f);
    FILE* o = fopen(a.Str("out").c_str(), "wb");
    if (!o) Die("cannot open --out");
    Wavelength wavelength;
    for (const LightSceneObject* lightObject : scene.GetLights())
    {
        const ILight& light = lightObject->GetLight();
        for (size_t i = 0; i + 5 <= samples.size(); i += 5)
        {
            const ILight::EmitParam param = { lightObject->GetTransform(0.0f), wavelength, Float3(samples[i], samples[i + 1], samples[i + 2]), Float2(samples[i + 3], samples[i + 4]) };
            ILight::EmitResult r;
            r.position = Vector4::Zero(); r.direction = Vector4::Zero(); r.directPdfA = r.emissionPdfW = r.cosAtLight = 0.0f;
            const RayColor color = light.Emit(param, r);
            const Vector4 rgb = color.ConvertToTristimulus(wavelength);
            const float out[12] = { r.position.x, r.position.y, r.position.z, r.direction.x, r.direction.y, r.direction.z,
                                    r.directPdfA, r.emissionPdfW, r.cosAtLight, rgb.x, rgb.y, rgb.z };
            WriteBytes(o, out, sizeof(out));
        }
    }
    fclose(o);
    return 0;
}

// illuminate --scene X.json --points F (15 floats each: position, tangent, bitangent, normal, sample) --out F: per light,
// per point 10 floats: direction to the light, distance, directPdfW, emissionPdfW, cosAtLight, radiance rgb
int CmdIlluminate(const Args& a)
{
    Scene scene; Camera camera;
    if (!LoadSceneFile(a, scene, camera, 64, 64)) return 1;
    FILE* f = fopen(a.Str("points").c_str(), "rb");
    if (!f) Die("cannot open --points");
    std::vector<float> pts;
    float v[15];
    while (fread(v, 4, 15, f) == 15) pts.insert(pts.end(), v, v + 15);
    fclose(f);
    FILE* o = fopen(a.Str("out").c_str(), "wb");
    if (!o) Die("cannot open --out");
    Wavelength wavelength;
    for (const LightSceneObject* lightObject : scene.GetLights())
    {
        const ILight& light = lightObject->GetLight();
        const Matrix4 lightToWorld = lightObject->GetTransform(0.0f);
        const Matrix4 worldToLight = lightObject->GetInverseTransform(0.0f);
        for (size_t i = 0; i + 15 <= pts.size(); i += 15)
        {
            const float* p = &pts[i];
            IntersectionData intersection;
            intersection.frame[0] = Vector4(p[3], p[4], p[5], 0.0f);
            intersection.frame[1] = Vector4(p[6], p[7], p[8], 0.0f);
            intersection.frame[2] = Vector4(p[9], p[10], p[11], 0.0f);
            intersection.frame[3] = Vector4(p[0], p[1], p[2], 1.0f);
            intersection.texCoord = Vector4::Zero();
            const ILight::IlluminateParam param = { worldToLight, lightToWorld, intersection, wavelength, Float3(p[12], p[13], p[14]), false };
            ILight::IlluminateResult r;
            const RayColor color = light.Illuminate(param, r);
            const Vector4 rgb = color.ConvertToTristimulus(wavelength);
            const float out[10] = { r.directionToLight.x, r.directionToLight.y, r.directionToLight.z, r.distance, r.directPdfW, r.emissionPdfW, r.cosAtLight, rgb.x, rgb.y, rgb.z };
            WriteBytes(o, out, sizeof(out));
        }
    }
    fclose(o);
    return 0;
}

// radiance --scene X.json --rays F (4 floats each: unit direction in light space, cosAtLight) --out F: per light that can be
// hit (everything but point / spot lights), per ray 5 floats: radiance rgb, directPdfA, emissionPdfW (-1 where not written)
int CmdRadiance(const Args& a)
{
    Scene scene; Camera camera;
    if (!LoadSceneFile(a, scene, camera, 64, 64)) return 1;
    FILE* f = fopen(a.Str("rays").c_str(), "rb");
    if (!f) Die("cannot open --rays");
    std::vector<float> rays;
    float v[4];
    while (fread(v, 4, 4, f) == 4) rays.insert(rays.end(), v, v + 4);
    fclose(f);
    FILE* o = fopen(a.Str("out").c_str(), "wb");
    if (!o) Die("cannot open --out");
    RenderingContext ctx;
    for (const LightSceneObject* lightObject : scene.GetLights())
    {
        const ILight& light = lightObject->GetLight();
        if (dynamic_cast<const PointLight*>(&light) || dynamic_cast<const SpotLight*>(&light)) continue;
        for (size_t i = 0; i + 4 <= rays.size(); i += 4)
        {
            Ray ray;
            ray.origin = Vector4::Zero();
            ray.dir = Vector4(rays[i], rays[i + 1], rays[i + 2], 0.0f);
            ray.invDir = Vector4::Zero(); ray.originDivDir = Vector4::Zero();
            const ILight::RadianceParam param = { ctx, ray, Vector4::Zero(), rays[i + 3], false };
            float directPdfA = -1.0f, emissionPdfW = -1.0f;
            const RayColor color = light.GetRadiance(param, &directPdfA, &emissionPdfW);
            const Vector4 rgb = color.ConvertToTristimulus(ctx.wavelength);
            const float out[5] = { rgb.x, rgb.y, rgb.z, directPdfA, emissionPdfW };
            WriteBytes(o, out, sizeof(out));
        }
    }
    fclose(o);
    return 0;
}

// bsdfpdf --scene X.json --dirs F (6 floats each: local outgoing xyz, local incoming xyz, both unit) --out F:
// per material (Scene flatten order = first use by an object), per pair: forward pdf, reverse pdf.  Material parameters
// are the constants of the material (no texture lookups).
int CmdBsdfPdf(const Args& a)
{
    Scene scene; Camera camera;
    if (!LoadSceneFile(a, scene, camera, 64, 64)) return 1;
    Flat flat;
    flat.Build(scene);
    FILE* f = fopen(a.Str("dirs").c_str(), "rb");
    if (!f) Die("cannot open --dirs");
    std::vector<float> dirs;
    float v[6];
    while (fread(v, 4, 6, f) == 6) dirs.insert(dirs.end(), v, v + 6);
    fclose(f);
    FILE* o = fopen(a.Str("out").c_str(), "wb");
    if (!o) Die("cannot open --out");
    Wavelength wavelength;
    std::vector<const Material*> materialPtrs(flat.materials.size(), nullptr);
    for (const auto& it : flat.materialIds) materialPtrs[it.second] = it.first;
    for (const Material* material : materialPtrs)
    {
        SampledMaterialParameters params;
        params.baseColor = RayColor::Resolve(wavelength, Spectrum(material->baseColor.baseValue));
        params.emissionColor = RayColor::Zero();
        params.roughness = material->roughness.baseValue;
        params.metalness = material->metalness.baseValue;
        params.IoR = material->IoR;
        for (size_t i = 0; i + 6 <= dirs.size(); i += 6)
        {
            const BSDF::EvaluationContext ctx = { *material, params, wavelength, Vector4(dirs[i], dirs[i + 1], dirs[i + 2], 0.0f), Vector4(dirs[i + 3], dirs[i + 4], dirs[i + 5], 0.0f) };
            const float out[2] = { material->GetBSDF()->Pdf(ctx, BSDF::ForwardPdf), material->GetBSDF()->Pdf(ctx, BSDF::ReversePdf) };
            WriteBytes(o, out, sizeof(out));
        }
    }
    fclose(o);
    return 0;
}

// camera --scene X.json --width W --height H --points F (3 floats each) --out F: per point 7 words: WorldToFilm's
// return value, film x, film y (0 when behind the camera), the unit direction camera -> point, Camera::PdfW(direction)
int CmdCamera(const Args& a)
{
    const uint32 w = (uint32)a.Int("width", 64), h = (uint32)a.Int("height", 64);
    Scene scene; Camera camera;
    if (!LoadSceneFile(a, scene, camera, w, h)) return 1;
    FILE* f = fopen(a.Str("points").c_str(), "rb");
    if (!f) Die("cannot open --points");
    FILE* o = fopen(a.Str("out").c_str(), "wb");
    if (!o) Die("cannot open --out");
    float p[3];
    while (fread(p, 4, 3, f) == 3)
    {
        const Vector4 point(p[0], p[1], p[2], 0.0f);
        Vector4 film = Vector4::Zero();
        const uint32 ok = camera.WorldToFilm(point, film) ? 1u : 0u;
        Vector4 dir = point - camera.GetTransform().GetTranslation();
        dir /= sqrtf(dir.SqrLength3());
        const float pdf = camera.PdfW(dir);
        WriteU32(o, ok);
        const float out[6] = { ok ? film.x : 0.0f, ok ? film.y : 0.0f, dir.x, dir.y, dir.z, pdf };
        WriteBytes(o, out, sizeof(out));
    }
    fclose(f); fclose(o);
    return 0;
}

// writeexr --out F.exr --width W --height H --channels 3|4 --type half|float --compression 0..4 [--seed S]
// (channel names A, B, G, R as tinyexr's own SaveEXR orders them; values: HDR-ish noise with zeros, smooth regions and
// large peaks so that every compressor sees runs, gradients and outliers)
int CmdWriteExr(const Args& a)
{
    const int w = (int)a.Int("width", 16), h = (int)a.Int("height", 16), channels = (int)a.Int("channels", 3);
    const bool half = a.Str("type", "half") == "half";
    uint64_t state = (uint64_t)a.Int("seed", 1) * 0x9E3779B97F4A7C15ull + 0x1234567ull;
    auto next = [&state]() { state ^= state << 13; state ^= state >> 7; state ^= state << 17; return (uint32_t)(state >> 11); };
    std::vector<std::vector<float>> planes(channels, std::vector<float>((size_t)w * h));
    for (int c = 0; c < channels; ++c)
        for (int y = 0; y < h; ++y)
            for (int x = 0; x < w; ++x)
            {
                const uint32_t r = next();
                float v = 0.25f * (float)((x / 3 + y / 2 + c) % 7);                  // plateaus
                if (r % 5u == 0u) v += (float)(r >> 8) * (1.0f / 16777216.0f);      // noise
                if (r % 97u == 0u) v *= 1000.0f;                                    // highlights
                if (r % 11u == 0u) v = 0.0f;
                planes[c][(size_t)y * w + x] = v;
            }
    EXRHeader header; InitEXRHeader(&header);
    EXRImage image; InitEXRImage(&image);
    image.num_channels = channels;
    std::vector<float*> ptrs(channels);
    for (int c = 0; c < channels; ++c) ptrs[c] = planes[c].data();
    image.images = (unsigned char**)ptrs.data();
    image.width = w; image.height = h;
    header.num_channels = channels;
    std::vector<EXRChannelInfo> infos(channels);
    const char* names3[] = { "B", "G", "R" };
    const char* names4[] = { "A", "B", "G", "R" };
    std::vector<int> pixelTypes(channels, TINYEXR_PIXELTYPE_FLOAT), requested(channels, half ? TINYEXR_PIXELTYPE_HALF : TINYEXR_PIXELTYPE_FLOAT);
    for (int c = 0; c < channels; ++c) { memset(&infos[c], 0, sizeof(EXRChannelInfo)); strncpy(infos[c].name, channels == 3 ? names3[c] : names4[c], 255); }
    header.channels = infos.data();
    header.pixel_types = pixelTypes.data();
    header.requested_pixel_types = requested.data();
    header.compression_type = (int)a.Int("compression", 0);
    const char* err = nullptr;
    const int ret = SaveEXRImageToFile(&image, &header, a.Str("out").c_str(), &err);
    if (ret != TINYEXR_SUCCESS) { fprintf(stderr, "rtref: SaveEXRImageToFile: %s\n", err ? err : "?"); return 1; }
    return 0;
}

int main(int argc, char** argv)
{
    signal(SIGSEGV, CrashHandler);
    signal(SIGILL, CrashHandler);
    SetFlushDenormalsToZero();
    InitMemory();
    const Args a = ParseArgs(argc, argv);
    if (a.cmd == "render") return CmdRender(a);
    if (a.cmd == "flat") return CmdFlat(a);
    if (a.cmd == "trace") return CmdTrace(a);
    if (a.cmd == "sampler") return CmdSampler(a);
    if (a.cmd == "texeval") return CmdTexEval(a);
    if (a.cmd == "writeexr") return CmdWriteExr(a);
    if (a.cmd == "emit") return CmdEmit(a);
    if (a.cmd == "camera") return CmdCamera(a);
    if (a.cmd == "bsdfpdf") return CmdBsdfPdf(a);
    if (a.cmd == "illuminate") return CmdIlluminate(a);
    if (a.cmd == "radiance") return CmdRadiance(a);
    if (a.cmd == "primary") return CmdPrimary(a);
    fprintf(stderr, "usage: rtref render|flat|trace|sampler|primary --scene X.json [options]\n");
    return 2;
}
