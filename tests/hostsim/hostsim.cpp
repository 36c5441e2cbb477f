// hostsim — g++ build of the DEVICE headers (raytracer_b200/csrc/rtb_*.h), one pixel at a time on the CPU.
//
// TEST TOOL ONLY.  The authoring container has no GPU, so the numerics of the device code (operation order, FMA
// placement, SSE min/max semantics ...) are debugged here against the oracle before a B200 run.  It is NOT part of
// the product: librtb200.so does not contain it, the raytracer_b200 package cannot reach it, and rtb_create() fails
// without a CUDA device.  Built by tests/hostsim/build.sh, used by tests/test_hostsim_parity.py.
//
//   hostsim render <scene.rtb> <passes.bin> <width> <height> <out.sum> [renderer 0|1|2] [maxDepth] [debugMode] [strategyAll]
//   hostsim trace  <scene.rtb> <rays.bin> <out.hits> [shadow tmax.bin]
//   hostsim sampler <passes.bin> <w> <h> <numDims> <out.u32> <bluenoise.dat|->
//   hostsim primary <scene.rtb> <passes.bin> <pass> <w> <h> <out.rays>
//   hostsim emit <scene.rtb> <samples.f32 (5 per sample)> <out.f32 (12 per light and sample)>
//   hostsim camera <scene.rtb> <w> <h> <points.f32 (3 per point)> <ref.bin (for the directions)> <out.bin (7 words per point)>
//   hostsim lighttrace <scene.rtb> <w> <h> <passes> <maxDepth> <seed> <out.sum>     (LightTracer: statistical parity only)
//   hostsim bsdfpdf <scene.rtb> <dirs.f32 (6 per pair)> <out.f32 (forward, reverse per material and pair)>
//   hostsim illuminate <scene.rtb> <points.f32 (15 per point)> <out.f32 (10 per light and point)>
//   hostsim radiance <scene.rtb> <rays.f32 (4 per ray)> <out.f32 (5 per hittable light and ray)>
//   hostsim vcm <scene.rtb> <passes.bin> <w> <h> <seed> <connection 0|1> <merging 0|1> <out.sum> [bluenoise]     (statistical parity only)
//   hostsim blockerrors <sum.bin> <sum2.bin> <w> <h> <numPasses> <blocks.bin> <out.f32>
//   hostsim front  <sum.bin> <w> <h> <numPasses> <exposure> <contrast> <saturation> <dither> <tonemapper> <r,g,b filter> <out.bgra> [bloomFactor]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <xmmintrin.h>
#include <pmmintrin.h>

#include "../../raytracer_b200/csrc/rtb_integrator.h"
#include "../../raytracer_b200/csrc/rtb_postprocess.h"
#include "../../raytracer_b200/csrc/rtb_bidir.h"
#include "hostsim_vcm.h"
#include "../../raytracer_b200/csrc/rtb_upload.h"
#include "../../raytracer_b200/csrc/rtb_scene_file.h"

using namespace rtb;

static void die(const std::string& m) { fprintf(stderr, "hostsim: %s\n", m.c_str()); exit(1); }

struct Passes
{
    uint32_t numPasses = 0, numDims = 0;
    std::vector<uint32_t> seeds;
    std::vector<float> offsets;
    void load(const char* path)
    {
        FILE* f = fopen(path, "rb");
        if (!f) die("cannot open passes");
        if (fread(&numPasses, 4, 1, f) != 1 || fread(&numDims, 4, 1, f) != 1) die("bad passes");
        seeds.resize((size_t)numPasses * numDims); offsets.resize((size_t)numPasses * 2);
        for (uint32_t p = 0; p < numPasses; ++p)
        {
            if (fread(&seeds[(size_t)p * numDims], 4, numDims, f) != numDims) die("bad passes");
            if (fread(&offsets[(size_t)p * 2], 4, 2, f) != 2) die("bad passes");
        }
        fclose(f);
    }
};

struct World
{
    SceneFile file;
    HostScene hs;
    std::vector<DMesh> meshViews;
    std::vector<uint16_t> blueNoise;
    SceneView S;

    void load(const char* path)
    {
        std::string err;
        if (!file.Load(path, err)) die(err);
        const RtbSceneDesc d = file.View();
        if (!build_host_scene(d, hs, err)) die(err);
        meshViews.resize(hs.meshes.size());
        for (size_t i = 0; i < hs.meshes.size(); ++i)
        {
            DMesh& v = meshViews[i];
            const HostMesh& m = hs.meshes[i];
            v.nodes = m.nodes.data(); v.tris = m.tris.data(); v.indices = m.indices.data(); v.shading = m.shading.data();
            v.materialIds = m.materialIds.data();
            v.numNodes = (uint32_t)m.nodes.size(); v.numTriangles = (uint32_t)m.tris.size();
            v.numVertices = (uint32_t)m.shading.size(); v.numMaterials = (uint32_t)m.materialIds.size();
            const bool wide = !m.wide.nodes.empty() && !getenv("HOSTSIM_NO_WIDE");
            v.wideNodes = wide ? m.wide.nodes.data() : nullptr; v.wideTris = wide ? m.wide.tris.data() : nullptr;
            v.wideMaxAbs = m.wide.maxAbs; v.numWideNodes = wide ? (uint32_t)m.wide.nodes.size() : 0u;
        }
        memset(&S, 0, sizeof(S));
        S.tlas = hs.tlas.data(); S.objTrace = hs.objTrace.data(); S.objShade = hs.objShade.data();
        S.shapes = hs.shapes.data(); S.meshes = meshViews.data(); S.materials = hs.materials.data();
        S.lights = hs.lights.data(); S.globalLights = hs.globalLights.data();
        S.numTlasNodes = (uint32_t)hs.tlas.size(); S.numObjects = (uint32_t)hs.objTrace.size();
        S.numLights = (uint32_t)hs.lights.size(); S.numGlobalLights = (uint32_t)hs.globalLights.size();
        S.blueNoise = nullptr;
        for (size_t i = 0; i < hs.textures.size(); ++i) hs.textures[i].texels = hs.texelSources[i];     // host pointers (owned by `file`)
        S.textures = hs.textures.empty() ? nullptr : hs.textures.data();
        S.materialTextures = hs.materialTextures.empty() ? nullptr : hs.materialTextures.data();
    }

    void loadBlueNoise(const char* path)
    {
        FILE* f = fopen(path, "rb");
        if (!f) return;
        blueNoise.resize(128 * 128 * 4);
        if (fread(blueNoise.data(), 2, blueNoise.size(), f) != blueNoise.size()) die("bad blue noise file");
        fclose(f);
        S.blueNoise = blueNoise.data();
    }
};

static void fillCamera(PassParams& P, const RtbCamera& c, uint32_t w, uint32_t h, float tanHalfFoV)
{
    P.camLocalToWorld = load_m4(c.localToWorld);
    P.aspectRatio = (float)w / (float)h;
    P.tanHalfFoV = tanHalfFoV;
    P.dofEnable = c.dofEnable; P.bokehShape = c.bokehShape; P.aperture = c.aperture; P.focalPlaneDistance = c.focalPlaneDistance;
    P.barrelConst = c.barrelDistortionConstFactor; P.barrelVar = c.barrelDistortionVariableFactor;
    P.width = w; P.height = h;
}

static void fillBokeh(PassParams& P, const SceneFile& file)
{
    P.bokehCdf = file.bokehCdf.empty() ? nullptr : file.bokehCdf.data();
    P.bokehWidth = file.bokehWidth; P.bokehHeight = file.bokehHeight;
}

// immediate-mode sink: shadow rays are traced on the spot, in light order
struct HostSink
{
    const SceneView& S;
    V3 result;
    V4 lightSum;
    uint64_t shadowRays = 0, shadowRaysHit = 0;
    explicit HostSink(const SceneView& s) : S(s) { result = v3s(0.0f); lightSum = v4s(0.0f); }
    void add(const V4& thr, const V4& c) { result = mul_add(xyz(thr), xyz(c), result); }
    void shadow(const Ray& ray, float tMax, const V4& contribution, uint32_t)
    {
        shadowRays++;
        const bool occluded = scene_traverse_shadow<false>(S, ray, tMax, nullptr);
        if (!occluded) { shadowRaysHit++; lightSum = lightSum + contribution; }
        else lightSum = lightSum + v4s(0.0f);
    }
    void end_lights(const V4& thr)
    {
        result = mul_add(xyz(thr), xyz(lightSum), result);
        lightSum = v4s(0.0f);
    }
};

static int cmdRender(int argc, char** argv)
{
    if (argc < 7) die("usage: render scene passes w h out [renderer] [depth] [debugMode] [all] [bluenoise] [fovTan]");
    World world; world.load(argv[2]);
    Passes passes; passes.load(argv[3]);
    const uint32_t w = atoi(argv[4]), h = atoi(argv[5]);
    const uint32_t renderer = argc > 7 ? atoi(argv[7]) : 1;
    const uint32_t maxDepth = argc > 8 ? atoi(argv[8]) : 8;
    const uint32_t debugMode = argc > 9 ? atoi(argv[9]) : 1;
    const uint32_t all = argc > 10 ? atoi(argv[10]) : 1;
    if (argc > 11 && strcmp(argv[11], "-") != 0) world.loadBlueNoise(argv[11]);

    PassParams P;
    memset(&P, 0, sizeof(P));
    fillCamera(P, world.file.camera, w, h, tanf(world.file.fieldOfView * 0.5f));
    fillBokeh(P, world.file);
    P.renderer = renderer; P.debugMode = debugMode; P.maxRayDepth = maxDepth; P.minRussianRouletteDepth = 1;
    P.lightSamplingAll = all; P.useBlueNoise = world.S.blueNoise != nullptr; P.numSeedDims = passes.numDims;

    std::vector<float> sum((size_t)w * h * 3, 0.0f);
    uint64_t numRays = 0, numShadow = 0, numShadowHit = 0;
    int dbgX = -1, dbgY = -1;      // RTB_DEBUG_PIXEL=x,y prints every path vertex of that pixel
    if (const char* e = getenv("RTB_DEBUG_PIXEL")) sscanf(e, "%d,%d", &dbgX, &dbgY);
    for (uint32_t p = 0; p < passes.numPasses; ++p)
    {
        P.passIndex = p;
        P.seed = &passes.seeds[(size_t)p * passes.numDims];
        P.sampleOffsetX = passes.offsets[p * 2]; P.sampleOffsetY = passes.offsets[p * 2 + 1];
        for (uint32_t y = 0; y < h; ++y) for (uint32_t x = 0; x < w; ++x)
        {
            PathState ps;
            init_path(P, world.S.blueNoise, x, y, 0, ps);
            HostSink sink(world.S);
            for (;;)
            {
                Hit hit;
                hit.objectId = RTB_INVALID_OBJECT; hit.subObjectId = 0; hit.distance = INFINITY; hit.u = 0.0f; hit.v = 0.0f;
                scene_traverse<false>(world.S, ps.ray, hit, nullptr);
                bool alive = false;
                uint32_t ns = 0;
                if (renderer == RTB_RENDERER_PATH_TRACER_MIS) alive = shade_vertex_mis(world.S, P, ps, hit, sink, ns);
                else if (renderer == RTB_RENDERER_PATH_TRACER) alive = shade_vertex_pt(world.S, P, ps, hit, sink);
                else shade_debug(world.S, P, ps, hit, sink);
                if (dbgX == (int)x && dbgY == (int)y)
                    fprintf(stderr, "pass %u depth %u: hit obj %u sub %u dist %.9g u %.9g v %.9g -> result %.9g %.9g %.9g alive %d shadow %u thr %.9g %.9g %.9g\n", p, ps.depth, hit.objectId, hit.subObjectId,
                            hit.distance, hit.u, hit.v, sink.result.x, sink.result.y, sink.result.z, (int)alive, ns, ps.throughput.x, ps.throughput.y, ps.throughput.z);
                if (!alive) break;
            }
            numRays += (uint64_t)ps.depth + 1;
            numShadow += sink.shadowRays; numShadowHit += sink.shadowRaysHit;
            float* px = &sum[((size_t)y * w + x) * 3];
            px[0] = px[0] + sink.result.x; px[1] = px[1] + sink.result.y; px[2] = px[2] + sink.result.z;
        }
    }
    FILE* f = fopen(argv[6], "wb");
    fwrite(sum.data(), 4, sum.size(), f);
    fclose(f);
    printf("{\"numRays\": %llu, \"numShadowRays\": %llu, \"numShadowRaysHit\": %llu}\n",
           (unsigned long long)numRays, (unsigned long long)numShadow, (unsigned long long)numShadowHit);
    return 0;
}

static int cmdTexEval(int argc, char** argv)
{
    if (argc < 5) die("usage: texeval scene uvfile out");
    World world; world.load(argv[2]);
    std::vector<float> uv;
    FILE* f = fopen(argv[3], "rb");
    if (!f) die("cannot open uv file");
    float pair[2];
    while (fread(pair, 4, 2, f) == 2) { uv.push_back(pair[0]); uv.push_back(pair[1]); }
    fclose(f);
    FILE* of = fopen(argv[4], "wb");
    for (size_t t = 0; t < world.hs.textures.size(); ++t)
        for (size_t i = 0; i + 1 < uv.size(); i += 2)
        {
            const V4 c = texture_evaluate(world.S, (uint32_t)t, uv[i], uv[i + 1]);
            const float o[4] = { c.x, c.y, c.z, c.w };
            fwrite(o, 4, 4, of);
        }
    fclose(of);
    return 0;
}

static int cmdTrace(int argc, char** argv)
{
    if (argc < 5) die("usage: trace scene rays out [tmax]");
    World world; world.load(argv[2]);
    FILE* f = fopen(argv[3], "rb");
    if (!f) die("cannot open rays");
    fseek(f, 0, SEEK_END); const size_t n = ftell(f) / sizeof(RtbRay); fseek(f, 0, SEEK_SET);
    std::vector<RtbRay> rays(n);
    if (n && fread(rays.data(), sizeof(RtbRay), n, f) != n) die("read failed");
    fclose(f);
    const bool shadow = argc > 5;
    std::vector<float> tmax(n, INFINITY);
    if (shadow)
    {
        FILE* tf = fopen(argv[5], "rb");
        if (!tf || fread(tmax.data(), 4, n, tf) != n) die("cannot read tmax");
        fclose(tf);
    }
    std::vector<RtbHit> hits(n);
    TraceCounters cnt; cnt.boxTests = 0; cnt.triTests = 0;
    unsigned long long boxTests = 0, triTests = 0;
    const bool countTests = getenv("HOSTSIM_COUNT") != nullptr;
    for (size_t i = 0; i < n; ++i)
    {
        Ray r;
        r.origin = v3(rays[i].origin[0], rays[i].origin[1], rays[i].origin[2]);
        r.dir = v3(rays[i].dir[0], rays[i].dir[1], rays[i].dir[2]);
        r.invDir = v3(rays[i].invDir[0], rays[i].invDir[1], rays[i].invDir[2]);
        r.originDivDir = v3(rays[i].originDivDir[0], rays[i].originDivDir[1], rays[i].originDivDir[2]);
        Hit hit;
        hit.objectId = RTB_INVALID_OBJECT; hit.subObjectId = 0; hit.distance = INFINITY; hit.u = 0.0f; hit.v = 0.0f;
        RtbHit& o = hits[i];
        memset(&o, 0, sizeof(o));
        if (shadow) { hit.distance = tmax[i]; o.occluded = (countTests ? scene_traverse_shadow<true>(world.S, r, tmax[i], &cnt) : scene_traverse_shadow<false>(world.S, r, tmax[i], nullptr)) ? 1u : 0u; }
        else if (countTests) scene_traverse<true>(world.S, r, hit, &cnt);
        else scene_traverse<false>(world.S, r, hit, nullptr);
        boxTests += cnt.boxTests; triTests += cnt.triTests; cnt.boxTests = 0; cnt.triTests = 0;
        o.objectId = hit.objectId; o.subObjectId = hit.subObjectId; o.distance = hit.distance; o.u = hit.u; o.v = hit.v;
    }
    FILE* of = fopen(argv[4], "wb");
    fwrite(hits.data(), sizeof(RtbHit), n, of);
    fclose(of);
    if (countTests)
    {
        size_t wideNodes = 0, wideTris = 0; uint32_t depth = 0;
        for (const HostMesh& m : world.hs.meshes) { wideNodes += m.wide.nodes.size(); wideTris += m.wide.tris.size(); depth = std::max(depth, m.wide.depth); }
        printf("{\"rays\": %zu, \"boxTests\": %llu, \"triTests\": %llu, \"wideNodes\": %zu, \"wideTris\": %zu, \"wideDepth\": %u}\n", n, boxTests, triTests, wideNodes, wideTris, depth);
    }
    return 0;
}

static int cmdSampler(int argc, char** argv)
{
    if (argc < 8) die("usage: sampler passes w h numDims out bluenoise|-");
    Passes passes; passes.load(argv[2]);
    const uint32_t w = atoi(argv[3]), h = atoi(argv[4]), nd = atoi(argv[5]);
    std::vector<uint16_t> bn;
    if (strcmp(argv[7], "-") != 0)
    {
        FILE* f = fopen(argv[7], "rb");
        if (!f) die("cannot open blue noise");
        bn.resize(128 * 128 * 4);
        if (fread(bn.data(), 2, bn.size(), f) != bn.size()) die("bad blue noise");
        fclose(f);
    }
    FILE* of = fopen(argv[6], "wb");
    PassParams P;
    memset(&P, 0, sizeof(P));
    P.numSeedDims = passes.numDims; P.useBlueNoise = !bn.empty();
    for (uint32_t p = 0; p < passes.numPasses; ++p)
    {
        P.seed = &passes.seeds[(size_t)p * passes.numDims];
        for (uint32_t y = 0; y < h; ++y) for (uint32_t x = 0; x < w; ++x)
        {
            Sampler s;
            sampler_reset_pixel(s, x, y);
            for (uint32_t d = 0; d < nd; ++d)
            {
                const uint32_t v = sampler_get_int(s, P, bn.empty() ? nullptr : bn.data());
                fwrite(&v, 4, 1, of);
            }
        }
    }
    fclose(of);
    return 0;
}

static int cmdPrimary(int argc, char** argv)
{
    if (argc < 8) die("usage: primary scene passes pass w h out");
    World world; world.load(argv[2]);
    Passes passes; passes.load(argv[3]);
    const uint32_t pass = atoi(argv[4]), w = atoi(argv[5]), h = atoi(argv[6]);
    PassParams P;
    memset(&P, 0, sizeof(P));
    fillCamera(P, world.file.camera, w, h, tanf(world.file.fieldOfView * 0.5f));
    fillBokeh(P, world.file);
    P.numSeedDims = passes.numDims;
    P.seed = &passes.seeds[(size_t)pass * passes.numDims];
    P.sampleOffsetX = passes.offsets[pass * 2]; P.sampleOffsetY = passes.offsets[pass * 2 + 1];
    std::vector<RtbRay> rays((size_t)w * h);
    for (uint32_t y = 0; y < h; ++y) for (uint32_t x = 0; x < w; ++x)
    {
        Sampler s;
        sampler_reset_pixel(s, x, y);
        const Ray r = generate_primary_ray(P, x, y, P.sampleOffsetX, P.sampleOffsetY, s, nullptr);
        RtbRay& o = rays[(size_t)y * w + x];
        memset(&o, 0, sizeof(o));
        o.origin[0] = r.origin.x; o.origin[1] = r.origin.y; o.origin[2] = r.origin.z;
        o.dir[0] = r.dir.x; o.dir[1] = r.dir.y; o.dir[2] = r.dir.z;
        o.invDir[0] = r.invDir.x; o.invDir[1] = r.invDir.y; o.invDir[2] = r.invDir.z;
        o.originDivDir[0] = r.originDivDir.x; o.originDivDir[1] = r.originDivDir.y; o.originDivDir[2] = r.originDivDir.z;
    }
    FILE* of = fopen(argv[7], "wb");
    fwrite(rays.data(), sizeof(RtbRay), rays.size(), of);
    fclose(of);
    return 0;
}

// ILight::Emit for every light (in Scene::GetLights order = the order of the flattened light array) and every sample
static int cmdEmit(int argc, char** argv)
{
    if (argc < 5) die("emit: arguments");
    World world; world.load(argv[2]);
    FILE* f = fopen(argv[3], "rb");
    if (!f) die("cannot open samples");
    std::vector<float> samples;
    float v[5];
    while (fread(v, 4, 5, f) == 5) samples.insert(samples.end(), v, v + 5);
    fclose(f);
    FILE* of = fopen(argv[4], "wb");
    for (uint32_t l = 0; l < world.S.numLights; ++l)
        for (size_t i = 0; i + 5 <= samples.size(); i += 5)
        {
            EmitResult r;
            memset(&r, 0, sizeof(r));
            const V4 c = light_emit<true>(world.S, world.S.lights[l], samples[i], samples[i + 1], samples[i + 2], samples[i + 3], samples[i + 4], r);
            const float out[12] = { r.position.x, r.position.y, r.position.z, r.direction.x, r.direction.y, r.direction.z, r.directPdfA, r.emissionPdfW, r.cosAtLight, c.x, c.y, c.z };
            fwrite(out, 4, 12, of);
        }
    fclose(of);
    return 0;
}

// Camera::WorldToFilm / PdfW; the direction handed to PdfW is read from the reference's record (word 3..5) so that both
// sides evaluate the same input
static int cmdCamera(int argc, char** argv)
{
    if (argc < 8) die("camera: arguments");
    World world; world.load(argv[2]);
    const uint32_t w = atoi(argv[3]), h = atoi(argv[4]);
    PassParams P;
    memset(&P, 0, sizeof(P));
    fillCamera(P, world.file.camera, w, h, tanf(world.file.fieldOfView * 0.5f));
    const M4 worldToScreen = camera_world_to_screen(P.camLocalToWorld, P.aspectRatio, P.tanHalfFoV);
    FILE* f = fopen(argv[5], "rb");
    FILE* rf = fopen(argv[6], "rb");
    if (!f || !rf) die("cannot open points / reference records");
    FILE* of = fopen(argv[7], "wb");
    float p[3], ref[7];
    while (fread(p, 4, 3, f) == 3 && fread(ref, 4, 7, rf) == 7)
    {
        float fx = 0.0f, fy = 0.0f;
        const uint32_t ok = camera_world_to_film(worldToScreen, v3(p[0], p[1], p[2]), fx, fy) ? 1u : 0u;
        const float pdf = camera_pdf_w(P.camLocalToWorld, P.tanHalfFoV, P.aspectRatio, v3(ref[3], ref[4], ref[5]));
        fwrite(&ok, 4, 1, of);
        const float out[6] = { ok ? fx : 0.0f, ok ? fy : 0.0f, ref[3], ref[4], ref[5], pdf };
        fwrite(out, 4, 6, of);
    }
    fclose(f); fclose(rf); fclose(of);
    return 0;
}

// LightTracer (Rendering/LightTracer.cpp): one light sub-path per pixel and pass, camera connections splatted onto the
// film.  The reference draws from per-thread generators, so only the statistics of the image can be compared.
struct HostRng
{
    uint64_t state;
    uint64_t next() { state ^= state >> 12; state ^= state << 25; state ^= state >> 27; return state * 0x2545F4914F6CDD1DULL; }
    uint32_t next_uint() { return (uint32_t)(next() >> 32); }
    float next_float() { return bits2f(((uint32_t)(next() >> 41)) | 0x3f800000u) - 1.0f; }
};

static int cmdLightTrace(int argc, char** argv)
{
    if (argc < 9) die("lighttrace: arguments");
    World world; world.load(argv[2]);
    const uint32_t w = atoi(argv[3]), h = atoi(argv[4]), passes = atoi(argv[5]), maxDepth = atoi(argv[6]);
    HostRng rng{ strtoull(argv[7], nullptr, 10) * 0x9E3779B97F4A7C15ULL + 1u };
    PassParams P;
    memset(&P, 0, sizeof(P));
    fillCamera(P, world.file.camera, w, h, tanf(world.file.fieldOfView * 0.5f));
    CameraConnection camera;
    camera.localToWorld = P.camLocalToWorld; camera.tanHalfFoV = P.tanHalfFoV; camera.aspectRatio = P.aspectRatio;
    camera.worldToScreen = camera_world_to_screen(P.camLocalToWorld, P.aspectRatio, P.tanHalfFoV);
    std::vector<float> sum((size_t)w * h * 3, 0.0f);
    uint64_t numRays = 0, numSplats = 0;
    // Film::AccumulateColor(pos, value, rng), Film.cpp:41-77: nearest pixel with a random +1 step in x by the fractional
    // part (the y step is computed but the row comes from the unjittered coordinate, as in the reference)
    auto splat = [&](float filmX, float filmY, const V3& value)
    {
        const float cx = filmX * (float)w + 0.0f, cy = filmY * (float)h + 0.5f;
        int32_t ix = cvt_rn(cx);
        const float fractionX = cx - (float)ix;
        const float ux = rng.next_float();
        rng.next_float(); rng.next_float(); rng.next_float();       // GetVector4: four lanes drawn
        if (ux < fractionX) ix++;
        const int32_t x = ix, y = (int32_t)(h - 1) - (int32_t)cy;
        if ((uint32_t)x < w && (uint32_t)y < h)
        {
            float* px = &sum[((size_t)y * w + x) * 3];
            px[0] += value.x; px[1] += value.y; px[2] += value.z;
            numSplats++;
        }
    };
    for (uint32_t p = 0; p < passes; ++p)
        for (uint32_t i = 0; i < w * h; ++i) numRays += light_trace_path<true>(world.S, camera, maxDepth, rng, splat);
    FILE* f = fopen(argv[8], "wb");
    fwrite(sum.data(), 4, sum.size(), f);
    fclose(f);
    printf("{\"numRays\": %llu, \"numSplats\": %llu}\n", (unsigned long long)numRays, (unsigned long long)numSplats);
    return 0;
}

// BSDF::Pdf, forward and reverse, of every material (constant parameters) for a list of local direction pairs
static int cmdBsdfPdf(int argc, char** argv)
{
    if (argc < 5) die("bsdfpdf: arguments");
    World world; world.load(argv[2]);
    FILE* f = fopen(argv[3], "rb");
    if (!f) die("cannot open dirs");
    std::vector<float> dirs;
    float v[6];
    while (fread(v, 4, 6, f) == 6) dirs.insert(dirs.end(), v, v + 6);
    fclose(f);
    FILE* of = fopen(argv[4], "wb");
    for (size_t m = 0; m < world.file.materials.size(); ++m)
    {
        const RtbMaterial& mat = world.file.materials[m];
        Vertex vx;
        memset(&vx, 0, sizeof(vx));
        vx.bsdf = mat.bsdf; vx.baseColor = v4(mat.baseColor[0], mat.baseColor[1], mat.baseColor[2], mat.baseColor[3]);
        vx.roughness = mat.roughness; vx.metalness = mat.metalness; vx.IoR = mat.IoR; vx.K = mat.K;
        for (size_t i = 0; i + 6 <= dirs.size(); i += 6)
        {
            const V3 out = v3(dirs[i], dirs[i + 1], dirs[i + 2]), in = v3(dirs[i + 3], dirs[i + 4], dirs[i + 5]);
            const float r[2] = { bsdf_pdf(vx, out, in, false), bsdf_pdf(vx, out, in, true) };
            fwrite(r, 4, 2, of);
        }
    }
    fclose(of);
    return 0;
}

// ILight::Illuminate without solid-angle sampling (the bidirectional renderers' call) for every light and shading point
static int cmdIlluminate(int argc, char** argv)
{
    if (argc < 5) die("illuminate: arguments");
    World world; world.load(argv[2]);
    FILE* f = fopen(argv[3], "rb");
    if (!f) die("cannot open points");
    std::vector<float> pts;
    float v[15];
    while (fread(v, 4, 15, f) == 15) pts.insert(pts.end(), v, v + 15);
    fclose(f);
    FILE* of = fopen(argv[4], "wb");
    for (uint32_t l = 0; l < world.S.numLights; ++l)
        for (size_t i = 0; i + 15 <= pts.size(); i += 15)
        {
            const float* p = &pts[i];
            Vertex vx;
            memset(&vx, 0, sizeof(vx));
            vx.position = v3(p[0], p[1], p[2]); vx.tangent = v3(p[3], p[4], p[5]); vx.bitangent = v3(p[6], p[7], p[8]); vx.normal = v3(p[9], p[10], p[11]);
            IlluminateResultBidir r;
            const V4 c = light_illuminate_bidir<true>(world.S, world.S.lights[l], vx, p[12], p[13], p[14], r);
            const float out[10] = { r.directionToLight.x, r.directionToLight.y, r.directionToLight.z, r.distance, r.directPdfW, r.emissionPdfW, r.cosAtLight, c.x, c.y, c.z };
            fwrite(out, 4, 10, of);
        }
    fclose(of);
    return 0;
}

// ILight::GetRadiance with both pdf outputs for every light a ray can hit
static int cmdRadiance(int argc, char** argv)
{
    if (argc < 5) die("radiance: arguments");
    World world; world.load(argv[2]);
    FILE* f = fopen(argv[3], "rb");
    if (!f) die("cannot open rays");
    std::vector<float> rays;
    float v[4];
    while (fread(v, 4, 4, f) == 4) rays.insert(rays.end(), v, v + 4);
    fclose(f);
    FILE* of = fopen(argv[4], "wb");
    for (uint32_t l = 0; l < world.S.numLights; ++l)
    {
        const DLight& light = world.S.lights[l];
        if (light.type == RTB_LIGHT_POINT || light.type == RTB_LIGHT_SPOT) continue;
        for (size_t i = 0; i + 4 <= rays.size(); i += 4)
        {
            float directPdfA = -1.0f, emissionPdfW = -1.0f;
            const V4 c = light_radiance_bidir<true>(world.S, light, v3(rays[i], rays[i + 1], rays[i + 2]), rays[i + 3], directPdfA, emissionPdfW);
            const float out[5] = { c.x, c.y, c.z, directPdfA, emissionPdfW };
            fwrite(out, 4, 5, of);
        }
    }
    fclose(of);
    return 0;
}

// VCM (hostsim_vcm.h): camera rays and camera-path samples from the pass script, light paths from the host generator
static int cmdVcm(int argc, char** argv)
{
    if (argc < 10) die("vcm: arguments");
    World world; world.load(argv[2]);
    Passes passes; passes.load(argv[3]);
    const uint32_t w = atoi(argv[4]), h = atoi(argv[5]);
    HostRng rng{ strtoull(argv[6], nullptr, 10) * 0x9E3779B97F4A7C15ULL + 1u };
    VcmSettings settings;
    settings.useVertexConnection = atoi(argv[7]) != 0; settings.useVertexMerging = atoi(argv[8]) != 0;
    if (argc > 10 && strcmp(argv[10], "-") != 0) world.loadBlueNoise(argv[10]);
    PassParams P;
    memset(&P, 0, sizeof(P));
    fillCamera(P, world.file.camera, w, h, tanf(world.file.fieldOfView * 0.5f));
    fillBokeh(P, world.file);
    P.useBlueNoise = world.S.blueNoise != nullptr; P.numSeedDims = passes.numDims;
    CameraConnection camera;
    camera.localToWorld = P.camLocalToWorld; camera.tanHalfFoV = P.tanHalfFoV; camera.aspectRatio = P.aspectRatio;
    camera.worldToScreen = camera_world_to_screen(P.camLocalToWorld, P.aspectRatio, P.tanHalfFoV);
    std::vector<float> sum((size_t)w * h * 3, 0.0f);
    auto splat = [&](float filmX, float filmY, const V3& value)
    {
        const float cx = filmX * (float)w + 0.0f, cy = filmY * (float)h + 0.5f;
        int32_t ix = cvt_rn(cx);
        const float ux = rng.next_float();
        rng.next_float(); rng.next_float(); rng.next_float();
        if (ux < cx - (float)ix) ix++;
        const int32_t x = ix, y = (int32_t)(h - 1) - (int32_t)cy;
        if ((uint32_t)x < w && (uint32_t)y < h) { float* px = &sum[((size_t)y * w + x) * 3]; px[0] += value.x; px[1] += value.y; px[2] += value.z; }
    };
    // the camera path draws from the pixel's sampler like the reference's (ctx.sampler: Halton + blue noise - its low
    // discrepancy is what keeps the reference inside its own furnace-test tolerances); the light path from the generator
    PathState* current = nullptr;
    auto sampler3 = [&]() { return sampler_get_float(current->sampler, P, world.S.blueNoise); };
    VcmModel<HostRng> vcm(world.S, camera, settings, rng, w, h);
    for (uint32_t p = 0; p < passes.numPasses; ++p)
    {
        P.passIndex = p;
        P.seed = &passes.seeds[(size_t)p * passes.numDims];
        P.sampleOffsetX = passes.offsets[p * 2]; P.sampleOffsetY = passes.offsets[p * 2 + 1];
        vcm.BeginPass(p);
        for (uint32_t y = 0; y < h; ++y) for (uint32_t x = 0; x < w; ++x)
        {
            PathState ps;
            init_path(P, world.S.blueNoise, x, y, 0, ps);
            current = &ps;
            const V4 c = vcm.RenderPixel(ps.ray, sampler3, splat);
            float* px = &sum[((size_t)y * w + x) * 3];
            px[0] += c.x; px[1] += c.y; px[2] += c.z;
        }
    }
    FILE* f = fopen(argv[9], "wb");
    fwrite(sum.data(), 4, sum.size(), f);
    fclose(f);
    return 0;
}

// the arithmetic of rtb_block_errors on raster sum buffers (the three kernels, one thread after the other)
static int cmdBlockErrors(int argc, char** argv)
{
    if (argc < 9) die("blockerrors: arguments");
    const uint32_t w = atoi(argv[4]), h = atoi(argv[5]), passes = atoi(argv[6]);
    std::vector<float> sum((size_t)w * h * 3), sum2((size_t)w * h * 3);
    FILE* f = fopen(argv[2], "rb");
    if (!f || fread(sum.data(), 4, sum.size(), f) != sum.size()) die("cannot read sum");
    fclose(f);
    f = fopen(argv[3], "rb");
    if (!f || fread(sum2.data(), 4, sum2.size(), f) != sum2.size()) die("cannot read sum2");
    fclose(f);
    std::vector<float> pixelErrors((size_t)w * h);
    for (size_t i = 0; i < pixelErrors.size(); ++i)
        pixelErrors[i] = pixel_error(v3(sum[3 * i], sum[3 * i + 1], sum[3 * i + 2]), v3(sum2[3 * i], sum2[3 * i + 1], sum2[3 * i + 2]), 1.0f / (float)passes);
    f = fopen(argv[7], "rb");
    if (!f) die("cannot read blocks");
    FILE* of = fopen(argv[8], "wb");
    RtbBlock b;
    while (fread(&b, sizeof(b), 1, f) == 1)
    {
        std::vector<float> rows;
        for (uint32_t y = b.minY; y < b.maxY; ++y) rows.push_back(row_error(pixelErrors.data() + (size_t)y * w, b.minX, b.maxX));
        const float e = block_error(rows.data(), (uint32_t)rows.size(), b.maxX - b.minX, w * h);
        fwrite(&e, 4, 1, of);
    }
    fclose(f); fclose(of);
    return 0;
}

// the arithmetic of rtb_read_front on a raster sum buffer
static int cmdFront(int argc, char** argv)
{
    if (argc < 13) die("front: arguments");
    const uint32_t w = atoi(argv[3]), h = atoi(argv[4]);
    std::vector<float> sum((size_t)w * h * 3);
    FILE* f = fopen(argv[2], "rb");
    if (!f || fread(sum.data(), 4, sum.size(), f) != sum.size()) die("cannot read sum");
    fclose(f);
    PostParams P;
    float fr = 1.0f, fg = 1.0f, fb = 1.0f;
    if (sscanf(argv[11], "%f,%f,%f", &fr, &fg, &fb) != 3) die("front: filter");
    const float exposureScale = powf(2.0f, (float)atof(argv[6]));
    P.colorScale = v3(fr * exposureScale, fg * exposureScale, fb * exposureScale);
    P.pixelScaling = 1.0f / (float)(uint32_t)atoi(argv[5]);
    P.contrast = (float)atof(argv[7]); P.saturation = (float)atof(argv[8]); P.ditheringStrength = (float)atof(argv[9]);
    P.bloomFactor = argc > 13 ? (float)atof(argv[13]) : 0.0f; P.tonemapper = atoi(argv[10]); P.ditherSeed = 1;
    const size_t levelStride = (size_t)w * h * 3;
    std::vector<float> blurred;
    if (P.bloomFactor > 0.0f)
    {
        // the launches of rtb_read_front, one "thread" after the other
        blurred.resize(levelStride * RTB_BLOOM_LEVELS);
        std::vector<float> scratch(levelStride * 2);
        float sigma = 2.0f;
        for (uint32_t l = 0; l < RTB_BLOOM_LEVELS; ++l, sigma *= 2.5f)
        {
            const BlurSchedule schedule = blur_schedule(sigma);
            const float* from = l == 0 ? sum.data() : blurred.data() + (l - 1) * levelStride;
            float* level = blurred.data() + l * levelStride;
            for (uint32_t t = 0; t < h * 3; ++t)
                blur_line(from + (size_t)(t / 3) * w * 3 + t % 3, level + (size_t)(t / 3) * w * 3 + t % 3, 3, scratch.data() + t, scratch.data() + levelStride + t,
                          (size_t)h * 3, schedule, RTB_BLUR_KEPT_HORIZONTAL, w);
            for (uint32_t t = 0; t < w * 3; ++t)
                blur_line(level + t, level + t, (size_t)w * 3, scratch.data() + t, scratch.data() + levelStride + t, (size_t)w * 3, schedule, RTB_BLUR_KEPT_VERTICAL, h);
        }
    }
    std::vector<uint32_t> out((size_t)w * h);
    for (uint32_t i = 0; i < w * h; ++i)
    {
        const V3 bloom = P.bloomFactor > 0.0f ? bloom_color(blurred.data(), levelStride, i) : v3s(0.0f);
        out[i] = postprocess_pixel(P, v3(sum[3 * i], sum[3 * i + 1], sum[3 * i + 2]), bloom, i, P.pixelScaling);
    }
    FILE* of = fopen(argv[12], "wb");
    fwrite(out.data(), 4, out.size(), of);
    fclose(of);
    return 0;
}

int main(int argc, char** argv)
{
    // the reference runs with FTZ + DAZ (Core/Math/Math.cpp:27-34); the CUDA build uses -ftz=true
    _MM_SET_FLUSH_ZERO_MODE(_MM_FLUSH_ZERO_ON);
    _MM_SET_DENORMALS_ZERO_MODE(_MM_DENORMALS_ZERO_ON);
    if (argc < 2) die("usage: hostsim render|trace|sampler|primary ...");
    const std::string cmd = argv[1];
    if (cmd == "render") return cmdRender(argc, argv);
    if (cmd == "trace") return cmdTrace(argc, argv);
    if (cmd == "texeval") return cmdTexEval(argc, argv);
    if (cmd == "sampler") return cmdSampler(argc, argv);
    if (cmd == "primary") return cmdPrimary(argc, argv);
    if (cmd == "front") return cmdFront(argc, argv);
    if (cmd == "emit") return cmdEmit(argc, argv);
    if (cmd == "vcm") return cmdVcm(argc, argv);
    if (cmd == "radiance") return cmdRadiance(argc, argv);
    if (cmd == "illuminate") return cmdIlluminate(argc, argv);
    if (cmd == "bsdfpdf") return cmdBsdfPdf(argc, argv);
    if (cmd == "lighttrace") return cmdLightTrace(argc, argv);
    if (cmd == "camera") return cmdCamera(argc, argv);
    if (cmd == "blockerrors") return cmdBlockErrors(argc, argv);
    die("unknown command");
    return 2;
}
