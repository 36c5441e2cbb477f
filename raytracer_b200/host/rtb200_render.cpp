// rtb200_render — headless command line renderer over the host object model (the reference ships an interactive
// viewer, Demo/Main.cpp:6-46; its flags are mirrored where they make sense: -w -h -s --renderer --data).
//
//   rtb200_render -s scene.json --data DIR -w 1920 -h 1080 --spp 64 --renderer "Path Tracer MIS" --depth 8
//                 --dims 128 --strategy all --seed 12345 --device 0 --out sum.f32 [--flat scene.rtb]
//                 [--gpus N [--check-gather 1]]      tiles over N devices of this box from this one process
//                 [--front image.bmp --exposure E --contrast C --saturation S --dither D --bloom B --tonemapper 0..3]
//                 [--adaptive 1 --adaptive-initial 10 --adaptive-min-block 4 --adaptive-max-block 256
//                  --adaptive-subdivision 0.005 --adaptive-convergence 0.0001]      AdaptiveRenderingSettings, Context.h:35-43
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "rt_loader.h"
#include "rt_viewport.h"

int main(int argc, char** argv)
{
    std::map<std::string, std::string> a;
    for (int i = 1; i < argc; ++i)
    {
        std::string k = argv[i];
        if (k[0] == '-' && i + 1 < argc) { a[k] = argv[++i]; }
    }
    auto get = [&](const char* k, const char* d) { auto it = a.find(k); return it == a.end() ? std::string(d) : it->second; };
    const std::string scenePath = get("-s", get("--scene", "").c_str());
    if (scenePath.empty())
    {
        fprintf(stderr, "usage: rtb200_render -s scene.json [--data DIR] [-w W] [-h H] [--spp N] [--renderer NAME] [--depth D] [--dims K]\n"
                        "                     [--strategy all|single] [--seed S] [--device I] [--bluenoise FILE] [--out sum.f32] [--flat scene.rtb]\n"
                        "                     [--adaptive 1] [--adaptive-initial N] [--adaptive-min-block N] [--adaptive-max-block N] [--adaptive-subdivision T] [--adaptive-convergence T]\n"
                        "                     [--front image.bmp] [--exposure E] [--contrast C] [--saturation S] [--dither D] [--bloom B] [--tonemapper 0..3]\n");
        return 2;
    }
    const uint32_t w = atoi(get("-w", "1280").c_str()), h = atoi(get("-h", "720").c_str()), spp = atoi(get("--spp", "16").c_str());
    rt::Scene scene; rt::Camera camera;
    std::string err;
    if (!rt::helpers::LoadScene(scenePath, get("--data", ""), scene, camera, err)) { fprintf(stderr, "[ERROR] %s\n", err.c_str()); return 1; }
    if (!scene.BuildBVH()) { fprintf(stderr, "[ERROR] BuildBVH failed\n"); return 1; }
    camera.SetPerspective((float)w / (float)h, camera.mFieldOfView);
    if (a.count("--flat"))
    {
        rtb::SceneFile flat;
        scene.Flatten(flat);
        camera.Describe(flat.camera);
        camera.DescribeBokeh(flat);
        flat.fieldOfView = camera.mFieldOfView;
        if (!flat.Save(a["--flat"])) { fprintf(stderr, "[ERROR] cannot write %s\n", a["--flat"].c_str()); return 1; }
        if (spp == 0) return 0;
    }
    const int numGpus = atoi(get("--gpus", "1").c_str());
    if (numGpus > 1)
    {
        // the frame's tiles over several devices of this box (tile t -> device t % N), one process
        std::vector<int> devices;
        for (int i = 0; i < numGpus; ++i) devices.push_back(atoi(get("--device", "0").c_str()) + i);
        rt::ViewportGroupB200 group;
        rt::RenderingParams p;
        p.maxRayDepth = atoi(get("--depth", "8").c_str());
        p.dimensions = atoi(get("--dims", "128").c_str());
        p.lightSamplingStrategy = get("--strategy", "all") == "single" ? RTB_LIGHT_SAMPLING_SINGLE : RTB_LIGHT_SAMPLING_ALL;
        auto bail = [&]() { fprintf(stderr, "[ERROR] %s\n", group.GetLastError().c_str()); return 1; };
        if (!group.Create(devices, 0, get("--bluenoise", ""))) return bail();
        if (!group.SetScene(scene, &camera)) return bail();
        group.SetRenderingParams(p);
        group.SetSeed(strtoull(get("--seed", "12345").c_str(), nullptr, 10));
        if (!group.SetRenderer(get("--renderer", "Path Tracer MIS"))) return bail();
        if (!group.Resize(w, h)) return bail();
        const auto t0 = std::chrono::steady_clock::now();
        if (!group.RenderPasses(camera, spp)) return bail();
        if (!group.Synchronize()) return bail();
        const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        RtbCounters c;
        if (!group.GetCounters(c)) return bail();
        std::vector<float> img((size_t)w * h * 3);
        if (!group.GetSumBuffer(img.data())) return bail();
        if (atoi(get("--check-gather", "0").c_str()))
        {
            // the same frame through the library's NCCL all-gather: every device ends up with the whole film
            std::vector<void*> gathered;
            if (!group.FinishFrame(gathered)) return bail();
            for (int i = 0; i < numGpus; ++i)
            {
                std::vector<float> viaNccl((size_t)w * h * 3);
                if (rtb_unshard_film(group.GetViewport(i).GetContext(), gathered[i], viaNccl.data()) != RTB_OK) { fprintf(stderr, "[ERROR] %s\n", rtb_last_error()); return 1; }
                if (memcmp(viaNccl.data(), img.data(), img.size() * sizeof(float)) != 0) { fprintf(stderr, "[ERROR] device %d: the all-gathered film differs from the scattered frame\n", i); return 1; }
            }
            fprintf(stderr, "all-gathered film on %d devices == scattered host frame\n", numGpus);
        }
        printf("{\"width\": %u, \"height\": %u, \"passes\": %u, \"gpus\": %d, \"seconds\": %.6f, \"numPrimaryRays\": %llu, \"numRays\": %llu, \"numShadowRays\": %llu, "
               "\"numShadowRaysHit\": %llu, \"mrays_per_s\": %.3f, \"msamples_per_s\": %.3f}\n", w, h, spp, numGpus, s,
               (unsigned long long)c.numPrimaryRays, (unsigned long long)c.numRays, (unsigned long long)c.numShadowRays, (unsigned long long)c.numShadowRaysHit,
               (c.numRays + c.numShadowRays) / s * 1e-6, c.numPrimaryRays / s * 1e-6);
        if (a.count("--out"))
        {
            FILE* f = fopen(a["--out"].c_str(), "wb");
            if (!f) { fprintf(stderr, "[ERROR] cannot write %s\n", a["--out"].c_str()); return 1; }
            fwrite(img.data(), 4, img.size(), f);
            fclose(f);
        }
        return 0;
    }
    rt::ViewportB200 vp;
    if (!vp.Create(atoi(get("--device", "0").c_str()), 1, 0, 0, get("--bluenoise", ""))) { fprintf(stderr, "[ERROR] %s\n", vp.GetLastError().c_str()); return 1; }
    if (!vp.SetScene(scene, &camera)) { fprintf(stderr, "[ERROR] %s\n", vp.GetLastError().c_str()); return 1; }
    rt::RenderingParams p;
    p.maxRayDepth = atoi(get("--depth", "8").c_str());
    p.dimensions = atoi(get("--dims", "128").c_str());
    p.lightSamplingStrategy = get("--strategy", "all") == "single" ? RTB_LIGHT_SAMPLING_SINGLE : RTB_LIGHT_SAMPLING_ALL;
    p.adaptiveSettings.enable = atoi(get("--adaptive", "0").c_str()) != 0;
    p.adaptiveSettings.numInitialPasses = atoi(get("--adaptive-initial", "10").c_str());
    p.adaptiveSettings.minBlockSize = atoi(get("--adaptive-min-block", "4").c_str());
    p.adaptiveSettings.maxBlockSize = atoi(get("--adaptive-max-block", "256").c_str());
    p.adaptiveSettings.subdivisionTreshold = (float)atof(get("--adaptive-subdivision", "0.005").c_str());
    p.adaptiveSettings.convergenceTreshold = (float)atof(get("--adaptive-convergence", "0.0001").c_str());
    vp.SetRenderingParams(p);
    vp.SetSeed(strtoull(get("--seed", "12345").c_str(), nullptr, 10));
    if (!vp.SetRenderer(get("--renderer", "Path Tracer MIS"))) { fprintf(stderr, "[ERROR] %s\n", vp.GetLastError().c_str()); return 1; }
    if (!vp.Resize(w, h)) { fprintf(stderr, "[ERROR] %s\n", vp.GetLastError().c_str()); return 1; }
    const auto t0 = std::chrono::steady_clock::now();
    // all passes in one call: the device traces them in shared wavefronts, the image is the one `spp` Render() calls give
    if (!vp.RenderPasses(camera, spp)) { fprintf(stderr, "[ERROR] %s\n", vp.GetLastError().c_str()); return 1; }
    vp.Synchronize();
    const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    RtbCounters c;
    vp.GetCounters(c);
    rt::RenderingProgress progress;
    if (vp.GetProgress(progress))
        fprintf(stderr, "progress: %u passes, %u active blocks, %u active pixels, converged %.4f, average error %g\n", progress.passesFinished,
                progress.activeBlocks, progress.activePixels, progress.converged, progress.averageError);
    printf("{\"width\": %u, \"height\": %u, \"passes\": %u, \"seconds\": %.6f, \"numPrimaryRays\": %llu, \"numRays\": %llu, \"numShadowRays\": %llu, "
           "\"numShadowRaysHit\": %llu, \"mrays_per_s\": %.3f, \"msamples_per_s\": %.3f}\n", w, h, spp, s,
           (unsigned long long)c.numPrimaryRays, (unsigned long long)c.numRays, (unsigned long long)c.numShadowRays, (unsigned long long)c.numShadowRaysHit,
           (c.numRays + c.numShadowRays) / s * 1e-6, c.numPrimaryRays / s * 1e-6);
    if (a.count("--out"))
    {
        std::vector<float> img((size_t)w * h * 3);
        if (!vp.GetSumBuffer(img.data())) { fprintf(stderr, "[ERROR] %s\n", vp.GetLastError().c_str()); return 1; }
        FILE* f = fopen(a["--out"].c_str(), "wb");
        fwrite(img.data(), 4, img.size(), f);
        fclose(f);
    }
    if (a.count("--front"))
    {
        // Viewport::GetFrontBuffer as a top-down 32-bit BMP
        rt::PostprocessParams pp;
        pp.exposure = (float)atof(get("--exposure", "0").c_str());
        pp.contrast = (float)atof(get("--contrast", "0.8").c_str());
        pp.saturation = (float)atof(get("--saturation", "0.98").c_str());
        pp.ditheringStrength = (float)atof(get("--dither", "0.005").c_str());
        pp.bloomFactor = (float)atof(get("--bloom", "0").c_str());
        pp.tonemapper = atoi(get("--tonemapper", "3").c_str());
        vp.SetPostprocessParams(pp);
        std::vector<uint32_t> img((size_t)w * h);
        if (!vp.GetFrontBuffer(img.data())) { fprintf(stderr, "[ERROR] %s\n", vp.GetLastError().c_str()); return 1; }
        FILE* f = fopen(a["--front"].c_str(), "wb");
        if (!f) { fprintf(stderr, "[ERROR] cannot write %s\n", a["--front"].c_str()); return 1; }
        const uint32_t dataBytes = w * h * 4, fileHeader[3] = { 54 + dataBytes, 0, 54 };
        const int32_t topDown = -(int32_t)h;
        const uint32_t info[10] = { 40, w, (uint32_t)topDown, 1u | (32u << 16), 0, dataBytes, 2835, 2835, 0, 0 };
        fwrite("BM", 1, 2, f); fwrite(fileHeader, 4, 3, f); fwrite(info, 4, 10, f);
        fwrite(img.data(), 4, img.size(), f);
        fclose(f);
    }
    return 0;
}
