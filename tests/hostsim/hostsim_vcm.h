// hostsim_vcm.h — CPU model of the reference's VCM renderer (Rendering/VertexConnectionAndMerging.cpp): the control flow
// of RenderPixel / TraceLightPath around the per-vertex functions of raytracer_b200/csrc/rtb_bidir.h (bidir_*), which
// are written for host and device alike.
//
// TEST TOOL ONLY (see hostsim.cpp).  It exists to pin the ALGORITHM - MIS bookkeeping (dVCM / dVC / dVM), light vertex
// storage, camera connection, vertex connection, vertex merging with a one-pass delay - statistically against the
// reference before a device version is written; the reference draws its light paths from per-thread
// generators, so its VCM images are reproducible only statistically.  One thread, one pixel at a time, like the
// reference's RenderPixel; the photon search is a plain uniform grid (any exact radius search returns the same set).
#pragma once

#include <unordered_map>
#include <vector>

#include "../../raytracer_b200/csrc/rtb_integrator.h"
#include "../../raytracer_b200/csrc/rtb_bidir.h"

namespace rtb {

struct VcmSettings
{
    bool useVertexConnection = true, useVertexMerging = true;
    uint32_t maxPathLength = 10;                                                   // VertexConnectionAndMerging.cpp:55-62
    float initialMergingRadius = 0.02f, minMergingRadius = 0.02f, mergingRadiusMultiplier = 1.0f;
};

template <typename Rng>
class VcmModel
{
public:
    VcmModel(const SceneView& s, const CameraConnection& c, const VcmSettings& st, Rng& r, uint32_t width, uint32_t height)
        : S(s), camera(c), settings(st), rng(r), mWidth(width), mHeight(height)
    {
        memset(&pass, 0, sizeof(pass));
        pass.useVertexConnection = st.useVertexConnection; pass.useVertexMerging = st.useVertexMerging; pass.maxPathLength = st.maxPathLength;
    }

    // PreRender + PreRenderGlobal, :80-160: radii, MIS factors, the photons of the previous pass into the grid
    void BeginPass(uint32_t passNumber)
    {
        bidir_begin_pass(pass, passNumber, (float)(mWidth * mHeight), settings.initialMergingRadius, settings.minMergingRadius, settings.mergingRadiusMultiplier);
        mPhotons.swap(mNewPhotons);
        mNewPhotons.clear();
        mGrid.clear();
        if (pass.useVertexMerging)
            for (uint32_t i = 0; i < mPhotons.size(); ++i) mGrid[CellKey(mPhotons[i].position, 0, 0, 0)].push_back(i);
    }

    // RenderPixel, :170-290.  `sampler3()` supplies the camera path's numbers (ctx.sampler), `splat` receives the light
    // path's camera connections.
    template <typename Sampler3, typename Splat>
    V4 RenderPixel(const Ray& primaryRay, Sampler3& sampler3, Splat& splat)
    {
        const auto occluded = [this](const Ray& ray, float tMax) { return scene_traverse_shadow<false>(S, ray, tMax, nullptr); };
        TraceLightPath(splat);
        V4 result = v4s(0.0f);
        BidirPath path;
        bidir_begin_camera_path(path, primaryRay, camera);
        Vertex vx;
        for (;;)
        {
            Hit hit;
            hit.objectId = RTB_INVALID_OBJECT; hit.subObjectId = 0; hit.distance = INFINITY; hit.u = 0.0f; hit.v = 0.0f;
            scene_traverse<false>(S, path.ray, hit, nullptr);
            if (hit.distance == INFINITY)
            {
                V4 sum = v4s(0.0f);
                for (uint32_t i = 0; i < S.numGlobalLights; ++i) sum = sum + bidir_evaluate_light<true>(S, pass, S.lights[S.globalLights[i]], nullptr, path);
                result = MulAdd(path.throughput, sum, result);
                break;
            }
            evaluate_intersection<true>(S, path.ray, hit, vx, RTB_NO_MATERIAL);
            bidir_arrive(path, vx, hit.distance, false);
            if (hit.subObjectId == RTB_LIGHT_OBJECT)
            {
                const DLight& light = S.lights[S.objShade[hit.objectId].lightIndex];
                result = MulAdd(path.throughput, bidir_evaluate_light<true>(S, pass, light, &vx, path), result);
                break;
            }
            vx.outgoingDirWorld = neg(path.ray.dir);
            evaluate_shading_data<true>(S, vx);
            result = MulAdd(path.throughput, vx.emission, result);
            if (path.length >= pass.maxPathLength) break;
            const bool isDeltaBsdf = bidir_is_delta_bsdf(vx);
            if (!isDeltaBsdf && pass.useVertexConnection)
            {
                V4 sum = v4s(0.0f);
                for (uint32_t l = 0; l < S.numLights; ++l)
                {
                    const float u0 = sampler3(), u1 = sampler3(), u2 = sampler3();
                    sum = sum + bidir_sample_light<true>(S, pass, S.lights[l], vx, path, u0, u1, u2, occluded);
                }
                result = MulAdd(path.throughput, sum, result);
            }
            if (!isDeltaBsdf && pass.useVertexConnection && !mLightVertices.empty())
            {
                V4 sum = v4s(0.0f);
                for (const BidirLightVertex& lv : mLightVertices)
                {
                    if (lv.pathLength + path.length + 1u > pass.maxPathLength) break;
                    sum = MulAdd(lv.throughput, bidir_connect_vertices(pass, path, vx, lv, occluded), sum);
                }
                result = MulAdd(path.throughput, sum, result);
            }
            if (!isDeltaBsdf && pass.useVertexMerging && pass.iteration > 0)
                result = MulAdd(path.throughput * MergeVertices(path, vx), v4s(pass.mergingNormalization), result);
            if (path.length > pass.maxPathLength) break;
            const float s0 = sampler3(), s1 = sampler3(), s2 = sampler3();
            if (!bidir_advance_path(pass, path, vx, s0, s1, s2)) break;
        }
        return result;
    }

private:
    static V4 MulAdd(const V4& a, const V4& b, const V4& c) { return v4(fma_(a.x, b.x, c.x), fma_(a.y, b.y, c.y), fma_(a.z, b.z, c.z), fma_(a.w, b.w, c.w)); }

    uint64_t CellKey(const V3& p, int dx, int dy, int dz) const
    {
        const float inv = 1.0f / pass.radiusVM;
        const int64_t x = (int64_t)floorf(p.x * inv) + dx, y = (int64_t)floorf(p.y * inv) + dy, z = (int64_t)floorf(p.z * inv) + dz;
        return (uint64_t)(x * 73856093LL) ^ (uint64_t)(y * 19349663LL) ^ (uint64_t)(z * 83492791LL);
    }

    // TraceLightPath, :292-428
    template <typename Splat>
    void TraceLightPath(Splat& splat)
    {
        const auto occluded = [this](const Ray& ray, float tMax) { return scene_traverse_shadow<false>(S, ray, tMax, nullptr); };
        mLightVertices.clear();
        BidirPath path;
        path.throughput = v4(1.0f, 1.0f, 1.0f, 1.0f); path.dVC = path.dVM = path.dVCM = 0.0f;
        const uint32_t pick = S.numLights ? rng.next_uint() : 0u;
        const float p0 = S.numLights ? rng.next_float() : 0.0f, p1 = S.numLights ? rng.next_float() : 0.0f, p2 = S.numLights ? rng.next_float() : 0.0f;
        const float d0 = S.numLights ? rng.next_float() : 0.0f, d1 = S.numLights ? rng.next_float() : 0.0f;
        if (!bidir_generate_light_sample<true>(S, pass, pick, p0, p1, p2, d0, d1, path)) return;
        for (;;)
        {
            Hit hit;
            hit.objectId = RTB_INVALID_OBJECT; hit.subObjectId = 0; hit.distance = INFINITY; hit.u = 0.0f; hit.v = 0.0f;
            scene_traverse<false>(S, path.ray, hit, nullptr);
            if (hit.distance == INFINITY || hit.subObjectId == RTB_LIGHT_OBJECT) break;
            BidirLightVertex lv;
            evaluate_intersection<true>(S, path.ray, hit, lv.vx, RTB_NO_MATERIAL);
            lv.vx.outgoingDirWorld = neg(path.ray.dir);
            evaluate_shading_data<true>(S, lv.vx);
            bidir_arrive(path, lv.vx, hit.distance, true);
            if (!bidir_is_delta_bsdf(lv.vx))
            {
                if (pass.useVertexConnection && mLightVertices.size() < 256)               // g_MaxLightVertices, :30
                {
                    lv.pathLength = path.length; lv.throughput = path.throughput; lv.dVC = path.dVC; lv.dVM = path.dVM; lv.dVCM = path.dVCM;
                    mLightVertices.push_back(lv);
                    float filmX, filmY;
                    V3 value;
                    if (bidir_connect_to_camera(pass, camera, lv, occluded, filmX, filmY, value)) splat(filmX, filmY, value);
                }
                if (pass.useVertexMerging)
                {
                    BidirPhoton ph;
                    ph.position = lv.vx.position; ph.direction = lv.vx.outgoingDirWorld; ph.throughput = xyz(path.throughput);
                    ph.dVM = path.dVM; ph.dVCM = path.dVCM;
                    mNewPhotons.push_back(ph);
                }
            }
            if (path.length + 2 > pass.maxPathLength) break;
            const float s0 = rng.next_float(), s1 = rng.next_float(), s2 = rng.next_float();
            if (!bidir_advance_path(pass, path, lv.vx, s0, s1, s2)) break;
        }
    }

    // MergeVertices, :808-906: every photon of the previous pass within radiusVM of the camera vertex
    V4 MergeVertices(const BidirPath& cameraPath, const Vertex& vx) const
    {
        V4 contribution = v4s(0.0f);
        const float radiusSqr = s_sqr(pass.radiusVM);
        uint64_t visited[27];
        int numVisited = 0;
        for (int dz = -1; dz <= 1; ++dz) for (int dy = -1; dy <= 1; ++dy) for (int dx = -1; dx <= 1; ++dx)
        {
            const uint64_t key = CellKey(vx.position, dx, dy, dz);
            bool seen = false;
            for (int k = 0; k < numVisited; ++k) seen = seen || visited[k] == key;      // two neighbour cells with one hash
            if (seen) continue;
            visited[numVisited++] = key;
            const auto it = mGrid.find(key);
            if (it == mGrid.end()) continue;
            for (const uint32_t index : it->second)
            {
                const BidirPhoton& photon = mPhotons[index];
                if (sqr_length3(vx.position - photon.position) > radiusSqr) continue;
                bidir_merge_photon(pass, cameraPath, vx, photon, contribution);
            }
        }
        return contribution;
    }

    const SceneView& S;
    CameraConnection camera;
    VcmSettings settings;
    Rng& rng;
    uint32_t mWidth, mHeight;
    BidirPass pass;
    std::vector<BidirLightVertex> mLightVertices;
    std::vector<BidirPhoton> mPhotons, mNewPhotons;
    std::unordered_map<uint64_t, std::vector<uint32_t>> mGrid;
};

} // namespace rtb
