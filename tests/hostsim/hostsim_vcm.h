// hostsim_vcm.h — CPU model of the reference's VCM renderer (Rendering/VertexConnectionAndMerging.cpp), assembled from
// the checked parts in raytracer_b200/csrc/rtb_bidir.h and the path tracers' traversal / BSDF code.
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
    uint32_t maxPathLength = 10;
    float initialMergingRadius = 0.02f, minMergingRadius = 0.02f, mergingRadiusMultiplier = 1.0f;
};

template <typename Rng>
class VcmModel
{
public:
    struct PathStateVcm
    {
        Ray ray;
        V4 throughput = v4(1.0f, 1.0f, 1.0f, 1.0f);
        float dVC = 0.0f, dVM = 0.0f, dVCM = 0.0f;
        uint32_t length = 1;
        bool lastSpecular = true, isFiniteLight = false;
    };
    struct LightVertex { Vertex vx; V4 throughput; float dVC, dVM, dVCM; uint32_t pathLength; };
    struct Photon { V3 position, direction, throughput; float dVM, dVCM; };

    VcmModel(const SceneView& s, const CameraConnection& c, const VcmSettings& st, Rng& r, uint32_t width, uint32_t height)
        : S(s), camera(c), settings(st), rng(r), mWidth(width), mHeight(height) {}

    // VertexConnectionAndMerging::PreRender + PreRenderGlobal, :80-160: radii, MIS factors, photons of the previous pass
    void BeginPass(uint32_t passNumber)
    {
        mIteration = passNumber;
        mLightPathsCount = (float)(mWidth * mHeight);
        if (passNumber == 0) mRadiusVC = mRadiusVM = settings.initialMergingRadius;
        else
        {
            mRadiusVM = mRadiusVC;
            mRadiusVC *= settings.mergingRadiusMultiplier;
            mRadiusVC = s_max(mRadiusVC, settings.minMergingRadius);
        }
        mMergingNormalization = 1.0f / (s_sqr(mRadiusVM) * RTB_PI * mLightPathsCount);
        {
            const float etaVCM = RTB_PI * s_sqr(mRadiusVC) * mLightPathsCount;
            mMisVmWeightVC = (settings.useVertexMerging && passNumber > 0) ? etaVCM : 0.0f;
            mMisVcWeightVC = settings.useVertexConnection ? 1.0f / etaVCM : 0.0f;
        }
        {
            const float etaVCM = RTB_PI * s_sqr(mRadiusVM) * mLightPathsCount;
            mMisVmWeightVM = settings.useVertexMerging ? etaVCM : 0.0f;
            mMisVcWeightVM = settings.useVertexConnection ? 1.0f / etaVCM : 0.0f;
        }
        mPhotons.swap(mNewPhotons);
        mNewPhotons.clear();
        mGrid.clear();
        if (settings.useVertexMerging)
            for (uint32_t i = 0; i < mPhotons.size(); ++i) mGrid[CellKey(mPhotons[i].position, 0, 0, 0)].push_back(i);
    }

    // VertexConnectionAndMerging::RenderPixel, :170-290.  `sampler3()` supplies the camera path's numbers (ctx.sampler),
    // `splat` receives the light path's camera connections.
    template <typename Sampler3, typename Splat>
    V4 RenderPixel(const Ray& primaryRay, Sampler3& sampler3, Splat& splat)
    {
        TraceLightPath(splat);
        V4 result = v4s(0.0f);
        PathStateVcm path;
        path.ray = primaryRay;
        path.dVC = 0.0f; path.dVM = 0.0f;
        path.dVCM = 1.0f / camera_pdf_w(camera.localToWorld, camera.tanHalfFoV, camera.aspectRatio, primaryRay.dir);
        path.lastSpecular = true;
        Vertex vx;
        for (;;)
        {
            Hit hit;
            hit.objectId = RTB_INVALID_OBJECT; hit.subObjectId = 0; hit.distance = INFINITY; hit.u = 0.0f; hit.v = 0.0f;
            scene_traverse<false>(S, path.ray, hit, nullptr);
            if (hit.distance == INFINITY)
            {
                V4 sum = v4s(0.0f);
                for (uint32_t i = 0; i < S.numGlobalLights; ++i) sum = sum + EvaluateLight(S.lights[S.globalLights[i]], nullptr, path);
                result = mul_add(path.throughput, sum, result);
                break;
            }
            evaluate_intersection<true>(S, path.ray, hit, vx, RTB_NO_MATERIAL);
            {
                const float cosTheta = dot3(path.ray.dir, vx.normal);
                const float invMis = 1.0f / s_abs(cosTheta);
                path.dVCM *= s_sqr(hit.distance);
                path.dVCM *= invMis; path.dVC *= invMis; path.dVM *= invMis;
            }
            if (hit.subObjectId == RTB_LIGHT_OBJECT)
            {
                const DLight& light = S.lights[S.objShade[hit.objectId].lightIndex];
                result = mul_add(path.throughput, EvaluateLight(light, &vx, path), result);
                break;
            }
            vx.outgoingDirWorld = neg(path.ray.dir);
            evaluate_shading_data<true>(S, vx);
            result = mul_add(path.throughput, vx.emission, result);
            if (path.length >= settings.maxPathLength) break;
            const bool isDeltaBsdf = vx.bsdf == RTB_BSDF_METAL || vx.bsdf == RTB_BSDF_DIELECTRIC;
            if (!isDeltaBsdf && settings.useVertexConnection)
            {
                V4 sum = v4s(0.0f);
                for (uint32_t l = 0; l < S.numLights; ++l) sum = sum + SampleLight(S.lights[l], vx, path, sampler3);
                result = mul_add(path.throughput, sum, result);
            }
            if (!isDeltaBsdf && settings.useVertexConnection && !mLightVertices.empty())
            {
                V4 sum = v4s(0.0f);
                for (const LightVertex& lv : mLightVertices)
                {
                    if (lv.pathLength + path.length + 1u > settings.maxPathLength) break;
                    sum = mul_add(lv.throughput, ConnectVertices(path, vx, lv), sum);
                }
                result = mul_add(path.throughput, sum, result);
            }
            if (!isDeltaBsdf && settings.useVertexMerging && mIteration > 0)
            {
                const V4 merged = MergeVertices(path, vx);
                result = mul_add(path.throughput * merged, v4s(mMergingNormalization), result);
            }
            if (path.length > settings.maxPathLength) break;
            const float s0 = sampler3(), s1 = sampler3(), s2 = sampler3();
            if (!AdvancePath(path, vx, s0, s1, s2)) break;
        }
        return result;
    }

    size_t NumPhotons() const { return mPhotons.size(); }

private:
    static V4 mul_add(const V4& a, const V4& b, const V4& c) { return v4(fma_(a.x, b.x, c.x), fma_(a.y, b.y, c.y), fma_(a.z, b.z, c.z), fma_(a.w, b.w, c.w)); }
    static float CosTheta(const Vertex& vx, const V3& dir) { return dot3(vx.normal, dir); }

    // Material::Evaluate with both pdfs
    V4 Evaluate(const Vertex& vx, const V3& incomingDirWorld, float& pdfW, float& revPdfW) const
    {
        const V3 out = world_to_local(vx, vx.outgoingDirWorld), in = world_to_local(vx, incomingDirWorld);
        pdfW = 0.0f;
        const V4 value = bsdf_evaluate(vx, out, in, pdfW);
        revPdfW = bsdf_pdf(vx, out, in, true);
        return value;
    }

    uint64_t CellKey(const V3& p, int dx, int dy, int dz) const
    {
        const float inv = 1.0f / mRadiusVM;
        const int64_t x = (int64_t)floorf(p.x * inv) + dx, y = (int64_t)floorf(p.y * inv) + dy, z = (int64_t)floorf(p.z * inv) + dz;
        return (uint64_t)(x * 73856093LL) ^ (uint64_t)(y * 19349663LL) ^ (uint64_t)(z * 83492791LL);
    }

    // GenerateLightSample, :430-485
    bool GenerateLightSample(PathStateVcm& outPath)
    {
        if (S.numLights == 0) return false;
        const float lightPickProbability = 1.0f / (float)S.numLights;
        const DLight& light = S.lights[rng.next_uint() % S.numLights];
        const float p0 = rng.next_float(), p1 = rng.next_float(), p2 = rng.next_float(), d0 = rng.next_float(), d1 = rng.next_float();
        EmitResult emit;
        const V4 throughput = light_emit<true>(S, light, p0, p1, p2, d0, d1, emit);
        if (almost_zero(throughput)) return false;
        emit.directPdfA *= lightPickProbability;
        emit.emissionPdfW *= lightPickProbability;
        const float emissionInvPdfW = 1.0f / emit.emissionPdfW;
        emit.position = emit.position + emit.direction * 0.0005f;
        outPath.ray = make_ray(emit.position, emit.direction);
        outPath.throughput = throughput * emissionInvPdfW;
        outPath.isFiniteLight = (light.flags & RTB_LIGHT_FLAG_FINITE) != 0;
        outPath.dVCM = emit.directPdfA * emissionInvPdfW;
        if ((light.flags & RTB_LIGHT_FLAG_DELTA) == 0) outPath.dVC = (outPath.isFiniteLight ? emit.cosAtLight : 1.0f) * emissionInvPdfW;
        else outPath.dVC = 0.0f;
        outPath.dVM = outPath.dVC * mMisVcWeightVC;
        return true;
    }

    // TraceLightPath, :292-428
    template <typename Splat>
    void TraceLightPath(Splat& splat)
    {
        mLightVertices.clear();
        PathStateVcm path;
        if (!GenerateLightSample(path)) return;
        for (;;)
        {
            Hit hit;
            hit.objectId = RTB_INVALID_OBJECT; hit.subObjectId = 0; hit.distance = INFINITY; hit.u = 0.0f; hit.v = 0.0f;
            scene_traverse<false>(S, path.ray, hit, nullptr);
            if (hit.distance == INFINITY || hit.subObjectId == RTB_LIGHT_OBJECT) break;
            LightVertex lv;
            evaluate_intersection<true>(S, path.ray, hit, lv.vx, RTB_NO_MATERIAL);
            lv.vx.outgoingDirWorld = neg(path.ray.dir);
            evaluate_shading_data<true>(S, lv.vx);
            {
                if (path.length > 1 || path.isFiniteLight) path.dVCM *= s_sqr(hit.distance);
                const float cosTheta = dot3(path.ray.dir, lv.vx.normal);
                const float invMis = 1.0f / s_abs(cosTheta);
                path.dVCM *= invMis; path.dVC *= invMis; path.dVM *= invMis;
            }
            const bool isDeltaBsdf = lv.vx.bsdf == RTB_BSDF_METAL || lv.vx.bsdf == RTB_BSDF_DIELECTRIC;
            if (!isDeltaBsdf)
            {
                if (settings.useVertexConnection && mLightVertices.size() < 256)
                {
                    lv.pathLength = path.length; lv.throughput = path.throughput; lv.dVC = path.dVC; lv.dVM = path.dVM; lv.dVCM = path.dVCM;
                    mLightVertices.push_back(lv);
                    ConnectToCamera(lv, splat);
                }
                if (settings.useVertexMerging)
                {
                    Photon ph;
                    ph.position = lv.vx.position; ph.direction = lv.vx.outgoingDirWorld; ph.throughput = xyz(path.throughput);
                    ph.dVM = path.dVM; ph.dVCM = path.dVCM;
                    mNewPhotons.push_back(ph);
                }
            }
            if (path.length + 2 > settings.maxPathLength) break;
            const float s0 = rng.next_float(), s1 = rng.next_float(), s2 = rng.next_float();
            if (!AdvancePath(path, lv.vx, s0, s1, s2)) break;
        }
    }

    // AdvancePath, :487-575
    bool AdvancePath(PathStateVcm& path, const Vertex& vx, float s0, float s1, float s2)
    {
        BsdfSample bs;
        const V3 outLocal = world_to_local(vx, vx.outgoingDirWorld);
        if (!bsdf_sample(vx, outLocal, s0, s1, s2, bs)) return false;
        const V3 incomingDirWorld = local_to_world(vx, bs.incomingDir);
        const float cosThetaOut = s_abs(dot3(incomingDirWorld, vx.normal));
        path.throughput = path.throughput * bs.color;
        if (almost_zero(path.throughput)) return false;
        path.ray = make_ray(vx.position, incomingDirWorld);
        path.ray.origin = path.ray.origin + path.ray.dir * 0.001f;
        path.length++;
        if (bs.event & EV_SPECULAR)
        {
            path.dVC *= cosThetaOut; path.dVM *= cosThetaOut; path.dVCM = 0.0f;
            path.lastSpecular = true;
        }
        else
        {
            const float bsdfRevPdf = bsdf_pdf(vx, outLocal, neg(world_to_local(vx, incomingDirWorld)), true);
            const float invBsdfDirPdf = 1.0f / bs.pdf;
            path.dVC = (cosThetaOut * invBsdfDirPdf) * (path.dVC * bsdfRevPdf + path.dVCM + mMisVmWeightVC);
            path.dVM = (cosThetaOut * invBsdfDirPdf) * (path.dVM * bsdfRevPdf + path.dVCM * mMisVcWeightVC + 1.0f);
            path.dVCM = invBsdfDirPdf;
            path.lastSpecular = false;
        }
        return true;
    }

    // EvaluateLight, :577-626 (`hitVertex` null: a global light seen by a ray that left the scene)
    V4 EvaluateLight(const DLight& light, const Vertex* hitVertex, const PathStateVcm& path) const
    {
        const Ray lightSpaceRay = transform_ray_unsafe(light.invTransform, path.ray);
        const float cosAtLight = hitVertex ? -CosTheta(*hitVertex, path.ray.dir) : 1.0f;
        float directPdfA = 0.0f, emissionPdfW = 0.0f;
        V4 contribution = light_radiance_bidir<true>(S, light, lightSpaceRay.dir, cosAtLight, directPdfA, emissionPdfW);
        if (almost_zero(contribution)) return v4s(0.0f);
        if (path.length > 1)
        {
            const bool useVertexMerging = settings.useVertexMerging && mIteration > 0;
            if (useVertexMerging && !settings.useVertexConnection)
            {
                if (!path.lastSpecular) return v4s(0.0f);
            }
            else
            {
                const float wCamera = directPdfA * path.dVCM + emissionPdfW * path.dVC;
                contribution = contribution * (1.0f / (1.0f + wCamera));
            }
        }
        return contribution;
    }

    // SampleLight, :628-716
    template <typename Sampler3>
    V4 SampleLight(const DLight& light, const Vertex& vx, const PathStateVcm& path, Sampler3& sampler3) const
    {
        const float u0 = sampler3(), u1 = sampler3(), u2 = sampler3();
        IlluminateResultBidir ir;
        const V4 radiance = light_illuminate_bidir<true>(S, light, vx, u0, u1, u2, ir);
        if (almost_zero(radiance)) return v4s(0.0f);
        float bsdfPdfW, bsdfRevPdfW;
        const V4 bsdfFactor = Evaluate(vx, neg(ir.directionToLight), bsdfPdfW, bsdfRevPdfW);
        if (almost_zero(bsdfFactor)) return v4s(0.0f);
        Ray shadowRay = make_ray(vx.position, ir.directionToLight);
        shadowRay.origin = shadowRay.origin + shadowRay.dir * 0.0001f;
        if (scene_traverse_shadow<false>(S, shadowRay, ir.distance * 0.999f, nullptr)) return v4s(0.0f);
        const bool isDeltaLight = (light.flags & RTB_LIGHT_FLAG_DELTA) != 0;
        bsdfPdfW *= isDeltaLight ? 0.0f : 1.0f;
        const float cosToLight = dot3(vx.normal, ir.directionToLight);
        if (cosToLight <= FLT_EPSILON) return v4s(0.0f);
        const float wLight = bsdfPdfW / ir.directPdfW;
        const float wCamera = (ir.emissionPdfW * cosToLight / (ir.directPdfW * ir.cosAtLight)) * (mMisVmWeightVC + path.dVCM + path.dVC * bsdfRevPdfW);
        const float misWeight = 1.0f / (wLight + 1.0f + wCamera);
        return (radiance * bsdfFactor) * (misWeight / ir.directPdfW);
    }

    // ConnectVertices, :738-806
    V4 ConnectVertices(const PathStateVcm& cameraPath, const Vertex& vx, const LightVertex& lv) const
    {
        V3 lightDir = lv.vx.position - vx.position;
        const float distanceSqr = sqr_length3(lightDir);
        const float distance = sqrtf(distanceSqr);
        lightDir = lightDir / distance;
        const float cosCameraVertex = CosTheta(vx, lightDir);
        const float cosLightVertex = CosTheta(lv.vx, neg(lightDir));
        if (cosCameraVertex <= 0.0f || cosLightVertex <= 0.0f) return v4s(0.0f);
        const float geometryTerm = 1.0f / distanceSqr;
        float cameraBsdfPdfW, cameraBsdfRevPdfW, lightBsdfPdfW, lightBsdfRevPdfW;
        const V4 cameraFactor = Evaluate(vx, neg(lightDir), cameraBsdfPdfW, cameraBsdfRevPdfW);
        if (almost_zero(cameraFactor)) return v4s(0.0f);
        const V4 lightFactor = Evaluate(lv.vx, lightDir, lightBsdfPdfW, lightBsdfRevPdfW);
        if (almost_zero(lightFactor)) return v4s(0.0f);
        Ray shadowRay = make_ray(vx.position, lightDir);
        shadowRay.origin = shadowRay.origin + shadowRay.dir * 0.0001f;
        if (scene_traverse_shadow<false>(S, shadowRay, distance * 0.999f, nullptr)) return v4s(0.0f);
        const float cameraBsdfPdfA = cameraBsdfPdfW * s_abs(cosLightVertex) / s_sqr(distance);
        const float lightBsdfPdfA = lightBsdfPdfW * s_abs(cosCameraVertex) / s_sqr(distance);
        const float wLight = cameraBsdfPdfA * (mMisVmWeightVC + lv.dVCM + lv.dVC * lightBsdfRevPdfW);
        const float wCamera = lightBsdfPdfA * (mMisVmWeightVC + cameraPath.dVCM + cameraPath.dVC * cameraBsdfRevPdfW);
        const float misWeight = 1.0f / (wLight + 1.0f + wCamera);
        return (cameraFactor * lightFactor) * (geometryTerm * misWeight);
    }

    // MergeVertices, :808-906
    V4 MergeVertices(const PathStateVcm& cameraPath, const Vertex& vx) const
    {
        V4 contribution = v4s(0.0f);
        const float radiusSqr = s_sqr(mRadiusVM);
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
                const Photon& photon = mPhotons[index];
                if (sqr_length3(vx.position - photon.position) > radiusSqr) continue;
                const float cosToLight = CosTheta(vx, photon.direction);
                if (cosToLight < FLT_EPSILON) continue;
                float cameraBsdfDirPdfW, cameraBsdfRevPdfW;
                const V4 cameraBsdfFactor = Evaluate(vx, neg(photon.direction), cameraBsdfDirPdfW, cameraBsdfRevPdfW);
                if (almost_zero(cameraBsdfFactor)) continue;
                const float wLight = photon.dVCM * mMisVcWeightVM + photon.dVM * cameraBsdfDirPdfW;
                const float wCamera = cameraPath.dVCM * mMisVcWeightVM + cameraPath.dVM * cameraBsdfRevPdfW;
                const float weight = (1.0f / (wLight + 1.0f + wCamera)) / cosToLight;
                contribution = mul_add(cameraBsdfFactor * v4(photon.throughput.x, photon.throughput.y, photon.throughput.z, 0.0f), v4s(weight), contribution);
            }
        }
        return contribution;
    }

    // ConnectToCamera, :908-970
    template <typename Splat>
    void ConnectToCamera(const LightVertex& lv, Splat& splat) const
    {
        const V3 cameraPos = xyz(camera.localToWorld.r3);
        V3 dirToCamera = cameraPos - lv.vx.position;
        const float cameraDistanceSqr = sqr_length3(dirToCamera);
        const float cameraDistance = sqrtf(cameraDistanceSqr);
        dirToCamera = dirToCamera / cameraDistance;
        float bsdfPdfW, bsdfRevPdfW;
        const V4 cameraFactor = Evaluate(lv.vx, neg(dirToCamera), bsdfPdfW, bsdfRevPdfW);
        if (almost_zero(cameraFactor)) return;
        float filmX, filmY;
        if (!camera_world_to_film(camera.worldToScreen, lv.vx.position, filmX, filmY)) return;
        Ray shadowRay = make_ray(lv.vx.position, dirToCamera);
        shadowRay.origin = shadowRay.origin + shadowRay.dir * 0.0001f;
        if (scene_traverse_shadow<false>(S, shadowRay, cameraDistance * 0.999f, nullptr)) return;
        const float cosToCamera = dot3(dirToCamera, lv.vx.normal);
        if (cosToCamera <= FLT_EPSILON) return;
        const float cameraPdfW = camera_pdf_w(camera.localToWorld, camera.tanHalfFoV, camera.aspectRatio, neg(dirToCamera));
        const float cameraPdfA = cameraPdfW * cosToCamera / cameraDistanceSqr;
        const float wLight = cameraPdfA * (mMisVmWeightVC + lv.dVCM + lv.dVC * bsdfRevPdfW);
        const float misWeight = 1.0f / (wLight + 1.0f);
        const V4 contribution = (cameraFactor * lv.throughput) * (misWeight * cameraPdfA / cosToCamera);
        splat(filmX, filmY, xyz(contribution));
    }

    const SceneView& S;
    CameraConnection camera;
    VcmSettings settings;
    Rng& rng;
    uint32_t mWidth, mHeight, mIteration = 0;
    float mLightPathsCount = 0.0f, mRadiusVC = 0.0f, mRadiusVM = 0.0f, mMergingNormalization = 0.0f;
    float mMisVmWeightVC = 0.0f, mMisVcWeightVC = 0.0f, mMisVmWeightVM = 0.0f, mMisVcWeightVM = 0.0f;
    std::vector<LightVertex> mLightVertices;
    std::vector<Photon> mPhotons, mNewPhotons;
    std::unordered_map<uint64_t, std::vector<uint32_t>> mGrid;
};

} // namespace rtb
