// rtb_integrator.h — one path vertex of the integrator loop, as a pure per-path device function.
//
// The reference runs `for(;;) { Traverse; shade; }` per pixel on a CPU thread (PathTracerMIS.cpp:270-396).  Here the
// loop is cut at the traversal calls: the wavefront kernels run closest-hit traversal for every live path, then
// shade_vertex() for every live path, then any-hit traversal for the shadow rays shade_vertex() emitted.  What
// shade_vertex() does between two traversals is, statement for statement, the body of the reference's loop:
//
//   PathTracerMIS::RenderPixel / SampleLights / SampleLight / EvaluateLight / EvaluateGlobalLights   PathTracerMIS.cpp:43-415
//   PathTracer::RenderPixel / EvaluateLight / EvaluateGlobalLights                                   PathTracer.cpp:25-172
//   DebugRenderer::RenderPixel                                                                       DebugRenderer.cpp:26-178
//
// Radiance bookkeeping goes through a Sink (see rtb_kernels.cu / tests/hostsim):
//   sink.add(thr, c)                     resultColor.MulAndAccumulate(thr, c)
//   sink.shadow(ray, tMax, contribution, light) one NEE sample whose visibility is still unknown (light order preserved)
//   sink.end_lights(thr)                 resultColor.MulAndAccumulate(thr, sum of the visible contributions)
#pragma once

#include "rtb_shading.h"

namespace rtb {

struct PathState
{
    Ray      ray;
    V4       throughput;
    Sampler  sampler;
    uint32_t pixel;         // pixel slot of this shard (the wavefront's film / pass-colour index)
    uint32_t rasterPixel;   // y * width + x: what shard-invariant decisions (LightSamplingStrategy::Single) key on
    uint32_t passSlot;      // which pass of the batch in flight (0 when passes are rendered one at a time)
    uint32_t depth;
    float    lastPdfW;
    uint32_t lastSpecular;
    // material of the vertex this path bounced from (RTB_NO_MATERIAL on the camera ray).  The reference reuses one
    // IntersectionData for the whole path (PathTracerMIS.cpp:259, 286) and a light's EvaluateIntersection does not set
    // its material (SceneObject_Light.cpp:63-74), so Scene::EvaluateIntersection (Scene.cpp:328) applies the PREVIOUS
    // vertex' normal map to the frame of a directly hit area light - which changes cosAtLight and the MIS weight.
    uint32_t lastMaterial;
};

// CombineMis / PdfAtoW, PathTracerMIS.cpp:16-29 (FastDivide is a true division in the exact-math oracle)
RTB_HD float combine_mis(const float samplePdf, const float otherPdf) { return samplePdf / (samplePdf + otherPdf); }
RTB_HD float pdf_a_to_w(const float pdfA, const float distance, const float cosThere) { return (pdfA * s_sqr(distance)) / s_abs(cosThere); }

RTB_HD void init_path(const PassParams& P, const uint16_t* blueNoise, uint32_t x, uint32_t y, uint32_t passSlot, PathState& ps)
{
    sampler_reset_pixel(ps.sampler, x, y);
    ps.sampler.seedBase = passSlot * P.numSeedDims;
    ps.passSlot = passSlot;
    const float offX = P.sampleOffsets ? P.sampleOffsets[2 * passSlot] : P.sampleOffsetX;
    const float offY = P.sampleOffsets ? P.sampleOffsets[2 * passSlot + 1] : P.sampleOffsetY;
    ps.ray = generate_primary_ray(P, x, y, offX, offY, ps.sampler, blueNoise);
    ps.throughput = v4s(1.0f);      // RayColor::One() = (1, 1, 1, 1)
    ps.pixel = y * P.width + x;
    ps.rasterPixel = y * P.width + x;
    ps.depth = 0;
    ps.lastPdfW = 1.0f;
    ps.lastSpecular = 1u;
    ps.lastMaterial = RTB_NO_MATERIAL;
}

// generate the secondary ray: Ray(position, dir); ray.origin += ray.dir * 0.001 (originDivDir is NOT refreshed)
RTB_HD void spawn_ray(PathState& ps, const V3& position, const V3& dirWorld)
{
    ps.ray = make_ray(position, dirWorld);
    ps.ray.origin = ps.ray.origin + ps.ray.dir * 0.001f;
}

// returns true if the path continues (ps holds the next ray), false if it terminated
template <bool kTex = true, typename Sink>
RTB_HD bool shade_vertex_mis(const SceneView& S, const PassParams& P, PathState& ps, const Hit& hit, Sink& sink, uint32_t& numShadowRays)
{
    const float lightPickProbability = P.lightSamplingAll ? 1.0f : 1.0f / (float)S.numLights;   // GetLightPickingProbability
    const Ray& ray = ps.ray;

    // ray missed - return background light color                         PathTracerMIS.cpp:277-282, 214-252
    if (hit.objectId == RTB_INVALID_OBJECT)
    {
        V4 result = v4s(0.0f);
        for (uint32_t i = 0; i < S.numGlobalLights; ++i)
        {
            const DLight& light = S.lights[S.globalLights[i]];
            const Ray lightSpaceRay = transform_ray_unsafe(light.invTransform, ray);
            float directPdfW = 0.0f;
            const V4 lightContribution = light_radiance<kTex>(S, light, lightSpaceRay, v3s(0.0f), 1.0f, directPdfW);
            if (!almost_zero(lightContribution))
            {
                float misWeight = 1.0f;
                if (ps.depth > 0 && !ps.lastSpecular) misWeight = combine_mis(ps.lastPdfW, directPdfW * lightPickProbability);
                result = mul_add(lightContribution, misWeight, result);
            }
        }
        sink.add(ps.throughput, result);
        return false;
    }

    Vertex vx;
    if (hit.distance < FLT_MAX) evaluate_intersection<kTex>(S, ray, hit, vx, ps.lastMaterial);

    // we hit a light directly                                            PathTracerMIS.cpp:290-302, 174-212
    if (hit.subObjectId == RTB_LIGHT_OBJECT)
    {
        const DLight& light = S.lights[S.objShade[hit.objectId].lightIndex];
        const Ray lightSpaceRay = transform_ray_unsafe(light.invTransform, ray);
        const V3 lightSpaceHitPoint = transform_point(light.invTransform, vx.position);
        const float cosAtLight = -dot3(vx.normal, ray.dir);
        float directPdfA = 0.0f;
        V4 lightContribution = light_radiance<kTex>(S, light, lightSpaceRay, lightSpaceHitPoint, cosAtLight, directPdfA);
        if (!almost_zero(lightContribution))
        {
            float misWeight = 1.0f;
            if (ps.depth > 0 && !ps.lastSpecular)
            {
                const float directPdfW = pdf_a_to_w(directPdfA, hit.distance, cosAtLight);
                misWeight = combine_mis(ps.lastPdfW, directPdfW * lightPickProbability);
            }
            lightContribution = lightContribution * misWeight;
        }
        else
        {
            lightContribution = v4s(0.0f);
        }
        sink.add(ps.throughput, lightContribution);
        return false;
    }

    vx.outgoingDirWorld = neg(ray.dir);
    evaluate_shading_data<kTex>(S, vx);
    const V3 outLocal = world_to_local(vx, vx.outgoingDirWorld);

    // accumulate emission color                                          PathTracerMIS.cpp:309-317
    sink.add(ps.throughput, vx.emission);

    // sample lights directly (next event estimation)                     PathTracerMIS.cpp:320, 43-155
    if (S.numLights > 0)
    {
        uint32_t first = 0, last = S.numLights;
        if (!P.lightSamplingAll)
        {
            // the reference picks with a per-thread RNG (PathTracerMIS.cpp:136); here: a counter-based hash of the
            // raster pixel, pass and depth: reproducible, independent of how the frame is sharded, but only statistically
            // equivalent to the reference
            first = (uint32_t)(hash64(((uint64_t)ps.rasterPixel << 32) ^ ((uint64_t)(P.passIndex + ps.passSlot) << 8) ^ ps.depth ^ 0x9E3779B97F4A7C15ULL) % S.numLights);
            last = first + 1;
        }
        const bool isLastPathSegment = ps.depth >= P.maxRayDepth;
        for (uint32_t li = first; li < last; ++li)
        {
            const DLight& light = S.lights[li];
            const float u0 = sampler_get_float(ps.sampler, P, S.blueNoise);
            const float u1 = sampler_get_float(ps.sampler, P, S.blueNoise);
            const float u2 = sampler_get_float(ps.sampler, P, S.blueNoise);
            IlluminateResult ir;
            const V4 radiance = light_illuminate<kTex>(S, light, vx, u0, u1, u2, ir);
            if (almost_zero(radiance)) continue;
            float bsdfPdfW = 0.0f;
            const V4 factor = bsdf_evaluate(vx, outLocal, world_to_local(vx, neg(ir.directionToLight)), bsdfPdfW);
            if (almost_zero(factor)) continue;

            // shadow ray                                                  PathTracerMIS.cpp:83-101
            Ray shadowRay = make_ray(vx.position, ir.directionToLight);
            shadowRay.origin = shadowRay.origin + shadowRay.dir * 0.0001f;
            const float tMax = ir.distance * 0.999f;

            float weight = 1.0f;
            if (!(light.flags & RTB_LIGHT_FLAG_DELTA) && !isLastPathSegment)
            {
                bsdfPdfW *= 1.0f;
                weight = combine_mis(ir.directPdfW * lightPickProbability, bsdfPdfW);
            }
            const V4 contribution = (radiance * factor) * (weight / (lightPickProbability * ir.directPdfW));
            sink.shadow(shadowRay, tMax, contribution, li);
            numShadowRays++;
        }
        sink.end_lights(ps.throughput);
    }

    // check if the ray depth won't be exceeded in the next iteration     PathTracerMIS.cpp:323-327
    if (ps.depth >= P.maxRayDepth) return false;

    // Russian roulette                                                   PathTracerMIS.cpp:330-347
    if (ps.depth >= P.minRussianRouletteDepth)
    {
        const float minColorValue = 0.125f;
        const float threshold = minColorValue + (1.0f - minColorValue) * color_max(vx.baseColor);
        if (sampler_get_float(ps.sampler, P, S.blueNoise) > threshold) return false;
        ps.throughput = ps.throughput * (1.0f / threshold);
    }

    // sample BSDF                                                        PathTracerMIS.cpp:350-373
    const float s0 = sampler_get_float(ps.sampler, P, S.blueNoise);
    const float s1 = sampler_get_float(ps.sampler, P, S.blueNoise);
    const float s2 = sampler_get_float(ps.sampler, P, S.blueNoise);
    BsdfSample bs;
    if (!bsdf_sample(vx, outLocal, s0, s1, s2, bs)) return false;      // NullEvent
    ps.throughput = ps.throughput * bs.color;
    if (almost_zero(ps.throughput)) return false;
    ps.lastSpecular = (bs.event & EV_SPECULAR) != 0 ? 1u : 0u;
    ps.lastPdfW = bs.pdf;

    // generate secondary ray                                             PathTracerMIS.cpp:392-395
    spawn_ray(ps, vx.position, local_to_world(vx, bs.incomingDir));
    ps.depth++;
    ps.lastMaterial = vx.materialIndex;
    return true;
}

// PathTracer::RenderPixel loop body (BSDF sampling only), PathTracer.cpp:74-172
template <bool kTex = true, typename Sink>
RTB_HD bool shade_vertex_pt(const SceneView& S, const PassParams& P, PathState& ps, const Hit& hit, Sink& sink)
{
    const Ray& ray = ps.ray;
    // miss: hitPoint.distance == DefaultDistance                         PathTracer.cpp:88-93, 48-72
    if (hit.distance == INFINITY)
    {
        V4 result = v4s(0.0f);
        for (uint32_t i = 0; i < S.numGlobalLights; ++i)
        {
            const DLight& light = S.lights[S.globalLights[i]];
            const Ray lightSpaceRay = transform_ray_unsafe(light.invTransform, ray);
            float pdf = 0.0f;
            result = result + light_radiance<kTex>(S, light, lightSpaceRay, v3s(0.0f), 1.0f, pdf);
        }
        sink.add(ps.throughput, result);
        return false;
    }
    Vertex vx;
    evaluate_intersection<kTex>(S, ray, hit, vx, ps.lastMaterial);
    vx.outgoingDirWorld = neg(ray.dir);
    if (hit.subObjectId == RTB_LIGHT_OBJECT)
    {
        // PathTracer::EvaluateLight, PathTracer.cpp:25-46
        const DLight& light = S.lights[S.objShade[hit.objectId].lightIndex];
        const float cosAtLight = -dot3(vx.normal, ray.dir);
        const Ray lightSpaceRay = transform_ray_unsafe(light.invTransform, ray);
        const V3 lightSpaceHitPoint = transform_point(light.invTransform, vx.position);
        float pdf = 0.0f;
        sink.add(ps.throughput, light_radiance<kTex>(S, light, lightSpaceRay, lightSpaceHitPoint, cosAtLight, pdf));
        return false;
    }
    evaluate_shading_data<kTex>(S, vx);
    sink.add(ps.throughput, vx.emission);
    if (ps.depth >= P.maxRayDepth) return false;
    if (ps.depth >= P.minRussianRouletteDepth)
    {
        const float minColorValue = 0.125f;
        const float threshold = minColorValue + (1.0f - minColorValue) * color_max(vx.baseColor);
        if (sampler_get_float(ps.sampler, P, S.blueNoise) > threshold) return false;
        ps.throughput = ps.throughput * (1.0f / threshold);
    }
    const float s0 = sampler_get_float(ps.sampler, P, S.blueNoise);
    const float s1 = sampler_get_float(ps.sampler, P, S.blueNoise);
    const float s2 = sampler_get_float(ps.sampler, P, S.blueNoise);
    BsdfSample bs;
    const V3 outLocal = world_to_local(vx, vx.outgoingDirWorld);
    if (!bsdf_sample(vx, outLocal, s0, s1, s2, bs)) return false;      // Material::Sample returns Zero -> throughput AlmostZero
    ps.throughput = ps.throughput * bs.color;
    if (almost_zero(ps.throughput)) return false;
    spawn_ray(ps, vx.position, local_to_world(vx, bs.incomingDir));
    ps.depth++;
    ps.lastMaterial = vx.materialIndex;
    return true;
}

// HSVtoRGB, Core/Color/ColorHelpers.h:134-156
RTB_HD V4 hsv_to_rgb(const float hue, const float saturation, const float value)
{
    const int h_i = (int)(hue * 6.0f);
    const float f = hue * 6 - h_i;
    const float p = value * (1 - saturation);
    const float q = value * (1 - f * saturation);
    const float t = value * (1 - (1 - f) * saturation);
    if (h_i == 0) return v4(value, t, p, 0.0f);
    else if (h_i == 1) return v4(q, value, p, 0.0f);
    else if (h_i == 2) return v4(p, value, t, 0.0f);
    else if (h_i == 3) return v4(p, q, value, 0.0f);
    else if (h_i == 4) return v4(t, p, value, 0.0f);
    else if (h_i == 5) return v4(value, p, q, 0.0f);
    return v4s(0.0f);
}

// DebugRenderer::RenderPixel, DebugRenderer.cpp:26-178 (primary hit visualisation; never continues)
template <bool kTex = true, typename Sink>
RTB_HD void shade_debug(const SceneView& S, const PassParams& P, PathState& ps, const Hit& hit, Sink& sink)
{
    const V4 one = v4s(1.0f);
    if (hit.distance == INFINITY) return;                                               // background: zero
    if (hit.subObjectId == RTB_LIGHT_OBJECT) { sink.add(one, v4(1.0f, 1.0f, 0.0f, 0.0f)); return; }
    Vertex vx;
    vx.outgoingDirWorld = neg(ps.ray.dir);
    if (P.debugMode != RTB_DEBUG_TRIANGLE_ID && P.debugMode != RTB_DEBUG_DEPTH)
    {
        if (hit.distance < FLT_MAX) evaluate_intersection<kTex>(S, ps.ray, hit, vx, RTB_NO_MATERIAL);
        evaluate_shading_data<kTex>(S, vx);
    }
    V4 c = v4s(0.0f);
    switch (P.debugMode)
    {
    case RTB_DEBUG_CAMERA_LIGHT:
    {
        const float NdotL = dot3(ps.ray.dir, vx.normal);
        c = vx.baseColor * s_abs(NdotL);
        break;
    }
    case RTB_DEBUG_DEPTH:
    {
        const float invDepth = 1.0f - 1.0f / (1.0f + fdiv(hit.distance, 10.0f));
        c = v4s(invDepth);
        break;
    }
    case RTB_DEBUG_TRIANGLE_ID:
    {
        const uint64_t h = hash64((uint64_t)hit.objectId | ((uint64_t)hit.subObjectId << 32));
        const float hue = (float)(uint32_t)h / (float)UINT32_MAX;
        const float saturation = 0.5f + 0.5f * (float)(uint32_t)(h >> 32) / (float)UINT32_MAX;
        c = hsv_to_rgb(hue, saturation, 1.0f);
        break;
    }
    // Saturate(BipolarToUnipolar(v)) = min(1, max(0, fmadd(v, 0.5, 0.5)))
    case RTB_DEBUG_TANGENTS:   c = v4(sse_min(1.0f, sse_max(0.0f, fma_(vx.tangent.x, 0.5f, 0.5f))), sse_min(1.0f, sse_max(0.0f, fma_(vx.tangent.y, 0.5f, 0.5f))), sse_min(1.0f, sse_max(0.0f, fma_(vx.tangent.z, 0.5f, 0.5f))), 0.0f); break;
    case RTB_DEBUG_BITANGENTS: c = v4(sse_min(1.0f, sse_max(0.0f, fma_(vx.bitangent.x, 0.5f, 0.5f))), sse_min(1.0f, sse_max(0.0f, fma_(vx.bitangent.y, 0.5f, 0.5f))), sse_min(1.0f, sse_max(0.0f, fma_(vx.bitangent.z, 0.5f, 0.5f))), 0.0f); break;
    case RTB_DEBUG_NORMALS:    c = v4(sse_min(1.0f, sse_max(0.0f, fma_(vx.normal.x, 0.5f, 0.5f))), sse_min(1.0f, sse_max(0.0f, fma_(vx.normal.y, 0.5f, 0.5f))), sse_min(1.0f, sse_max(0.0f, fma_(vx.normal.z, 0.5f, 0.5f))), 0.0f); break;
    case RTB_DEBUG_POSITION:   c = v4(sse_max(0.0f, vx.position.x), sse_max(0.0f, vx.position.y), sse_max(0.0f, vx.position.z), 0.0f); break;
    case RTB_DEBUG_TEXCOORDS:  c = v4(vx.texU - floorf(vx.texU), vx.texV - floorf(vx.texV), 0.0f, 0.0f); break;
    case RTB_DEBUG_BASE_COLOR: c = vx.baseColor; break;
    case RTB_DEBUG_EMISSION:   c = vx.emission; break;
    case RTB_DEBUG_ROUGHNESS:  c = v4s(vx.roughness); break;
    case RTB_DEBUG_METALNESS:  c = v4s(vx.metalness); break;
    case RTB_DEBUG_IOR:        c = v4s(vx.IoR); break;
    default: break;
    }
    sink.add(one, c);
}

} // namespace rtb
