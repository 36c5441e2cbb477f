// rtb_bidir.h — the per-vertex arithmetic of the reference's bidirectional renderers, written once for host and device:
// ILight::Emit for the five light types, ILight::Illuminate / GetRadiance as VCM calls them, BSDF::Pdf, the camera
// connection (Camera::WorldToFilm / PdfW), LightTracer::RenderPixel, and VertexConnectionAndMerging's building blocks
// (bidir_*: pass weights, light sample generation, the dVCM / dVC / dVM updates, AdvancePath, EvaluateLight, SampleLight,
// ConnectVertices, the per-photon merge, ConnectToCamera).  (Rendering/LightTracer.cpp:27-183,
// Rendering/VertexConnectionAndMerging.cpp:80-125, 292-970.)
//
// Used by the VCM kernels of rtb_vcm.cuh (the product, included from rtb_kernels.cu) and by the CPU model of the same
// renderer (tests/hostsim/hostsim_vcm.h), which pinned every function here against the reference's own
// ILight / Camera / BSDF bit for bit and the assembled renderer statistically (tests/test_hostsim_parity.py).
#pragma once

#include "rtb_shading.h"
#include "rtb_trace.h"

namespace rtb {

struct EmitResult
{
    V3    position;        // world space
    V3    direction;       // world space (the spot light's stays in light space, as in the reference)
    float directPdfA;
    float emissionPdfW;
    float cosAtLight;
};

#define RTB_SCENE_RADIUS 30.0f       // BackgroundLight.cpp:16, DirectionalLight.cpp ("TODO" constants of the reference)

// IShape::Sample(u, &normal): a point on the surface, uniform by area (SphereShape.cpp:48-63, BoxShape.cpp:127-178,
// RectShape.cpp:51-64).  Box / rect: same arithmetic as the area-sampling half of shape_sample (rtb_shading.h), kept
// separate so that the path-tracing kernels' code is untouched.
RTB_HD void shape_sample_point(const RtbShape& s, float u0, float u1, float u2, V3& position, V3& normal)
{
    if (s.type == RTB_SHAPE_SPHERE)
    {
        const V3 point = sample_sphere(u0, u1);
        normal = point;
        position = point * s.size[0];
    }
    else if (s.type == RTB_SHAPE_BOX)
    {
        float v = u2;
        uint32_t zAxis;
        v *= s.aux[2];
        if (v < s.aux[0]) { v /= s.aux[0]; zAxis = 0; }
        else if (v < s.aux[1]) { v = (v - s.aux[0]) / (s.aux[1] - s.aux[0]); zAxis = 1; }
        else { v = (v - s.aux[1]) / (s.aux[2] - s.aux[1]); zAxis = 2; }
        const uint32_t xAxis = (zAxis + 1) % 3u, yAxis = (zAxis + 2) % 3u;
        float n[3] = { 0.0f, 0.0f, 0.0f }, p[3] = { 0.0f, 0.0f, 0.0f };
        n[zAxis] = v < 0.5f ? -1.0f : 1.0f;
        p[xAxis] = (2.0f * u0 - 1.0f) * s.size[xAxis];
        p[yAxis] = (2.0f * u1 - 1.0f) * s.size[yAxis];
        p[zAxis] = n[zAxis] * s.size[zAxis];
        position = v3(p[0], p[1], p[2]); normal = v3(n[0], n[1], n[2]);
    }
    else
    {
        position = v3(s.size[0] * (2.0f * u0 - 1.0f), s.size[1] * (2.0f * u1 - 1.0f), 0.0f * (2.0f * 0.0f - 1.0f));
        normal = v3(0.0f, 0.0f, 1.0f);
    }
}

// ILight::Emit: positionSample = (p0, p1, p2), directionSample = (d0, d1); returns the emitted radiance (already
// multiplied by cosAtLight for area lights, as the reference does)
template <bool kTex = true>
RTB_HD V4 light_emit(const SceneView& S, const DLight& light, float p0, float p1, float p2, float d0, float d1, EmitResult& out)
{
    switch (light.type)
    {
    case RTB_LIGHT_AREA:
    {
        // AreaLight::Emit, AreaLight.cpp:149-182
        V3 positionLocal, normalLocal;
        shape_sample_point(S.shapes[light.shapeIndex], p0, p1, p2, positionLocal, normalLocal);
        out.position = transform_point(light.transform, positionLocal);
        V3 tangent, bitangent;
        build_orthonormal_basis(normalLocal, tangent, bitangent);
        const V3 randomDir = hemisphere_cos(d0, d1);
        const V3 dirLocal = (tangent * randomDir.x + bitangent * randomDir.y) + normalLocal * randomDir.z;
        out.direction = transform_vector(light.transform, dirLocal);
        const float cosAtLight = randomDir.z;
        const float invArea = 1.0f / shape_surface_area(S.shapes[light.shapeIndex]);
        out.cosAtLight = cosAtLight;
        out.directPdfA = invArea;
        out.emissionPdfW = invArea * cosAtLight * RTB_INV_PI;
        return light.color * cosAtLight;
    }
    case RTB_LIGHT_POINT:
        // PointLight::Emit, PointLight.cpp:51-62
        out.position = xyz(light.transform.r3);
        out.direction = sample_sphere(d0, d1);
        out.emissionPdfW = RTB_INV_PI / 4.0f;
        out.directPdfA = 1.0f;
        out.cosAtLight = 1.0f;
        return light.color;
    case RTB_LIGHT_SPOT:
    {
        // SpotLight::Emit, SpotLight.cpp:63-93 (the direction is NOT taken to world space there)
        if (light.isDelta)
        {
            out.emissionPdfW = 1.0f;
            out.direction = v3(0.0f, 0.0f, 1.0f);
        }
        else
        {
            const float phi = RTB_2PI * d1;
            float sinPhi, cosPhi;
            sin_cos(phi, sinPhi, cosPhi);
            const float cosTheta = s_lerp(light.cosAngle, 1.0f, d0);
            const float sinThetaSqr = 1.0f - s_sqr(cosTheta);
            const float sinTheta = sqrtf(sinThetaSqr);
            out.direction = normalize3(v3(sinTheta * sinPhi, sinTheta * cosPhi, cosTheta));
            out.emissionPdfW = sphere_cap_pdf(light.cosAngle);
        }
        out.position = xyz(light.transform.r3);
        out.directPdfA = 1.0f;
        out.cosAtLight = 1.0f;
        return light.color;
    }
    case RTB_LIGHT_DIRECTIONAL:
    {
        // DirectionalLight::Emit, DirectionalLight.cpp:120-134 (+ SampleDirection 48-73); the origin is not transformed
        V3 d;
        if (light.isDelta)
        {
            out.directPdfA = 1.0f;
            d = v3(0.0f, 0.0f, 1.0f);
        }
        else
        {
            out.directPdfA = sphere_cap_pdf(light.cosAngle);
            const float phi = RTB_2PI * d1;
            float sinPhi, cosPhi;
            sin_cos(phi, sinPhi, cosPhi);
            const float cosTheta = s_lerp(light.cosAngle, 1.0f, d0);
            const float sinThetaSqr = 1.0f - s_sqr(cosTheta);
            const float sinTheta = sqrtf(sinThetaSqr);
            d = normalize3(v3(sinTheta * sinPhi, sinTheta * cosPhi, cosTheta));
        }
        out.direction = transform_vector(light.transform, v3(-d.x, -d.y, -d.z));      // TransformVector(-d): a sign flip, then the plain transform
        float cx, cy;
        sample_circle(p0, p1, cx, cy);
        out.position = v3(cx * RTB_SCENE_RADIUS, cy * RTB_SCENE_RADIUS, -1.0f * RTB_SCENE_RADIUS);
        out.cosAtLight = 1.0f;
        out.emissionPdfW = out.directPdfA * (1.0f / (RTB_PI * s_sqr(RTB_SCENE_RADIUS)));
        return light.color;
    }
    default:
    {
        // BackgroundLight::Emit, BackgroundLight.cpp:92-113
        out.direction = sample_sphere(d0, d1);
        float cx, cy;
        sample_circle(p0, p1, cx, cy);
        V3 u, v;
        build_orthonormal_basis(out.direction, u, v);
        out.position = ((u * cx + v * cy) - out.direction) * RTB_SCENE_RADIUS;
        out.directPdfA = uniform_hemisphere_pdf();
        out.emissionPdfW = (RTB_INV_PI / 4.0f) * (1.0f / (RTB_PI * s_sqr(RTB_SCENE_RADIUS)));
        out.cosAtLight = 1.0f;
        return background_color<kTex>(S, light, v3(-out.direction.x, -out.direction.y, -out.direction.z));
    }
    }
}

// ---- ILight::Illuminate as the bidirectional renderers call it --------------------------------------------------------------
// VCM passes rendererSupportsSolidAngleSampling = false (VertexConnectionAndMerging.cpp:649) and reads emissionPdfW as
// well: area lights then take the area-sampling branch of AreaLight::Illuminate (AreaLight.cpp:78-102), the other light
// types behave as for the path tracers (light_illuminate, rtb_shading.h) and add their emission pdf.
struct IlluminateResultBidir { V3 directionToLight; float distance, directPdfW, emissionPdfW, cosAtLight; };

template <bool kTex = true>
RTB_HD V4 light_illuminate_bidir(const SceneView& S, const DLight& light, const Vertex& vx, float u0, float u1, float u2, IlluminateResultBidir& out)
{
    if (light.type == RTB_LIGHT_AREA)
    {
        out.directionToLight = v3s(0.0f); out.distance = -1.0f; out.directPdfW = -1.0f; out.emissionPdfW = -1.0f; out.cosAtLight = -1.0f;
        V3 positionLocal, normalLocal;
        shape_sample_point(S.shapes[light.shapeIndex], u0, u1, u2, positionLocal, normalLocal);
        const V3 lightPointWorld = transform_point(light.transform, positionLocal);
        const V3 normalWorld = transform_point(light.transform, normalLocal);          // TransformPoint on the normal, as in the reference
        out.directionToLight = lightPointWorld - vx.position;
        const float sqrDistance = sqr_length3(out.directionToLight);
        out.distance = sqrtf(sqrDistance);
        out.directionToLight = out.directionToLight / out.distance;
        const float cosNormalDir = dot3(v3(-normalWorld.x, -normalWorld.y, -normalWorld.z), out.directionToLight);
        if (cosNormalDir < RTB_EPSILON) return v4s(0.0f);
        const float invArea = 1.0f / shape_surface_area(S.shapes[light.shapeIndex]);
        out.cosAtLight = cosNormalDir;
        out.directPdfW = invArea * sqrDistance / cosNormalDir;
        out.emissionPdfW = cosNormalDir * invArea * RTB_INV_PI;
        return light.color;
    }
    IlluminateResult r;
    r.directionToLight = v3s(0.0f); r.distance = -1.0f; r.directPdfW = -1.0f; r.cosAtLight = -1.0f;
    const V4 color = light_illuminate<kTex>(S, light, vx, u0, u1, u2, r);
    out.directionToLight = r.directionToLight; out.distance = r.distance; out.directPdfW = r.directPdfW; out.cosAtLight = r.cosAtLight;
    const float circlePdf = 1.0f / (RTB_PI * s_sqr(RTB_SCENE_RADIUS));
    switch (light.type)
    {
    case RTB_LIGHT_POINT: out.emissionPdfW = RTB_INV_PI / 4.0f; break;                                   // PointLight.cpp:34-49
    case RTB_LIGHT_SPOT: out.emissionPdfW = light.isDelta ? 1.0f : sphere_cap_pdf(light.cosAngle); break;  // SpotLight.cpp:41-61
    case RTB_LIGHT_DIRECTIONAL: out.emissionPdfW = r.directPdfW * circlePdf; break;                        // DirectionalLight.cpp:80-91
    default: out.emissionPdfW = (RTB_INV_PI / 4.0f) * circlePdf; break;                                    // BackgroundLight.cpp:63-76
    }
    return color;
}

// ILight::GetRadiance with rendererSupportsSolidAngleSampling = false and both pdf outputs (VCM's EvaluateLight,
// VertexConnectionAndMerging.cpp:576-626): `lightRayDir` is the direction of the ray in light space.  Point / spot lights
// cannot be hit.  Returns the radiance; the pdfs keep the caller's values where the reference does not write them.
template <bool kTex = true>
RTB_HD V4 light_radiance_bidir(const SceneView& S, const DLight& light, const V3& lightRayDir, float cosAtLight, float& directPdfA, float& emissionPdfW)
{
    const float circlePdf = 1.0f / (RTB_PI * s_sqr(RTB_SCENE_RADIUS));
    switch (light.type)
    {
    case RTB_LIGHT_AREA:
    {
        // AreaLight::GetRadiance, AreaLight.cpp:107-147
        if (cosAtLight < RTB_EPSILON) return v4s(0.0f);
        const float invArea = 1.0f / shape_surface_area(S.shapes[light.shapeIndex]);
        directPdfA = invArea;
        emissionPdfW = cosAtLight * invArea * RTB_INV_PI;
        return light.color;
    }
    case RTB_LIGHT_BACKGROUND:
        // BackgroundLight::GetRadiance, BackgroundLight.cpp:78-90
        directPdfA = uniform_hemisphere_pdf();
        emissionPdfW = (RTB_INV_PI / 4.0f) * circlePdf;
        return background_color<kTex>(S, light, lightRayDir);
    case RTB_LIGHT_DIRECTIONAL:
    {
        // DirectionalLight::GetRadiance, DirectionalLight.cpp:92-118
        if (light.isDelta) return v4s(0.0f);
        if (dot3(lightRayDir, v3(0.0f, 0.0f, 1.0f)) > -light.cosAngle) return v4s(0.0f);
        const float directPdf = sphere_cap_pdf(light.cosAngle);
        directPdfA = directPdf;
        emissionPdfW = directPdf * circlePdf;
        return light.color;
    }
    default:
        return v4s(0.0f);
    }
}

// ---- BSDF::Pdf(ctx, ForwardPdf | ReversePdf) for the nine BSDFs ------------------------------------------------------------
// (what VCM's vertex connection asks for, VertexConnectionAndMerging.cpp:553-563).  `out` / `in` are the local-space
// outgoing and incoming directions of BSDF::EvaluationContext (incoming points towards the surface).
RTB_HD float bsdf_pdf(const Vertex& vx, const V3& out, const V3& in, bool reverse)
{
    const float NdotV = out.z, NdotL = -in.z;
    switch (vx.bsdf)
    {
    case RTB_BSDF_DIFFUSE:            // DiffuseBSDF.cpp:56-74
    case RTB_BSDF_ROUGH_DIFFUSE:      // RoughDiffuseBSDF.cpp:75-93
        if (NdotV > RTB_COS_EPSILON && NdotL > RTB_COS_EPSILON) return (reverse ? NdotV : NdotL) * RTB_INV_PI;
        return 0.0f;
    case RTB_BSDF_ROUGH_METAL:
    {
        // RoughMetalBSDF::Pdf, RoughMetalBSDF.cpp:109-137 (the same value in both directions)
        const float roughness = vx.roughness;
        if (roughness < RTB_SPECULAR_ROUGHNESS_THRESHOLD) return 0.0f;
        const V3 m = normalize3(out - in);
        const float VdotH = dot3(m, out);
        if (NdotV < RTB_COS_EPSILON || NdotL < RTB_COS_EPSILON || VdotH < RTB_COS_EPSILON) return 0.0f;
        const Microfacet microfacet(roughness * roughness);
        return microfacet.Pdf(m) / (4.0f * VdotH);
    }
    case RTB_BSDF_ROUGH_DIELECTRIC:
    {
        // RoughDielectricBSDF::Pdf, RoughDielectricBSDF.cpp:195-253 ("TODO is this correct for reverse direction?")
        if (s_abs(NdotV) < RTB_COS_EPSILON || s_abs(NdotL) < RTB_COS_EPSILON) return 0.0f;
        const float roughness = vx.roughness;
        if (roughness < RTB_SPECULAR_ROUGHNESS_THRESHOLD) return 0.0f;
        const float ior = vx.IoR;
        const float eta = NdotV < 0.0f ? ior : 1.0f / ior;
        const bool reflection = NdotV * NdotL >= 0.0f;
        V3 m;
        if (reflection) m = out - in;
        else m = eta * out - in;
        const float sg = m.z > 0.0f ? 1.0f : (m.z < 0.0f ? -1.0f : 0.0f);
        m = m * sg;
        m = normalize3(m);
        if (s_abs(m.z) < RTB_COS_EPSILON) return 0.0f;
        const float VdotH = dot3(m, out);
        const float LdotH = dot3(m, neg(in));
        const Microfacet microfacet(roughness * roughness);
        const float F = fresnel_dielectric(VdotH, ior);
        if (reflection) return F * microfacet.Pdf(m) / (4.0f * s_abs(VdotH));
        const float denom = s_sqr(eta * VdotH + LdotH);
        return (1.0f - F) * microfacet.Pdf(m) * s_abs(LdotH) / denom;
    }
    case RTB_BSDF_PLASTIC:
    case RTB_BSDF_ROUGH_PLASTIC:
    {
        // PlasticBSDF::Pdf, PlasticBSDF.cpp:101-127; RoughPlasticBSDF::Pdf, RoughPlasticBSDF.cpp:160-212
        if (NdotV < RTB_COS_EPSILON || NdotL < RTB_COS_EPSILON) return 0.0f;
        const float Fi = fresnel_dielectric(NdotV, vx.IoR);
        const float specularWeight = Fi;
        const float diffuseWeight = (1.0f - Fi) * color_max(vx.baseColor);
        const float specularProbability = specularWeight / (specularWeight + diffuseWeight);
        const float diffuseProbability = 1.0f - specularProbability;
        if (vx.bsdf == RTB_BSDF_PLASTIC || vx.roughness < RTB_SPECULAR_ROUGHNESS_THRESHOLD)
            return (reverse ? NdotV : NdotL) * RTB_INV_PI * diffuseProbability;
        const float diffusePdf = (reverse ? NdotV : NdotL) * RTB_INV_PI;
        float specularPdf = 0.0f;
        const V3 m = normalize3(out - in);
        const float VdotH = dot3(m, out);
        if (VdotH >= RTB_COS_EPSILON)
        {
            const Microfacet microfacet(vx.roughness * vx.roughness);
            specularPdf = microfacet.Pdf(m) / (4.0f * VdotH);
        }
        return diffusePdf * diffuseProbability + specularPdf * specularProbability;
    }
    default:
        return 0.0f;      // Metal, Dielectric (Dirac deltas) and Null: MetalBSDF.cpp:56-63, DielectricBSDF.cpp:124-131, NullBSDF.cpp:27-33
    }
}

// ---- camera connection -------------------------------------------------------------------------------------------------

RTB_HD V4 row_times_matrix(const V4& a, const M4& b)          // one row of Matrix4::operator*, Matrix4.cpp:35-75: a mul, then three FMAs
{
    V4 c = v4(a.x * b.r0.x, a.x * b.r0.y, a.x * b.r0.z, a.x * b.r0.w);
    c = v4(fma_(a.y, b.r1.x, c.x), fma_(a.y, b.r1.y, c.y), fma_(a.y, b.r1.z, c.z), fma_(a.y, b.r1.w, c.w));
    c = v4(fma_(a.z, b.r2.x, c.x), fma_(a.z, b.r2.y, c.y), fma_(a.z, b.r2.z, c.z), fma_(a.z, b.r2.w, c.w));
    c = v4(fma_(a.w, b.r3.x, c.x), fma_(a.w, b.r3.y, c.y), fma_(a.w, b.r3.z, c.z), fma_(a.w, b.r3.w, c.w));
    return c;
}

// Camera::mWorldToScreen as Camera::SetPerspective builds it (Camera.cpp:30-47): FastInverseNoScale() * MakePerspective.
// All four lanes are carried the way the reference's SSE code carries them: Vector4::Transpose3 (Vector4ImplSSE.h:578-585)
// leaves lanes of the third row in the w lanes of the transposed rows, and the inverse that is multiplied is the one
// WITHOUT the w clean-up of Camera.cpp:38-42 (that clean-up is applied to a copy nobody uses) - both only reach the
// z lane of the product, which Camera::WorldToFilm uses for its "in front of the camera" test.
RTB_HD M4 camera_world_to_screen(const M4& localToWorld, float aspectRatio, float tanHalfFoV)
{
    const V4 a = localToWorld.r0, b = localToWorld.r1, c = localToWorld.r2, t = localToWorld.r3;
    M4 inv;
    inv.r0 = v4(a.x, b.x, c.x, c.y);
    inv.r1 = v4(a.y, b.y, c.y, c.w);
    inv.r2 = v4(a.z, b.z, c.z, c.w);
    // rows[3] = TransformVectorNeg(original row 3): x * r0, then NegMulAndSub(y, r1, .), then NegMulAndAdd(z, r2, .)
    V4 r = v4(t.x * inv.r0.x, t.x * inv.r0.y, t.x * inv.r0.z, t.x * inv.r0.w);
    r = v4(fma_(-t.y, inv.r1.x, -r.x), fma_(-t.y, inv.r1.y, -r.y), fma_(-t.y, inv.r1.z, -r.z), fma_(-t.y, inv.r1.w, -r.w));
    r = v4(fma_(-t.z, inv.r2.x, r.x), fma_(-t.z, inv.r2.y, r.y), fma_(-t.z, inv.r2.z, r.z), fma_(-t.z, inv.r2.w, r.w));
    inv.r3 = r;
    // Matrix4::MakePerspective(aspect, fov, 0.01, 1000), Matrix4.cpp:15-24
    const float nearZ = 0.01f, farZ = 1000.0f;
    const float yScale = fdiv(1.0f, tanHalfFoV), xScale = fdiv(yScale, aspectRatio);
    M4 proj;
    proj.r0 = v4(xScale, 0.0f, 0.0f, 0.0f);
    proj.r1 = v4(0.0f, yScale, 0.0f, 0.0f);
    proj.r2 = v4(0.0f, 0.0f, fdiv(farZ, farZ - nearZ), 1.0f);
    proj.r3 = v4(0.0f, 0.0f, fdiv(-nearZ * farZ, farZ - nearZ), 0.0f);
    M4 out;
    out.r0 = row_times_matrix(inv.r0, proj); out.r1 = row_times_matrix(inv.r1, proj);
    out.r2 = row_times_matrix(inv.r2, proj); out.r3 = row_times_matrix(inv.r3, proj);
    return out;
}

// Camera::WorldToFilm, Camera.cpp:120-134: false behind the camera; film coordinates in [0, 1]^2 when inside the frame
RTB_HD bool camera_world_to_film(const M4& worldToScreen, const V3& p, float& filmX, float& filmY)
{
    const M4& m = worldToScreen;
    V4 t = v4(fma_(p.x, m.r0.x, m.r3.x), fma_(p.x, m.r0.y, m.r3.y), fma_(p.x, m.r0.z, m.r3.z), fma_(p.x, m.r0.w, m.r3.w));
    t = v4(fma_(p.y, m.r1.x, t.x), fma_(p.y, m.r1.y, t.y), fma_(p.y, m.r1.z, t.z), fma_(p.y, m.r1.w, t.w));
    t = v4(fma_(p.z, m.r2.x, t.x), fma_(p.z, m.r2.y, t.y), fma_(p.z, m.r2.z, t.z), fma_(p.z, m.r2.w, t.w));
    if (!(t.z > 0.0f)) return false;
    filmX = fma_(fdiv(t.x, t.w), 0.5f, 0.5f);          // BipolarToUnipolar(cameraSpacePosition / w)
    filmY = fma_(fdiv(t.y, t.w), 0.5f, 0.5f);
    return true;
}

// Camera::PdfW, Camera.cpp:136-148
RTB_HD float camera_pdf_w(const M4& localToWorld, float tanHalfFoV, float aspectRatio, const V3& direction)
{
    const float cosAtCamera = dot3(xyz(localToWorld.r2), direction);
    const float pdf = fdiv(0.25f, (tanHalfFoV * tanHalfFoV) * ((cosAtCamera * cosAtCamera) * cosAtCamera) * aspectRatio);
    return s_max(0.0f, pdf);
}

// ---- LightTracer::RenderPixel, Rendering/LightTracer.cpp:27-183 ---------------------------------------------------------
//
// One light sub-path: pick a light, emit, and at every surface vertex connect to the camera and splat onto the film.
// The reference draws ALL its random numbers here from the per-thread generator (ctx.randomGenerator), so its image is
// reproducible only statistically; `Rng` supplies uniform floats in [0, 1) / integers, `Splat` receives
// (filmX, filmY, rgb) for Film::AccumulateColor(pos, value, rng) (Film.cpp:41-77).  Returns the number of rays traced.
struct CameraConnection { M4 localToWorld, worldToScreen; float tanHalfFoV, aspectRatio; };

template <bool kTex = true, typename Rng, typename Splat>
RTB_HD uint32_t light_trace_path(const SceneView& S, const CameraConnection& camera, uint32_t maxRayDepth, Rng& rng, Splat& splat, SharedStack sstack = 0,
                                 unsigned long long* shadowRays = nullptr, unsigned long long* visible = nullptr)
{
    if (S.numLights == 0) return 0;
    const float lightPickingProbability = 1.0f / (float)S.numLights;
    const uint32_t lightIndex = rng.next_uint() % S.numLights;
    const DLight& light = S.lights[lightIndex];

    const float p0 = rng.next_float(), p1 = rng.next_float(), p2 = rng.next_float();
    const float d0 = rng.next_float(), d1 = rng.next_float();
    EmitResult emit;
    V4 throughput = light_emit<kTex>(S, light, p0, p1, p2, d0, d1, emit);
    if (almost_zero(throughput)) return 0;
    emit.emissionPdfW *= lightPickingProbability;
    emit.position = emit.position + emit.direction * 0.0005f;
    Ray ray = make_ray(emit.position, emit.direction);
    throughput = throughput * (1.0f / emit.emissionPdfW);

    uint32_t depth = 0;
    Vertex vx;
    for (;;)
    {
        Hit hit;
        hit.objectId = RTB_INVALID_OBJECT; hit.subObjectId = 0; hit.distance = INFINITY; hit.u = 0.0f; hit.v = 0.0f;
        scene_traverse<false>(S, ray, hit, nullptr, sstack);
        if (hit.distance == INFINITY) break;                       // ray missed
        if (hit.subObjectId == RTB_LIGHT_OBJECT) break;         // we hit a light directly
        evaluate_intersection<kTex>(S, ray, hit, vx, RTB_NO_MATERIAL);
        vx.outgoingDirWorld = neg(ray.dir);
        evaluate_shading_data<kTex>(S, vx);
        if (depth >= maxRayDepth) break;

        // connect to camera                                       LightTracer.cpp:117-151
        {
            const V3 cameraPos = xyz(camera.localToWorld.r3);
            V3 dirToCamera = cameraPos - vx.position;
            const float cameraDistanceSqr = sqr_length3(dirToCamera);
            const float cameraDistance = sqrtf(cameraDistanceSqr);
            dirToCamera = dirToCamera / cameraDistance;
            float bsdfPdfW = 0.0f;
            const V3 outLocal = world_to_local(vx, vx.outgoingDirWorld);
            const V4 cameraFactor = bsdf_evaluate(vx, outLocal, world_to_local(vx, neg(dirToCamera)), bsdfPdfW);
            if (!almost_zero(cameraFactor))
            {
                float filmX, filmY;
                if (camera_world_to_film(camera.worldToScreen, vx.position, filmX, filmY))
                {
                    const Ray shadowRay = make_ray(vx.position + vx.normal * 0.0001f, dirToCamera);
                    if (shadowRays) (*shadowRays)++;
                    if (!scene_traverse_shadow<false>(S, shadowRay, cameraDistance * 0.999f, nullptr, sstack))
                    {
                        if (visible) (*visible)++;
                        const float cameraPdfA = camera_pdf_w(camera.localToWorld, camera.tanHalfFoV, camera.aspectRatio, neg(dirToCamera)) / cameraDistanceSqr;
                        const V4 contribution = (cameraFactor * throughput) * cameraPdfA;
                        splat(filmX, filmY, xyz(contribution));
                    }
                }
            }
        }

        // sample BSDF                                             LightTracer.cpp:154-172
        const float s0 = rng.next_float(), s1 = rng.next_float(), s2 = rng.next_float();
        BsdfSample bs;
        if (!bsdf_sample(vx, world_to_local(vx, vx.outgoingDirWorld), s0, s1, s2, bs)) break;
        throughput = throughput * bs.color;
        if (almost_zero(throughput)) break;
        ray = make_ray(vx.position, local_to_world(vx, bs.incomingDir));
        ray.origin = ray.origin + ray.dir * 0.001f;
        depth++;
    }
    return depth + 1u;
}

// ---- VCM: VertexConnectionAndMerging.cpp -------------------------------------------------------------------------------
//
// The per-vertex arithmetic of the reference's VCM renderer as free functions on plain structs, so that the CPU model
// (tests/hostsim/hostsim_vcm.h, which pins them statistically against the reference) and a future device version call
// the same code.  `Occluded(ray, tMax)` is the visibility test (Scene::Traverse_Shadow); Mis() is the identity there.

struct BidirPath
{
    Ray      ray;
    V4       throughput;
    float    dVC, dVM, dVCM;          // MIS quantities (VertexConnectionAndMerging.h:98-112)
    uint32_t length;
    uint32_t lastSpecular, isFiniteLight;
};
struct BidirLightVertex { Vertex vx; V4 throughput; float dVC, dVM, dVCM; uint32_t pathLength; };
struct BidirPhoton { V3 position, direction, throughput; float dVM, dVCM; };

struct BidirPass
{
    uint32_t useVertexConnection, useVertexMerging, maxPathLength, iteration;
    float    radiusVC, radiusVM, mergingNormalization;
    float    vmWeightVC, vcWeightVC, vmWeightVM, vcWeightVM;     // mMisVertex{Merging,Connection}WeightFactor{VC,VM}
};

// PreRender(passNumber, film), :80-125.  `pass` carries radiusVC from the previous call (any value for pass 0).
RTB_HD void bidir_begin_pass(BidirPass& pass, uint32_t passNumber, float lightPathsCount, float initialRadius, float minRadius, float radiusMultiplier)
{
    pass.iteration = passNumber;
    if (passNumber == 0) pass.radiusVC = pass.radiusVM = initialRadius;
    else
    {
        pass.radiusVM = pass.radiusVC;
        pass.radiusVC *= radiusMultiplier;
        pass.radiusVC = s_max(pass.radiusVC, minRadius);
    }
    pass.mergingNormalization = 1.0f / (s_sqr(pass.radiusVM) * RTB_PI * lightPathsCount);
    {
        const float etaVCM = RTB_PI * s_sqr(pass.radiusVC) * lightPathsCount;
        pass.vmWeightVC = (pass.useVertexMerging && passNumber > 0) ? etaVCM : 0.0f;
        pass.vcWeightVC = pass.useVertexConnection ? 1.0f / etaVCM : 0.0f;
    }
    {
        const float etaVCM = RTB_PI * s_sqr(pass.radiusVM) * lightPathsCount;
        pass.vmWeightVM = pass.useVertexMerging ? etaVCM : 0.0f;
        pass.vcWeightVM = pass.useVertexConnection ? 1.0f / etaVCM : 0.0f;
    }
}

RTB_HD bool bidir_is_delta_bsdf(const Vertex& vx) { return vx.bsdf == RTB_BSDF_METAL || vx.bsdf == RTB_BSDF_DIELECTRIC; }      // BSDF::IsDelta

// Material::Evaluate with both pdfs
RTB_HD V4 bidir_evaluate(const Vertex& vx, const V3& incomingDirWorld, float& pdfW, float& revPdfW)
{
    const V3 out = world_to_local(vx, vx.outgoingDirWorld), in = world_to_local(vx, incomingDirWorld);
    pdfW = 0.0f;
    const V4 value = bsdf_evaluate(vx, out, in, pdfW);
    revPdfW = bsdf_pdf(vx, out, in, true);
    return value;
}

// the camera path's start, RenderPixel :176-184
RTB_HD void bidir_begin_camera_path(BidirPath& path, const Ray& primaryRay, const CameraConnection& camera)
{
    path.ray = primaryRay;
    path.throughput = v4(1.0f, 1.0f, 1.0f, 1.0f);
    path.dVC = 0.0f; path.dVM = 0.0f;
    path.dVCM = 1.0f / camera_pdf_w(camera.localToWorld, camera.tanHalfFoV, camera.aspectRatio, primaryRay.dir);
    path.length = 1; path.lastSpecular = 1; path.isFiniteLight = 0;
}

// GenerateLightSample, :430-485: `pick` chooses the light, (p0, p1, p2, d0, d1) are the emission samples
template <bool kTex = true>
RTB_HD bool bidir_generate_light_sample(const SceneView& S, const BidirPass& pass, uint32_t pick, float p0, float p1, float p2, float d0, float d1, BidirPath& outPath)
{
    if (S.numLights == 0) return false;
    const float lightPickProbability = 1.0f / (float)S.numLights;
    const DLight& light = S.lights[pick % S.numLights];
    EmitResult emit;
    const V4 throughput = light_emit<kTex>(S, light, p0, p1, p2, d0, d1, emit);
    if (almost_zero(throughput)) return false;
    emit.directPdfA *= lightPickProbability;
    emit.emissionPdfW *= lightPickProbability;
    const float emissionInvPdfW = 1.0f / emit.emissionPdfW;
    emit.position = emit.position + emit.direction * 0.0005f;
    outPath.ray = make_ray(emit.position, emit.direction);
    outPath.throughput = throughput * emissionInvPdfW;
    outPath.length = 1; outPath.lastSpecular = 1;
    outPath.isFiniteLight = (light.flags & RTB_LIGHT_FLAG_FINITE) != 0 ? 1u : 0u;
    outPath.dVCM = emit.directPdfA * emissionInvPdfW;
    if ((light.flags & RTB_LIGHT_FLAG_DELTA) == 0) outPath.dVC = (outPath.isFiniteLight ? emit.cosAtLight : 1.0f) * emissionInvPdfW;
    else outPath.dVC = 0.0f;
    outPath.dVM = outPath.dVC * pass.vcWeightVC;
    return true;
}

// the MIS update when a path arrives at a surface: camera paths :203-211, light paths :352-364
RTB_HD void bidir_arrive(BidirPath& path, const Vertex& vx, float hitDistance, bool lightPath)
{
    if (!lightPath || path.length > 1 || path.isFiniteLight) path.dVCM *= s_sqr(hitDistance);
    const float cosTheta = dot3(path.ray.dir, vx.normal);
    const float invMis = 1.0f / s_abs(cosTheta);
    path.dVCM *= invMis; path.dVC *= invMis; path.dVM *= invMis;
}

// AdvancePath, :487-575
RTB_HD bool bidir_advance_path(const BidirPass& pass, BidirPath& path, const Vertex& vx, float s0, float s1, float s2)
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
        path.lastSpecular = 1;
    }
    else
    {
        const float bsdfRevPdf = bsdf_pdf(vx, outLocal, neg(world_to_local(vx, incomingDirWorld)), true);
        const float invBsdfDirPdf = 1.0f / bs.pdf;
        path.dVC = (cosThetaOut * invBsdfDirPdf) * (path.dVC * bsdfRevPdf + path.dVCM + pass.vmWeightVC);
        path.dVM = (cosThetaOut * invBsdfDirPdf) * (path.dVM * bsdfRevPdf + path.dVCM * pass.vcWeightVC + 1.0f);
        path.dVCM = invBsdfDirPdf;
        path.lastSpecular = 0;
    }
    return true;
}

// EvaluateLight, :577-626 (`hitVertex` null: a global light seen by a ray that left the scene)
template <bool kTex = true>
RTB_HD V4 bidir_evaluate_light(const SceneView& S, const BidirPass& pass, const DLight& light, const Vertex* hitVertex, const BidirPath& path)
{
    const Ray lightSpaceRay = transform_ray_unsafe(light.invTransform, path.ray);
    const float cosAtLight = hitVertex ? -dot3(hitVertex->normal, path.ray.dir) : 1.0f;
    float directPdfA = 0.0f, emissionPdfW = 0.0f;
    V4 contribution = light_radiance_bidir<kTex>(S, light, lightSpaceRay.dir, cosAtLight, directPdfA, emissionPdfW);
    if (almost_zero(contribution)) return v4s(0.0f);
    if (path.length > 1)
    {
        const bool useVertexMerging = pass.useVertexMerging && pass.iteration > 0;
        if (useVertexMerging && !pass.useVertexConnection)
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
template <bool kTex = true, typename Occluded>
RTB_HD V4 bidir_sample_light(const SceneView& S, const BidirPass& pass, const DLight& light, const Vertex& vx, const BidirPath& path, float u0, float u1, float u2,
                             const Occluded& occluded)
{
    IlluminateResultBidir ir;
    const V4 radiance = light_illuminate_bidir<kTex>(S, light, vx, u0, u1, u2, ir);
    if (almost_zero(radiance)) return v4s(0.0f);
    float bsdfPdfW, bsdfRevPdfW;
    const V4 bsdfFactor = bidir_evaluate(vx, neg(ir.directionToLight), bsdfPdfW, bsdfRevPdfW);
    if (almost_zero(bsdfFactor)) return v4s(0.0f);
    Ray shadowRay = make_ray(vx.position, ir.directionToLight);
    shadowRay.origin = shadowRay.origin + shadowRay.dir * 0.0001f;
    if (occluded(shadowRay, ir.distance * 0.999f)) return v4s(0.0f);
    const bool isDeltaLight = (light.flags & RTB_LIGHT_FLAG_DELTA) != 0;
    bsdfPdfW *= isDeltaLight ? 0.0f : 1.0f;
    const float cosToLight = dot3(vx.normal, ir.directionToLight);
    if (cosToLight <= FLT_EPSILON) return v4s(0.0f);
    const float wLight = bsdfPdfW / ir.directPdfW;
    const float wCamera = (ir.emissionPdfW * cosToLight / (ir.directPdfW * ir.cosAtLight)) * (pass.vmWeightVC + path.dVCM + path.dVC * bsdfRevPdfW);
    const float misWeight = 1.0f / (wLight + 1.0f + wCamera);
    return (radiance * bsdfFactor) * (misWeight / ir.directPdfW);
}

// ConnectVertices, :738-806
template <typename Occluded>
RTB_HD V4 bidir_connect_vertices(const BidirPass& pass, const BidirPath& cameraPath, const Vertex& vx, const BidirLightVertex& lv, const Occluded& occluded)
{
    V3 lightDir = lv.vx.position - vx.position;
    const float distanceSqr = sqr_length3(lightDir);
    const float distance = sqrtf(distanceSqr);
    lightDir = lightDir / distance;
    const float cosCameraVertex = dot3(vx.normal, lightDir);
    const float cosLightVertex = dot3(lv.vx.normal, neg(lightDir));
    if (cosCameraVertex <= 0.0f || cosLightVertex <= 0.0f) return v4s(0.0f);
    const float geometryTerm = 1.0f / distanceSqr;
    float cameraBsdfPdfW, cameraBsdfRevPdfW, lightBsdfPdfW, lightBsdfRevPdfW;
    const V4 cameraFactor = bidir_evaluate(vx, neg(lightDir), cameraBsdfPdfW, cameraBsdfRevPdfW);
    if (almost_zero(cameraFactor)) return v4s(0.0f);
    const V4 lightFactor = bidir_evaluate(lv.vx, lightDir, lightBsdfPdfW, lightBsdfRevPdfW);
    if (almost_zero(lightFactor)) return v4s(0.0f);
    Ray shadowRay = make_ray(vx.position, lightDir);
    shadowRay.origin = shadowRay.origin + shadowRay.dir * 0.0001f;
    if (occluded(shadowRay, distance * 0.999f)) return v4s(0.0f);
    const float cameraBsdfPdfA = cameraBsdfPdfW * s_abs(cosLightVertex) / s_sqr(distance);        // PdfWtoA, :23-26
    const float lightBsdfPdfA = lightBsdfPdfW * s_abs(cosCameraVertex) / s_sqr(distance);
    const float wLight = cameraBsdfPdfA * (pass.vmWeightVC + lv.dVCM + lv.dVC * lightBsdfRevPdfW);
    const float wCamera = lightBsdfPdfA * (pass.vmWeightVC + cameraPath.dVCM + cameraPath.dVC * cameraBsdfRevPdfW);
    const float misWeight = 1.0f / (wLight + 1.0f + wCamera);
    return (cameraFactor * lightFactor) * (geometryTerm * misWeight);
}

// one photon of MergeVertices' range query, :822-851 (the caller has checked the distance); adds to `contribution`
RTB_HD void bidir_merge_photon(const BidirPass& pass, const BidirPath& cameraPath, const Vertex& vx, const BidirPhoton& photon, V4& contribution)
{
    const float cosToLight = dot3(vx.normal, photon.direction);
    if (cosToLight < FLT_EPSILON) return;
    float cameraBsdfDirPdfW, cameraBsdfRevPdfW;
    const V4 cameraBsdfFactor = bidir_evaluate(vx, neg(photon.direction), cameraBsdfDirPdfW, cameraBsdfRevPdfW);
    if (almost_zero(cameraBsdfFactor)) return;
    const float wLight = photon.dVCM * pass.vcWeightVM + photon.dVM * cameraBsdfDirPdfW;
    const float wCamera = cameraPath.dVCM * pass.vcWeightVM + cameraPath.dVM * cameraBsdfRevPdfW;
    const float weight = (1.0f / (wLight + 1.0f + wCamera)) / cosToLight;
    const V4 t = cameraBsdfFactor * v4(photon.throughput.x, photon.throughput.y, photon.throughput.z, 0.0f);
    contribution = v4(fma_(t.x, weight, contribution.x), fma_(t.y, weight, contribution.y), fma_(t.z, weight, contribution.z), fma_(t.w, weight, contribution.w));
}

// ConnectToCamera, :908-970: true with the film position and the value to splat
template <typename Occluded>
RTB_HD bool bidir_connect_to_camera(const BidirPass& pass, const CameraConnection& camera, const BidirLightVertex& lv, const Occluded& occluded,
                                    float& filmX, float& filmY, V3& value)
{
    const V3 cameraPos = xyz(camera.localToWorld.r3);
    V3 dirToCamera = cameraPos - lv.vx.position;
    const float cameraDistanceSqr = sqr_length3(dirToCamera);
    const float cameraDistance = sqrtf(cameraDistanceSqr);
    dirToCamera = dirToCamera / cameraDistance;
    float bsdfPdfW, bsdfRevPdfW;
    const V4 cameraFactor = bidir_evaluate(lv.vx, neg(dirToCamera), bsdfPdfW, bsdfRevPdfW);
    if (almost_zero(cameraFactor)) return false;
    if (!camera_world_to_film(camera.worldToScreen, lv.vx.position, filmX, filmY)) return false;
    Ray shadowRay = make_ray(lv.vx.position, dirToCamera);
    shadowRay.origin = shadowRay.origin + shadowRay.dir * 0.0001f;
    if (occluded(shadowRay, cameraDistance * 0.999f)) return false;
    const float cosToCamera = dot3(dirToCamera, lv.vx.normal);
    if (cosToCamera <= FLT_EPSILON) return false;
    const float cameraPdfW = camera_pdf_w(camera.localToWorld, camera.tanHalfFoV, camera.aspectRatio, neg(dirToCamera));
    const float cameraPdfA = cameraPdfW * cosToCamera / cameraDistanceSqr;
    const float wLight = cameraPdfA * (pass.vmWeightVC + lv.dVCM + lv.dVC * bsdfRevPdfW);
    const float misWeight = 1.0f / (wLight + 1.0f);
    value = xyz((cameraFactor * lv.throughput) * (misWeight * cameraPdfA / cosToCamera));
    return true;
}

} // namespace rtb
