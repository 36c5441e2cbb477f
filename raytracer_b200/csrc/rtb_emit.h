// rtb_emit.h — ILight::Emit for the five light types (the first step of every light sub-path of the
// reference's Light Tracer and VCM renderers, Rendering/LightTracer.cpp:44-66, VertexConnectionAndMerging.cpp:320-360).
//
// STATUS: building block.  No kernel launches this yet (the bidirectional renderers are not built); the
// functions are restated here in the device headers' arithmetic and pinned bit for bit against the reference's own
// ILight::Emit through tests/hostsim (`emit` command, tests/test_hostsim_parity.py::test_hostsim_light_emission_*), so
// that the renderers can be built on checked parts.  rtb_kernels.cu does not include this file.
#pragma once

#include "rtb_shading.h"

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

} // namespace rtb
