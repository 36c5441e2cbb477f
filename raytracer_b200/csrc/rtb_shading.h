// rtb_shading.h — intersection evaluation, materials (9 BSDFs), lights (5 types).  Per-vertex device functions.
//
//   Scene::EvaluateIntersection + shape EvaluateIntersection   Scene.cpp:305-365, SphereShape.cpp:159-175,
//                                                              BoxShape.cpp:181-191, RectShape.cpp:124-132, MeshShape.cpp:283-328
//   Material::EvaluateShadingData / Sample / Evaluate          Material.cpp:151-232, Intersection.h:15-36
//   BSDFs                                                      Core/Material/BSDF/*.cpp, Microfacet.h
//   IShape::Sample / Pdf, SphereShape / RectShape / BoxShape   Shape.cpp:65-101, SphereShape.cpp:48-125, RectShape.cpp:51-95, BoxShape.cpp:127-178
//   ILight::Illuminate / GetRadiance (area, background, directional, point, spot)   Core/Scene/Light/*.cpp
#pragma once

#include "rtb_trace.h"
#include "rtb_sampler.h"
#include "rtb_texture.h"

namespace rtb {

// BSDF::EventType, BSDF.h:25-43
enum : uint32_t
{
    EV_NULL = 0, EV_DIFFUSE_REFLECTION = 1, EV_DIFFUSE_TRANSMISSION = 2, EV_GLOSSY_REFLECTION = 4,
    EV_GLOSSY_REFRACTION = 8, EV_SPECULAR_REFLECTION = 16, EV_SPECULAR_REFRACTION = 32,
    EV_SPECULAR = EV_SPECULAR_REFLECTION | EV_SPECULAR_REFRACTION
};

#define RTB_COS_EPSILON 1.0e-5f                 // BSDF::CosEpsilon, BSDF.h:53
#define RTB_SPECULAR_ROUGHNESS_THRESHOLD 0.005f // BSDF::SpecularEventRoughnessTreshold, BSDF.h:57
#define RTB_EPSILON 0.000001f                   // RT_EPSILON, Math.h:14

// RayColor::AlmostZero: all FOUR lanes |x| < 1e-6   RayColor.h:100-103, Vector4Impl.h:126-129
RTB_HD bool almost_zero(const V4& c)
{
    return (bit_abs(c.x - 0.0f) < RTB_EPSILON) && (bit_abs(c.y - 0.0f) < RTB_EPSILON) &&
           (bit_abs(c.z - 0.0f) < RTB_EPSILON) && (bit_abs(c.w - 0.0f) < RTB_EPSILON);
}
// RayColor::Max (w masked out), RayColor.h:105-112 + Vector4::HorizontalMax
RTB_HD float color_max(const V4& c) { return sse_max(sse_max(c.x, c.y), sse_max(c.z, 0.0f)); }

// IntersectionData (Intersection.h:9-37) + ShadingData (ShadingData.h:12-32)
struct Vertex
{
    V3 tangent, bitangent, normal, position;   // frame[0..3], world space
    float texU, texV;
    uint32_t materialIndex;
    // SampledMaterialParameters
    V4 baseColor, emission;
    float roughness, metalness, IoR, K;
    uint32_t bsdf;
    V3 outgoingDirWorld;
};

// IntersectionData::WorldToLocal, Intersection.h:21-31
RTB_HD V3 world_to_local(const Vertex& vx, const V3& w)
{
    V3 r = v3(vx.tangent.x * w.x, vx.bitangent.x * w.x, vx.normal.x * w.x);
    r = v3(fma_(vx.tangent.y, w.y, r.x), fma_(vx.bitangent.y, w.y, r.y), fma_(vx.normal.y, w.y, r.z));
    r = v3(fma_(vx.tangent.z, w.z, r.x), fma_(vx.bitangent.z, w.z, r.y), fma_(vx.normal.z, w.z, r.z));
    return r;
}
// IntersectionData::LocalToWorld = frame.TransformVector, Intersection.h:16-19
RTB_HD V3 local_to_world(const Vertex& vx, const V3& l)
{
    V3 t = vx.tangent * l.x;
    t = mul_add(vx.bitangent, l.y, t);
    t = mul_add(vx.normal, l.z, t);
    return t;
}

// helper::ConvertXYZtoCubeUV, BoxShape.cpp:25-91 (returns the face index; uv only when requested)
RTB_HD int32_t cube_side(const V3& p, float& outU, float& outV)
{
    const float ax = bit_abs(p.x), ay = bit_abs(p.y), az = bit_abs(p.z);
    const int32_t isXPositive = p.x > 0 ? 1 : 0, isYPositive = p.y > 0 ? 1 : 0, isZPositive = p.z > 0 ? 1 : 0;
    float maxAxis, uc, vc;
    int32_t side;
    if (ax >= ay && ax >= az)
    {
        uc = isXPositive ? -p.z : p.z;
        side = isXPositive; maxAxis = ax; vc = p.y;
    }
    else if (ay >= ax && ay >= az)
    {
        vc = isYPositive ? -p.z : p.z;
        side = isYPositive + 2; maxAxis = ay; uc = p.x;
    }
    else
    {
        uc = isZPositive ? p.x : -p.x;
        side = isZPositive + 4; maxAxis = az; vc = p.y;
    }
    outU = uc / (2.0f * maxAxis) + 0.5f;
    outV = vc / (2.0f * maxAxis) + 0.5f;
    return side;
}

// Scene::EvaluateIntersection (+ per-shape EvaluateIntersection + Material::EvaluateShadingData).
// `wantMaterial` is false for light hits (LightSceneObject::EvaluateIntersection leaves the material untouched).
// kTex = false compiles the texture lookups out (scenes without textures run a k_shade without the calls)
template <bool kTex = true>
RTB_HD void evaluate_intersection(const SceneView& S, const Ray& ray, const Hit& hit, Vertex& vx, uint32_t staleMaterial)
{
    const DObjectShade& obj = S.objShade[hit.objectId];
    const M4& transform = obj.transform;
    const M4 invTransform = fast_inverse_no_scale(transform);      // Scene.cpp:313-314 (NOT the stored inverse)
    const V3 worldPosition = ray_at(ray, hit.distance);
    const V3 localPos = transform_point(invTransform, worldPosition);

    V3 tangent, normal;
    vx.texU = 0.0f; vx.texV = 0.0f;
    vx.materialIndex = obj.materialIndex;
    const RtbShape& shape = S.shapes[obj.shapeIndex];
    if (shape.type == RTB_SHAPE_SPHERE)
    {
        // SphereShape::EvaluateIntersection, SphereShape.cpp:159-175
        cartesian_to_spherical(neg(localPos), vx.texU, vx.texV);
        const float invRadius = shape.invSize[0];
        normal = localPos * invRadius;
        tangent = v3(-normal.z, 0.0f, normal.x);     // (n.zxxx & mask<1,0,1,0>).ChangeSign<1,0,0,0>
        tangent = normalize3(tangent);
        normal = normalize3(normal);
    }
    else if (shape.type == RTB_SHAPE_BOX)
    {
        // BoxShape::EvaluateIntersection, BoxShape.cpp:181-191 + g_faceFrames 15-23
        const V3 p = v3(localPos.x * shape.invSize[0], localPos.y * shape.invSize[1], localPos.z * shape.invSize[2]);
        const int32_t side = cube_side(p, vx.texU, vx.texV);
        switch (side)
        {
        case 0:  tangent = v3(0.0f, 0.0f, 1.0f);  normal = v3(-1.0f, 0.0f, 0.0f); break;
        case 1:  tangent = v3(0.0f, 0.0f, -1.0f); normal = v3(1.0f, 0.0f, 0.0f);  break;
        case 2:  tangent = v3(1.0f, 0.0f, 0.0f);  normal = v3(0.0f, -1.0f, 0.0f); break;
        case 3:  tangent = v3(1.0f, 0.0f, 0.0f);  normal = v3(0.0f, 1.0f, 0.0f);  break;
        case 4:  tangent = v3(-1.0f, 0.0f, 0.0f); normal = v3(0.0f, 0.0f, -1.0f); break;
        default: tangent = v3(1.0f, 0.0f, 0.0f);  normal = v3(0.0f, 0.0f, 1.0f);  break;
        }
    }
    else if (shape.type == RTB_SHAPE_RECT)
    {
        // RectShape::EvaluateIntersection, RectShape.cpp:124-132
        vx.texU = localPos.x * shape.aux[0]; vx.texV = localPos.y * shape.aux[1];
        tangent = v3(1.0f, 0.0f, 0.0f); normal = v3(0.0f, 0.0f, 1.0f);
    }
    else
    {
        // MeshShape::EvaluateIntersection, MeshShape.cpp:283-328
        const DMesh& mesh = S.meshes[shape.meshIndex];
        const DVertexIndices idx = mesh.indices[hit.subObjectId];
        if (idx.materialIndex != 0xFFFFFFFFu) vx.materialIndex = mesh.materialIds[idx.materialIndex];
        const DVertexShading a = mesh.shading[idx.i0], b = mesh.shading[idx.i1], c = mesh.shading[idx.i2];
        const float coeff1 = hit.u, coeff2 = hit.v;
        const float coeff0 = 1.0f - (coeff1 + coeff2);
        vx.texU = fma_(coeff0, a.u, fma_(coeff2, c.u, coeff1 * b.u));
        vx.texV = fma_(coeff0, a.v, fma_(coeff2, c.v, coeff1 * b.v));
        V3 t = v3(b.tx, b.ty, b.tz) * coeff1;
        t = mul_add(v3(c.tx, c.ty, c.tz), coeff2, t);
        t = mul_add(v3(a.tx, a.ty, a.tz), coeff0, t);
        tangent = normalize3(t);                       // FastNormalize3 (exact in the oracle)
        V3 n = v3(b.nx, b.ny, b.nz) * coeff1;
        n = mul_add(v3(c.nx, c.ny, c.nz), coeff2, n);
        n = mul_add(v3(a.nx, a.ny, a.nz), coeff0, n);
        normal = normalize3(n);
    }

    // normal mapping, Scene.cpp:323-337.  A light object does not set IntersectionData::material, so the reference
    // applies the normal map of the material the path bounced from (see PathState::lastMaterial), with the light
    // shape's own texture coordinates; on a camera ray the material is still null and nothing is applied.
    const uint32_t normalMapMaterial = obj.kind == RTB_OBJECT_SHAPE ? vx.materialIndex : staleMaterial;
    if (kTex && S.materialTextures != nullptr && normalMapMaterial != RTB_NO_MATERIAL)
    {
        const RtbMaterialTextures& mt = S.materialTextures[normalMapMaterial];
        if (mt.normalMap != RTB_NO_TEXTURE)
        {
            const V3 bitangent = cross3(tangent, normal);
            const V3 localNormal = material_normal_vector(S, mt, vx.texU, vx.texV);
            V3 newNormal = tangent * localNormal.x;
            newNormal = mul_add(bitangent, localNormal.y, newNormal);
            newNormal = mul_add(normal, localNormal.z, newNormal);
            normal = normalize3(newNormal);             // FastNormalized3 (exact in the oracle)
        }
    }
    // Scene.cpp:339-348
    tangent = normalize3(orthogonalize(tangent, normal));
    vx.normal = transform_vector(transform, normal);
    vx.tangent = transform_vector(transform, tangent);
    vx.bitangent = cross3(vx.tangent, vx.normal);
    vx.position = worldPosition;
}

// Material::EvaluateShadingData, Material.cpp:151-158: every parameter is baseValue * texture(uv) when it has a texture
// (MaterialParameter::Evaluate, MaterialParameter.h:22-32; a float parameter takes the texture's x lane)
template <bool kTex = true>
RTB_HD void evaluate_shading_data(const SceneView& S, Vertex& vx)
{
    const RtbMaterial& m = S.materials[vx.materialIndex];
    vx.baseColor = v4(m.baseColor[0], m.baseColor[1], m.baseColor[2], m.baseColor[3]);
    vx.emission = v4(m.emission[0], m.emission[1], m.emission[2], m.emission[3]);
    vx.roughness = m.roughness; vx.metalness = m.metalness; vx.IoR = m.IoR; vx.K = m.K;
    vx.bsdf = m.bsdf;
    if (kTex && S.materialTextures != nullptr)
    {
        const RtbMaterialTextures& mt = S.materialTextures[vx.materialIndex];
        if (mt.baseColor != RTB_NO_TEXTURE) vx.baseColor = vx.baseColor * texture_evaluate(S, mt.baseColor, vx.texU, vx.texV);
        if (mt.emission != RTB_NO_TEXTURE) vx.emission = vx.emission * texture_evaluate(S, mt.emission, vx.texU, vx.texV);
        if (mt.roughness != RTB_NO_TEXTURE) vx.roughness = vx.roughness * texture_evaluate(S, mt.roughness, vx.texU, vx.texV).x;
        if (mt.metalness != RTB_NO_TEXTURE) vx.metalness = vx.metalness * texture_evaluate(S, mt.metalness, vx.texU, vx.texV).x;
    }
}

// ---- Microfacet (GGX), Microfacet.h:10-60 -----------------------------------------------------------------------
struct Microfacet
{
    float alphaSqr;
    RTB_HD explicit Microfacet(float alpha) : alphaSqr(alpha * alpha) {}
    RTB_HD float D(const V3& m) const
    {
        const float NdotH = m.z;
        const float cosThetaSq = s_sqr(NdotH);
        const float tanThetaSq = s_max(1.0f - cosThetaSq, 0.0f) / cosThetaSq;
        const float cosThetaQu = cosThetaSq * cosThetaSq;
        return alphaSqr * RTB_INV_PI / (cosThetaQu * s_sqr(alphaSqr + tanThetaSq));
    }
    RTB_HD float Pdf(const V3& m) const { return D(m) * s_abs(m.z); }
    RTB_HD float G(const float NdotV, const float NdotL) const
    {
        const float tanThetaSqV = (1.0f - NdotV * NdotV) / (NdotV * NdotV);
        const float tanThetaSqL = (1.0f - NdotL * NdotL) / (NdotL * NdotL);
        return 4.0f / ((1.0f + sqrtf(1.0f + alphaSqr * tanThetaSqV)) * (1.0f + sqrtf(1.0f + alphaSqr * tanThetaSqL)));
    }
    RTB_HD V3 Sample(float ux, float uy) const
    {
        const float cosThetaSqr = (1.0f - ux) / (1.0f + (alphaSqr - 1.0f) * ux);
        const float cosTheta = sqrtf(cosThetaSqr);
        const float sinTheta = sqrtf(1.0f - cosThetaSqr);
        const float phi = RTB_2PI * uy;
        float s, c;
        sin_cos(phi, s, c);
        return v3(sinTheta * s, sinTheta * c, cosTheta);
    }
};

struct BsdfSample
{
    V4 color;
    V3 incomingDir;     // local space
    float pdf;
    uint32_t event;
};

// RoughDiffuseBSDF::Evaluate_Internal, RoughDiffuseBSDF.cpp:14-24
RTB_HD float oren_nayar(const float NdotL, const float NdotV, const float LdotV, const float roughness)
{
    const float s2 = roughness * roughness;
    const float A = 1.0f - 0.50f * s2 / (0.33f + s2);
    const float B = 0.45f * s2 / (0.09f + s2);
    const float s = LdotV - NdotL * NdotV;
    const float stinv = s > 0.0f ? s / s_max(NdotL, NdotV) : 0.0f;
    return s_max(A + B * stinv, 0.0f);
}

RTB_HD bool sample_diffuse(const Vertex& vx, const V3& out, float u0, float u1, BsdfSample& r)
{
    // DiffuseBSDF::Sample, DiffuseBSDF.cpp:14-29
    if (out.z < RTB_COS_EPSILON) return false;
    r.incomingDir = hemisphere_cos(u0, u1);
    r.pdf = r.incomingDir.z * RTB_INV_PI;
    r.color = vx.baseColor;
    r.event = EV_DIFFUSE_REFLECTION;
    return true;
}

RTB_HD bool sample_metal(const Vertex& vx, const V3& out, BsdfSample& r)
{
    // MetalBSDF::Sample, MetalBSDF.cpp:15-33
    const float NdotV = out.z;
    if (NdotV < RTB_COS_EPSILON) return false;
    const float F = fresnel_metal(NdotV, vx.IoR, vx.K);
    r.color = vx.baseColor * v4s(F);
    r.incomingDir = neg(reflect3(out, v3(0.0f, 0.0f, 1.0f)));
    r.pdf = 1.0f;
    r.event = EV_SPECULAR_REFLECTION;
    return true;
}

RTB_HD bool sample_dielectric(const Vertex& vx, const V3& out, float u0, BsdfSample& r)
{
    // DielectricBSDF::Sample, DielectricBSDF.cpp:15-103 (no dispersion: spectral rendering is compiled out)
    const float NdotV = out.z;
    if (s_abs(NdotV) < RTB_COS_EPSILON) return false;
    const float ior = vx.IoR;
    const float F = fresnel_dielectric(NdotV, ior);
    const float minReflectionProbability = 0.25f;
    const float reflectionProbability = minReflectionProbability + (1.0f - minReflectionProbability) * F;
    const float refractionProbability = 1.0f - reflectionProbability;
    const bool reflection = (reflectionProbability >= 1.0f) || u0 < reflectionProbability;
    if (reflection)
    {
        r.incomingDir = neg(reflect3(out, v3(0.0f, 0.0f, 1.0f)));
        r.event = EV_SPECULAR_REFLECTION;
    }
    else
    {
        r.incomingDir = refract3(neg(out), v3(0.0f, 0.0f, 1.0f), ior);
        r.event = EV_SPECULAR_REFRACTION;
    }
    const float NdotL = r.incomingDir.z;
    if ((NdotV * NdotL > 0.0f) != reflection) return false;
    if (reflection)
    {
        r.pdf = reflectionProbability;
        r.color = v4s(1.0f) * (F / reflectionProbability);
    }
    else
    {
        r.pdf = refractionProbability;
        r.color = vx.baseColor * ((1.0f - F) / refractionProbability);
    }
    return true;
}

RTB_HD bool sample_plastic(const Vertex& vx, const V3& out, float u0, float u1, float u2, BsdfSample& r)
{
    // PlasticBSDF::Sample, PlasticBSDF.cpp:15-58
    const float NdotV = out.z;
    if (NdotV < RTB_COS_EPSILON) return false;
    const float ior = vx.IoR;
    const float Fi = fresnel_dielectric(NdotV, ior);
    const float minSpecularWeight = 0.25f;
    const float specularWeight = minSpecularWeight + Fi * (1.0f - minSpecularWeight);
    const float diffuseWeight = (1.0f - Fi) * color_max(vx.baseColor);
    const float specularProbability = specularWeight / (specularWeight + diffuseWeight);
    const float diffuseProbability = 1.0f - specularProbability;
    const bool specular = (specularProbability >= 1.0f) || (u2 < specularProbability);
    if (specular)
    {
        r.color = v4s(Fi / specularProbability);
        r.incomingDir = neg(reflect3(out, v3(0.0f, 0.0f, 1.0f)));
        r.pdf = specularProbability;
        r.event = EV_SPECULAR_REFLECTION;
    }
    else
    {
        r.incomingDir = hemisphere_cos(u0, u1);
        const float NdotL = r.incomingDir.z;
        r.pdf = r.incomingDir.z * RTB_INV_PI * diffuseProbability;
        const float Fo = fresnel_dielectric(NdotL, ior);
        r.color = vx.baseColor * ((1.0f - Fi) * (1.0f - Fo) / diffuseProbability);
        r.event = EV_DIFFUSE_REFLECTION;
    }
    return true;
}

// Material::Sample -> BSDF::Sample.  `out` is the outgoing direction in local space.
RTB_HD bool bsdf_sample(const Vertex& vx, const V3& out, float u0, float u1, float u2, BsdfSample& r)
{
    r.color = v4s(0.0f); r.incomingDir = v3s(0.0f); r.pdf = 0.0f; r.event = EV_NULL;
    switch (vx.bsdf)
    {
    case RTB_BSDF_DIFFUSE:
        return sample_diffuse(vx, out, u0, u1, r);

    case RTB_BSDF_ROUGH_DIFFUSE:
    {
        // RoughDiffuseBSDF::Sample, RoughDiffuseBSDF.cpp:26-46
        const float NdotV = out.z;
        if (NdotV < RTB_COS_EPSILON) return false;
        r.incomingDir = hemisphere_cos(u0, u1);
        const float NdotL = r.incomingDir.z;
        const float LdotV = s_max(0.0f, dot3(out, neg(r.incomingDir)));
        const float value = oren_nayar(NdotL, NdotV, LdotV, vx.roughness);
        r.pdf = NdotL * RTB_INV_PI;
        r.color = vx.baseColor * value;
        r.event = EV_DIFFUSE_REFLECTION;
        return true;
    }

    case RTB_BSDF_METAL:
        return sample_metal(vx, out, r);

    case RTB_BSDF_ROUGH_METAL:
    {
        // RoughMetalBSDF::Sample, RoughMetalBSDF.cpp:17-60
        const float roughness = vx.roughness;
        if (roughness < RTB_SPECULAR_ROUGHNESS_THRESHOLD) return sample_metal(vx, out, r);
        const float NdotV = out.z;
        if (NdotV < RTB_COS_EPSILON) return false;
        const Microfacet microfacet(roughness * roughness);
        const V3 m = microfacet.Sample(u0, u1);
        r.incomingDir = neg(reflect3(out, m));
        if (r.incomingDir.z < RTB_COS_EPSILON) return false;
        const float NdotL = r.incomingDir.z;
        const float VdotH = dot3(m, out);
        const float pdf = microfacet.Pdf(m);
        const float D = microfacet.D(m);
        const float G = microfacet.G(NdotV, NdotL);
        const float F = fresnel_metal(VdotH, vx.IoR, vx.K);
        r.pdf = pdf / (4.0f * VdotH);
        r.color = vx.baseColor * v4s(VdotH * F * G * D / (pdf * NdotV));
        r.event = EV_GLOSSY_REFLECTION;
        return true;
    }

    case RTB_BSDF_DIELECTRIC:
        return sample_dielectric(vx, out, u0, r);

    case RTB_BSDF_ROUGH_DIELECTRIC:
    {
        // RoughDielectricBSDF::Sample, RoughDielectricBSDF.cpp:17-110
        const float NdotV = out.z;
        if (s_abs(NdotV) < RTB_COS_EPSILON) return false;
        const float ior = vx.IoR;
        const float roughness = vx.roughness;
        if (roughness < RTB_SPECULAR_ROUGHNESS_THRESHOLD) return sample_dielectric(vx, out, u0, r);
        const Microfacet microfacet(roughness * roughness);
        const V3 m = microfacet.Sample(u0, u1);
        const float microfacetPdf = microfacet.Pdf(m);
        const float VdotH = dot3(m, out);
        const float F = fresnel_dielectric(VdotH, ior);
        const bool reflection = u2 < F;
        if (reflection)
        {
            r.incomingDir = neg(reflect3(out, m));
            r.event = EV_GLOSSY_REFLECTION;
        }
        else
        {
            r.incomingDir = refract3(neg(out), m, ior);
            r.event = EV_GLOSSY_REFRACTION;
        }
        const float NdotL = r.incomingDir.z;
        const float LdotH = dot3(m, r.incomingDir);
        if ((NdotV * NdotL > 0.0f) != reflection) return false;
        const float D = microfacet.D(m);
        const float G = microfacet.G(NdotV, NdotL);
        r.color = v4s(s_abs(VdotH) * G * D / (microfacetPdf * s_abs(NdotV)));
        if (reflection)
        {
            r.pdf = F * microfacetPdf / (4.0f * s_abs(VdotH));
        }
        else
        {
            const float eta = NdotV < 0.0f ? ior : 1.0f / ior;
            const float denom = s_sqr(eta * VdotH + LdotH);
            r.pdf = (1.0f - F) * microfacetPdf * s_abs(LdotH) / denom;
            r.color = r.color * vx.baseColor;
        }
        return true;
    }

    case RTB_BSDF_PLASTIC:
        return sample_plastic(vx, out, u0, u1, u2, r);

    case RTB_BSDF_ROUGH_PLASTIC:
    {
        // RoughPlasticBSDF::Sample, RoughPlasticBSDF.cpp:18-88
        const float NdotV = out.z;
        if (NdotV < RTB_COS_EPSILON) return false;
        const float roughness = vx.roughness;
        if (roughness < RTB_SPECULAR_ROUGHNESS_THRESHOLD) return sample_plastic(vx, out, u0, u1, u2, r);
        const float ior = vx.IoR;
        const float Fi = fresnel_dielectric(NdotV, ior);
        const float specularWeight = Fi;
        const float diffuseWeight = (1.0f - Fi) * color_max(vx.baseColor);
        const float specularProbability = specularWeight / (specularWeight + diffuseWeight);
        const float diffuseProbability = 1.0f - specularProbability;
        const bool specular = u2 < specularProbability;
        if (specular)
        {
            const Microfacet microfacet(roughness * roughness);
            const V3 m = microfacet.Sample(u0, u1);
            r.incomingDir = neg(reflect3(out, m));
            const float NdotL = r.incomingDir.z;
            const float VdotH = dot3(m, out);
            if (NdotL < RTB_COS_EPSILON || VdotH < RTB_COS_EPSILON) return false;
            const float pdf = microfacet.Pdf(m);
            const float D = microfacet.D(m);
            const float G = microfacet.G(NdotV, NdotL);
            const float F = fresnel_dielectric(VdotH, vx.IoR);
            r.pdf = pdf / (4.0f * VdotH) * specularProbability;
            r.color = v4s(VdotH * F * G * D / (pdf * NdotV * specularProbability));
            r.event = EV_GLOSSY_REFLECTION;
        }
        else
        {
            r.incomingDir = hemisphere_cos(u0, u1);
            const float NdotL = r.incomingDir.z;
            r.pdf = r.incomingDir.z * RTB_INV_PI * diffuseProbability;
            const float Fo = fresnel_dielectric(NdotL, ior);
            r.color = vx.baseColor * ((1.0f - Fi) * (1.0f - Fo) / diffuseProbability);
            r.event = EV_DIFFUSE_REFLECTION;
        }
        return true;
    }

    default:   // RTB_BSDF_NULL: NullBSDF::Sample returns false
        return false;
    }
}

RTB_HD V4 eval_plastic(const Vertex& vx, const V3& out, const V3& in, float& pdfW)
{
    // PlasticBSDF::Evaluate, PlasticBSDF.cpp:60-97
    const float NdotV = out.z;
    const float NdotL = -in.z;
    if (NdotV < RTB_COS_EPSILON || NdotL < RTB_COS_EPSILON) return v4s(0.0f);
    const float ior = vx.IoR;
    const float Fi = fresnel_dielectric(NdotV, ior);
    const float Fo = fresnel_dielectric(NdotL, ior);
    const float specularWeight = Fi;
    const float diffuseWeight = (1.0f - Fi) * color_max(vx.baseColor);
    const float specularProbability = specularWeight / (specularWeight + diffuseWeight);
    const float diffuseProbability = 1.0f - specularProbability;
    pdfW = NdotL * RTB_INV_PI * diffuseProbability;
    return vx.baseColor * (NdotL * RTB_INV_PI * (1.0f - Fi) * (1.0f - Fo));
}

// Material::Evaluate -> BSDF::Evaluate.  `in` is the incoming light direction in local space (pointing towards the
// surface); pdfW is written only on the paths where the reference writes it (the caller presets it like the
// reference's uninitialised local would not matter: every path that returns non-zero sets it).
RTB_HD V4 bsdf_evaluate(const Vertex& vx, const V3& out, const V3& in, float& pdfW)
{
    switch (vx.bsdf)
    {
    case RTB_BSDF_DIFFUSE:
    {
        // DiffuseBSDF::Evaluate, DiffuseBSDF.cpp:31-54
        const float NdotV = out.z, NdotL = -in.z;
        if (NdotV > RTB_COS_EPSILON && NdotL > RTB_COS_EPSILON)
        {
            pdfW = NdotL * RTB_INV_PI;
            return vx.baseColor * v4s(NdotL * RTB_INV_PI);
        }
        return v4s(0.0f);
    }
    case RTB_BSDF_ROUGH_DIFFUSE:
    {
        // RoughDiffuseBSDF::Evaluate, RoughDiffuseBSDF.cpp:48-73
        const float NdotV = out.z, NdotL = -in.z;
        if (NdotV > RTB_COS_EPSILON && NdotL > RTB_COS_EPSILON)
        {
            pdfW = NdotL * RTB_INV_PI;
            const float LdotV = s_max(0.0f, dot3(out, neg(in)));
            const float value = NdotL * RTB_INV_PI * oren_nayar(NdotL, NdotV, LdotV, vx.roughness);
            return vx.baseColor * value;
        }
        return v4s(0.0f);
    }
    case RTB_BSDF_ROUGH_METAL:
    {
        // RoughMetalBSDF::Evaluate, RoughMetalBSDF.cpp:62-107
        const float roughness = vx.roughness;
        if (roughness < RTB_SPECULAR_ROUGHNESS_THRESHOLD) { pdfW = 0.0f; return v4s(0.0f); }
        const V3 m = normalize3(out - in);
        const float NdotV = out.z, NdotL = -in.z;
        const float VdotH = dot3(m, out);
        if (NdotV < RTB_COS_EPSILON || NdotL < RTB_COS_EPSILON || VdotH < RTB_COS_EPSILON) return v4s(0.0f);
        const Microfacet microfacet(roughness * roughness);
        const float D = microfacet.D(m);
        const float G = microfacet.G(NdotV, NdotL);
        const float F = fresnel_metal(VdotH, vx.IoR, vx.K);
        pdfW = microfacet.Pdf(m) / (4.0f * VdotH);
        return vx.baseColor * v4s(F * G * D / (4.0f * NdotV));
    }
    case RTB_BSDF_ROUGH_DIELECTRIC:
    {
        // RoughDielectricBSDF::Evaluate, RoughDielectricBSDF.cpp:112-193
        const float NdotV = out.z, NdotL = -in.z;
        if (s_abs(NdotV) < RTB_COS_EPSILON || s_abs(NdotL) < RTB_COS_EPSILON) return v4s(0.0f);
        const float roughness = vx.roughness;
        if (roughness < RTB_SPECULAR_ROUGHNESS_THRESHOLD) return v4s(0.0f);
        const float ior = vx.IoR;
        const float eta = NdotV < 0.0f ? ior : 1.0f / ior;
        const bool reflection = NdotV * NdotL >= 0.0f;
        V3 m;
        if (reflection) m = out - in;
        else m = eta * out - in;
        const float sg = m.z > 0.0f ? 1.0f : (m.z < 0.0f ? -1.0f : 0.0f);   // Signum, Math.h:147-162
        m = m * sg;
        m = normalize3(m);
        if (s_abs(m.z) < RTB_COS_EPSILON) return v4s(0.0f);
        const float VdotH = dot3(m, out);
        const float LdotH = dot3(m, neg(in));
        const Microfacet microfacet(roughness * roughness);
        const float F = fresnel_dielectric(VdotH, ior);
        const float D = microfacet.D(m);
        const float G = microfacet.G(NdotV, NdotL);
        if (reflection)
        {
            pdfW = F * microfacet.Pdf(m) / (4.0f * s_abs(VdotH));
            return v4s(F * G * D / (4.0f * s_abs(NdotV)));
        }
        const float denom = s_sqr(eta * VdotH + LdotH);
        pdfW = (1.0f - F) * microfacet.Pdf(m) * s_abs(LdotH) / denom;
        return v4s(s_abs(VdotH * LdotH) * (1.0f - F) * G * D / (denom * s_abs(NdotV)));
    }
    case RTB_BSDF_PLASTIC:
        return eval_plastic(vx, out, in, pdfW);

    case RTB_BSDF_ROUGH_PLASTIC:
    {
        // RoughPlasticBSDF::Evaluate, RoughPlasticBSDF.cpp:90-158
        const float roughness = vx.roughness;
        if (roughness < RTB_SPECULAR_ROUGHNESS_THRESHOLD) return eval_plastic(vx, out, in, pdfW);
        const float NdotV = out.z, NdotL = -in.z;
        if (NdotV < RTB_COS_EPSILON || NdotL < RTB_COS_EPSILON) return v4s(0.0f);
        const float ior = vx.IoR;
        const float Fi = fresnel_dielectric(NdotV, ior);
        const float Fo = fresnel_dielectric(NdotL, ior);
        const float specularWeight = Fi;
        const float diffuseWeight = (1.0f - Fi) * color_max(vx.baseColor);
        const float specularProbability = specularWeight / (specularWeight + diffuseWeight);
        const float diffuseProbability = 1.0f - specularProbability;
        const float diffusePdf = NdotL * RTB_INV_PI;
        float specularPdf = 0.0f;
        const V4 diffuseTerm = vx.baseColor * (NdotL * RTB_INV_PI * (1.0f - Fi) * (1.0f - Fo));
        V4 specularTerm = v4s(0.0f);
        {
            const V3 m = normalize3(out - in);
            const float VdotH = dot3(m, out);
            if (VdotH >= RTB_COS_EPSILON)
            {
                const Microfacet microfacet(roughness * roughness);
                const float D = microfacet.D(m);
                const float G = microfacet.G(NdotV, NdotL);
                const float F = fresnel_dielectric(VdotH, vx.IoR);
                specularPdf = microfacet.Pdf(m) / (4.0f * VdotH);
                specularTerm = v4s(F * G * D / (4.0f * NdotV));
            }
        }
        pdfW = diffusePdf * diffuseProbability + specularPdf * specularProbability;
        return diffuseTerm + specularTerm;
    }
    default:
        // Null / Metal / Dielectric: Dirac delta or nothing -> zero, pdf 0 (MetalBSDF.cpp:35-54, DielectricBSDF.cpp:105-124)
        pdfW = 0.0f;
        return v4s(0.0f);
    }
}

// ---- shapes as emitters: IShape::Sample(ref, u) / Pdf ---------------------------------------------------------------

struct ShapeSample { V3 direction; float distance, pdf, cosAtSurface; };

RTB_HD float shape_surface_area(const RtbShape& s)
{
    if (s.type == RTB_SHAPE_SPHERE) return 4.0f * RTB_PI * s_sqr(s.size[0]);                                  // SphereShape.cpp:24-27
    if (s.type == RTB_SHAPE_BOX) return 8.0f * (s.size[0] * (s.size[1] + s.size[2]) + s.size[1] * s.size[2]);   // BoxShape.cpp:113-116
    return 4.0f * s.size[0] * s.size[1];                                                                        // RectShape.cpp:27-30
}

RTB_HD bool shape_sample(const RtbShape& s, const V3& ref, float u0, float u1, float u2, ShapeSample& result)
{
    if (s.type == RTB_SHAPE_SPHERE)
    {
        // SphereShape::Sample(ref, u, result), SphereShape.cpp:65-108
        const float radius = s.size[0];
        const V3 centerDir = neg(ref);
        const float centerDistSqr = sqr_length3(centerDir);
        const float centerDist = sqrtf(centerDistSqr);
        if (centerDistSqr < s_sqr(radius)) return false;
        const float phi = RTB_2PI * u1;
        float sinPhi, cosPhi;
        sin_cos(phi, sinPhi, cosPhi);
        const float sinThetaMaxSqr = s_sqr(radius) / centerDistSqr;
        const float cosThetaMax = sqrtf(1.0f - s_clamp(sinThetaMaxSqr, 0.0f, 1.0f));
        const float cosTheta = s_lerp(cosThetaMax, 1.0f, u0);
        const float sinThetaSqr = 1.0f - s_sqr(cosTheta);
        const float sinTheta = sqrtf(sinThetaSqr);
        const V3 w = centerDir / centerDist;
        V3 tangent, bitangent;
        build_orthonormal_basis(w, tangent, bitangent);
        V3 d = (tangent * cosPhi + bitangent * sinPhi) * sinTheta + w * cosTheta;
        result.direction = normalize3(d);
        result.distance = centerDist * cosTheta - sqrtf(s_max(0.0f, s_sqr(radius) - centerDistSqr * sinThetaSqr));
        result.cosAtSurface = cosTheta;
        result.pdf = cosThetaMax > 0.999999f ? FLT_MAX : sphere_cap_pdf(cosThetaMax);
        return true;
    }

    // area sampling + IShape::Sample(ref, u, result), Shape.cpp:65-91
    V3 position, normal;
    if (s.type == RTB_SHAPE_BOX)
    {
        // BoxShape::Sample(u), BoxShape.cpp:127-178
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
        // RectShape::Sample(u), RectShape.cpp:51-64: size * (2 * u - 1), lanes z/w of the Float2 vectors are zero
        position = v3(s.size[0] * (2.0f * u0 - 1.0f), s.size[1] * (2.0f * u1 - 1.0f), 0.0f * (2.0f * 0.0f - 1.0f));
        normal = v3(0.0f, 0.0f, 1.0f);
    }
    V3 dir = ref - position;
    const float sqrDistance = sqr_length3(dir);
    if (sqrDistance > FLT_EPSILON * FLT_EPSILON)
    {
        const float distance = sqrtf(sqrDistance);
        dir = dir / distance;
        const float cosNormalDir = dot3(normal, dir);
        if (cosNormalDir > FLT_EPSILON)
        {
            const float invArea = 1.0f / shape_surface_area(s);
            result.pdf = invArea * sqrDistance / cosNormalDir;
            result.distance = distance;
            result.cosAtSurface = cosNormalDir;
            result.direction = neg(dir);
            return true;
        }
    }
    return false;
}

// IShape::Pdf / SphereShape::Pdf, Shape.cpp:93-99, SphereShape.cpp:110-125
RTB_HD float shape_pdf(const RtbShape& s, const V3& ref, const V3& point)
{
    if (s.type == RTB_SHAPE_SPHERE)
    {
        const float radius = s.size[0];
        const V3 rayDir = normalize3(point - ref);
        const V3 centerDir = neg(ref);
        const float centerDistSqr = sqr_length3(centerDir);
        const V3 normal = normalize3(point);
        const float cosAtLight = s_max(0.0f, dot3(neg(rayDir), normal));
        const float sinThetaMaxSqr = s_clamp(s_sqr(radius) / centerDistSqr, 0.0f, 1.0f);
        const float cosThetaMax = sqrtf(1.0f - sinThetaMaxSqr);
        const float pdfW = sphere_cap_pdf(cosThetaMax);
        return pdfW * cosAtLight / sqr_length3(point - ref);
    }
    return 1.0f / shape_surface_area(s);
}

// ---- lights -------------------------------------------------------------------------------------------------------------

struct IlluminateResult { V3 directionToLight; float distance, directPdfW, cosAtLight; };

// BackgroundLight::GetBackgroundColor, BackgroundLight.cpp:45-61: colour * max(0, environment map at the spherical
// coordinates of `dir`)
template <bool kTex>
RTB_HD V4 background_color(const SceneView& S, const DLight& light, const V3& dir)
{
    if (!kTex || light.textureIndex == RTB_NO_TEXTURE) return light.color;
    float u, v;
    cartesian_to_spherical(dir, u, v);
    const V4 t = texture_evaluate(S, light.textureIndex, u, v);
    return light.color * v4(sse_max(0.0f, t.x), sse_max(0.0f, t.y), sse_max(0.0f, t.z), sse_max(0.0f, t.w));
}

// ILight::Illuminate for every light type; returns the radiance (RayColor, 4 lanes)
template <bool kTex = true>
RTB_HD V4 light_illuminate(const SceneView& S, const DLight& light, const Vertex& vx, float u0, float u1, float u2, IlluminateResult& out)
{
    out.directionToLight = v3s(0.0f); out.distance = -1.0f; out.directPdfW = -1.0f; out.cosAtLight = -1.0f;
    switch (light.type)
    {
    case RTB_LIGHT_AREA:
    {
        // AreaLight::Illuminate (solid-angle branch), AreaLight.cpp:43-77
        const V3 ref = transform_point(light.invTransform, vx.position);
        ShapeSample ss;
        if (!shape_sample(S.shapes[light.shapeIndex], ref, u0, u1, u2, ss)) return v4s(0.0f);
        out.directionToLight = transform_vector(light.transform, ss.direction);
        out.distance = ss.distance;
        out.cosAtLight = ss.cosAtSurface;
        out.directPdfW = ss.pdf;
        return light.color;
    }
    case RTB_LIGHT_BACKGROUND:
    {
        // BackgroundLight::Illuminate, BackgroundLight.cpp:63-76
        const V3 randomDirLocalSpace = sample_hemisphere(u0, u1);
        out.directionToLight = local_to_world(vx, randomDirLocalSpace);
        out.directPdfW = uniform_hemisphere_pdf();
        out.distance = FLT_MAX;
        out.cosAtLight = 1.0f;
        return background_color<kTex>(S, light, out.directionToLight);      // world-space direction ("TODO include light rotation")
    }
    case RTB_LIGHT_DIRECTIONAL:
    {
        // DirectionalLight::SampleDirection + Illuminate, DirectionalLight.cpp:48-91
        V3 d;
        if (light.isDelta)
        {
            out.directPdfW = 1.0f;
            d = v3(0.0f, 0.0f, 1.0f);
        }
        else
        {
            out.directPdfW = sphere_cap_pdf(light.cosAngle);
            const float phi = RTB_2PI * u1;
            float sinPhi, cosPhi;
            sin_cos(phi, sinPhi, cosPhi);
            const float cosTheta = s_lerp(light.cosAngle, 1.0f, u0);
            const float sinThetaSqr = 1.0f - s_sqr(cosTheta);
            const float sinTheta = sqrtf(sinThetaSqr);
            d = normalize3(v3(sinTheta * sinPhi, sinTheta * cosPhi, cosTheta));
        }
        out.directionToLight = transform_vector_neg(light.transform, d);
        out.cosAtLight = 1.0f;
        out.distance = FLT_MAX;
        return light.color;
    }
    case RTB_LIGHT_POINT:
    case RTB_LIGHT_SPOT:
    {
        // PointLight::Illuminate, PointLight.cpp:35-49; SpotLight::Illuminate, SpotLight.cpp:41-61
        V3 d = xyz(light.transform.r3) - vx.position;
        const float sqrDistance = sqr_length3(d);
        out.directPdfW = sqrDistance;
        out.distance = sqrtf(sqrDistance);
        d = d / out.distance;
        out.directionToLight = d;
        out.cosAtLight = 1.0f;
        if (light.type == RTB_LIGHT_SPOT)
        {
            const float angle = dot3(d, v3(0.0f - 0.0f, 0.0f - 0.0f, 0.0f - 1.0f));
            if (angle < light.cosAngle) return v4s(0.0f);
        }
        return light.color;
    }
    default:
        return v4s(0.0f);
    }
}

// ILight::GetRadiance for a ray that hit the light (area) or escaped (background / directional).
// `lightRay` is the ray in light space, `lightSpaceHit` the hit point in light space.
template <bool kTex = true>
RTB_HD V4 light_radiance(const SceneView& S, const DLight& light, const Ray& lightRay, const V3& lightSpaceHit, float cosAtLight, float& outDirectPdfA)
{
    switch (light.type)
    {
    case RTB_LIGHT_AREA:
        // AreaLight::GetRadiance, AreaLight.cpp:109-147
        if (cosAtLight < RTB_EPSILON) return v4s(0.0f);
        outDirectPdfA = shape_pdf(S.shapes[light.shapeIndex], lightRay.origin, lightSpaceHit);
        return light.color;
    case RTB_LIGHT_BACKGROUND:
        // BackgroundLight::GetRadiance, BackgroundLight.cpp:78-90
        outDirectPdfA = uniform_hemisphere_pdf();
        return background_color<kTex>(S, light, lightRay.dir);               // param.ray = the light-space ray
    case RTB_LIGHT_DIRECTIONAL:
        // DirectionalLight::GetRadiance, DirectionalLight.cpp:93-118
        if (light.isDelta) return v4s(0.0f);
        if (dot3(lightRay.dir, v3(0.0f, 0.0f, 1.0f)) > -light.cosAngle) return v4s(0.0f);
        outDirectPdfA = sphere_cap_pdf(light.cosAngle);
        return light.color;
    default:
        return v4s(0.0f);
    }
}

} // namespace rtb
