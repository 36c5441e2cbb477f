// rtb_trace.h — two-level BVH ray/scene intersection: closest hit (Scene::Traverse) and any hit
// (Scene::Traverse_Shadow).  Per-ray device functions; the wavefront kernels in rtb_kernels.cu call them with one
// ray per thread.
//
// Mirrors, operation for operation:
//   Intersect_BoxRay / Intersect_BoxRay_TwoSided / Intersect_TriangleRay   Core/Math/Geometry.h:57-189 (SSE branch)
//   GenericTraverse / GenericTraverse_Shadow                               Core/Traversal/Traversal_Single.h:15-179
//   Scene::Traverse / Traverse_Shadow / Traverse_Object(_Shadow) / Traverse_Leaf(_Shadow)  Core/Scene/Scene.cpp:128-261
//   IShape::Traverse / Traverse_Shadow                                     Core/Shapes/Shape.cpp:19-57
//   SphereShape / BoxShape / RectShape ::Intersect                         SphereShape.cpp:29-46, BoxShape.cpp:118-125, RectShape.cpp:32-49
//   MeshShape::Traverse_Leaf / Traverse_Leaf_Shadow                        Core/Shapes/MeshShape.cpp:134-207
//   LightSceneObject::Traverse / Traverse_Shadow, AreaLight::TestRayHit    SceneObject_Light.cpp:27-53, AreaLight.cpp:30-41
#pragma once

#include "rtb_scene.h"

namespace rtb {

// HitPoint (Core/Traversal/HitPoint.h:14-51)
struct Hit
{
    uint32_t objectId;
    uint32_t subObjectId;
    float    distance;
    float    u, v;
};

struct TraceCounters { uint32_t boxTests, triTests; };

#define RTB_STACK_SIZE 96      // BVH::MaxDepth is 128 (BVH.h:20); upload verifies the real depth of every tree fits

// Intersect_BoxRay, Geometry.h:57-86
RTB_HD bool box_ray(const Ray& ray, float minx, float miny, float minz, float maxx, float maxy, float maxz, float& outDistance)
{
    const float t1x = fma_(minx, ray.invDir.x, -ray.originDivDir.x);
    const float t1y = fma_(miny, ray.invDir.y, -ray.originDivDir.y);
    const float t1z = fma_(minz, ray.invDir.z, -ray.originDivDir.z);
    const float t2x = fma_(maxx, ray.invDir.x, -ray.originDivDir.x);
    const float t2y = fma_(maxy, ray.invDir.y, -ray.originDivDir.y);
    const float t2z = fma_(maxz, ray.invDir.z, -ray.originDivDir.z);
    const float lminx = sse_min(t1x, t2x), lminy = sse_min(t1y, t2y), lminz = sse_min(t1z, t2z);
    const float lmaxx = sse_max(t1x, t2x), lmaxy = sse_max(t1y, t2y), lmaxz = sse_max(t1z, t2z);
    const float tNear = sse_max(lminx, sse_max(lminy, lminz));
    const float tFar = sse_min(lmaxx, sse_min(lmaxy, lmaxz));
    outDistance = tNear;
    return (tFar >= tNear) && (tFar >= 0.0f);
}

// Intersect_BoxRay_TwoSided, Geometry.h:100-130
RTB_HD bool box_ray_two_sided(const Ray& ray, const float* size, float& outNear, float& outFar)
{
    // Box(-mSize, mSize): operator-() is 0 - v
    const float minx = 0.0f - size[0], miny = 0.0f - size[1], minz = 0.0f - size[2];
    const float t1x = fma_(minx, ray.invDir.x, -ray.originDivDir.x);
    const float t1y = fma_(miny, ray.invDir.y, -ray.originDivDir.y);
    const float t1z = fma_(minz, ray.invDir.z, -ray.originDivDir.z);
    const float t2x = fma_(size[0], ray.invDir.x, -ray.originDivDir.x);
    const float t2y = fma_(size[1], ray.invDir.y, -ray.originDivDir.y);
    const float t2z = fma_(size[2], ray.invDir.z, -ray.originDivDir.z);
    const float lminx = sse_min(t1x, t2x), lminy = sse_min(t1y, t2y), lminz = sse_min(t1z, t2z);
    const float lmaxx = sse_max(t1x, t2x), lmaxy = sse_max(t1y, t2y), lmaxz = sse_max(t1z, t2z);
    outNear = sse_max(lminx, sse_max(lminy, lminz));
    outFar = sse_min(lmaxx, sse_min(lmaxy, lmaxz));
    return outNear < outFar;
}

// Intersect_TriangleRay (SSE branch), Geometry.h:132-167
RTB_HD bool triangle_ray(const Ray& ray, const V3& v0, const V3& e1, const V3& e2, float& outU, float& outV, float& outT)
{
    const V3 tvec = ray.origin - v0;
    const V3 pvec = cross3(ray.dir, e2);
    const V3 qvec = cross3(tvec, e1);
    const float det = dot3(e1, pvec);
    const float u = dot3(tvec, pvec);
    const float v = dot3(ray.dir, qvec);
    const float t = dot3(e2, qvec);
    const float upv = (u + v) / det;
    outU = u / det;
    outV = v / det;
    outT = t / det;
    // mask 0xE: (u+v > 1) false, v > 0, t > 0, u > 0
    return !(upv > 1.0f) && (outV > 0.0f) && (outT > 0.0f) && (outU > 0.0f);
}

struct ShapeIntersection { float nearDist, farDist; uint32_t subObjectId; };

// SphereShape::Intersect (double precision quadratic), SphereShape.cpp:29-46
RTB_HD bool sphere_intersect(const Ray& ray, float radius, ShapeIntersection& out)
{
    const double radiusD = (double)radius;
    const double v = (double)dot3(ray.dir, neg(ray.origin));
    const double det = radiusD * radiusD - (double)sqr_length3(ray.origin) + v * v;
    if (det <= 0.0) return false;
    const double sqrtDet = sqrt(det);
    out.nearDist = (float)(v - sqrtDet);
    out.farDist = (float)(v + sqrtDet);
    out.subObjectId = 0;
    return out.farDist > out.nearDist;
}

// RectShape::Intersect, RectShape.cpp:32-49
RTB_HD bool rect_intersect(const Ray& ray, const float* size, ShapeIntersection& out)
{
    const float t = -ray.origin.z * ray.invDir.z;
    if (t > FLT_EPSILON)
    {
        const float px = fma_(ray.dir.x, t, ray.origin.x);
        const float py = fma_(ray.dir.y, t, ray.origin.y);
        if (bit_abs(px) < size[0] && bit_abs(py) < size[1])
        {
            out.nearDist = t;
            out.farDist = t;
            return true;
        }
    }
    return false;
}

RTB_HD bool shape_intersect(uint32_t shapeType, const float* size, const Ray& ray, ShapeIntersection& out)
{
    out.subObjectId = 0xFFFFFFFFu;   // ShapeIntersection default, Shape.h:25-30
    if (shapeType == RTB_SHAPE_SPHERE) return sphere_intersect(ray, size[0], out);
    if (shapeType == RTB_SHAPE_BOX) { out.subObjectId = 0; return box_ray_two_sided(ray, size, out.nearDist, out.farDist); }
    if (shapeType == RTB_SHAPE_RECT) return rect_intersect(ray, size, out);
    return false;
}

RTB_HD void load_tri(const DTri* tris, uint32_t index, V3& v0, V3& e1, V3& e2)
{
#if defined(__CUDA_ARCH__)
    const float4* p = reinterpret_cast<const float4*>(tris + index);
    const float4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
    v0 = v3(a.x, a.y, a.z); e1 = v3(b.x, b.y, b.z); e2 = v3(c.x, c.y, c.z);
#else
    const DTri& t = tris[index];
    v0 = v3(t.v0x, t.v0y, t.v0z); e1 = v3(t.e1x, t.e1y, t.e1z); e2 = v3(t.e2x, t.e2y, t.e2z);
#endif
}

RTB_HD DNode load_node(const DNode* nodes, uint32_t index)
{
#if defined(__CUDA_ARCH__)
    const float4* p = reinterpret_cast<const float4*>(nodes + index);
    const float4 a = __ldg(p), b = __ldg(p + 1);
    DNode n;
    n.minx = a.x; n.miny = a.y; n.minz = a.z; n.childIndex = __float_as_uint(a.w);
    n.maxx = b.x; n.maxy = b.y; n.maxz = b.z; n.leafInfo = __float_as_uint(b.w);
    return n;
#else
    return nodes[index];
#endif
}

#define RTB_NUM_LEAVES(info) ((info) & 0x3FFFFFFFu)

// MeshShape::Traverse -> GenericTraverse<MeshShape> + MeshShape::Traverse_Leaf
template <bool kCount>
RTB_HD void mesh_traverse(const DMesh& mesh, const Ray& ray, uint32_t objectID, Hit& hit, uint32_t* stack, TraceCounters* cnt)
{
    if (mesh.numNodes == 0) return;
    uint32_t stackSize = 0;
    DNode node = load_node(mesh.nodes, 0);
    for (;;)
    {
        const uint32_t numLeaves = RTB_NUM_LEAVES(node.leafInfo);
        if (numLeaves != 0)
        {
            if (kCount) cnt->triTests += numLeaves;
            for (uint32_t i = 0; i < numLeaves; ++i)
            {
                const uint32_t triangleIndex = node.childIndex + i;
                V3 v0, e1, e2;
                load_tri(mesh.tris, triangleIndex, v0, e1, e2);
                float u, v, distance;
                if (triangle_ray(ray, v0, e1, e2, u, v, distance))
                {
                    if (distance < hit.distance)
                    {
                        hit.distance = distance;
                        hit.subObjectId = triangleIndex;
                        hit.objectId = objectID;
                        hit.u = u;
                        hit.v = v;
                    }
                }
            }
        }
        else
        {
            const uint32_t childA = node.childIndex;
            const DNode a = load_node(mesh.nodes, childA);
            const DNode b = load_node(mesh.nodes, childA + 1);
            float distanceA, distanceB;
            bool hitA = box_ray(ray, a.minx, a.miny, a.minz, a.maxx, a.maxy, a.maxz, distanceA);
            bool hitB = box_ray(ray, b.minx, b.miny, b.minz, b.maxx, b.maxy, b.maxz, distanceB);
            hitA &= (distanceA < hit.distance);
            hitB &= (distanceB < hit.distance);
            if (kCount) cnt->boxTests += 2;
            if (hitA && hitB)
            {
                // visit the nearer child first (Traversal_Single.h:65-75)
                if (distanceB < distanceA) { node = b; stack[stackSize++] = childA; }
                else                       { node = a; stack[stackSize++] = childA + 1; }
                continue;
            }
            if (hitA) { node = a; continue; }
            if (hitB) { node = b; continue; }
        }
        if (stackSize == 0) break;
        node = load_node(mesh.nodes, stack[--stackSize]);
    }
}

// MeshShape::Traverse_Shadow -> GenericTraverse_Shadow<MeshShape> + MeshShape::Traverse_Leaf_Shadow
template <bool kCount>
RTB_HD bool mesh_traverse_shadow(const DMesh& mesh, const Ray& ray, float maxDistance, uint32_t* stack, TraceCounters* cnt)
{
    if (mesh.numNodes == 0) return false;
    uint32_t stackSize = 0;
    DNode node = load_node(mesh.nodes, 0);
    for (;;)
    {
        const uint32_t numLeaves = RTB_NUM_LEAVES(node.leafInfo);
        if (numLeaves != 0)
        {
            if (kCount) cnt->triTests += numLeaves;
            for (uint32_t i = 0; i < numLeaves; ++i)
            {
                V3 v0, e1, e2;
                load_tri(mesh.tris, node.childIndex + i, v0, e1, e2);
                float u, v, distance;
                if (triangle_ray(ray, v0, e1, e2, u, v, distance))
                {
                    if (distance < maxDistance) return true;
                }
            }
        }
        else
        {
            const uint32_t childA = node.childIndex;
            const DNode a = load_node(mesh.nodes, childA);
            const DNode b = load_node(mesh.nodes, childA + 1);
            float distanceA, distanceB;
            bool hitA = box_ray(ray, a.minx, a.miny, a.minz, a.maxx, a.maxy, a.maxz, distanceA);
            bool hitB = box_ray(ray, b.minx, b.miny, b.minz, b.maxx, b.maxy, b.maxz, distanceB);
            hitA &= (distanceA < maxDistance);
            hitB &= (distanceB < maxDistance);
            if (kCount) cnt->boxTests += 2;
            // no distance ordering for shadow rays (Traversal_Single.h:151-156)
            if (hitA && hitB) { node = a; stack[stackSize++] = childA + 1; continue; }
            if (hitA) { node = a; continue; }
            if (hitB) { node = b; continue; }
        }
        if (stackSize == 0) break;
        node = load_node(mesh.nodes, stack[--stackSize]);
    }
    return false;
}

// Scene::Traverse_Object, Scene.cpp:128-145
template <bool kCount>
RTB_HD void traverse_object(const SceneView& S, const Ray& worldRay, uint32_t objectID, Hit& hit, uint32_t* stack, TraceCounters* cnt)
{
    const DObjectTrace& obj = S.objTrace[objectID];
    const Ray ray = transform_ray_unsafe(obj.invTransform, worldRay);
    if (obj.kind == RTB_OBJECT_SHAPE)
    {
        if (obj.shapeType == RTB_SHAPE_MESH)
        {
            mesh_traverse<kCount>(S.meshes[obj.meshIndex], ray, objectID, hit, stack, cnt);
            return;
        }
        // IShape::Traverse, Shape.cpp:19-45
        ShapeIntersection isect;
        if (shape_intersect(obj.shapeType, obj.size, ray, isect))
        {
            if (isect.nearDist > 0.0f && isect.nearDist < hit.distance)
            {
                hit.distance = isect.nearDist; hit.objectId = objectID; hit.subObjectId = isect.subObjectId;
                return;
            }
            if (isect.farDist > 0.0f && isect.farDist < hit.distance)
            {
                hit.distance = isect.farDist; hit.objectId = objectID; hit.subObjectId = isect.subObjectId;
                return;
            }
        }
    }
    else
    {
        // LightSceneObject::Traverse + AreaLight::TestRayHit (point/spot lights never report a hit)
        if (obj.shapeType == RTB_NO_SHAPE) return;
        ShapeIntersection isect;
        if (shape_intersect(obj.shapeType, obj.size, ray, isect))
        {
            const float lightDistance = isect.nearDist;
            if (lightDistance > 0.0f && lightDistance < hit.distance)
            {
                hit.distance = lightDistance; hit.objectId = objectID; hit.subObjectId = RTB_LIGHT_OBJECT;
            }
        }
    }
}

// Scene::Traverse_Object_Shadow, Scene.cpp:147-165
template <bool kCount>
RTB_HD bool traverse_object_shadow(const SceneView& S, const Ray& worldRay, uint32_t objectID, float maxDistance, uint32_t* stack, TraceCounters* cnt)
{
    const DObjectTrace& obj = S.objTrace[objectID];
    const Ray ray = transform_ray_unsafe(obj.invTransform, worldRay);
    if (obj.kind == RTB_OBJECT_SHAPE)
    {
        if (obj.shapeType == RTB_SHAPE_MESH)
        {
            return mesh_traverse_shadow<kCount>(S.meshes[obj.meshIndex], ray, maxDistance, stack, cnt);
        }
        // IShape::Traverse_Shadow, Shape.cpp:47-57
        ShapeIntersection isect;
        if (!shape_intersect(obj.shapeType, obj.size, ray, isect)) return false;
        return isect.farDist > 0.0f && isect.nearDist < maxDistance;
    }
    // LightSceneObject::Traverse_Shadow, SceneObject_Light.cpp:40-53
    if (obj.shapeType == RTB_NO_SHAPE) return false;
    ShapeIntersection isect;
    if (shape_intersect(obj.shapeType, obj.size, ray, isect))
    {
        if (isect.nearDist < maxDistance) return true;
    }
    return false;
}

// Scene::Traverse, Scene.cpp:219-243.  `hit` must be preset (objectId = RTB_INVALID_OBJECT, distance = +inf).
// Note: the top-level box tests use ray.originDivDir exactly as stored - the integrator offsets ray.origin AFTER
// constructing the ray without refreshing originDivDir (PathTracerMIS.cpp:392-393) and that stale value is part of
// the reference's behaviour; Traverse_Object rebuilds the ray in object space.
template <bool kCount>
RTB_HD void scene_traverse(const SceneView& S, const Ray& ray, Hit& hit, TraceCounters* cnt)
{
    if (S.numObjects == 0) return;
    uint32_t stack[RTB_STACK_SIZE];
    if (S.numObjects == 1)
    {
        traverse_object<kCount>(S, ray, 0, hit, stack, cnt);
        return;
    }
    if (S.numTlasNodes == 0) return;
    // GenericTraverse<Scene>: the object-level stack lives in the lower part, mesh traversal uses the part above it
    uint32_t stackSize = 0;
    DNode node = load_node(S.tlas, 0);
    for (;;)
    {
        const uint32_t numLeaves = RTB_NUM_LEAVES(node.leafInfo);
        if (numLeaves != 0)
        {
            for (uint32_t i = 0; i < numLeaves; ++i)
            {
                traverse_object<kCount>(S, ray, node.childIndex + i, hit, stack + stackSize, cnt);
            }
        }
        else
        {
            const uint32_t childA = node.childIndex;
            const DNode a = load_node(S.tlas, childA);
            const DNode b = load_node(S.tlas, childA + 1);
            float distanceA, distanceB;
            bool hitA = box_ray(ray, a.minx, a.miny, a.minz, a.maxx, a.maxy, a.maxz, distanceA);
            bool hitB = box_ray(ray, b.minx, b.miny, b.minz, b.maxx, b.maxy, b.maxz, distanceB);
            hitA &= (distanceA < hit.distance);
            hitB &= (distanceB < hit.distance);
            if (kCount) cnt->boxTests += 2;
            if (hitA && hitB)
            {
                if (distanceB < distanceA) { node = b; stack[stackSize++] = childA; }
                else                       { node = a; stack[stackSize++] = childA + 1; }
                continue;
            }
            if (hitA) { node = a; continue; }
            if (hitB) { node = b; continue; }
        }
        if (stackSize == 0) break;
        node = load_node(S.tlas, stack[--stackSize]);
    }
}

// Scene::Traverse_Shadow, Scene.cpp:245-261
template <bool kCount>
RTB_HD bool scene_traverse_shadow(const SceneView& S, const Ray& ray, float maxDistance, TraceCounters* cnt)
{
    if (S.numObjects == 0) return false;
    uint32_t stack[RTB_STACK_SIZE];
    if (S.numObjects == 1) return traverse_object_shadow<kCount>(S, ray, 0, maxDistance, stack, cnt);
    if (S.numTlasNodes == 0) return false;
    uint32_t stackSize = 0;
    DNode node = load_node(S.tlas, 0);
    for (;;)
    {
        const uint32_t numLeaves = RTB_NUM_LEAVES(node.leafInfo);
        if (numLeaves != 0)
        {
            for (uint32_t i = 0; i < numLeaves; ++i)
            {
                if (traverse_object_shadow<kCount>(S, ray, node.childIndex + i, maxDistance, stack + stackSize, cnt)) return true;
            }
        }
        else
        {
            const uint32_t childA = node.childIndex;
            const DNode a = load_node(S.tlas, childA);
            const DNode b = load_node(S.tlas, childA + 1);
            float distanceA, distanceB;
            bool hitA = box_ray(ray, a.minx, a.miny, a.minz, a.maxx, a.maxy, a.maxz, distanceA);
            bool hitB = box_ray(ray, b.minx, b.miny, b.minz, b.maxx, b.maxy, b.maxz, distanceB);
            hitA &= (distanceA < maxDistance);
            hitB &= (distanceB < maxDistance);
            if (kCount) cnt->boxTests += 2;
            if (hitA && hitB) { node = a; stack[stackSize++] = childA + 1; continue; }
            if (hitA) { node = a; continue; }
            if (hitB) { node = b; continue; }
        }
        if (stackSize == 0) break;
        node = load_node(S.tlas, stack[--stackSize]);
    }
    return false;
}

} // namespace rtb
