/*
 * rt_restate.c — plain C restatement of the integer / traversal core of the reference hot path.
 *
 * TEST INFRASTRUCTURE ONLY (part of the oracle): nothing in the product links, imports or executes this file.
 * It shares no code with raytracer_b200/ - every function is written from the reference's algorithm and cites the
 * reference file:line it follows.  PARITY IS PINNED: tests/test_oracle_restate.py checks every output against the
 * golden vectors generated from the real reference (oracle/_ref/exact, see tests/golden/make_golden.py):
 *   sampler   GenericSampler streams                      (sampler_bn.u32 / sampler_nobn.u32)
 *   primary   Camera::GenerateRay for every pixel         (<scene>.primary)
 *   trace     Scene::Traverse / Traverse_Shadow           (<scene>.hits / .shadow_hits)
 *   debug     DebugRenderer TriangleID / Depth sum buffer (<scene>_dbgtri.sum / _dbgdepth.sum)
 *   texeval   ITexture::Evaluate (bitmap / checkerboard / noise / mix) at a list of coordinates (textures.texeval)
 *   front     Viewport::PostProcessTile with the default PostprocessParams, no dithering   (<scene>_mis.front)
 *   blockerr  Viewport::ComputeBlockError for a list of blocks                             (<scene>_mis.blockerr)
 *   passes    the host prologue of Viewport::Render under the oracle's fixed seed: Halton seed vectors and
 *             anti-aliasing offsets                                                        (<scene>.passes)
 * The full integrator (BSDFs, lights, MIS) is NOT restated here: for it the oracle is the reference itself
 * (oracle/_ref), whose sum buffers are the committed fixtures.
 *
 * Build: gcc -O2 -std=c11 -ffp-contract=off -mfma -msse4.1 (oracle/restate/build.sh).  FTZ/DAZ on, like the
 * reference (Core/Math/Math.cpp:27-34).
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <xmmintrin.h>
#include <pmmintrin.h>

#include "../../include/rtb200.h"

typedef struct { float x, y, z; } vec3;
typedef struct { vec3 o, d, inv, odd; } ray_t;           /* math::Ray, Core/Math/Ray.h:13-28 */
typedef struct { uint32_t object, sub; float t, u, v; } hit_t;   /* HitPoint, Core/Traversal/HitPoint.h:14-51 */

/* ---- scene file (RTBS, see raytracer_b200/csrc/rtb_scene_file.h for the layout; read independently here) --------- */
typedef struct
{
    uint32_t numNodes, numTris, numVerts, numMats;
    RtbBvhNode* nodes; float* tris; uint32_t* indices; float* shading; uint32_t* matIds;
} mesh_t;

typedef struct
{
    uint32_t numTlas, numObjects, numShapes, numMeshes, numMaterials, numLights, numGlobal;
    RtbBvhNode* tlas; RtbObject* objects; RtbShape* shapes; mesh_t* meshes; RtbMaterial* materials; RtbLight* lights; uint32_t* global;
    RtbCamera camera; float fov;
    /* RTBS version 2: textures */
    uint32_t numTextures;
    struct tex_s* textures;
} scene_t;

typedef struct tex_s
{
    uint32_t type, width, height, octaves, a, b, w;
    float colorA[4], colorB[4];
    float* texels;       /* width * height * 4, what Bitmap::GetPixel returns */
} tex_t;

static void die(const char* m) { fprintf(stderr, "rt_restate: %s\n", m); exit(1); }

static void* read_block(FILE* f, size_t bytes)
{
    void* p = malloc(bytes ? bytes : 1);
    if (bytes && fread(p, 1, bytes, f) != bytes) die("truncated scene file");
    return p;
}
static uint32_t read_u32(FILE* f) { uint32_t v; if (fread(&v, 4, 1, f) != 1) die("truncated scene file"); return v; }

static void load_scene(const char* path, scene_t* s)
{
    FILE* f = fopen(path, "rb");
    if (!f) die("cannot open scene");
    if (read_u32(f) != 0x53425452u) die("not an RTBS file");
    const uint32_t version = read_u32(f);
    if (version != 1 && version != 2) die("not an RTBS v1 / v2 file");
    s->numTlas = read_u32(f); s->tlas = read_block(f, (size_t)s->numTlas * sizeof(RtbBvhNode));
    s->numObjects = read_u32(f); s->objects = read_block(f, (size_t)s->numObjects * sizeof(RtbObject));
    s->numShapes = read_u32(f); s->shapes = read_block(f, (size_t)s->numShapes * sizeof(RtbShape));
    s->numMeshes = read_u32(f);
    s->meshes = calloc(s->numMeshes ? s->numMeshes : 1, sizeof(mesh_t));
    for (uint32_t i = 0; i < s->numMeshes; ++i)
    {
        mesh_t* m = &s->meshes[i];
        m->numNodes = read_u32(f); m->numTris = read_u32(f); m->numVerts = read_u32(f); m->numMats = read_u32(f);
        m->nodes = read_block(f, (size_t)m->numNodes * 32); m->tris = read_block(f, (size_t)m->numTris * 36);
        m->indices = read_block(f, (size_t)m->numTris * 16); m->shading = read_block(f, (size_t)m->numVerts * 32);
        m->matIds = read_block(f, (size_t)m->numMats * 4);
    }
    s->numMaterials = read_u32(f); s->materials = read_block(f, (size_t)s->numMaterials * sizeof(RtbMaterial));
    s->numLights = read_u32(f); s->lights = read_block(f, (size_t)s->numLights * sizeof(RtbLight));
    s->numGlobal = read_u32(f); s->global = read_block(f, (size_t)s->numGlobal * 4);
    if (fread(&s->camera, sizeof(RtbCamera), 1, f) != 1 || fread(&s->fov, 4, 1, f) != 1) die("truncated scene file");
    s->numTextures = 0; s->textures = NULL;
    if (version >= 2)
    {
        s->numTextures = read_u32(f);
        s->textures = calloc(s->numTextures ? s->numTextures : 1, sizeof(tex_t));
        for (uint32_t i = 0; i < s->numTextures; ++i)
        {
            tex_t* t = &s->textures[i];
            t->type = read_u32(f); t->width = read_u32(f); t->height = read_u32(f); t->octaves = read_u32(f);
            t->a = read_u32(f); t->b = read_u32(f); t->w = read_u32(f); (void)read_u32(f);
            if (fread(t->colorA, 4, 4, f) != 4 || fread(t->colorB, 4, 4, f) != 4) die("truncated texture");
            if (t->type == RTB_TEXTURE_BITMAP) t->texels = read_block(f, (size_t)t->width * t->height * 16);
        }
    }
    fclose(f);
}

/* ---- SSE semantics, scalar (Core/Math/Vector4ImplSSE.h) -------------------------------------------------------- */
static float minps(float a, float b) { return a < b ? a : b; }     /* _mm_min_ps: second operand on NaN, :388-391 */
static float maxps(float a, float b) { return a > b ? a : b; }     /* _mm_max_ps, :393-396 */
/* _mm_dp_ps 0x7F (:451-454): (x*x' + y*y') + (z*z' + 0) */
static float dp3(vec3 a, vec3 b) { return (a.x * b.x + a.y * b.y) + (a.z * b.z + 0.0f); }
/* Cross3 (:476-484): a.yzx * b.zxy, then fnmadd(a.zxy, b.yzx, .) */
static vec3 cross(vec3 a, vec3 b)
{
    vec3 r = { fmaf(-a.z, b.y, a.y * b.z), fmaf(-a.x, b.z, a.z * b.x), fmaf(-a.y, b.x, a.x * b.y) };
    return r;
}
static float absbits(float x) { union { float f; uint32_t u; } c; c.f = x; c.u &= 0x7fffffffu; return c.f; }

/* Matrix4::TransformPoint / TransformVector (Core/Math/Matrix4.h:110-139) */
static vec3 xform_point(const RtbMat4* m, vec3 a)
{
    vec3 t = { fmaf(a.x, m->r[0][0], m->r[3][0]), fmaf(a.x, m->r[0][1], m->r[3][1]), fmaf(a.x, m->r[0][2], m->r[3][2]) };
    t.x = fmaf(a.y, m->r[1][0], t.x); t.y = fmaf(a.y, m->r[1][1], t.y); t.z = fmaf(a.y, m->r[1][2], t.z);
    t.x = fmaf(a.z, m->r[2][0], t.x); t.y = fmaf(a.z, m->r[2][1], t.y); t.z = fmaf(a.z, m->r[2][2], t.z);
    return t;
}
static vec3 xform_vector(const RtbMat4* m, vec3 a)
{
    vec3 t = { a.x * m->r[0][0], a.x * m->r[0][1], a.x * m->r[0][2] };
    t.x = fmaf(a.y, m->r[1][0], t.x); t.y = fmaf(a.y, m->r[1][1], t.y); t.z = fmaf(a.y, m->r[1][2], t.z);
    t.x = fmaf(a.z, m->r[2][0], t.x); t.y = fmaf(a.z, m->r[2][1], t.y); t.z = fmaf(a.z, m->r[2][2], t.z);
    return t;
}

/* Ray(origin, direction) via Vector4::InvNormalized (Ray.h:23-28, Vector4Impl.h:90-98) */
static ray_t ray_make(vec3 o, vec3 dir)
{
    ray_t r;
    const float len = sqrtf(dp3(dir, dir));
    const float ix = 1.0f / dir.x, iy = 1.0f / dir.y, iz = 1.0f / dir.z, il = 1.0f / len;
    r.o = o;
    r.inv.x = len * ix; r.inv.y = len * iy; r.inv.z = len * iz;
    r.d.x = dir.x * il; r.d.y = dir.y * il; r.d.z = dir.z * il;
    r.odd.x = o.x * r.inv.x; r.odd.y = o.y * r.inv.y; r.odd.z = o.z * r.inv.z;
    return r;
}
/* Matrix4::TransformRay_Unsafe -> Ray::BuildUnsafe (Matrix4.h:245-250, Ray.h:31-39) */
static ray_t ray_to_local(const RtbMat4* inv, const ray_t* w)
{
    ray_t r;
    r.o = xform_point(inv, w->o);
    r.d = xform_vector(inv, w->d);
    r.inv.x = 1.0f / r.d.x; r.inv.y = 1.0f / r.d.y; r.inv.z = 1.0f / r.d.z;
    r.odd.x = r.o.x * r.inv.x; r.odd.y = r.o.y * r.inv.y; r.odd.z = r.o.z * r.inv.z;
    return r;
}

/* ---- intersectors (Core/Math/Geometry.h) ------------------------------------------------------------------------- */
/* Intersect_BoxRay, :57-86: slab distances with one fmsub per plane, near = max of mins, far = min of maxes */
static int box_ray(const ray_t* r, const float* bmin, const float* bmax, float* outNear, float* outFar)
{
    const float a0 = fmaf(bmin[0], r->inv.x, -r->odd.x), a1 = fmaf(bmin[1], r->inv.y, -r->odd.y), a2 = fmaf(bmin[2], r->inv.z, -r->odd.z);
    const float b0 = fmaf(bmax[0], r->inv.x, -r->odd.x), b1 = fmaf(bmax[1], r->inv.y, -r->odd.y), b2 = fmaf(bmax[2], r->inv.z, -r->odd.z);
    const float lo0 = minps(a0, b0), lo1 = minps(a1, b1), lo2 = minps(a2, b2);
    const float hi0 = maxps(a0, b0), hi1 = maxps(a1, b1), hi2 = maxps(a2, b2);
    const float tn = maxps(lo0, maxps(lo1, lo2));
    const float tf = minps(hi0, minps(hi1, hi2));
    *outNear = tn;
    if (outFar) *outFar = tf;
    return (tf >= tn) && (tf >= 0.0f);
}

/* Intersect_TriangleRay, SSE branch, :132-167: strict > 0 on u, v, t and NOT (u + v > 1); true division */
static int tri_ray(const ray_t* r, const float* tri, float* u, float* v, float* t)
{
    const vec3 v0 = { tri[0], tri[1], tri[2] }, e1 = { tri[3], tri[4], tri[5] }, e2 = { tri[6], tri[7], tri[8] };
    const vec3 tvec = { r->o.x - v0.x, r->o.y - v0.y, r->o.z - v0.z };
    const vec3 pvec = cross(r->d, e2);
    const vec3 qvec = cross(tvec, e1);
    const float det = dp3(e1, pvec), uu = dp3(tvec, pvec), vv = dp3(r->d, qvec), tt = dp3(e2, qvec);
    const float sum = (uu + vv) / det;
    *u = uu / det; *v = vv / det; *t = tt / det;
    return !(sum > 1.0f) && (*v > 0.0f) && (*t > 0.0f) && (*u > 0.0f);
}

/* shape Intersect: SphereShape.cpp:29-46 (double), BoxShape.cpp:118-125 + Geometry.h:100-130, RectShape.cpp:32-49 */
static int shape_intersect(const RtbShape* s, const ray_t* r, float* nearD, float* farD, uint32_t* sub)
{
    *sub = 0xFFFFFFFFu;
    if (s->type == RTB_SHAPE_SPHERE)
    {
        const vec3 no = { 0.0f - r->o.x, 0.0f - r->o.y, 0.0f - r->o.z };
        const double rd = (double)s->size[0];
        const double v = (double)dp3(r->d, no);
        const double det = rd * rd - (double)dp3(r->o, r->o) + v * v;
        if (det <= 0.0) return 0;
        const double sq = sqrt(det);
        *nearD = (float)(v - sq); *farD = (float)(v + sq); *sub = 0;
        return *farD > *nearD;
    }
    if (s->type == RTB_SHAPE_BOX)
    {
        const float bmin[3] = { 0.0f - s->size[0], 0.0f - s->size[1], 0.0f - s->size[2] };
        *sub = 0;
        box_ray(r, bmin, s->size, nearD, farD);
        return *nearD < *farD;
    }
    if (s->type == RTB_SHAPE_RECT)
    {
        const float t = -r->o.z * r->inv.z;
        if (t > 1.1920929e-7f)
        {
            const float px = fmaf(r->d.x, t, r->o.x), py = fmaf(r->d.y, t, r->o.y);
            if (absbits(px) < s->size[0] && absbits(py) < s->size[1]) { *nearD = t; *farD = t; return 1; }
        }
    }
    return 0;
}

#define LEAVES(n) ((n)->leafInfo & 0x3FFFFFFFu)

/* GenericTraverse<MeshShape> + MeshShape::Traverse_Leaf (Traversal_Single.h:15-96, MeshShape.cpp:134-168) */
static void mesh_closest(const mesh_t* m, const ray_t* r, uint32_t objectId, hit_t* h)
{
    const RtbBvhNode* stack[128];
    uint32_t sp = 0;
    if (m->numNodes == 0) return;
    for (const RtbBvhNode* cur = m->nodes;;)
    {
        if (LEAVES(cur) != 0)
        {
            for (uint32_t i = 0; i < LEAVES(cur); ++i)
            {
                float u, v, t;
                const uint32_t ti = cur->childIndex + i;
                if (tri_ray(r, m->tris + (size_t)ti * 9, &u, &v, &t) && t < h->t) { h->t = t; h->sub = ti; h->object = objectId; h->u = u; h->v = v; }
            }
        }
        else
        {
            const RtbBvhNode* a = m->nodes + cur->childIndex;
            const RtbBvhNode* b = a + 1;
            float da, db;
            int ha = box_ray(r, a->min, a->max, &da, NULL), hb = box_ray(r, b->min, b->max, &db, NULL);
            ha &= (da < h->t); hb &= (db < h->t);
            if (ha && hb) { if (db < da) { const RtbBvhNode* t = a; a = b; b = t; } cur = a; stack[sp++] = b; continue; }
            if (ha) { cur = a; continue; }
            if (hb) { cur = b; continue; }
        }
        if (sp == 0) break;
        cur = stack[--sp];
    }
}

/* GenericTraverse_Shadow<MeshShape> + Traverse_Leaf_Shadow (Traversal_Single.h:98-179, MeshShape.cpp:175-207) */
static int mesh_shadow(const mesh_t* m, const ray_t* r, float tmax)
{
    const RtbBvhNode* stack[128];
    uint32_t sp = 0;
    if (m->numNodes == 0) return 0;
    for (const RtbBvhNode* cur = m->nodes;;)
    {
        if (LEAVES(cur) != 0)
        {
            for (uint32_t i = 0; i < LEAVES(cur); ++i)
            {
                float u, v, t;
                if (tri_ray(r, m->tris + (size_t)(cur->childIndex + i) * 9, &u, &v, &t) && t < tmax) return 1;
            }
        }
        else
        {
            const RtbBvhNode* a = m->nodes + cur->childIndex;
            const RtbBvhNode* b = a + 1;
            float da, db;
            int ha = box_ray(r, a->min, a->max, &da, NULL), hb = box_ray(r, b->min, b->max, &db, NULL);
            ha &= (da < tmax); hb &= (db < tmax);
            if (ha && hb) { cur = a; stack[sp++] = b; continue; }
            if (ha) { cur = a; continue; }
            if (hb) { cur = b; continue; }
        }
        if (sp == 0) break;
        cur = stack[--sp];
    }
    return 0;
}

/* Scene::Traverse_Object (Scene.cpp:128-145) + IShape::Traverse (Shape.cpp:19-45) + LightSceneObject::Traverse
 * (SceneObject_Light.cpp:27-38) */
static void object_closest(const scene_t* s, const ray_t* world, uint32_t id, hit_t* h)
{
    const RtbObject* o = &s->objects[id];
    const ray_t r = ray_to_local(&o->invTransform, world);
    if (o->shapeIndex == 0xFFFFFFFFu) return;                       /* point / spot light: TestRayHit is false */
    const RtbShape* sh = &s->shapes[o->shapeIndex];
    if (o->kind == RTB_OBJECT_SHAPE && sh->type == RTB_SHAPE_MESH) { mesh_closest(&s->meshes[sh->meshIndex], &r, id, h); return; }
    float n, f; uint32_t sub;
    if (!shape_intersect(sh, &r, &n, &f, &sub)) return;
    if (o->kind == RTB_OBJECT_LIGHT) { if (n > 0.0f && n < h->t) { h->t = n; h->object = id; h->sub = RTB_LIGHT_OBJECT; } return; }
    if (n > 0.0f && n < h->t) { h->t = n; h->object = id; h->sub = sub; return; }
    if (f > 0.0f && f < h->t) { h->t = f; h->object = id; h->sub = sub; }
}

/* Scene::Traverse_Object_Shadow (Scene.cpp:147-165) + IShape::Traverse_Shadow (Shape.cpp:47-57) +
 * LightSceneObject::Traverse_Shadow (SceneObject_Light.cpp:40-53) */
static int object_shadow(const scene_t* s, const ray_t* world, uint32_t id, float tmax)
{
    const RtbObject* o = &s->objects[id];
    const ray_t r = ray_to_local(&o->invTransform, world);
    if (o->shapeIndex == 0xFFFFFFFFu) return 0;
    const RtbShape* sh = &s->shapes[o->shapeIndex];
    if (o->kind == RTB_OBJECT_SHAPE && sh->type == RTB_SHAPE_MESH) return mesh_shadow(&s->meshes[sh->meshIndex], &r, tmax);
    float n, f; uint32_t sub;
    if (!shape_intersect(sh, &r, &n, &f, &sub)) return 0;
    if (o->kind == RTB_OBJECT_LIGHT) return n < tmax;
    return f > 0.0f && n < tmax;
}

/* Scene::Traverse (Scene.cpp:219-243): 1-object bypass, else GenericTraverse<Scene> + Scene::Traverse_Leaf */
static void scene_closest(const scene_t* s, const ray_t* r, hit_t* h)
{
    const RtbBvhNode* stack[128];
    uint32_t sp = 0;
    if (s->numObjects == 0) return;
    if (s->numObjects == 1) { object_closest(s, r, 0, h); return; }
    if (s->numTlas == 0) return;
    for (const RtbBvhNode* cur = s->tlas;;)
    {
        if (LEAVES(cur) != 0)
        {
            for (uint32_t i = 0; i < LEAVES(cur); ++i) object_closest(s, r, cur->childIndex + i, h);
        }
        else
        {
            const RtbBvhNode* a = s->tlas + cur->childIndex;
            const RtbBvhNode* b = a + 1;
            float da, db;
            int ha = box_ray(r, a->min, a->max, &da, NULL), hb = box_ray(r, b->min, b->max, &db, NULL);
            ha &= (da < h->t); hb &= (db < h->t);
            if (ha && hb) { if (db < da) { const RtbBvhNode* t = a; a = b; b = t; } cur = a; stack[sp++] = b; continue; }
            if (ha) { cur = a; continue; }
            if (hb) { cur = b; continue; }
        }
        if (sp == 0) break;
        cur = stack[--sp];
    }
}

/* Scene::Traverse_Shadow (Scene.cpp:245-261) */
static int scene_shadow(const scene_t* s, const ray_t* r, float tmax)
{
    const RtbBvhNode* stack[128];
    uint32_t sp = 0;
    if (s->numObjects == 0) return 0;
    if (s->numObjects == 1) return object_shadow(s, r, 0, tmax);
    if (s->numTlas == 0) return 0;
    for (const RtbBvhNode* cur = s->tlas;;)
    {
        if (LEAVES(cur) != 0)
        {
            for (uint32_t i = 0; i < LEAVES(cur); ++i) if (object_shadow(s, r, cur->childIndex + i, tmax)) return 1;
        }
        else
        {
            const RtbBvhNode* a = s->tlas + cur->childIndex;
            const RtbBvhNode* b = a + 1;
            float da, db;
            int ha = box_ray(r, a->min, a->max, &da, NULL), hb = box_ray(r, b->min, b->max, &db, NULL);
            ha &= (da < tmax); hb &= (db < tmax);
            if (ha && hb) { cur = a; stack[sp++] = b; continue; }
            if (ha) { cur = a; continue; }
            if (hb) { cur = b; continue; }
        }
        if (sp == 0) break;
        cur = stack[--sp];
    }
    return 0;
}

/* ---- sampler (Core/Sampling/GenericSampler.cpp:56-112) + MurmurHash3 finalizer (Core/Math/Math.h:269-277) ------- */
static uint64_t murmur64(uint64_t h) { h ^= h >> 33; h *= 0xff51afd7ed558ccdULL; h ^= h >> 33; h *= 0xc4ceb9fe1a85ec53ULL; h ^= h >> 33; return h; }
typedef struct { uint32_t salt, generated, bx, by; } sampler_t;
static void sampler_reset(sampler_t* s, uint32_t x, uint32_t y) { s->bx = x & 127u; s->by = y & 127u; s->salt = (uint32_t)murmur64((uint64_t)(x | (y << 16))); s->generated = 0; }
static uint32_t sampler_int(sampler_t* s, const uint32_t* seed, uint32_t dims, const uint16_t* blue)
{
    uint32_t v = s->generated < dims ? seed[s->generated] : 0u;
    if (blue && s->generated < 4u) v += (uint32_t)blue[4u * (128u * s->by + s->bx) + s->generated] << 16;
    else { uint32_t x = s->salt; v += x; x ^= x << 13; x ^= x >> 17; x ^= x << 5; s->salt = x; }
    s->generated++;
    return v;
}

/* Viewport::RenderTile coordinates (Viewport.cpp:297-321) + Camera::GenerateRay without DOF (Camera.cpp:81-118) */
static ray_t primary_ray(const RtbCamera* c, uint32_t w, uint32_t h, uint32_t x, uint32_t y, float offx, float offy)
{
    const float aspect = (float)w / (float)h;
    const float cx = ((float)(int32_t)x + offx) * (1.0f / (float)(int32_t)w);
    const float cy = ((float)(int32_t)(h - 1u - y) + offy) * (1.0f / (float)(int32_t)h);
    const float ox = fmaf(cx, 2.0f, -1.0f), oy = fmaf(cy, 2.0f, -1.0f);          /* UnipolarToBipolar = fmsub(x, 2, 1) */
    const float sx = ox * aspect;
    const float (*m)[4] = c->localToWorld.r;
    vec3 dir;
    dir.x = fmaf(fmaf(m[0][0], sx, m[1][0] * oy), c->tanHalfFoV, m[2][0]);
    dir.y = fmaf(fmaf(m[0][1], sx, m[1][1] * oy), c->tanHalfFoV, m[2][1]);
    dir.z = fmaf(fmaf(m[0][2], sx, m[1][2] * oy), c->tanHalfFoV, m[2][2]);
    const vec3 o = { m[3][0], m[3][1], m[3][2] };
    return ray_make(o, dir);
}

/* HSVtoRGB (Core/Color/ColorHelpers.h:134-156) + DebugRenderer TriangleID / Depth (DebugRenderer.cpp:94-108) */
static void debug_color(uint32_t mode, const hit_t* h, float* rgb)
{
    rgb[0] = rgb[1] = rgb[2] = 0.0f;
    if (h->t == INFINITY) return;
    if (h->sub == RTB_LIGHT_OBJECT) { rgb[0] = 1.0f; rgb[1] = 1.0f; return; }
    if (mode == 2) { const float d = 1.0f - 1.0f / (1.0f + h->t / 10.0f); rgb[0] = rgb[1] = rgb[2] = d; return; }
    const uint64_t k = murmur64((uint64_t)h->object | ((uint64_t)h->sub << 32));
    const float hue = (float)(uint32_t)k / (float)UINT32_MAX;
    const float sat = 0.5f + 0.5f * (float)(uint32_t)(k >> 32) / (float)UINT32_MAX;
    const float val = 1.0f;
    const int hi = (int)(hue * 6.0f);
    const float f = hue * 6 - hi, p = val * (1 - sat), q = val * (1 - f * sat), t = val * (1 - (1 - f) * sat);
    switch (hi)
    {
    case 0: rgb[0] = val; rgb[1] = t; rgb[2] = p; break;
    case 1: rgb[0] = q; rgb[1] = val; rgb[2] = p; break;
    case 2: rgb[0] = p; rgb[1] = val; rgb[2] = t; break;
    case 3: rgb[0] = p; rgb[1] = q; rgb[2] = val; break;
    case 4: rgb[0] = t; rgb[1] = p; rgb[2] = val; break;
    case 5: rgb[0] = val; rgb[1] = p; rgb[2] = q; break;
    default: break;
    }
}

/* ---- pass script (seeds + AA offsets, written by rtref render --dump-passes) ------------------------------------ */
typedef struct { uint32_t passes, dims; uint32_t* seeds; float* offsets; } passes_t;
static void load_passes(const char* path, passes_t* p)
{
    FILE* f = fopen(path, "rb");
    if (!f) die("cannot open pass script");
    p->passes = read_u32(f); p->dims = read_u32(f);
    p->seeds = malloc((size_t)p->passes * p->dims * 4 + 4); p->offsets = malloc((size_t)p->passes * 8 + 4);
    for (uint32_t i = 0; i < p->passes; ++i)
    {
        if (fread(p->seeds + (size_t)i * p->dims, 4, p->dims, f) != p->dims || fread(p->offsets + 2 * i, 4, 2, f) != 2) die("bad pass script");
    }
    fclose(f);
}

static ray_t ray_from_file(const RtbRay* r)
{
    ray_t o;
    o.o.x = r->origin[0]; o.o.y = r->origin[1]; o.o.z = r->origin[2];
    o.d.x = r->dir[0]; o.d.y = r->dir[1]; o.d.z = r->dir[2];
    o.inv.x = r->invDir[0]; o.inv.y = r->invDir[1]; o.inv.z = r->invDir[2];
    o.odd.x = r->originDivDir[0]; o.odd.y = r->originDivDir[1]; o.odd.z = r->originDivDir[2];
    return o;
}

/* ---- textures (Core/Textures/*.cpp) ---------------------------------------------------------------------------- */

/* Vector4::Lerp = MulAndAdd(b - a, w, a), Vector4Impl.h:58-66 (a fused multiply-add in the reference build) */
static float lerp_fma(float a, float b, float w) { return fmaf(b - a, w, a); }

/* Ken Perlin's permutation, NoiseTexture.cpp:10-27 */
static const uint8_t perm[256] = {
    151,160,137,91,90,15,131,13,201,95,96,53,194,233,7,225,140,36,103,30,69,142,8,99,37,240,21,10,23,190,6,148,247,120,234,75,0,26,197,62,94,252,219,203,117,35,11,32,57,177,33,
    88,237,149,56,87,174,20,125,136,171,168,68,175,74,165,71,134,139,48,27,166,77,146,158,231,83,111,229,122,60,211,133,230,220,105,92,41,55,46,245,40,244,
    102,143,54,65,25,63,161,1,216,80,73,209,76,132,187,208,89,18,169,200,196,135,130,116,188,159,86,164,100,109,198,173,186,3,64,52,217,226,250,124,123,
    5,202,38,147,118,126,255,82,85,212,207,206,59,227,47,16,58,17,182,189,28,42,223,183,170,213,119,248,152,2,44,154,163,70,221,153,101,155,167,43,172,9,
    129,22,39,253,19,98,108,110,79,113,224,232,178,185,112,104,218,246,97,228,251,34,242,193,238,210,144,12,191,179,162,241,81,51,145,235,249,14,239,107,
    49,192,214,31,181,199,106,157,184,84,204,176,115,121,50,45,127,4,150,254,138,236,205,93,222,114,67,29,24,72,243,141,128,195,78,66,215,61,156,180 };

static int floor_to_int(float v) { const int i = (int)v; return v < (float)i ? i - 1 : i; }       /* FloorInt, Math.h:100-104 */

static float corner(int gi, float x, float y)            /* one simplex corner: falloff^4 * gradient, NoiseTexture.cpp:34-40, 104-135 */
{
    float t = 0.5f - x * x - y * y;
    if (t < 0.0f) return 0.0f;
    const int h = gi & 0x3F;
    const float u = h < 4 ? x : y, v = h < 4 ? y : x;
    const float g = ((h & 1) ? -u : u) + ((h & 2) ? -2.0f * v : 2.0f * v);
    t *= t;
    return t * t * g;
}

static float simplex2(float x, float y)                   /* NoiseTexture::EvaluateInternal, NoiseTexture.cpp:59-139 */
{
    const float F2 = 0.366025403f, G2 = 0.211324865f;
    const float s = (x + y) * F2;
    const int i = floor_to_int(x + s), j = floor_to_int(y + s);
    const float t = (float)(i + j) * G2;
    const float x0 = x - ((float)i - t), y0 = y - ((float)j - t);
    const int i1 = x0 > y0 ? 1 : 0, j1 = 1 - i1;
    const float x1 = x0 - (float)i1 + G2, y1 = y0 - (float)j1 + G2;
    const float x2 = x0 - 1.0f + 2.0f * G2, y2 = y0 - 1.0f + 2.0f * G2;
    const int gi0 = perm[(uint8_t)(i + perm[(uint8_t)j])];
    const int gi1 = perm[(uint8_t)(i + i1 + perm[(uint8_t)(j + j1)])];
    const int gi2 = perm[(uint8_t)(i + 1 + perm[(uint8_t)(j + 1)])];
    const float n0 = corner(gi0, x0, y0), n1 = corner(gi1, x1, y1), n2 = corner(gi2, x2, y2);
    const float v = 0.5f + 22.615325f * (n0 + n1 + n2);
    return v > 1.0f ? 1.0f : (v < 0.0f ? 0.0f : v);
}

static void tex_eval(const scene_t* s, uint32_t index, float u, float v, float out[4])
{
    const tex_t* t = &s->textures[index];
    if (t->type == RTB_TEXTURE_MIX)
    {
        /* MixTexture::Evaluate, MixTexture.cpp:24-31 */
        float a[4], b[4], w[4];
        tex_eval(s, t->a, u, v, a); tex_eval(s, t->b, u, v, b); tex_eval(s, t->w, u, v, w);
        for (int k = 0; k < 4; ++k) out[k] = lerp_fma(a[k], b[k], w[k]);
        return;
    }
    if (t->type == RTB_TEXTURE_CHECKERBOARD)
    {
        /* CheckerboardTexture::Evaluate, CheckerboardTexture.cpp:32-41 */
        const float fu = u - floorf(u), fv = v - floorf(v);
        const float* c = ((fu > 0.5f) ^ (fv > 0.5f)) ? t->colorA : t->colorB;
        memcpy(out, c, 16);
        return;
    }
    if (t->type == RTB_TEXTURE_NOISE)
    {
        /* NoiseTexture::Evaluate, NoiseTexture.cpp:141-155 */
        float value = 0.0f, vs = 0.5f, cs = 1.0f;
        for (uint32_t o = 0; o < t->octaves; ++o) { value += vs * simplex2(u * cs, v * cs); vs *= 0.5f; cs *= 2.0f; }
        for (int k = 0; k < 4; ++k) out[k] = lerp_fma(t->colorA[k], t->colorB[k], value);
        return;
    }
    /* BitmapTexture::Evaluate, BitmapTexture.cpp:32-93: wrap, floor, smooth-step bilinear */
    const int w = (int)t->width, h = (int)t->height;
    const float su = (u - floorf(u)) * (float)w, sv = (v - floorf(v)) * (float)h;
    const int iu = (int)lrintf(floorf(su)), iv = (int)lrintf(floorf(sv));        /* _mm_cvtps_epi32 of an integral value */
    int x0 = iu, y0 = iv;
    if (!(iu < w)) x0 -= w;
    if (!(iv < h)) y0 -= h;
    if (iu < 0) x0 += w;
    if (iv < 0) y0 += h;
    int x1 = x0 + 1, y1 = y0 + 1;
    if (!(x1 < w)) x1 -= w;
    if (!(y1 < h)) y1 -= h;
    const float* c00 = t->texels + 4 * ((size_t)y0 * w + x0), *c10 = t->texels + 4 * ((size_t)y0 * w + x1);
    const float* c01 = t->texels + 4 * ((size_t)y1 * w + x0), *c11 = t->texels + 4 * ((size_t)y1 * w + x1);
    float wx = su - (float)iu, wy = sv - (float)iv;
    wx = wx * wx * (3.0f - wx * 2.0f); wy = wy * wy * (3.0f - wy * 2.0f);              /* SmoothStep, Math.h:176-179 */
    for (int k = 0; k < 4; ++k) out[k] = lerp_fma(lerp_fma(c00[k], c01[k], wy), lerp_fma(c10[k], c11[k], wy), wx);
}

/* ---- front buffer: Viewport::PostProcessTile, Rendering/Viewport.cpp:492-552 (bloom off) ---------------------------- */

static uint32_t as_u32(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static float as_f32(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

/* FastLog(Vector4), Math/Transcendental.cpp:216-232: exponent split off with integer arithmetic, log1p polynomial */
static float fast_log(float a)
{
    const uint32_t bits = as_u32(a);
    const uint32_t e = (bits - 0x3f2aaaabu) & 0xff800000u;
    const float m = as_f32(bits - e);
    const float i = (float)(int32_t)e * 1.19209290e-7f;
    const float f = m - 1.0f, s = f * f;
    float r = fmaf(f, 0.230836749f, -0.279208571f);
    const float t = fmaf(f, 0.331826031f, -0.498910338f);
    r = fmaf(r, s, t);
    r = fmaf(r, s, f);
    return fmaf(i, 0.693147182f, r);
}

/* _mm_cvtps_epi32 / _mm_cvttps_epi32: "integer indefinite" outside the int32 range */
static int32_t cvt_nearest(float v) { return (v >= -2147483648.0f && v < 2147483648.0f) ? (int32_t)lrintf(v) : INT32_MIN; }
static int32_t cvt_truncate(float v) { return (v >= -2147483648.0f && v < 2147483648.0f) ? (int32_t)v : INT32_MIN; }

/* FastExp(Vector4), Math/Transcendental.cpp:147-165 */
static float fast_exp(float a)
{
    const float t = a * 1.442695041f, fi = floorf(t), f = t - fi;
    const int32_t i = cvt_nearest(fi);
    float y = fmaf(f, 0.3371894346f, 0.657636276f);
    y = fmaf(f, y, 1.00172476f);
    y = as_f32(as_u32(y) + ((uint32_t)i << 23));
    if (-a >= 87.0f) y = 0.0f;
    if (a >= 87.0f) y = INFINITY;
    return y;
}

/* Convert_Linear_To_sRGB, Color/ColorHelpers.h:30-45 */
static float to_srgb(float c)
{
    const float s1 = sqrtf(c), s2 = sqrtf(s1), s3 = sqrtf(s2);
    float r = 0.585122381f * s1;
    r = fmaf(s2, 0.783140355f, r);
    r = fmaf(s3, -0.368262736f, r);
    return minps(1.0f, maxps(0.0f, r));
}

/* ToneMap, Tonemapper::ACES (the default), Color/ColorHelpers.h:111-122 */
static float aces(float c)
{
    const float t0 = c * fmaf(c, 2.51f, 0.03f);
    const float t1 = fmaf(c, 2.43f, 0.59f);
    const float t2 = fmaf(c, t1, 0.14f);
    return to_srgb(t0 * (1.0f / t2));
}

/* one pixel with PostprocessParams' defaults (PostProcess.cpp:6-14): colorFilter 1, exposure 0, contrast 0.8,
 * saturation 0.98, no bloom; dithering strength 0 (MulAndAdd(random, 0, x) == x); Vector4::ToBGR, Vector4ImplSSE.h:89-110 */
static uint32_t front_pixel(const float* sum, float pixelScaling)
{
    const float contrast = 0.8f, saturation = 0.98f, colorScale = 1.0f * powf(2.0f, 0.0f);
    float c[3];
    for (int k = 0; k < 3; ++k) c[k] = sum[k] * pixelScaling;
    const float gray = (c[0] * 0.2126f + c[1] * 0.7152f) + (c[2] * 0.0722f + 0.0f);            /* Dot3 = dpps */
    uint32_t out = 0;
    for (int k = 0; k < 3; ++k)
    {
        float v = maxps(0.0f, fmaf(c[k] - gray, saturation, gray));                               /* Max(0, Lerp(gray, rgb, saturation)) */
        v = fast_exp(fast_log(v) * contrast);
        v = aces(v * colorScale);
        int32_t q = cvt_truncate(v * 255.0f);
        q = q < 0 ? 0 : (q > 255 ? 255 : q);
        out |= (uint32_t)q << (16 - 8 * k);                                                       /* x -> R (bits 16..23), z -> B */
    }
    return out;
}

/* ---- Viewport::ComputeBlockError, Rendering/Viewport.cpp:552-582 ----------------------------------------------------- */

static float block_error(const float* sum, const float* sum2, uint32_t w, uint32_t h, uint32_t passes, const uint32_t* b /* minX maxX minY maxY */)
{
    const float scale = 1.0f / (float)passes;
    float total = 0.0f;
    for (uint32_t y = b[2]; y < b[3]; ++y)
    {
        float row = 0.0f;
        for (uint32_t x = b[0]; x < b[1]; ++x)
        {
            const float* p = sum + ((size_t)y * w + x) * 3;
            const float* q = sum2 + ((size_t)y * w + x) * 3;
            float a[3], d[3];
            for (int k = 0; k < 3; ++k) { a[k] = scale * p[k]; d[k] = absbits(a[k] - (2.0f * scale) * q[k]); }
            row += (d[0] + 2.0f * d[1] + d[2]) / sqrtf(0.000001f + a[0] + 2.0f * a[1] + a[2]);
        }
        total += row;
    }
    const uint32_t area = (b[1] - b[0]) * (b[3] - b[2]);
    return total * sqrtf((float)area / (float)(w * h)) / (float)area;
}

/* ---- the host prologue of Viewport::Render under RT_FIXED_SEED --------------------------------------------------------- */

/* oracle/patch_ref.py P8: Entropy::GetInt = upper half of a splitmix64 step */
static uint64_t entropyState;
static uint32_t entropy_int(void)
{
    entropyState += 0x9E3779B97F4A7C15ULL;
    uint64_t z = entropyState;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return (uint32_t)((z ^ (z >> 31)) >> 32);
}

/* math::Random, Math/Random.cpp:19-126: xoroshiro128+ (scalar) and a 2 x 64-bit xorshift128+ (the "SIMD4" generator) */
typedef struct { uint64_t s[2]; uint64_t v4[2][2]; } rng_t;

static void rng_seed(rng_t* r)                      /* Random::Reset: 2 + 4 + 8 integers per half; GCC evaluates the */
{                                                   /* VectorInt4(x, y, z, w) arguments right to left                */
    for (int i = 0; i < 2; ++i)
    {
        const uint64_t hi = entropy_int(), lo = entropy_int();
        r->s[i] = (hi << 32) | lo;
        uint32_t lane[4];
        for (int k = 3; k >= 0; --k) lane[k] = entropy_int();
        r->v4[i][0] = ((uint64_t)lane[1] << 32) | lane[0];
        r->v4[i][1] = ((uint64_t)lane[3] << 32) | lane[2];
        for (int k = 0; k < 8; ++k) entropy_int();   /* mSeedSimd8 */
    }
}

static uint64_t rng_long(rng_t* r)                  /* Random::GetLong, :33-47 */
{
    const uint64_t s0 = r->s[0];
    uint64_t s1 = r->s[1];
    const uint64_t result = s0 + s1;
    s1 ^= s0;
    r->s[0] = ((s0 << 24) | (s0 >> 40)) ^ s1 ^ (s1 << 16);
    r->s[1] = (s1 << 37) | (s1 >> 27);
    return result;
}

static double rng_double(rng_t* r) { return (double)rng_long(r) / (double)UINT64_MAX; }           /* :60-63 */

static void rng_float2(rng_t* r, float* x, float* y)   /* GetFloat2 = lanes 0, 1 of GetVector4, :65-68, 83-126 */
{
    uint64_t out = 0;
    for (int l = 0; l < 2; ++l)
    {
        const uint64_t s0 = r->v4[1][l];
        uint64_t s1 = r->v4[0][l];
        if (l == 0) out = s0 + s1;
        s1 <<= 23;
        r->v4[0][l] = s0;
        r->v4[1][l] = (s0 ^ s1) ^ ((s0 >> 5) ^ (s1 >> 18));
    }
    *x = as_f32(((uint32_t)out & 0x007fffffu) | 0x3f800000u) - 1.0f;
    *y = as_f32(((uint32_t)(out >> 32) & 0x007fffffu) | 0x3f800000u) - 1.0f;
}

/* HaltonSequence, Sampling/HaltonSampler.cpp:37-230 (scrambled Halton with a random start, double precision) */
#define HALTON_WIDTH 64
typedef struct
{
    uint32_t dims;
    uint32_t* base; uint64_t* start; uint64_t** perm;
    uint64_t (*digit)[HALTON_WIDTH + 1]; uint64_t (*power)[HALTON_WIDTH]; double (*rnd)[HALTON_WIDTH + 1];
    rng_t rng;
} halton_t;

static void halton_init(halton_t* h, uint32_t dims)
{
    h->dims = dims;
    h->base = calloc(dims + 1, sizeof(uint32_t)); h->start = calloc(dims + 1, sizeof(uint64_t)); h->perm = calloc(dims + 1, sizeof(uint64_t*));
    h->digit = calloc(dims + 1, sizeof(*h->digit)); h->power = calloc(dims + 1, sizeof(*h->power)); h->rnd = calloc(dims + 1, sizeof(*h->rnd));
    /* InitPrimes, :135-154: the first `dims` primes by trial division */
    for (uint32_t m = 0, candidate = 2; m < dims; ++candidate)
    {
        int isPrime = 1;
        for (uint64_t i = 2; (double)i <= sqrt((double)candidate); ++i) if (candidate % i == 0) { isPrime = 0; break; }
        if (isPrime) h->base[m++] = candidate;
    }
    /* InitStart, :156-183: a random start index, built digit by digit from one uniform double */
    for (uint32_t i = 0; i < dims; ++i)
    {
        double r = rng_double(&h->rng);
        const uint64_t base = h->base[i];
        uint64_t z = 0, b = base;
        while (r > 1.0e-16)
        {
            if (r >= 1.0 / b)
            {
                const uint64_t cnt = (uint64_t)floor(r * b);
                r = r - cnt * 1.0 / b;
                z += cnt * b / base;
            }
            b *= base;
        }
        h->start[i] = z;
    }
    for (uint32_t i = 0; i < dims; ++i)              /* InitPowerBuffer */
        for (uint32_t j = 0; j < HALTON_WIDTH; ++j) h->power[i][j] = j ? h->power[i][j - 1] * h->base[i] : h->base[i];
    /* InitPermutation, :113-133: random transpositions of the digits */
    for (uint32_t i = 0; i < dims; ++i)
    {
        h->perm[i] = malloc(h->base[i] * sizeof(uint64_t));
        for (uint64_t j = 0; j < h->base[i]; ++j) h->perm[i][j] = j;
        for (uint64_t j = 1; j < h->base[i]; ++j)
        {
            const uint64_t other = (uint64_t)floor(rng_double(&h->rng) * h->base[i]);
            if (other != 0 && other < h->base[i]) { const uint64_t k = h->perm[i][j]; h->perm[i][j] = h->perm[i][other]; h->perm[i][other] = k; }
        }
    }
    /* InitExpansion: digits of start - 1, radical inverse accumulated from the most significant digit down */
    for (uint32_t i = 0; i < dims; ++i)
    {
        uint64_t n = h->start[i] - 1;
        int j = 0;
        while (n > 0 && j < HALTON_WIDTH) { h->digit[i][j] = n % h->base[i]; n /= h->base[i]; ++j; }
        for (--j; j >= 0; --j) h->rnd[i][j] = h->rnd[i][j + 1] + h->perm[i][h->digit[i][j]] * 1.0 / h->power[i][j];
    }
}

static void halton_next(halton_t* h)                 /* NextSample, :84-106: increment with carry, refresh the sums */
{
    for (uint32_t i = 0; i < h->dims; ++i)
    {
        int j = 0;
        while (h->digit[i][j] + 1 >= h->base[i] && j + 1 < HALTON_WIDTH) ++j;
        h->digit[i][j]++;
        h->rnd[i][j] = h->rnd[i][j + 1] + h->perm[i][h->digit[i][j]] * 1.0 / h->power[i][j];
        for (--j; j >= 0; --j)
        {
            h->digit[i][j] = 0;
            h->rnd[i][j] = h->rnd[i][j + 1] + h->perm[i][0] * 1.0 / h->power[i][j];
        }
    }
}

/* Sin(Vector4) lane, Math/Transcendental.cpp:54-76 (coefficients :26-34); SinCos(x) = Sin((x, x + pi/2)), TranscendentalImpl.h:22-26 */
static float sin_lane(float a)
{
    const int32_t i = cvt_nearest(a * (1.0f / 3.14159265359f));
    const float x = fmaf(-(float)i, 3.14159265359f, a);            /* NegMulAndAdd = fnmadd */
    const float x2 = x * x;
    float y = fmaf(-2.0532988642e-08f, x2, 2.7040521217e-06f);
    y = fmaf(y, x2, -1.9812576647e-04f);
    y = fmaf(y, x2, 8.3325579762e-03f);
    y = fmaf(y, x2, -1.6666577756e-01f);
    y = fmaf(y, x2, 9.9999970197e-01f);
    y *= x;
    return as_f32(as_u32(y) ^ ((uint32_t)i << 31));
}

/* scalar FastLog, Math/Transcendental.cpp:194-214: the same polynomial as the vector one, without fused operations */
static float fast_log_scalar(float x)
{
    const uint32_t bits = as_u32(x);
    const uint32_t e = (bits - 0x3f2aaaabu) & 0xff800000u;
    const float m = as_f32(bits - e);
    const float i = 1.19209290e-7f * (float)(int32_t)e;
    const float f = m - 1.0f, s = f * f;
    float r = 0.230836749f * f - 0.279208571f;
    const float t = 0.331826031f * f - 0.498910338f;
    r = r * s + t;
    r = r * s + f;
    return i * 0.693147182f + r;
}

int main(int argc, char** argv)
{
    _MM_SET_FLUSH_ZERO_MODE(_MM_FLUSH_ZERO_ON);
    _MM_SET_DENORMALS_ZERO_MODE(_MM_DENORMALS_ZERO_ON);
    if (argc < 2) die("usage: rt_restate trace|sampler|primary|debug|texeval|front|blockerr|passes ...");

    if (!strcmp(argv[1], "trace") && argc >= 5)
    {
        /* trace <scene.rtb> <rays> <out.hits> [tmax file -> shadow mode] */
        scene_t s; load_scene(argv[2], &s);
        FILE* f = fopen(argv[3], "rb"); if (!f) die("cannot open rays");
        fseek(f, 0, SEEK_END); const size_t n = (size_t)ftell(f) / sizeof(RtbRay); fseek(f, 0, SEEK_SET);
        RtbRay* rays = malloc(n * sizeof(RtbRay) + 1);
        if (fread(rays, sizeof(RtbRay), n, f) != n) die("read failed");
        fclose(f);
        float* tmax = NULL;
        if (argc > 5) { tmax = malloc(n * 4 + 4); FILE* t = fopen(argv[5], "rb"); if (!t || fread(tmax, 4, n, t) != n) die("cannot read tmax"); fclose(t); }
        RtbHit* hits = calloc(n ? n : 1, sizeof(RtbHit));
        for (size_t i = 0; i < n; ++i)
        {
            const ray_t r = ray_from_file(&rays[i]);
            hit_t h = { RTB_INVALID_OBJECT, 0, INFINITY, 0.0f, 0.0f };
            if (tmax) { h.t = tmax[i]; hits[i].occluded = (uint32_t)scene_shadow(&s, &r, tmax[i]); }
            else scene_closest(&s, &r, &h);
            hits[i].objectId = h.object; hits[i].subObjectId = h.sub; hits[i].distance = h.t; hits[i].u = h.u; hits[i].v = h.v;
        }
        FILE* o = fopen(argv[4], "wb"); fwrite(hits, sizeof(RtbHit), n, o); fclose(o);
        return 0;
    }
    if (!strcmp(argv[1], "texeval") && argc >= 5)
    {
        /* texeval <scene.rtb> <uv pairs f32> <out f32x4 per texture per point> */
        scene_t s; load_scene(argv[2], &s);
        FILE* f = fopen(argv[3], "rb"); if (!f) die("cannot open uv file");
        fseek(f, 0, SEEK_END); const size_t n = (size_t)ftell(f) / 8; fseek(f, 0, SEEK_SET);
        float* uv = malloc(n * 8 + 8);
        if (fread(uv, 8, n, f) != n) die("read failed");
        fclose(f);
        FILE* o = fopen(argv[4], "wb");
        for (uint32_t t = 0; t < s.numTextures; ++t)
            for (size_t i = 0; i < n; ++i) { float c[4]; tex_eval(&s, t, uv[2 * i], uv[2 * i + 1], c); fwrite(c, 4, 4, o); }
        fclose(o);
        return 0;
    }
    if (!strcmp(argv[1], "sampler") && argc >= 8)
    {
        /* sampler <passes> <w> <h> <numDims> <out> <interleaved blue noise file | -> */
        passes_t p; load_passes(argv[2], &p);
        const uint32_t w = (uint32_t)atoi(argv[3]), h = (uint32_t)atoi(argv[4]), nd = (uint32_t)atoi(argv[5]);
        uint16_t* blue = NULL;
        if (strcmp(argv[7], "-")) { blue = malloc(128 * 128 * 4 * 2); FILE* b = fopen(argv[7], "rb"); if (!b || fread(blue, 2, 128 * 128 * 4, b) != 128 * 128 * 4) die("bad blue noise"); fclose(b); }
        FILE* o = fopen(argv[6], "wb");
        for (uint32_t i = 0; i < p.passes; ++i)
            for (uint32_t y = 0; y < h; ++y) for (uint32_t x = 0; x < w; ++x)
            {
                sampler_t s; sampler_reset(&s, x, y);
                for (uint32_t d = 0; d < nd; ++d) { const uint32_t v = sampler_int(&s, p.seeds + (size_t)i * p.dims, p.dims, blue); fwrite(&v, 4, 1, o); }
            }
        fclose(o);
        return 0;
    }
    if ((!strcmp(argv[1], "primary") || !strcmp(argv[1], "debug")) && argc >= 8)
    {
        /* primary <scene> <passes> <pass> <w> <h> <out.rays>      debug <scene> <passes> <mode 1|2> <w> <h> <out.sum> */
        scene_t s; load_scene(argv[2], &s);
        passes_t p; load_passes(argv[3], &p);
        const uint32_t arg = (uint32_t)atoi(argv[4]), w = (uint32_t)atoi(argv[5]), h = (uint32_t)atoi(argv[6]);
        FILE* o = fopen(argv[7], "wb");
        if (!strcmp(argv[1], "primary"))
        {
            for (uint32_t y = 0; y < h; ++y) for (uint32_t x = 0; x < w; ++x)
            {
                const ray_t r = primary_ray(&s.camera, w, h, x, y, p.offsets[2 * arg], p.offsets[2 * arg + 1]);
                RtbRay out; memset(&out, 0, sizeof(out));
                out.origin[0] = r.o.x; out.origin[1] = r.o.y; out.origin[2] = r.o.z;
                out.dir[0] = r.d.x; out.dir[1] = r.d.y; out.dir[2] = r.d.z;
                out.invDir[0] = r.inv.x; out.invDir[1] = r.inv.y; out.invDir[2] = r.inv.z;
                out.originDivDir[0] = r.odd.x; out.originDivDir[1] = r.odd.y; out.originDivDir[2] = r.odd.z;
                fwrite(&out, sizeof(out), 1, o);
            }
        }
        else
        {
            /* DebugRenderer over all passes of the script, Film::AccumulateColor sum (Film.cpp:31-39) */
            float* sum = calloc((size_t)w * h * 3, 4);
            for (uint32_t i = 0; i < p.passes; ++i)
                for (uint32_t y = 0; y < h; ++y) for (uint32_t x = 0; x < w; ++x)
                {
                    const ray_t r = primary_ray(&s.camera, w, h, x, y, p.offsets[2 * i], p.offsets[2 * i + 1]);
                    hit_t hit = { RTB_INVALID_OBJECT, 0, INFINITY, 0.0f, 0.0f };
                    scene_closest(&s, &r, &hit);
                    float rgb[3];
                    debug_color(arg, &hit, rgb);
                    float* px = sum + ((size_t)y * w + x) * 3;
                    px[0] = px[0] + rgb[0]; px[1] = px[1] + rgb[1]; px[2] = px[2] + rgb[2];
                }
            fwrite(sum, 4, (size_t)w * h * 3, o);
        }
        fclose(o);
        return 0;
    }
    if (!strcmp(argv[1], "front") && argc >= 7)
    {
        /* front <sum raster> <w> <h> <passes> <out.bgra> */
        const uint32_t w = (uint32_t)atoi(argv[3]), h = (uint32_t)atoi(argv[4]), passes = (uint32_t)atoi(argv[5]);
        float* sum = malloc((size_t)w * h * 12 + 4);
        FILE* f = fopen(argv[2], "rb"); if (!f || fread(sum, 12, (size_t)w * h, f) != (size_t)w * h) die("cannot read sum");
        fclose(f);
        FILE* o = fopen(argv[6], "wb");
        const float pixelScaling = 1.0f / (float)passes;                /* 1 / (1 + passesFinished), Viewport.cpp:499 */
        for (size_t i = 0; i < (size_t)w * h; ++i) { const uint32_t px = front_pixel(sum + 3 * i, pixelScaling); fwrite(&px, 4, 1, o); }
        fclose(o);
        return 0;
    }
    if (!strcmp(argv[1], "blockerr") && argc >= 9)
    {
        /* blockerr <sum> <sum2> <w> <h> <passes> <blocks u32x4> <out f32> */
        const uint32_t w = (uint32_t)atoi(argv[4]), h = (uint32_t)atoi(argv[5]), passes = (uint32_t)atoi(argv[6]);
        float* sum = malloc((size_t)w * h * 12 + 4); float* sum2 = malloc((size_t)w * h * 12 + 4);
        FILE* f = fopen(argv[2], "rb"); if (!f || fread(sum, 12, (size_t)w * h, f) != (size_t)w * h) die("cannot read sum");
        fclose(f);
        f = fopen(argv[3], "rb"); if (!f || fread(sum2, 12, (size_t)w * h, f) != (size_t)w * h) die("cannot read sum2");
        fclose(f);
        f = fopen(argv[7], "rb"); if (!f) die("cannot read blocks");
        FILE* o = fopen(argv[8], "wb");
        uint32_t b[4];
        while (fread(b, 4, 4, f) == 4) { const float e = block_error(sum, sum2, w, h, passes, b); fwrite(&e, 4, 1, o); }
        fclose(f); fclose(o);
        return 0;
    }
    if (!strcmp(argv[1], "passes") && argc >= 8)
    {
        /* passes <seed> <dims> <w> <h> <numPasses> <out.passes>: what the oracle harness' Viewport draws under RT_FIXED_SEED=seed.
         * Construction order (Viewport.h:106-107): mRandomGenerator, then the Halton sequence's generator.  Resize draws one
         * Vector4 per pixel from mRandomGenerator (Viewport.cpp:99-102) and resets; the harness resets once more. */
        entropyState = strtoull(argv[2], NULL, 10);
        const uint32_t dims = (uint32_t)atoi(argv[3]), w = (uint32_t)atoi(argv[4]), h = (uint32_t)atoi(argv[5]), n = (uint32_t)atoi(argv[6]);
        rng_t viewport; halton_t halton;
        rng_seed(&viewport); rng_seed(&halton.rng);
        float x, y;
        for (uint64_t i = 0; i < (uint64_t)w * h; ++i) rng_float2(&viewport, &x, &y);
        halton_init(&halton, dims);
        halton_init(&halton, dims);                  /* (the first initialisation is leaked: this is a one-shot tool) */
        FILE* o = fopen(argv[7], "wb");
        fwrite(&n, 4, 1, o); fwrite(&dims, 4, 1, o);
        for (uint32_t p = 0; p < n; ++p)
        {
            halton_next(&halton);
            for (uint32_t d = 0; d < dims; ++d) { const uint32_t v = (uint32_t)(halton.rnd[d][0] * (double)UINT32_MAX); fwrite(&v, 4, 1, o); }   /* GetInt, HaltonSampler.h:26 */
            /* GetFloatNormal2(GetFloat2()) * antiAliasingSpread (0.5), SamplingHelpers.cpp:145-149, Viewport.cpp:235-242 */
            rng_float2(&viewport, &x, &y);
            const float radius = sqrtf(-2.0f * fast_log_scalar(x)), angle = 2.0f * 3.14159265359f * y;
            const float off[2] = { radius * sin_lane(angle + 0.0f) * 0.5f, radius * sin_lane(angle + 3.14159265359f / 2.0f) * 0.5f };
            fwrite(off, 4, 2, o);
        }
        fclose(o);
        return 0;
    }
    die("bad arguments");
    return 2;
}
