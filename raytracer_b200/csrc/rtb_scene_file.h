// rtb_scene_file.h — "RTBS" flat scene files: a serialized RtbSceneDesc (+ the scene's camera).
//
// Writers: the host scene model (raytracer_b200/host, Scene::Flatten) and the oracle harness (oracle/rtref.cpp, which
// flattens the REFERENCE's own Scene object) — comparing the two files is the host-side parity test.
//
//   u32 magic 'RTBS', u32 version
//   u32 n, RtbBvhNode[n]        top-level BVH
//   u32 n, RtbObject[n]
//   u32 n, RtbShape[n]
//   u32 n, per mesh { u32 numNodes, numTriangles, numVertices, numMaterials; nodes; tris 9f; indices 4u; shading 8f; materialIds }
//   u32 n, RtbMaterial[n]
//   u32 n, RtbLight[n]
//   u32 n, u32[n]               global lights
//   RtbCamera, f32 fieldOfView
#pragma once

#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/rtb200.h"

namespace rtb {

struct SceneFile
{
    std::vector<RtbBvhNode> tlas;
    std::vector<RtbObject> objects;
    std::vector<RtbShape> shapes;
    struct Mesh
    {
        std::vector<RtbBvhNode> nodes;
        std::vector<float> triangles;
        std::vector<uint32_t> indices;
        std::vector<float> shading;
        std::vector<uint32_t> materialIds;
        uint32_t numVertices = 0;
    };
    std::vector<Mesh> meshes;
    std::vector<RtbMesh> meshViews;
    std::vector<RtbMaterial> materials;
    std::vector<RtbLight> lights;
    std::vector<uint32_t> globalLights;
    RtbCamera camera;
    float fieldOfView = 0.0f;

    SceneFile() { memset(&camera, 0, sizeof(camera)); }

    // pointers stay valid while this object is alive and unmodified
    RtbSceneDesc View()
    {
        meshViews.resize(meshes.size());
        for (size_t i = 0; i < meshes.size(); ++i)
        {
            RtbMesh& v = meshViews[i];
            const Mesh& m = meshes[i];
            v.numNodes = (uint32_t)m.nodes.size();
            v.numTriangles = (uint32_t)(m.triangles.size() / 9);
            v.numVertices = m.numVertices;
            v.numMaterials = (uint32_t)m.materialIds.size();
            v.nodes = m.nodes.data(); v.triangles = m.triangles.data(); v.indices = m.indices.data();
            v.shading = m.shading.data(); v.materialIds = m.materialIds.data();
        }
        RtbSceneDesc d;
        d.numTlasNodes = (uint32_t)tlas.size();           d.tlasNodes = tlas.data();
        d.numObjects = (uint32_t)objects.size();          d.objects = objects.data();
        d.numShapes = (uint32_t)shapes.size();            d.shapes = shapes.data();
        d.numMeshes = (uint32_t)meshViews.size();         d.meshes = meshViews.data();
        d.numMaterials = (uint32_t)materials.size();      d.materials = materials.data();
        d.numLights = (uint32_t)lights.size();            d.lights = lights.data();
        d.numGlobalLights = (uint32_t)globalLights.size(); d.globalLights = globalLights.data();
        return d;
    }

    template <typename T> static bool ReadVec(FILE* f, std::vector<T>& v, size_t n)
    {
        v.resize(n);
        return n == 0 || fread(v.data(), sizeof(T), n, f) == n;
    }
    static bool ReadU32(FILE* f, uint32_t& v) { return fread(&v, 4, 1, f) == 1; }

    bool Load(const std::string& path, std::string& err)
    {
        FILE* f = fopen(path.c_str(), "rb");
        if (!f) { err = "cannot open " + path; return false; }
        bool ok = LoadFrom(f, err);
        fclose(f);
        return ok;
    }

    bool LoadFrom(FILE* f, std::string& err)
    {
        uint32_t magic = 0, version = 0, n = 0;
        if (!ReadU32(f, magic) || !ReadU32(f, version) || magic != 0x53425452u || version != 1) { err = "not an RTBS v1 file"; return false; }
        if (!ReadU32(f, n) || !ReadVec(f, tlas, n)) { err = "truncated (tlas)"; return false; }
        if (!ReadU32(f, n) || !ReadVec(f, objects, n)) { err = "truncated (objects)"; return false; }
        if (!ReadU32(f, n) || !ReadVec(f, shapes, n)) { err = "truncated (shapes)"; return false; }
        if (!ReadU32(f, n)) { err = "truncated (meshes)"; return false; }
        meshes.resize(n);
        for (Mesh& m : meshes)
        {
            uint32_t nn, nt, nv, nm;
            if (!ReadU32(f, nn) || !ReadU32(f, nt) || !ReadU32(f, nv) || !ReadU32(f, nm)) { err = "truncated (mesh header)"; return false; }
            m.numVertices = nv;
            if (!ReadVec(f, m.nodes, nn) || !ReadVec(f, m.triangles, (size_t)nt * 9) || !ReadVec(f, m.indices, (size_t)nt * 4) ||
                !ReadVec(f, m.shading, (size_t)nv * 8) || !ReadVec(f, m.materialIds, nm)) { err = "truncated (mesh data)"; return false; }
        }
        if (!ReadU32(f, n) || !ReadVec(f, materials, n)) { err = "truncated (materials)"; return false; }
        if (!ReadU32(f, n) || !ReadVec(f, lights, n)) { err = "truncated (lights)"; return false; }
        if (!ReadU32(f, n) || !ReadVec(f, globalLights, n)) { err = "truncated (global lights)"; return false; }
        if (fread(&camera, sizeof(camera), 1, f) != 1 || fread(&fieldOfView, 4, 1, f) != 1) { err = "truncated (camera)"; return false; }
        return true;
    }

    template <typename T> static void WriteVec(FILE* f, const std::vector<T>& v) { if (!v.empty()) fwrite(v.data(), sizeof(T), v.size(), f); }
    static void WriteU32(FILE* f, uint32_t v) { fwrite(&v, 4, 1, f); }

    bool Save(const std::string& path) const
    {
        FILE* f = fopen(path.c_str(), "wb");
        if (!f) return false;
        WriteU32(f, 0x53425452u); WriteU32(f, 1);
        WriteU32(f, (uint32_t)tlas.size()); WriteVec(f, tlas);
        WriteU32(f, (uint32_t)objects.size()); WriteVec(f, objects);
        WriteU32(f, (uint32_t)shapes.size()); WriteVec(f, shapes);
        WriteU32(f, (uint32_t)meshes.size());
        for (const Mesh& m : meshes)
        {
            WriteU32(f, (uint32_t)m.nodes.size()); WriteU32(f, (uint32_t)(m.triangles.size() / 9));
            WriteU32(f, m.numVertices); WriteU32(f, (uint32_t)m.materialIds.size());
            WriteVec(f, m.nodes); WriteVec(f, m.triangles); WriteVec(f, m.indices); WriteVec(f, m.shading); WriteVec(f, m.materialIds);
        }
        WriteU32(f, (uint32_t)materials.size()); WriteVec(f, materials);
        WriteU32(f, (uint32_t)lights.size()); WriteVec(f, lights);
        WriteU32(f, (uint32_t)globalLights.size()); WriteVec(f, globalLights);
        fwrite(&camera, sizeof(camera), 1, f);
        fwrite(&fieldOfView, 4, 1, f);
        fclose(f);
        return true;
    }
};

} // namespace rtb
