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
// version 2 (written only when the scene has textures, so texture-less files stay readable by v1 readers) appends
//   u32 n, per texture { u32 type, width, height, numOctaves, texA, texB, texWeight, pad; f32 colorA[4], colorB[4];
//                        f32 texels[width*height*4] for bitmaps }
//   u32 n (0 or numMaterials), RtbMaterialTextures[n]
//   u32 n (0 or numLights), u32[n]      light textures
//   optional: u32 bokehWidth, bokehHeight, f32[bokehWidth * bokehHeight + 1]   bokeh importance map (cumulative distribution)
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
    struct Texture
    {
        RtbTexture desc;              // desc.texels is filled by View()
        std::vector<float> texels;
    };
    std::vector<Texture> textures;
    std::vector<RtbTexture> textureViews;
    std::vector<RtbMaterialTextures> materialTextures;   // empty or one per material
    std::vector<uint32_t> lightTextures;                 // empty or one per light
    uint32_t bokehWidth = 0, bokehHeight = 0;
    std::vector<float> bokehCdf;                         // empty or bokehWidth * bokehHeight + 1 entries

    bool HasTextures() const { return !textures.empty() || !materialTextures.empty() || !lightTextures.empty() || !bokehCdf.empty(); }

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
        textureViews.resize(textures.size());
        for (size_t i = 0; i < textures.size(); ++i)
        {
            textureViews[i] = textures[i].desc;
            textureViews[i].texels = textures[i].texels.empty() ? nullptr : textures[i].texels.data();
        }
        d.numTextures = (uint32_t)textureViews.size();    d.textures = textureViews.empty() ? nullptr : textureViews.data();
        d.materialTextures = materialTextures.size() == materials.size() && !materials.empty() ? materialTextures.data() : nullptr;
        d.lightTextures = lightTextures.size() == lights.size() && !lights.empty() ? lightTextures.data() : nullptr;
        d.bokehWidth = bokehWidth; d.bokehHeight = bokehHeight;
        d.bokehCdf = bokehCdf.empty() ? nullptr : bokehCdf.data();
        return d;
    }

    template <typename T> static bool ReadVec(FILE* f, std::vector<T>& v, size_t n)
    {
        if (n == 0) { v.clear(); return true; }
        // a count read from the file must not size an allocation beyond what the file can still deliver
        const long here = ftell(f);
        if (here >= 0 && fseek(f, 0, SEEK_END) == 0)
        {
            const long size = ftell(f);
            fseek(f, here, SEEK_SET);
            if (size >= here && n > (size_t)(size - here) / sizeof(T)) return false;
        }
        v.resize(n);
        return fread(v.data(), sizeof(T), n, f) == n;
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
        if (!ReadU32(f, magic) || !ReadU32(f, version) || magic != 0x53425452u || (version != 1 && version != 2)) { err = "not an RTBS v1/v2 file"; return false; }
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
        textures.clear(); materialTextures.clear(); lightTextures.clear();
        if (version >= 2)
        {
            if (!ReadU32(f, n)) { err = "truncated (textures)"; return false; }
            textures.resize(n);
            for (Texture& t : textures)
            {
                memset(&t.desc, 0, sizeof(t.desc));
                uint32_t h[8];
                if (fread(h, 4, 8, f) != 8 || fread(t.desc.colorA, 4, 4, f) != 4 || fread(t.desc.colorB, 4, 4, f) != 4) { err = "truncated (texture header)"; return false; }
                t.desc.type = h[0]; t.desc.width = h[1]; t.desc.height = h[2]; t.desc.numOctaves = h[3];
                t.desc.texA = h[4]; t.desc.texB = h[5]; t.desc.texWeight = h[6];
                if (t.desc.type == RTB_TEXTURE_BITMAP && !ReadVec(f, t.texels, (size_t)t.desc.width * t.desc.height * 4)) { err = "truncated (texels)"; return false; }
            }
            if (!ReadU32(f, n) || !ReadVec(f, materialTextures, n)) { err = "truncated (material textures)"; return false; }
            if (!ReadU32(f, n) || !ReadVec(f, lightTextures, n)) { err = "truncated (light textures)"; return false; }
            bokehWidth = bokehHeight = 0; bokehCdf.clear();
            uint32_t bw = 0, bh = 0;
            if (ReadU32(f, bw) && ReadU32(f, bh) && bw && bh)
            {
                if (!ReadVec(f, bokehCdf, (size_t)bw * bh + 1)) { err = "truncated (bokeh distribution)"; return false; }
                bokehWidth = bw; bokehHeight = bh;
            }
        }
        return true;
    }

    template <typename T> static void WriteVec(FILE* f, const std::vector<T>& v) { if (!v.empty()) fwrite(v.data(), sizeof(T), v.size(), f); }
    static void WriteU32(FILE* f, uint32_t v) { fwrite(&v, 4, 1, f); }

    bool Save(const std::string& path) const
    {
        FILE* f = fopen(path.c_str(), "wb");
        if (!f) return false;
        WriteU32(f, 0x53425452u); WriteU32(f, HasTextures() ? 2 : 1);
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
        if (HasTextures())
        {
            WriteU32(f, (uint32_t)textures.size());
            for (const Texture& t : textures)
            {
                const uint32_t h[8] = { t.desc.type, t.desc.width, t.desc.height, t.desc.numOctaves, t.desc.texA, t.desc.texB, t.desc.texWeight, 0u };
                fwrite(h, 4, 8, f); fwrite(t.desc.colorA, 4, 4, f); fwrite(t.desc.colorB, 4, 4, f);
                if (t.desc.type == RTB_TEXTURE_BITMAP) WriteVec(f, t.texels);
            }
            WriteU32(f, (uint32_t)materialTextures.size()); WriteVec(f, materialTextures);
            WriteU32(f, (uint32_t)lightTextures.size()); WriteVec(f, lightTextures);
            if (!bokehCdf.empty()) { WriteU32(f, bokehWidth); WriteU32(f, bokehHeight); WriteVec(f, bokehCdf); }
        }
        fclose(f);
        return true;
    }
};

} // namespace rtb
