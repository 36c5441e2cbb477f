// rt_loader.h — JSON scene loader + OBJ mesh loader, same schema and same construction rules as the reference's
// Demo/SceneLoader.cpp (692-820; shapes 418-499, lights 501-618, camera 652-690, materials 364-416, transforms
// 180-207) and Demo/MeshLoader.cpp (120-377: vertex welding by (v, vn, vt) index triple, degenerate-triangle filter,
// per-face normals, Lengyel tangents).  Errors are reported the way the reference does: a message + `false`.
// Textures follow Demo/SceneLoader.cpp:218-242 (TryParseTextureName) and 269-362 (ParseTexture): a name is first looked
// up among the scene's "textures", else loaded as a bitmap file relative to the data path; a file that cannot be
// loaded leaves the parameter untextured, as in the reference.
#pragma once

#include <cstdio>
#include <fstream>
#include <sstream>
#include <unordered_map>

#include "rt_json.h"
#include "rt_scene.h"

namespace rt {
namespace helpers {

using MaterialsMap = std::map<std::string, MaterialPtr>;
using TexturesMap = std::map<std::string, TexturePtr>;

struct LoadError { std::string message; };

inline bool Fail(LoadError& e, const std::string& m) { if (e.message.empty()) e.message = m; return false; }

// ---- OBJ / MTL ------------------------------------------------------------------------------------------------------

// decimal parser of the OBJ reader the reference uses (External/tiny_obj_loader.h:567-680): digits are accumulated
// into a double mantissa, fractional digit k contributes digit * 10^-k; the result is then narrowed to float
inline bool ParseObjReal(const char* s, const char* sEnd, double& result)
{
    if (s >= sEnd) return false;
    double mantissa = 0.0;
    int exponent = 0;
    char sign = '+', expSign = '+';
    const char* curr = s;
    int read = 0;
    if (*curr == '+' || *curr == '-') { sign = *curr; curr++; }
    else if (!(*curr >= '0' && *curr <= '9')) return false;
    bool more = curr != sEnd;
    while (more && *curr >= '0' && *curr <= '9') { mantissa *= 10; mantissa += static_cast<int>(*curr - '0'); curr++; read++; more = curr != sEnd; }
    if (read == 0) return false;
    bool assemble = !more;
    if (!assemble)
    {
        if (*curr == '.')
        {
            static const double lut[] = { 1.0, 0.1, 0.01, 0.001, 0.0001, 0.00001, 0.000001, 0.0000001 };
            curr++; read = 1; more = curr != sEnd;
            while (more && *curr >= '0' && *curr <= '9')
            {
                mantissa += static_cast<int>(*curr - '0') * (read < 8 ? lut[read] : std::pow(10.0, -read));
                read++; curr++; more = curr != sEnd;
            }
        }
        else if (!(*curr == 'e' || *curr == 'E')) assemble = true;
        if (!assemble && more && (*curr == 'e' || *curr == 'E'))
        {
            curr++; more = curr != sEnd;
            if (more && (*curr == '+' || *curr == '-')) { expSign = *curr; curr++; }
            else if (!(more && *curr >= '0' && *curr <= '9')) return false;
            read = 0; more = curr != sEnd;
            while (more && *curr >= '0' && *curr <= '9') { exponent *= 10; exponent += static_cast<int>(*curr - '0'); curr++; read++; more = curr != sEnd; }
            exponent *= (expSign == '+' ? 1 : -1);
            if (read == 0) return false;
        }
    }
    result = (sign == '+' ? 1 : -1) * (exponent ? std::ldexp(mantissa * std::pow(5.0, exponent), exponent) : mantissa);
    return true;
}

inline float NextObjReal(const char*& tok)
{
    tok += strspn(tok, " \t");
    const char* e = tok + strcspn(tok, " \t\r");
    double v = 0.0;
    ParseObjReal(tok, e, v);
    tok = e;
    return static_cast<float>(v);
}

struct ObjMaterial
{
    std::string name; float diffuse[3] = { 0, 0, 0 }; float emission[3] = { 0, 0, 0 };
    std::string diffuseTexname, normalTexname, alphaTexname;     // map_Kd, norm, map_d (tiny_obj_loader material_t)
    bool hasTextureOptions = false;
};

struct ObjIndex { int v, vn, vt; bool operator==(const ObjIndex& o) const { return v == o.v && vn == o.vn && vt == o.vt; } };
struct ObjIndexHash { size_t operator()(const ObjIndex& k) const { return (size_t)k.v * 73856093u ^ (size_t)(k.vn + 1) * 19349663u ^ (size_t)(k.vt + 1) * 83492791u; } };

struct ObjData
{
    std::vector<float> vertices, normals, texcoords;
    std::vector<ObjIndex> faceIndices;     // 3 per triangle, file order
    std::vector<int> faceMaterial;         // per triangle, -1 = none
    std::vector<ObjMaterial> materials;
};

inline bool LoadMtl(const std::string& path, std::vector<ObjMaterial>& out, std::map<std::string, int>& ids)
{
    std::ifstream f(path);
    if (!f) return false;
    std::string line;
    ObjMaterial cur;
    bool have = false;
    while (std::getline(f, line))
    {
        const char* t = line.c_str();
        t += strspn(t, " \t");
        if (*t == '#' || *t == 0 || *t == '\r') continue;
        if (!strncmp(t, "newmtl", 6) && (t[6] == ' ' || t[6] == '\t'))
        {
            if (have) { ids[cur.name] = (int)out.size(); out.push_back(cur); }
            cur = ObjMaterial();
            std::string n(t + 7);
            while (!n.empty() && (n.back() == '\r' || n.back() == ' ' || n.back() == '\t')) n.pop_back();
            cur.name = n;
            have = true;
        }
        else if (t[0] == 'K' && t[1] == 'd' && (t[2] == ' ' || t[2] == '\t')) { t += 2; for (int i = 0; i < 3; ++i) cur.diffuse[i] = NextObjReal(t); }
        else if (t[0] == 'K' && t[1] == 'e' && (t[2] == ' ' || t[2] == '\t')) { t += 2; for (int i = 0; i < 3; ++i) cur.emission[i] = NextObjReal(t); }
        else if ((!strncmp(t, "map_Kd", 6) && (t[6] == ' ' || t[6] == '\t')) || (!strncmp(t, "map_d", 5) && (t[5] == ' ' || t[5] == '\t')) ||
                 (!strncmp(t, "norm", 4) && (t[4] == ' ' || t[4] == '\t')))
        {
            // tinyobj::ParseTextureNameAndOption (External/tiny_obj_loader.h:399, 962-977): without options the name is the rest of the line
            std::string& dst = t[0] == 'n' ? cur.normalTexname : (t[4] == 'K' ? cur.diffuseTexname : cur.alphaTexname);
            t += (t[0] == 'n' ? 5 : (t[4] == 'K' ? 7 : 6));
            t += strspn(t, " \t");
            std::string n(t);
            while (!n.empty() && (n.back() == '\r' || n.back() == '\n')) n.pop_back();
            if (!n.empty() && n[0] == '-') { cur.hasTextureOptions = true; n.clear(); }     // texture options (-blendu, -s ...): refused by LoadMesh
            dst = n;
        }
    }
    if (have) { ids[cur.name] = (int)out.size(); out.push_back(cur); }
    return true;
}

inline bool LoadObj(const std::string& path, const std::string& baseDir, ObjData& out, LoadError& err)
{
    std::ifstream f(path);
    if (!f) return Fail(err, "Failed to load mesh '" + path + "'");
    std::map<std::string, int> materialIds;
    int material = -1;
    std::string line;
    auto fixIndex = [](int idx, int n) { return idx > 0 ? idx - 1 : (idx == 0 ? 0 : n + idx); };
    while (std::getline(f, line))
    {
        const char* t = line.c_str();
        t += strspn(t, " \t");
        if (*t == '#' || *t == 0 || *t == '\r') continue;
        if (t[0] == 'v' && (t[1] == ' ' || t[1] == '\t')) { t += 2; for (int i = 0; i < 3; ++i) out.vertices.push_back(NextObjReal(t)); }
        else if (t[0] == 'v' && t[1] == 'n' && (t[2] == ' ' || t[2] == '\t')) { t += 3; for (int i = 0; i < 3; ++i) out.normals.push_back(NextObjReal(t)); }
        else if (t[0] == 'v' && t[1] == 't' && (t[2] == ' ' || t[2] == '\t')) { t += 3; for (int i = 0; i < 2; ++i) out.texcoords.push_back(NextObjReal(t)); }
        else if (t[0] == 'f' && (t[1] == ' ' || t[1] == '\t'))
        {
            t += 2;
            std::vector<ObjIndex> face;
            for (;;)
            {
                t += strspn(t, " \t");
                if (*t == 0 || *t == '\r' || *t == '\n') break;
                ObjIndex ix = { -1, -1, -1 };
                ix.v = fixIndex(atoi(t), (int)(out.vertices.size() / 3));
                t += strcspn(t, "/ \t\r");
                if (*t == '/')
                {
                    t++;
                    if (*t != '/') { ix.vt = fixIndex(atoi(t), (int)(out.texcoords.size() / 2)); t += strcspn(t, "/ \t\r"); }
                    if (*t == '/') { t++; ix.vn = fixIndex(atoi(t), (int)(out.normals.size() / 3)); t += strcspn(t, "/ \t\r"); }
                }
                face.push_back(ix);
            }
            if (face.size() < 3) continue;
            for (size_t k = 1; k + 1 < face.size(); ++k)     // triangle fan
            {
                out.faceIndices.push_back(face[0]); out.faceIndices.push_back(face[k]); out.faceIndices.push_back(face[k + 1]);
                out.faceMaterial.push_back(material);
            }
        }
        else if (!strncmp(t, "usemtl", 6) && (t[6] == ' ' || t[6] == '\t'))
        {
            std::string n(t + 7);
            while (!n.empty() && (n.back() == '\r' || n.back() == ' ' || n.back() == '\t')) n.pop_back();
            const auto it = materialIds.find(n);
            material = it != materialIds.end() ? it->second : -1;
        }
        else if (!strncmp(t, "mtllib", 6) && (t[6] == ' ' || t[6] == '\t'))
        {
            std::string n(t + 7);
            while (!n.empty() && (n.back() == '\r' || n.back() == ' ' || n.back() == '\t')) n.pop_back();
            LoadMtl(baseDir + n, out.materials, materialIds);
        }
    }
    for (const ObjIndex& ix : out.faceIndices)
        if (ix.v < 0 || (size_t)ix.v * 3 + 2 >= out.vertices.size() + 0 || (ix.vn >= 0 && (size_t)ix.vn * 3 + 2 >= out.normals.size()) || (ix.vt >= 0 && (size_t)ix.vt * 2 + 1 >= out.texcoords.size()))
            return Fail(err, "Mesh '" + path + "' has an out-of-range index");
    return true;
}

// MeshLoader::LoadMesh + ComputeTangentVectors + BuildMesh, Demo/MeshLoader.cpp:120-400
inline MeshShapePtr LoadMesh(const std::string& filePath, MaterialsMap& outMaterials, const float scale, LoadError& err)
{
    const std::string baseDir = filePath.substr(0, filePath.find_last_of("\\/")) + "/";
    ObjData obj;
    if (!LoadObj(filePath, baseDir, obj, err)) return nullptr;

    constexpr float MinEdgeLength = 0.001f;
    constexpr float MinEdgeLengthSqr = MinEdgeLength * MinEdgeLength;

    MeshDesc desc;
    desc.path = filePath;
    std::unordered_map<ObjIndex, uint32_t, ObjIndexHash> unique;
    const size_t numFaces = obj.faceIndices.size() / 3;
    for (size_t face = 0; face < numFaces; ++face)
    {
        const ObjIndex idx[3] = { obj.faceIndices[3 * face], obj.faceIndices[3 * face + 1], obj.faceIndices[3 * face + 2] };
        const bool hasNormals = idx[0].vn >= 0 && idx[1].vn >= 0 && idx[2].vn >= 0;
        const bool hasTexCoords = idx[0].vt >= 0 && idx[1].vt >= 0 && idx[2].vt >= 0;
        Vector4 verts[3];
        for (int i = 0; i < 3; ++i)
            verts[i] = Vector4(obj.vertices[3 * idx[i].v], obj.vertices[3 * idx[i].v + 1], obj.vertices[3 * idx[i].v + 2]) * scale;
        const Vector4 edge1 = verts[1] - verts[0], edge2 = verts[2] - verts[0], edge3 = verts[2] - verts[1];
        // TriangleSurfaceArea(edge0, edge1) = |cross(edge1, edge0)| / 2, Geometry.h:51-55
        if (edge1.SqrLength3() < MinEdgeLengthSqr || edge2.SqrLength3() < MinEdgeLengthSqr || edge3.SqrLength3() < MinEdgeLengthSqr ||
            Vector4::Cross3(edge2, edge1).Length3() * 0.5f < MinEdgeLengthSqr)
            continue;    // degenerate triangle
        const Vector4 faceNormal = Vector4::Cross3(verts[1] - verts[0], verts[2] - verts[0]).Normalized3();
        for (int i = 0; i < 3; ++i)
        {
            const auto it = unique.find(idx[i]);
            uint32_t uniqueIndex;
            if (it != unique.end()) uniqueIndex = it->second;
            else
            {
                uniqueIndex = (uint32_t)unique.size();
                unique[idx[i]] = uniqueIndex;
                desc.positions.push_back(verts[i].x); desc.positions.push_back(verts[i].y); desc.positions.push_back(verts[i].z);
                Vector4 n = faceNormal;
                if (hasNormals) n = Vector4(obj.normals[3 * idx[i].vn], obj.normals[3 * idx[i].vn + 1], obj.normals[3 * idx[i].vn + 2]).Normalized3();
                desc.normals.push_back(n.x); desc.normals.push_back(n.y); desc.normals.push_back(n.z);
                if (hasTexCoords) { desc.texCoords.push_back(obj.texcoords[2 * idx[i].vt]); desc.texCoords.push_back(obj.texcoords[2 * idx[i].vt + 1]); }
                else { desc.texCoords.push_back(0.0f); desc.texCoords.push_back(0.0f); }
            }
            desc.indices.push_back(uniqueIndex);
        }
        desc.materialIndices.push_back((uint32_t)obj.faceMaterial[face]);
    }

    // ComputeTangentVectors (Lengyel's method), MeshLoader.cpp:286-377
    {
        const uint32_t numVertices = (uint32_t)(desc.positions.size() / 3);
        std::vector<float> tan((size_t)numVertices * 3, 0.0f);
        std::vector<Vector4> bitangents(numVertices);
        auto P = [&](uint32_t i) { return Vector4(desc.positions[3 * i], desc.positions[3 * i + 1], desc.positions[3 * i + 2], 0.0f); };
        const uint32_t numTriangles = (uint32_t)(desc.indices.size() / 3);
        for (uint32_t i = 0; i < numTriangles; ++i)
        {
            const uint32_t i0 = desc.indices[3 * i], i1 = desc.indices[3 * i + 1], i2 = desc.indices[3 * i + 2];
            const Vector4 e1 = P(i1) - P(i0), e2 = P(i2) - P(i0);
            const float s1 = desc.texCoords[2 * i1] - desc.texCoords[2 * i0], t1 = desc.texCoords[2 * i1 + 1] - desc.texCoords[2 * i0 + 1];
            const float s2 = desc.texCoords[2 * i2] - desc.texCoords[2 * i0], t2 = desc.texCoords[2 * i2 + 1] - desc.texCoords[2 * i0 + 1];
            const float det = s1 * t2 - s2 * t1;
            const float adet = det < 0.0f ? -det : det;
            if (adet < 1.0e-10f) continue;
            const float r = 1.0f / det;
            const Vector4 sdir = (e1 * t2 - e2 * t1) * r;
            const Vector4 tdir = (e2 * s1 - e1 * s2) * r;
            for (uint32_t v : { i0, i1, i2 })
            {
                tan[3 * v] += sdir.x; tan[3 * v + 1] += sdir.y; tan[3 * v + 2] += sdir.z;
                bitangents[v] = bitangents[v] + tdir;
            }
        }
        desc.tangents.resize((size_t)numVertices * 3);
        for (uint32_t i = 0; i < numVertices; ++i)
        {
            Vector4 tangent(tan[3 * i], tan[3 * i + 1], tan[3 * i + 2], 0.0f);
            const Vector4 normal(desc.normals[3 * i], desc.normals[3 * i + 1], desc.normals[3 * i + 2], 0.0f);
            Vector4 bitangent = bitangents[i];
            bool tangentIsValid = false;
            if (tangent.SqrLength3() > 0.1f)
            {
                tangent = tangent.Normalized3();
                if (Vector4::Cross3(tangent, normal).SqrLength3() > 0.01f)
                {
                    tangent = Vector4::Orthogonalize(tangent, normal);
                    tangentIsValid = true;
                }
            }
            if (!tangentIsValid) math::BuildOrthonormalBasis(normal, tangent, bitangent);
            tangent = tangent.Normalized3();
            desc.tangents[3 * i] = tangent.x; desc.tangents[3 * i + 1] = tangent.y; desc.tangents[3 * i + 2] = tangent.z;
        }
    }

    // materials from the .mtl (LoadMaterial, MeshLoader.cpp:78-94) or the default material (96-109)
    for (const ObjMaterial& m : obj.materials)
    {
        if (m.hasTextureOptions) { Fail(err, "Mesh '" + filePath + "': .mtl texture options (-blendu, -s, ...) are not supported"); return nullptr; }
        auto material = std::make_shared<Material>();
        material->SetBsdf("diffuse");
        material->debugName = m.name;
        material->baseColor = Vector4(m.diffuse[0], m.diffuse[1], m.diffuse[2], 0.0f);
        material->emission = Vector4(m.emission[0], m.emission[1], m.emission[2], 0.0f);
        {
            // LoadMaterial, MeshLoader.cpp:86-88: textures relative to the mesh's directory
            const std::string meshBaseDir = filePath.substr(0, filePath.find_last_of("\\/")) + "/";
            std::string terr;
            material->baseColorTexture = LoadTexture(meshBaseDir, m.diffuseTexname, terr);
            material->normalMap = LoadTexture(meshBaseDir, m.normalTexname, terr);
            material->maskMap = LoadTexture(meshBaseDir, m.alphaTexname, terr);
            if (!terr.empty()) { Fail(err, "Mesh '" + filePath + "': " + terr); return nullptr; }
        }
        material->roughness = 0.075f;
        material->Compile();
        desc.materials.push_back(material);
        outMaterials[material->debugName] = material;
    }
    if (obj.materials.empty())
    {
        auto material = std::make_shared<Material>();
        material->debugName = "default";
        material->baseColor = Vector4(0.8f, 0.8f, 0.8f, 0.0f);
        material->emission = Vector4(0.0f, 0.0f, 0.0f, 0.0f);
        material->roughness = 0.75f;
        material->SetBsdf(Material::DefaultBsdfName);
        material->Compile();
        outMaterials[material->debugName] = material;
        desc.materials.push_back(material);
        for (uint32_t& index : desc.materialIndices) index = 0;
    }

    auto mesh = std::make_shared<MeshShape>();
    if (!mesh->Initialize(desc)) { Fail(err, "Failed to initialize mesh '" + filePath + "'"); return nullptr; }
    return mesh;
}

// ---- JSON scene -----------------------------------------------------------------------------------------------------

inline bool ParseVector(const json::Value& v, size_t n, Vector4& out, LoadError& err)
{
    if (!v.IsArray()) return Fail(err, n == 2 ? "2D vector description must be an array" : "3D vector description must be an array");
    if (v.Size() != n) return Fail(err, n == 2 ? "Invalid array size for 2D vector" : "Invalid array size for 3D vector");
    out = Vector4(v.array[0].GetFloat(), v.array[1].GetFloat(), n == 3 ? v.array[2].GetFloat() : 0.0f, 0.0f);
    return true;
}

inline bool TryParseBool(const json::Value& v, const char* name, bool optional, bool& out, LoadError& err)
{
    if (!v.HasMember(name)) return optional ? true : Fail(err, std::string("Missing '") + name + "' property");
    if (!v[name].IsBool()) return Fail(err, std::string("Property '") + name + "' must be a bool");
    out = v[name].boolean;
    return true;
}

inline bool TryParseFloat(const json::Value& v, const char* name, bool optional, float& out, LoadError& err)
{
    if (!v.HasMember(name)) return optional ? true : Fail(err, std::string("Missing '") + name + "' property");
    if (!v[name].IsFloat()) return Fail(err, std::string("Property '") + name + "' must be a float");
    out = v[name].GetFloat();
    return true;
}

inline bool TryParseVector(const json::Value& v, const char* name, size_t n, bool optional, Vector4& out, LoadError& err)
{
    if (!v.HasMember(name)) return optional ? true : Fail(err, std::string("Missing '") + name + "' property");
    return ParseVector(v[name], n, out, err);
}

// TryParseTransform, SceneLoader.cpp:180-207: translation + Euler orientation in degrees
inline bool TryParseTransform(const json::Value& parent, const char* name, Vector4& translation, Vector4& rotation, LoadError& err)
{
    translation = Vector4(0.0f);
    rotation = Vector4(0.0f, 0.0f, 0.0f, 1.0f);
    if (!parent.HasMember(name)) return true;
    const json::Value& v = parent[name];
    if (!v.IsObject()) return Fail(err, "Transform description must be a structure");
    if (!TryParseVector(v, "translation", 3, true, translation, err)) return false;
    Vector4 orientation(0.0f);
    if (!TryParseVector(v, "orientation", 3, true, orientation, err)) return false;
    orientation = orientation * (math::RT_PI_F / 180.0f);
    rotation = math::QuaternionFromEulerAngles(orientation.x, orientation.y, orientation.z);
    return true;
}

// TryParseTextureName, SceneLoader.cpp:218-242: a name defined under "textures", else a bitmap file under the data path;
// a file that cannot be loaded leaves outValue empty and is NOT an error
inline bool TryParseTextureName(const json::Value& v, const char* name, const std::string& dataPath, const TexturesMap& textures, TexturePtr& outValue, LoadError& err)
{
    if (!v.HasMember(name)) return true;
    if (!v[name].IsString()) return Fail(err, std::string("Texture path '") + name + "' must be a string");
    const std::string textureName = v[name].string;
    const auto it = textures.find(textureName);
    if (it != textures.end()) { outValue = it->second; return true; }
    std::string terr;
    outValue = LoadTexture(dataPath, textureName, terr);
    if (!terr.empty()) return Fail(err, terr);
    return true;
}

inline bool TryParseInt(const json::Value& v, const char* name, bool optional, int32_t& out, LoadError& err)
{
    if (!v.HasMember(name)) return optional ? true : Fail(err, std::string("Missing '") + name + "' property");
    if (!v[name].IsInt()) return Fail(err, std::string("Property '") + name + "' must be an integer");
    out = v[name].GetInt();
    return true;
}

// ParseTexture, SceneLoader.cpp:269-362
inline TexturePtr ParseTexture(const json::Value& v, const std::string& dataPath, const TexturesMap& textures, std::string& outName, LoadError& err)
{
    if (!v.IsObject()) { Fail(err, "Texture description must be a structure"); return nullptr; }
    if (!v.HasMember("name")) { Fail(err, "Texture is missing 'name' field"); return nullptr; }
    const std::string name = v["name"].string;
    if (name.empty()) { Fail(err, "Texture name cannot be empty"); return nullptr; }
    outName = name;
    if (!v.HasMember("type")) { Fail(err, "Texture is missing 'type' field"); return nullptr; }
    const std::string type = v["type"].string;
    if (type.empty()) { Fail(err, "Texture type cannot be empty"); return nullptr; }
    if (type == "bitmap")
    {
        if (!v.HasMember("path")) { Fail(err, "Texture is missing 'path' field"); return nullptr; }
        std::string terr;
        BitmapPtr bitmap = LoadBitmapObject(dataPath, v["path"].string, terr);
        if (!terr.empty()) { Fail(err, terr); return nullptr; }
        if (!bitmap || bitmap->GetWidth() == 0 || bitmap->GetHeight() == 0) { Fail(err, "Failed to load source image from file '" + dataPath + v["path"].string + "'"); return nullptr; }
        return std::make_shared<BitmapTexture>(bitmap);
    }
    if (type == "checkerboard" || type == "noise")
    {
        Vector4 colorA(0.0f), colorB(0.0f);
        if (!TryParseVector(v, "colorA", 3, false, colorA, err)) return nullptr;
        if (!TryParseVector(v, "colorB", 3, false, colorB, err)) return nullptr;
        if (type == "checkerboard") return std::make_shared<CheckerboardTexture>(colorA, colorB);
        int32_t numOctaves = 1;
        if (!TryParseInt(v, "octaves", true, numOctaves, err)) return nullptr;
        numOctaves = numOctaves > 20 ? 20 : (numOctaves < 1 ? 1 : numOctaves);
        return std::make_shared<NoiseTexture>(colorA, colorB, (uint32_t)numOctaves);
    }
    if (type == "mix")
    {
        TexturePtr texA, texB, texWeight;
        if (!TryParseTextureName(v, "textureA", dataPath, textures, texA, err)) return nullptr;
        if (!TryParseTextureName(v, "textureB", dataPath, textures, texB, err)) return nullptr;
        if (!TryParseTextureName(v, "weight", dataPath, textures, texWeight, err)) return nullptr;
        // MixTexture asserts all three (MixTexture.cpp:15-17); a release build of the reference would crash on the first lookup
        if (!texA || !texB || !texWeight) { Fail(err, "Mix texture '" + name + "' needs 'textureA', 'textureB' and 'weight'"); return nullptr; }
        return std::make_shared<MixTexture>(texA, texB, texWeight);
    }
    Fail(err, "Invalid texture type name: '" + type + "'");
    return nullptr;
}

inline MaterialPtr ParseMaterial(const json::Value& v, const std::string& dataPath, const TexturesMap& textures, LoadError& err)
{
    if (!v.IsObject()) { Fail(err, "Material description must be a structure"); return nullptr; }
    if (!v.HasMember("name")) { Fail(err, "Material is missing 'name' field"); return nullptr; }
    const std::string name = v["name"].string;
    if (name.empty()) { Fail(err, "Material name cannot be empty"); return nullptr; }
    std::string bsdfName = Material::DefaultBsdfName;
    if (v.HasMember("bsdf")) bsdfName = v["bsdf"].string;
    auto m = std::make_shared<Material>();
    m->debugName = name;
    if (!m->SetBsdf(bsdfName)) { Fail(err, "Unknown BSDF name: '" + bsdfName + "'"); return nullptr; }
    if (!TryParseBool(v, "dispersive", true, m->isDispersive, err)) return nullptr;
    if (!TryParseVector(v, "baseColor", 3, true, m->baseColor, err)) return nullptr;
    if (!TryParseVector(v, "emissionColor", 3, true, m->emission, err)) return nullptr;
    if (!TryParseFloat(v, "roughness", true, m->roughness, err)) return nullptr;
    if (!TryParseFloat(v, "metalness", true, m->metalness, err)) return nullptr;
    if (!TryParseTextureName(v, "baseColorTexture", dataPath, textures, m->baseColorTexture, err)) return nullptr;
    if (!TryParseTextureName(v, "emissionTexture", dataPath, textures, m->emissionTexture, err)) return nullptr;
    if (!TryParseTextureName(v, "roughnessTexture", dataPath, textures, m->roughnessTexture, err)) return nullptr;
    if (!TryParseTextureName(v, "metalnessTexture", dataPath, textures, m->metalnessTexture, err)) return nullptr;
    if (!TryParseTextureName(v, "normalMap", dataPath, textures, m->normalMap, err)) return nullptr;
    if (!TryParseTextureName(v, "maskMap", dataPath, textures, m->maskMap, err)) return nullptr;
    if (!TryParseFloat(v, "normalMapStrength", true, m->normalMapStrength, err)) return nullptr;
    if (!TryParseFloat(v, "IoR", true, m->IoR, err)) return nullptr;
    if (!TryParseFloat(v, "K", true, m->K, err)) return nullptr;
    m->Compile();
    return m;
}

inline ShapePtr ParseShape(const json::Value& v, const std::string& dataPath, MaterialsMap& materials, LoadError& err)
{
    if (!v.HasMember("type")) { Fail(err, "Object is missing 'type' field"); return nullptr; }
    const std::string type = v["type"].string;
    if (type == "sphere")
    {
        float radius = 1.0f;
        if (!TryParseFloat(v, "radius", false, radius, err)) return nullptr;
        return std::make_shared<SphereShape>(radius);
    }
    if (type == "box")
    {
        Vector4 size;
        if (!TryParseVector(v, "size", 3, false, size, err)) return nullptr;
        return std::make_shared<BoxShape>(size);
    }
    if (type == "rect" || type == "plane")
    {
        Vector4 size(FLT_MAX);
        if (!TryParseVector(v, "size", 2, false, size, err)) return nullptr;
        Vector4 textureScale(1.0f);
        if (!TryParseVector(v, "textureScale", 2, true, textureScale, err)) return nullptr;
        return std::make_shared<RectShape>(size.x, size.y, textureScale.x, textureScale.y);
    }
    if (type == "mesh")
    {
        if (!v.HasMember("path")) { Fail(err, "Missing 'path' property"); return nullptr; }
        if (!v["path"].IsString()) { Fail(err, "Mesh path must be a string"); return nullptr; }
        float scale = 1.0f;
        if (!TryParseFloat(v, "scale", true, scale, err)) return nullptr;
        return LoadMesh(dataPath + v["path"].string, materials, scale, err);
    }
    Fail(err, "Unknown scene object type: '" + type + "'" + (type == "csg" ? " (the reference's CSG shape is a placeholder and is not supported)" : ""));
    return nullptr;
}

inline bool ParseObject(const json::Value& v, const std::string& dataPath, Scene& scene, MaterialsMap& materials, LoadError& err)
{
    if (!v.IsObject()) return Fail(err, "Object description must be a structure");
    ShapePtr shape = ParseShape(v, dataPath, materials, err);
    if (!shape) return Fail(err, "invalid shape");
    auto object = std::make_unique<ShapeSceneObject>(shape);
    MaterialPtr material;
    if (v.HasMember("material"))
    {
        if (!v["material"].IsString()) return Fail(err, "Material name 'material' must be a string");
        const auto it = materials.find(v["material"].string);
        if (it == materials.end()) return Fail(err, "Material '" + v["material"].string + "' does not exist");
        material = it->second;
    }
    object->SetDefaultMaterial(material);
    Vector4 t, q;
    if (!TryParseTransform(v, "transform", t, q, err)) return false;
    object->SetTransform(math::TransformToMatrix4(t, q));
    scene.AddObject(std::move(object));
    return true;
}

inline bool ParseLight(const json::Value& v, const std::string& dataPath, Scene& scene, const TexturesMap& textures, LoadError& err)
{
    if (!v.IsObject()) return Fail(err, "Light description must be a structure");
    if (!v.HasMember("type")) return Fail(err, "Light is missing 'type' field");
    Vector4 color;
    if (!TryParseVector(v, "color", 3, false, color, err)) return false;
    LightPtr light;
    const std::string type = v["type"].string;
    if (type == "area")
    {
        if (!v.HasMember("shape")) return Fail(err, "Area light is missing 'shape' field");
        MaterialsMap dummy;
        ShapePtr shape = ParseShape(v["shape"], dataPath, dummy, err);
        if (!shape) return false;
        if (shape->GetType() == RTB_SHAPE_MESH) return Fail(err, "mesh area lights are not supported (MeshShape::Sample is unimplemented in the reference)");
        auto areaLight = std::make_unique<AreaLight>(std::move(shape), color);
        if (!TryParseTextureName(v, "texture", dataPath, textures, areaLight->mTexture, err)) return false;
        light = std::move(areaLight);
    }
    else if (type == "point") light = std::make_unique<PointLight>(color);
    else if (type == "spot")
    {
        float angle = 0.0f;
        if (!TryParseFloat(v, "angle", true, angle, err)) return false;
        light = std::make_unique<SpotLight>(color, angle / 180.0f * math::RT_PI_F);
    }
    else if (type == "directional")
    {
        float angle = 0.0f;
        if (!TryParseFloat(v, "angle", true, angle, err)) return false;
        light = std::make_unique<DirectionalLight>(color, math::DegToRad(angle));
    }
    else if (type == "background")
    {
        auto backgroundLight = std::make_unique<BackgroundLight>(color);
        if (!TryParseTextureName(v, "texture", dataPath, textures, backgroundLight->mTexture, err)) return false;
        light = std::move(backgroundLight);
    }
    else return Fail(err, "Unknown light type: '" + type + "'");   // incl. "sphere": that loader branch never creates a light (SceneLoader.cpp:587-597)
    auto object = std::make_unique<LightSceneObject>(std::move(light));
    Vector4 t, q;
    if (!TryParseTransform(v, "transform", t, q, err)) return false;
    object->SetTransform(math::TransformToMatrix4(t, q));
    scene.AddObject(std::move(object));
    return true;
}

inline bool ParseCamera(const json::Value& v, const std::string& dataPath, const TexturesMap& textures, Camera& camera, LoadError& err)
{
    if (!v.IsObject()) return Fail(err, "Camera description must be a structure");
    Vector4 t, q;
    if (!TryParseTransform(v, "transform", t, q, err)) return false;
    float fov = 60.0f;
    if (!TryParseFloat(v, "fieldOfView", true, fov, err)) return false;
    camera.SetTransform(t, q);
    camera.SetPerspective(1.0f, math::DegToRad(fov));
    if (!TryParseBool(v, "enableDOF", true, camera.mDOF.enable, err)) return false;
    if (!TryParseFloat(v, "aperture", true, camera.mDOF.aperture, err)) return false;
    if (!TryParseFloat(v, "focalPlaneDistance", true, camera.mDOF.focalPlaneDistance, err)) return false;
    if (!TryParseTextureName(v, "bokehTexture", dataPath, textures, camera.mDOF.bokehTexture, err)) return false;
    return true;
}

// helpers::LoadScene, SceneLoader.cpp:692-820
inline bool LoadScene(const std::string& path, const std::string& dataPath, Scene& scene, Camera& camera, std::string& errorMessage)
{
    LoadError err;
    std::ifstream f(path, std::ios::binary);
    if (!f) { errorMessage = "Failed to open file: " + path; return false; }
    std::stringstream ss;
    ss << f.rdbuf();
    json::Value d;
    json::Parser parser;
    std::string perr;
    if (!parser.Parse(ss.str(), d, perr) || !d.IsObject()) { errorMessage = "Failed to parse scene file '" + path + "': " + perr; return false; }

    MaterialsMap materials;
    bool ok = true;
    TexturesMap textures;
    if (d.HasMember("textures"))
    {
        const json::Value& a = d["textures"];
        if (!a.IsArray()) ok = Fail(err, "'textures' is expected to be an array");
        for (size_t i = 0; ok && i < a.Size(); ++i)
        {
            std::string name;
            TexturePtr t = ParseTexture(a[i], dataPath, textures, name, err);
            if (!t) { ok = Fail(err, "invalid texture"); break; }
            if (textures.count(name) > 0) { ok = Fail(err, "Duplicated texture: '" + name + "'"); break; }
            textures[name] = t;
        }
    }
    if (ok && d.HasMember("materials"))
    {
        const json::Value& a = d["materials"];
        if (!a.IsArray()) ok = Fail(err, "'materials' is expected to be an array");
        for (size_t i = 0; ok && i < a.Size(); ++i)
        {
            MaterialPtr m = ParseMaterial(a[i], dataPath, textures, err);
            if (!m) { ok = false; break; }
            if (materials.count(m->debugName) > 0) { ok = Fail(err, "Duplicated material: '" + m->debugName + "'"); break; }
            materials[m->debugName] = m;
        }
    }
    if (ok && d.HasMember("objects"))
    {
        const json::Value& a = d["objects"];
        if (!a.IsArray()) ok = Fail(err, "'objects' is expected to be an array");
        for (size_t i = 0; ok && i < a.Size(); ++i) ok = ParseObject(a[i], dataPath, scene, materials, err);
    }
    if (ok && d.HasMember("lights"))
    {
        const json::Value& a = d["lights"];
        if (!a.IsArray()) ok = Fail(err, "'lights' is expected to be an array");
        for (size_t i = 0; ok && i < a.Size(); ++i) ok = ParseLight(a[i], dataPath, scene, textures, err);
    }
    if (ok && d.HasMember("camera")) ok = ParseCamera(d["camera"], dataPath, textures, camera, err);
    if (!ok) errorMessage = err.message;
    return ok;
}

} // namespace helpers
} // namespace rt
