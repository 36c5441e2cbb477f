// rt_scene.h — the plugin surface: the reference's Scene / Camera / Material / Light / Shape object model, kept as the
// way inputs enter the B200 path (same class names, same construction rules), plus Scene::Flatten which produces the
// POD RtbSceneDesc that rtb_upload_scene consumes.
//
//   Material (+Compile)            Core/Material/Material.h:30-110, Material.cpp:110-121
//   IShape / Sphere / Box / Rect   Core/Shapes/*.cpp (constructors + GetBoundingBox)
//   MeshShape + VertexBuffer       Core/Shapes/MeshShape.cpp:34-112, Core/Shapes/Mesh/VertexBuffer.cpp:52-185
//   ILight x5                      Core/Scene/Light/*.cpp (constructors, GetBoundingBox, GetFlags)
//   ISceneObject / Shape / Light   Core/Scene/Object/SceneObject*.cpp
//   Scene::AddObject / BuildBVH    Core/Scene/Scene.cpp:26-126
//   Camera                         Core/Scene/Camera.cpp:12-49
#pragma once

#include <memory>
#include <string>
#include <vector>
#include <map>
#include <functional>

#include "rt_math.h"
#include "rt_bvh.h"
#include "rt_texture.h"
#include "../csrc/rtb_scene_file.h"

namespace rt {

using math::Vector4;
using math::Matrix4;
using math::Box;

// ---- Material --------------------------------------------------------------------------------------------------------

class Material
{
public:
    std::string debugName = "<unnamed>";
    Vector4 emission = Vector4(0.0f, 0.0f, 0.0f, 0.0f);
    Vector4 baseColor = Vector4(0.7f, 0.7f, 0.7f, 0.0f);
    float roughness = 0.1f;
    float metalness = 0.0f;
    float normalMapStrength = 1.0f;
    float IoR = 1.5f;
    float K = 4.0f;
    bool isDispersive = false;
    // MaterialParameter<T>::texture (MaterialParameter.h:14-15) + Material::maskMap / normalMap (Material.h:76-77)
    TexturePtr baseColorTexture, emissionTexture, roughnessTexture, metalnessTexture, normalMap, maskMap;

    static constexpr const char* DefaultBsdfName = "diffuse";

    // returns false for an unknown name (the reference logs an error and leaves the BSDF unset, Material.cpp:40-83)
    bool SetBsdf(const std::string& name)
    {
        static const char* names[] = { "null", "diffuse", "roughDiffuse", "dielectric", "roughDielectric", "metal", "roughMetal", "plastic", "roughPlastic" };
        for (uint32_t i = 0; i < 9; ++i) if (name == names[i]) { mBsdf = i; return true; }
        return false;
    }
    uint32_t GetBsdfId() const { return mBsdf; }

    // Material::Compile, Material.cpp:110-121
    void Compile()
    {
        emission = Vector4::Max(Vector4(0.0f), emission);
        baseColor = Vector4::Max(Vector4(0.0f), Vector4::Min(Vector4(1.0f), baseColor));
    }

    static std::shared_ptr<Material> GetDefaultMaterial()
    {
        static std::shared_ptr<Material> m = [] {
            auto p = std::make_shared<Material>();
            p->debugName = "default";
            p->SetBsdf(DefaultBsdfName);
            p->Compile();
            return p;
        }();
        return m;
    }

private:
    uint32_t mBsdf = RTB_BSDF_NULL;
};
using MaterialPtr = std::shared_ptr<Material>;

// ---- Shapes ----------------------------------------------------------------------------------------------------------

class IShape
{
public:
    virtual ~IShape() = default;
    virtual uint32_t GetType() const = 0;
    virtual Box GetBoundingBox() const = 0;
    virtual void Describe(RtbShape& out) const = 0;
};
using ShapePtr = std::shared_ptr<IShape>;

class SphereShape : public IShape
{
public:
    explicit SphereShape(float radius) : mRadius(radius), mInvRadius(1.0f / radius) {}
    uint32_t GetType() const override { return RTB_SHAPE_SPHERE; }
    Box GetBoundingBox() const override { return { Vector4(0.0f) - Vector4(mRadius), Vector4(0.0f) + Vector4(mRadius) }; }   // Box(center, radius)
    void Describe(RtbShape& o) const override { o.type = RTB_SHAPE_SPHERE; o.size[0] = mRadius; o.invSize[0] = mInvRadius; }
private:
    float mRadius, mInvRadius;
};

class BoxShape : public IShape
{
public:
    explicit BoxShape(const Vector4& size) : mSize(size), mInvSize(Vector4(1.0f) / size)
    {
        mSize.w = 0.0f; mInvSize.w = 0.0f;
        mFaceCdf[0] = mSize.y * mSize.z;
        mFaceCdf[1] = mFaceCdf[0] + mSize.z * mSize.x;
        mFaceCdf[2] = mFaceCdf[1] + mSize.x * mSize.y;
    }
    uint32_t GetType() const override { return RTB_SHAPE_BOX; }
    Box GetBoundingBox() const override { return { Vector4(0.0f) - mSize, mSize }; }
    void Describe(RtbShape& o) const override
    {
        o.type = RTB_SHAPE_BOX;
        for (int i = 0; i < 4; ++i) { o.size[i] = mSize[i]; o.invSize[i] = mInvSize[i]; }
        o.aux[0] = mFaceCdf[0]; o.aux[1] = mFaceCdf[1]; o.aux[2] = mFaceCdf[2];
    }
private:
    Vector4 mSize, mInvSize;
    float mFaceCdf[3];
};

class RectShape : public IShape
{
public:
    RectShape(float sx, float sy, float tu = 1.0f, float tv = 1.0f) : mSizeX(sx), mSizeY(sy), mTexU(tu), mTexV(tv) {}
    uint32_t GetType() const override { return RTB_SHAPE_RECT; }
    Box GetBoundingBox() const override { return { Vector4(-mSizeX, -mSizeY, 0.0f, 0.0f), Vector4(mSizeX, mSizeY, 0.0f, 0.0f) }; }
    void Describe(RtbShape& o) const override { o.type = RTB_SHAPE_RECT; o.size[0] = mSizeX; o.size[1] = mSizeY; o.aux[0] = mTexU; o.aux[1] = mTexV; }
private:
    float mSizeX, mSizeY, mTexU, mTexV;
};

// VertexBufferDesc (Core/Shapes/Mesh/VertexBufferDesc.h) with owned storage
struct MeshDesc
{
    std::vector<float> positions;       // 3 per vertex
    std::vector<float> normals;         // 3 per vertex
    std::vector<float> tangents;        // 3 per vertex
    std::vector<float> texCoords;       // 2 per vertex
    std::vector<uint32_t> indices;      // 3 per triangle
    std::vector<uint32_t> materialIndices;   // 1 per triangle (UINT32_MAX = object default)
    std::vector<MaterialPtr> materials;
    std::string path;
};

class MeshShape : public IShape
{
public:
    // MeshShape::Initialize, MeshShape.cpp:34-112 (+ VertexBuffer::Initialize, VertexBuffer.cpp:52-185)
    bool Initialize(const MeshDesc& desc)
    {
        const uint32_t numTriangles = (uint32_t)(desc.indices.size() / 3);
        const uint32_t numVertices = (uint32_t)(desc.positions.size() / 3);
        mBoundingBox = Box::Empty();
        std::vector<Box> boxes(numTriangles);
        auto pos = [&](uint32_t i) { return Vector4(desc.positions[3 * i], desc.positions[3 * i + 1], desc.positions[3 * i + 2], 0.0f); };
        for (uint32_t i = 0; i < numTriangles; ++i)
        {
            const Vector4 v0 = pos(desc.indices[3 * i]), v1 = pos(desc.indices[3 * i + 1]), v2 = pos(desc.indices[3 * i + 2]);
            Box b;
            b.min = Vector4::Min(v0, Vector4::Min(v1, v2));
            b.max = Vector4::Max(v0, Vector4::Max(v1, v2));
            boxes[i] = b;
            mBoundingBox = Box::Merge(mBoundingBox, b);
        }
        BVHBuilder::Indices order;
        BVHBuilder builder(mBVH);
        if (!builder.Build(boxes.data(), numTriangles, BvhBuildingParams(), order)) return false;

        // triangles in leaf order; {v0, v1 - v0, v2 - v0}
        mTriangles.resize((size_t)numTriangles * 9);
        mIndices.resize((size_t)numTriangles * 4);
        for (uint32_t i = 0; i < numTriangles; ++i)
        {
            const uint32_t src = order[i];
            const uint32_t i0 = desc.indices[3 * src], i1 = desc.indices[3 * src + 1], i2 = desc.indices[3 * src + 2];
            const Vector4 v0 = pos(i0), v1 = pos(i1), v2 = pos(i2);
            const Vector4 e1 = v1 - v0, e2 = v2 - v0;
            float* t = &mTriangles[(size_t)i * 9];
            t[0] = v0.x; t[1] = v0.y; t[2] = v0.z; t[3] = e1.x; t[4] = e1.y; t[5] = e1.z; t[6] = e2.x; t[7] = e2.y; t[8] = e2.z;
            uint32_t* ix = &mIndices[(size_t)i * 4];
            ix[0] = i0; ix[1] = i1; ix[2] = i2; ix[3] = desc.materialIndices[src];
        }
        mShading.resize((size_t)numVertices * 8, 0.0f);
        for (uint32_t v = 0; v < numVertices; ++v)
        {
            float* s = &mShading[(size_t)v * 8];
            if (!desc.normals.empty()) { s[0] = desc.normals[3 * v]; s[1] = desc.normals[3 * v + 1]; s[2] = desc.normals[3 * v + 2]; }
            if (!desc.tangents.empty()) { s[3] = desc.tangents[3 * v]; s[4] = desc.tangents[3 * v + 1]; s[5] = desc.tangents[3 * v + 2]; }
            if (!desc.texCoords.empty()) { s[6] = desc.texCoords[2 * v]; s[7] = desc.texCoords[2 * v + 1]; }
        }
        mMaterials = desc.materials;
        mNumVertices = numVertices;
        return true;
    }

    uint32_t GetType() const override { return RTB_SHAPE_MESH; }
    Box GetBoundingBox() const override { return mBoundingBox; }
    void Describe(RtbShape& o) const override { o.type = RTB_SHAPE_MESH; }

    BVH mBVH;
    std::vector<float> mTriangles;
    std::vector<uint32_t> mIndices;
    std::vector<float> mShading;
    std::vector<MaterialPtr> mMaterials;
    uint32_t mNumVertices = 0;
    Box mBoundingBox;
};
using MeshShapePtr = std::shared_ptr<MeshShape>;

// ---- Lights ----------------------------------------------------------------------------------------------------------

class ILight
{
public:
    enum Flags : uint32_t { Flag_None = 0, Flag_IsFinite = 1, Flag_IsDelta = 2 };
    explicit ILight(const Vector4& color) : mColor(color) {}
    virtual ~ILight() = default;
    virtual uint32_t GetType() const = 0;
    virtual Box GetBoundingBox() const = 0;
    virtual uint32_t GetFlags() const = 0;
    const Vector4& GetColor() const { return mColor; }
private:
    Vector4 mColor;
};
using LightPtr = std::unique_ptr<ILight>;

class AreaLight : public ILight
{
public:
    AreaLight(ShapePtr shape, const Vector4& color) : ILight(color), mShape(std::move(shape)) {}
    uint32_t GetType() const override { return RTB_LIGHT_AREA; }
    Box GetBoundingBox() const override { return mShape->GetBoundingBox(); }
    uint32_t GetFlags() const override { return Flag_IsFinite; }
    const ShapePtr& GetShape() const { return mShape; }
    TexturePtr mTexture;      // parsed like the reference does, never evaluated (AreaLight.cpp:49-58, 136-146 are commented out)
private:
    ShapePtr mShape;
};

class BackgroundLight : public ILight
{
public:
    explicit BackgroundLight(const Vector4& color) : ILight(color) {}
    uint32_t GetType() const override { return RTB_LIGHT_BACKGROUND; }
    Box GetBoundingBox() const override { return Box::Full(); }
    uint32_t GetFlags() const override { return Flag_None; }
    TexturePtr mTexture;      // environment map, BackgroundLight.cpp:45-61
};

class DirectionalLight : public ILight
{
public:
    static constexpr float CosEpsilon = 0.9999f;    // ILight::CosEpsilon, Light.h:25
    DirectionalLight(const Vector4& color, float angle) : ILight(color) { mCosAngle = cosf(angle); mIsDelta = mCosAngle > CosEpsilon; }
    uint32_t GetType() const override { return RTB_LIGHT_DIRECTIONAL; }
    Box GetBoundingBox() const override { return Box::Full(); }
    uint32_t GetFlags() const override { return mIsDelta ? Flag_IsDelta : Flag_None; }
    float mCosAngle;
    bool mIsDelta;
};

class PointLight : public ILight
{
public:
    explicit PointLight(const Vector4& color) : ILight(color) {}
    uint32_t GetType() const override { return RTB_LIGHT_POINT; }
    Box GetBoundingBox() const override { return { Vector4(0.0f) - Vector4(0.0f), Vector4(0.0f) + Vector4(0.0f) }; }
    uint32_t GetFlags() const override { return Flag_IsFinite | Flag_IsDelta; }
};

class SpotLight : public ILight
{
public:
    SpotLight(const Vector4& color, float angle) : ILight(color) { mCosAngle = cosf(angle); mIsDelta = mCosAngle > DirectionalLight::CosEpsilon; }
    uint32_t GetType() const override { return RTB_LIGHT_SPOT; }
    Box GetBoundingBox() const override { return { Vector4(0.0f) - Vector4(0.0f), Vector4(0.0f) + Vector4(0.0f) }; }
    uint32_t GetFlags() const override { return mIsDelta ? (Flag_IsFinite | Flag_IsDelta) : Flag_IsFinite; }
    float mCosAngle;
    bool mIsDelta;
};

// ---- Scene objects ---------------------------------------------------------------------------------------------------

class ISceneObject
{
public:
    enum class Type { Shape, Light };
    virtual ~ISceneObject() = default;
    virtual Type GetType() const = 0;
    virtual Box GetLocalBoundingBox() const = 0;
    // ISceneObject::SetTransform, SceneObject.cpp:16-24
    void SetTransform(const Matrix4& m) { mTransform = m; mInverseTransform = m.Inverse(); }
    const Matrix4& GetBaseTransform() const { return mTransform; }
    const Matrix4& GetBaseInverseTransform() const { return mInverseTransform; }
    // {TransformBox(local), TransformBox(local)} merged (SceneObject_Shape.cpp:27-31)
    Box GetBoundingBox() const
    {
        const Box local = GetLocalBoundingBox();
        return Box::Merge(mTransform.TransformBox(local), mTransform.TransformBox(local));
    }
private:
    Matrix4 mTransform = Matrix4::Identity();
    Matrix4 mInverseTransform = Matrix4::Identity();
};
using SceneObjectPtr = std::unique_ptr<ISceneObject>;

class ShapeSceneObject : public ISceneObject
{
public:
    explicit ShapeSceneObject(ShapePtr shape) : mShape(std::move(shape)) {}
    Type GetType() const override { return Type::Shape; }
    Box GetLocalBoundingBox() const override { return mShape->GetBoundingBox(); }
    void SetDefaultMaterial(const MaterialPtr& m) { mDefaultMaterial = m ? m : Material::GetDefaultMaterial(); }
    ShapePtr mShape;
    MaterialPtr mDefaultMaterial;
};

class LightSceneObject : public ISceneObject
{
public:
    explicit LightSceneObject(LightPtr light) : mLight(std::move(light)) {}
    Type GetType() const override { return Type::Light; }
    Box GetLocalBoundingBox() const override { return mLight->GetBoundingBox(); }
    const ILight& GetLight() const { return *mLight; }
private:
    LightPtr mLight;
};

// ---- Camera ----------------------------------------------------------------------------------------------------------

class Camera
{
public:
    float mAspectRatio = 1.0f;
    float mFieldOfView = math::DegToRad(20.0f);
    struct DOF
    {
        float focalPlaneDistance = 2.0f; float aperture = 0.1f; bool enable = false; uint32_t bokehShape = RTB_BOKEH_CIRCLE;
        TexturePtr bokehTexture;      // Camera.h:49; selects BokehShape::Texture (SceneLoader.cpp:680-687)
    } mDOF;
    float barrelDistortionConstFactor = 0.01f;
    float barrelDistortionVariableFactor = 0.0f;

    void SetTransform(const Vector4& translation, const Vector4& rotationQuat) { mLocalToWorld = math::TransformToMatrix4(translation, rotationQuat); }
    // Camera::SetPerspective, Camera.cpp:30-49
    void SetPerspective(float aspectRatio, float fov) { mAspectRatio = aspectRatio; mFieldOfView = fov; mTanHalfFoV = tanf(mFieldOfView * 0.5f); }
    const Matrix4& GetLocalToWorld() const { return mLocalToWorld; }

    void Describe(RtbCamera& c) const
    {
        memset(&c, 0, sizeof(c));
        for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) c.localToWorld.r[i][j] = mLocalToWorld.rows[i][j];
        c.aspectRatio = mAspectRatio; c.tanHalfFoV = mTanHalfFoV;
        c.dofEnable = mDOF.enable; c.bokehShape = mDOF.bokehShape; c.aperture = mDOF.aperture; c.focalPlaneDistance = mDOF.focalPlaneDistance;
        c.barrelDistortionConstFactor = barrelDistortionConstFactor; c.barrelDistortionVariableFactor = barrelDistortionVariableFactor;
        if (mDOF.bokehTexture)
        {
            // a bitmap is sampled through its importance map (DescribeBokeh); the procedural textures' Sample() returns the
            // sample itself as coordinates (CheckerboardTexture.cpp:43-55, NoiseTexture.cpp:157-169, MixTexture.cpp:33-45), and
            // UnipolarToBipolar of that is exactly the square bokeh (Camera.cpp:203-204)
            c.bokehShape = mDOF.bokehTexture->GetType() == RTB_TEXTURE_BITMAP ? RTB_BOKEH_TEXTURE : RTB_BOKEH_SQUARE;
        }
    }

    // the bokeh bitmap's importance map goes with the scene upload (RtbSceneDesc::bokehCdf)
    void DescribeBokeh(rtb::SceneFile& out) const
    {
        out.bokehWidth = out.bokehHeight = 0; out.bokehCdf.clear();
        if (!mDOF.bokehTexture || mDOF.bokehTexture->GetType() != RTB_TEXTURE_BITMAP) return;
        Bitmap& b = *static_cast<const BitmapTexture*>(mDOF.bokehTexture.get())->GetBitmap();
        out.bokehWidth = b.GetWidth(); out.bokehHeight = b.GetHeight();
        out.bokehCdf = b.GetImportanceCdf();
    }
private:
    float mTanHalfFoV = tanf(math::DegToRad(20.0f) * 0.5f);
    Matrix4 mLocalToWorld = Matrix4::Identity();
};

// ---- Scene -----------------------------------------------------------------------------------------------------------

class Scene
{
public:
    void AddObject(SceneObjectPtr object) { mAllObjects.push_back(std::move(object)); }

    // Scene::BuildBVH, Scene.cpp:36-126: classify lights, build the object BVH, reorder the traceable objects
    bool BuildBVH()
    {
        mTraceableObjects.clear(); mLights.clear(); mGlobalLights.clear();
        for (const auto& object : mAllObjects)
        {
            if (object->GetType() == ISceneObject::Type::Light)
            {
                const LightSceneObject* lo = static_cast<const LightSceneObject*>(object.get());
                mLights.push_back(lo);
                if (lo->GetLight().GetFlags() & ILight::Flag_IsFinite) mTraceableObjects.push_back(object.get());
                else mGlobalLights.push_back(lo);
            }
            else mTraceableObjects.push_back(object.get());
        }
        std::vector<Box> boxes;
        for (const ISceneObject* o : mTraceableObjects) boxes.push_back(o->GetBoundingBox());
        BVHBuilder::Indices order;
        BVHBuilder builder(mBVH);
        if (!builder.Build(boxes.data(), (uint32_t)boxes.size(), BvhBuildingParams(), order)) return false;
        std::vector<const ISceneObject*> reordered;
        for (uint32_t i = 0; i < mTraceableObjects.size(); ++i) reordered.push_back(mTraceableObjects[order[i]]);
        mTraceableObjects = std::move(reordered);
        return true;
    }

    const BVH& GetBVH() const { return mBVH; }
    const std::vector<const LightSceneObject*>& GetLights() const { return mLights; }
    const std::vector<const LightSceneObject*>& GetGlobalLights() const { return mGlobalLights; }
    size_t GetNumTraceableObjects() const { return mTraceableObjects.size(); }

    // POD view of the scene for rtb_upload_scene (what oracle/rtref.cpp's Flat does for the reference's Scene)
    void Flatten(rtb::SceneFile& out) const
    {
        out = rtb::SceneFile();
        out.tlas = mBVH.nodes;
        std::map<const IShape*, uint32_t> shapeIds;
        std::map<const Material*, uint32_t> materialIds;

        // textures are emitted children first, so a mix texture only references earlier entries
        std::map<const ITexture*, uint32_t> textureIds;
        std::function<uint32_t(const ITexture*)> addTexture = [&](const ITexture* t) -> uint32_t
        {
            if (!t) return RTB_NO_TEXTURE;
            auto it = textureIds.find(t);
            if (it != textureIds.end()) return it->second;
            rtb::SceneFile::Texture r;
            memset(&r.desc, 0, sizeof(r.desc));
            r.desc.type = t->GetType();
            if (t->GetType() == RTB_TEXTURE_BITMAP)
            {
                const Bitmap& b = *static_cast<const BitmapTexture*>(t)->GetBitmap();
                r.desc.width = b.GetWidth(); r.desc.height = b.GetHeight();
                r.texels = b.GetTexels();
            }
            else if (t->GetType() == RTB_TEXTURE_CHECKERBOARD)
            {
                const CheckerboardTexture* c = static_cast<const CheckerboardTexture*>(t);
                for (int i = 0; i < 4; ++i) { r.desc.colorA[i] = c->mColorA[i]; r.desc.colorB[i] = c->mColorB[i]; }
            }
            else if (t->GetType() == RTB_TEXTURE_NOISE)
            {
                const NoiseTexture* n = static_cast<const NoiseTexture*>(t);
                for (int i = 0; i < 4; ++i) { r.desc.colorA[i] = n->mColorA[i]; r.desc.colorB[i] = n->mColorB[i]; }
                r.desc.numOctaves = n->mNumOctaves;
            }
            else
            {
                const MixTexture* m = static_cast<const MixTexture*>(t);
                r.desc.texA = addTexture(m->mTextureA.get());
                r.desc.texB = addTexture(m->mTextureB.get());
                r.desc.texWeight = addTexture(m->mTextureMask.get());
            }
            const uint32_t id = (uint32_t)out.textures.size();
            out.textures.push_back(std::move(r));
            textureIds[t] = id;
            return id;
        };
        auto addMaterial = [&](const Material* m) -> uint32_t
        {
            auto it = materialIds.find(m);
            if (it != materialIds.end()) return it->second;
            RtbMaterial r;
            memset(&r, 0, sizeof(r));
            r.bsdf = m->GetBsdfId();
            if (m->baseColorTexture || m->emissionTexture || m->roughnessTexture || m->metalnessTexture || m->normalMap || m->maskMap) r.flags |= 1u;
            RtbMaterialTextures mt;
            memset(&mt, 0, sizeof(mt));
            mt.baseColor = addTexture(m->baseColorTexture.get());
            mt.emission = addTexture(m->emissionTexture.get());
            mt.roughness = addTexture(m->roughnessTexture.get());
            mt.metalness = addTexture(m->metalnessTexture.get());
            mt.normalMap = addTexture(m->normalMap.get());
            mt.maskMap = addTexture(m->maskMap.get());
            mt.normalMapStrength = m->normalMapStrength;
            out.materialTextures.push_back(mt);
            if (m->isDispersive) r.flags |= 2u;
            r.roughness = m->roughness; r.metalness = m->metalness;
            for (int i = 0; i < 4; ++i) { r.baseColor[i] = m->baseColor[i]; r.emission[i] = m->emission[i]; }
            r.IoR = m->IoR; r.K = m->K;
            const uint32_t id = (uint32_t)out.materials.size();
            out.materials.push_back(r);
            materialIds[m] = id;
            return id;
        };
        auto addShape = [&](const IShape* s) -> uint32_t
        {
            auto it = shapeIds.find(s);
            if (it != shapeIds.end()) return it->second;
            RtbShape r;
            memset(&r, 0, sizeof(r));
            r.meshIndex = 0xFFFFFFFFu;
            s->Describe(r);
            if (r.type == RTB_SHAPE_MESH)
            {
                const MeshShape* ms = static_cast<const MeshShape*>(s);
                r.meshIndex = (uint32_t)out.meshes.size();
                rtb::SceneFile::Mesh m;
                m.nodes = ms->mBVH.nodes;
                m.triangles = ms->mTriangles;
                m.indices = ms->mIndices;
                m.shading = ms->mShading;
                m.numVertices = ms->mNumVertices;
                for (const MaterialPtr& mp : ms->mMaterials) m.materialIds.push_back(addMaterial(mp.get()));
                out.meshes.push_back(std::move(m));
            }
            const uint32_t id = (uint32_t)out.shapes.size();
            out.shapes.push_back(r);
            shapeIds[s] = id;
            return id;
        };
        auto toMat = [](const Matrix4& m) { RtbMat4 r; for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) r.r[i][j] = m.rows[i][j]; return r; };

        std::map<const LightSceneObject*, uint32_t> lightIds;
        for (const LightSceneObject* lo : mLights)
        {
            const ILight& l = lo->GetLight();
            RtbLight r;
            memset(&r, 0, sizeof(r));
            r.type = l.GetType(); r.flags = l.GetFlags(); r.shapeIndex = 0xFFFFFFFFu;
            for (int i = 0; i < 4; ++i) r.color[i] = l.GetColor()[i];
            if (l.GetType() == RTB_LIGHT_AREA) r.shapeIndex = addShape(static_cast<const AreaLight&>(l).GetShape().get());
            out.lightTextures.push_back(l.GetType() == RTB_LIGHT_BACKGROUND ? addTexture(static_cast<const BackgroundLight&>(l).mTexture.get()) : RTB_NO_TEXTURE);
            if (l.GetType() == RTB_LIGHT_DIRECTIONAL) { r.cosAngle = static_cast<const DirectionalLight&>(l).mCosAngle; r.isDelta = static_cast<const DirectionalLight&>(l).mIsDelta; }
            if (l.GetType() == RTB_LIGHT_SPOT) { r.cosAngle = static_cast<const SpotLight&>(l).mCosAngle; r.isDelta = static_cast<const SpotLight&>(l).mIsDelta; }
            r.transform = toMat(lo->GetBaseTransform());
            r.invTransform = toMat(lo->GetBaseInverseTransform());
            lightIds[lo] = (uint32_t)out.lights.size();
            out.lights.push_back(r);
        }
        for (const LightSceneObject* lo : mGlobalLights) out.globalLights.push_back(lightIds.at(lo));

        for (const ISceneObject* obj : mTraceableObjects)
        {
            RtbObject r;
            memset(&r, 0, sizeof(r));
            r.shapeIndex = r.materialIndex = r.lightIndex = 0xFFFFFFFFu;
            r.transform = toMat(obj->GetBaseTransform());
            r.invTransform = toMat(obj->GetBaseInverseTransform());
            if (obj->GetType() == ISceneObject::Type::Shape)
            {
                const ShapeSceneObject* so = static_cast<const ShapeSceneObject*>(obj);
                r.kind = RTB_OBJECT_SHAPE;
                r.shapeIndex = addShape(so->mShape.get());
                r.materialIndex = addMaterial(so->mDefaultMaterial.get());
            }
            else
            {
                const LightSceneObject* lo = static_cast<const LightSceneObject*>(obj);
                r.kind = RTB_OBJECT_LIGHT;
                r.lightIndex = lightIds.at(lo);
                r.shapeIndex = out.lights[r.lightIndex].shapeIndex;
            }
            out.objects.push_back(r);
        }
        // a scene without any texture stays an RTBS v1 file / a texture-less RtbSceneDesc
        if (out.textures.empty()) { out.materialTextures.clear(); out.lightTextures.clear(); }
    }

private:
    std::vector<SceneObjectPtr> mAllObjects;
    std::vector<const LightSceneObject*> mLights, mGlobalLights;
    std::vector<const ISceneObject*> mTraceableObjects;
    BVH mBVH;
};

} // namespace rt
