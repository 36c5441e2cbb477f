// ref_flatten.h - walks the REFERENCE's Scene object (after Scene::BuildBVH) and Camera and fills the POD scene description
// of include/rtb200.h: the SceneFlattener of INTEGRATION.md section 2, working on the reference's real classes.
//
// TEST INFRASTRUCTURE (oracle/): compiled only into oracle/_ref/*/rtref (the oracle harness) and oracle/_ref/*/rtadapter
// (the reference-side adapter test), both built by oracle/build_ref.sh against the patched copy of the reference with
// -fno-access-control.  The including file provides the reference headers and `using namespace rt; using namespace rt::math;`.
#pragma once

void Die(const char* msg) { fprintf(stderr, "rtref: %s\n", msg); exit(1); }

template <typename T> void WritePod(FILE* f, const T& v) { if (fwrite(&v, sizeof(T), 1, f) != 1) Die("write failed"); }
void WriteU32(FILE* f, uint32_t v) { WritePod(f, v); }
void WriteBytes(FILE* f, const void* p, size_t n) { if (n && fwrite(p, 1, n, f) != n) Die("write failed"); }

RtbMat4 ToMat(const Matrix4& m)
{
    // Matrix4 is declared RT_ALIGN(32) but scene objects are allocated by Aligned<16>::operator new, so the compiler's
    // 32-byte aligned AVX copy faults here; hide the alignment from the optimizer and copy bytes.
    const char* p = reinterpret_cast<const char*>(&m);
    asm volatile("" : "+r"(p));
    RtbMat4 r;
    memcpy(&r, p, sizeof(r));
    return r;
}

void ToF4(const Vector4& v, float* out) { out[0] = v.x; out[1] = v.y; out[2] = v.z; out[3] = v.w; }

uint32_t BsdfId(const Material* m)
{
    const char* n = m->GetBSDF() ? m->GetBSDF()->GetName() : "null";
    const char* names[] = { "null", "diffuse", "roughDiffuse", "dielectric", "roughDielectric", "metal", "roughMetal", "plastic", "roughPlastic" };
    for (uint32_t i = 0; i < 9; ++i) if (!strcmp(n, names[i])) return i;
    Die("unknown BSDF");
    return 0;
}

void CopyNodes(const BVH& bvh, std::vector<RtbBvhNode>& out)
{
    static_assert(sizeof(BVH::Node) == sizeof(RtbBvhNode), "node layout");
    out.resize(bvh.GetNumNodes());
    if (!out.empty()) memcpy(out.data(), bvh.GetNodes(), sizeof(RtbBvhNode) * out.size());
    // node 1 is never written by BVHBuilder (root takes two slots, BVHBuilder.cpp:86) -> uninitialised memory
    if (out.size() > 1) memset(&out[1], 0, sizeof(RtbBvhNode));
}

// flattened scene, owning storage
struct Flat
{
    std::vector<RtbBvhNode> tlas;
    std::vector<RtbObject> objects;
    std::vector<RtbShape> shapes;
    std::vector<RtbMaterial> materials;
    std::vector<RtbLight> lights;
    std::vector<uint32_t> globalLights;
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

    std::map<const IShape*, uint32_t> shapeIds;
    std::map<const Material*, uint32_t> materialIds;

    // textures (RTBS version 2 section)
    struct Texture { RtbTexture desc; std::vector<float> texels; };
    std::vector<Texture> textures;
    std::vector<RtbMaterialTextures> materialTextures;
    std::vector<uint32_t> lightTextures;
    std::map<const ITexture*, uint32_t> textureIds;
    bool anyTexture = false;

    uint32_t AddTexture(const ITexture* t)
    {
        if (!t) return RTB_NO_TEXTURE;
        auto it = textureIds.find(t);
        if (it != textureIds.end()) return it->second;
        Texture r;
        memset(&r.desc, 0, sizeof(r.desc));
        if (const BitmapTexture* bt = dynamic_cast<const BitmapTexture*>(t))
        {
            const Bitmap* b = bt->mBitmap.get();
            if (!b) Die("bitmap texture without bitmap");
            if (bt->mFilter != BitmapTextureFilter::Bilinear_SmoothStep) Die("unexpected bitmap filter");
            r.desc.type = RTB_TEXTURE_BITMAP;
            r.desc.width = b->GetWidth(); r.desc.height = b->GetHeight();
            r.texels.resize((size_t)r.desc.width * r.desc.height * 4);
            // the texels are DEFINED as what the reference's own Bitmap::GetPixel returns
            for (uint32_t y = 0; y < r.desc.height; ++y)
                for (uint32_t x = 0; x < r.desc.width; ++x)
                    ToF4(b->GetPixel(x, y), &r.texels[4 * ((size_t)y * r.desc.width + x)]);
        }
        else if (const CheckerboardTexture* ct = dynamic_cast<const CheckerboardTexture*>(t))
        {
            r.desc.type = RTB_TEXTURE_CHECKERBOARD;
            ToF4(ct->mColorA, r.desc.colorA); ToF4(ct->mColorB, r.desc.colorB);
        }
        else if (const NoiseTexture* nt = dynamic_cast<const NoiseTexture*>(t))
        {
            r.desc.type = RTB_TEXTURE_NOISE;
            ToF4(nt->mColorA, r.desc.colorA); ToF4(nt->mColorB, r.desc.colorB);
            r.desc.numOctaves = nt->mNumOctaves;
        }
        else if (const MixTexture* mt = dynamic_cast<const MixTexture*>(t))
        {
            r.desc.type = RTB_TEXTURE_MIX;
            r.desc.texA = AddTexture(mt->mTextureA.get());
            r.desc.texB = AddTexture(mt->mTextureB.get());
            r.desc.texWeight = AddTexture(mt->mTextureMask.get());
        }
        else Die("unsupported texture class");
        const uint32_t id = (uint32_t)textures.size();
        textures.push_back(std::move(r));
        textureIds[t] = id;
        anyTexture = true;
        return id;
    }

    uint32_t AddMaterial(const Material* m)
    {
        auto it = materialIds.find(m);
        if (it != materialIds.end()) return it->second;
        RtbMaterial r;
        memset(&r, 0, sizeof(r));
        r.bsdf = BsdfId(m);
        if (m->baseColor.texture || m->emission.texture || m->roughness.texture || m->metalness.texture || m->normalMap || m->maskMap) r.flags |= 1u; // textured
        RtbMaterialTextures mt;
        memset(&mt, 0, sizeof(mt));
        mt.baseColor = AddTexture(m->baseColor.texture.get());
        mt.emission = AddTexture(m->emission.texture.get());
        mt.roughness = AddTexture(m->roughness.texture.get());
        mt.metalness = AddTexture(m->metalness.texture.get());
        mt.normalMap = AddTexture(m->normalMap.get());
        mt.maskMap = AddTexture(m->maskMap.get());
        mt.normalMapStrength = m->normalMapStrength;
        materialTextures.push_back(mt);
        if (m->isDispersive) r.flags |= 2u;
        r.roughness = m->roughness.baseValue;
        r.metalness = m->metalness.baseValue;
        ToF4(m->baseColor.baseValue, r.baseColor);
        ToF4(m->emission.baseValue, r.emission);
        r.IoR = m->IoR;
        r.K = m->K;
        const uint32_t id = (uint32_t)materials.size();
        materials.push_back(r);
        materialIds[m] = id;
        return id;
    }

    uint32_t AddShape(const IShape* s)
    {
        auto it = shapeIds.find(s);
        if (it != shapeIds.end()) return it->second;
        RtbShape r;
        memset(&r, 0, sizeof(r));
        r.meshIndex = UINT32_MAX;
        if (const SphereShape* sp = dynamic_cast<const SphereShape*>(s))
        {
            r.type = RTB_SHAPE_SPHERE;
            r.size[0] = sp->mRadius;
            r.invSize[0] = sp->mInvRadius;
        }
        else if (const BoxShape* bx = dynamic_cast<const BoxShape*>(s))
        {
            r.type = RTB_SHAPE_BOX;
            ToF4(bx->mSize, r.size);
            ToF4(bx->mInvSize, r.invSize);
            r.aux[0] = bx->mFaceCdf.x; r.aux[1] = bx->mFaceCdf.y; r.aux[2] = bx->mFaceCdf.z;
        }
        else if (const RectShape* rc = dynamic_cast<const RectShape*>(s))
        {
            r.type = RTB_SHAPE_RECT;
            r.size[0] = rc->mSize.x; r.size[1] = rc->mSize.y;
            r.aux[0] = rc->mTextureScale.x; r.aux[1] = rc->mTextureScale.y;
            if (rc->mEnableSolidAngleSampling) Die("rect solid angle sampling not supported");
        }
        else if (const MeshShape* ms = dynamic_cast<const MeshShape*>(s))
        {
            r.type = RTB_SHAPE_MESH;
            r.meshIndex = (uint32_t)meshes.size();
            Mesh m;
            CopyNodes(ms->mBVH, m.nodes);
            const VertexBuffer& vb = ms->mVertexBuffer;
            const uint32_t nt = vb.mNumTriangles, nv = vb.mNumVertices;
            m.numVertices = nv;
            m.triangles.resize((size_t)nt * 9);
            static_assert(sizeof(ProcessedTriangle) == 36, "tri layout");
            if (nt) memcpy(m.triangles.data(), vb.mPreprocessedTriangles, (size_t)nt * 36);
            m.indices.resize((size_t)nt * 4);
            if (nt) memcpy(m.indices.data(), vb.mBuffer + vb.mVertexIndexBufferOffset, (size_t)nt * 16);
            m.shading.resize((size_t)nv * 8);
            if (nv) memcpy(m.shading.data(), vb.mBuffer + vb.mShadingDataBufferOffset, (size_t)nv * 32);
            for (uint32_t i = 0; i < vb.mMaterials.Size(); ++i) m.materialIds.push_back(AddMaterial(vb.mMaterials[i].get()));
            meshes.push_back(std::move(m));
        }
        else
        {
            Die("unsupported shape type (csg?)");
        }
        const uint32_t id = (uint32_t)shapes.size();
        shapes.push_back(r);
        shapeIds[s] = id;
        return id;
    }

    void Build(const Scene& scene)
    {
        CopyNodes(scene.mTraceableObjectsBVH, tlas);

        std::map<const LightSceneObject*, uint32_t> lightIds;
        for (const LightSceneObject* lo : scene.mLights)
        {
            const ILight& l = lo->GetLight();
            RtbLight r;
            memset(&r, 0, sizeof(r));
            r.type = (uint32_t)l.GetType();
            r.flags = (uint32_t)l.GetFlags();
            r.shapeIndex = UINT32_MAX;
            ToF4(l.GetColor().rgbValues, r.color);
            if (const AreaLight* al = dynamic_cast<const AreaLight*>(&l)) { r.shapeIndex = AddShape(al->GetShape().get()); }
            uint32_t lightTexture = RTB_NO_TEXTURE;
            if (const BackgroundLight* bl = dynamic_cast<const BackgroundLight*>(&l)) lightTexture = AddTexture(bl->mTexture.get());
            lightTextures.push_back(lightTexture);     // (AreaLight::mTexture is never evaluated: AreaLight.cpp:49-58, 136-146)
            if (const DirectionalLight* dl = dynamic_cast<const DirectionalLight*>(&l)) { r.cosAngle = dl->mCosAngle; r.isDelta = dl->mIsDelta; }
            if (const SpotLight* sl = dynamic_cast<const SpotLight*>(&l)) { r.cosAngle = sl->mCosAngle; r.isDelta = sl->mIsDelta; }
            r.transform = ToMat(lo->GetBaseTransform());
            r.invTransform = ToMat(lo->GetBaseInverseTransform());
            lightIds[lo] = (uint32_t)lights.size();
            lights.push_back(r);
        }
        for (const LightSceneObject* lo : scene.mGlobalLights) globalLights.push_back(lightIds.at(lo));

        for (const ITraceableSceneObject* obj : scene.mTraceableObjects)
        {
            RtbObject r;
            memset(&r, 0, sizeof(r));
            r.shapeIndex = r.materialIndex = r.lightIndex = UINT32_MAX;
            r.transform = ToMat(obj->GetBaseTransform());
            r.invTransform = ToMat(obj->GetBaseInverseTransform());
            if (const ShapeSceneObject* so = dynamic_cast<const ShapeSceneObject*>(obj))
            {
                r.kind = RTB_OBJECT_SHAPE;
                r.shapeIndex = AddShape(so->mShape.get());
                r.materialIndex = AddMaterial(so->mDefaultMaterial.get());
            }
            else if (const LightSceneObject* lo = dynamic_cast<const LightSceneObject*>(obj))
            {
                r.kind = RTB_OBJECT_LIGHT;
                r.lightIndex = lightIds.at(lo);
                r.shapeIndex = lights[r.lightIndex].shapeIndex;
            }
            else Die("unsupported traceable object");
            objects.push_back(r);
        }
    }

    void Write(const char* path, const Camera& camera) const
    {
        FILE* f = fopen(path, "wb");
        if (!f) Die("cannot open output");
        WriteU32(f, 0x53425452u); // 'RTBS'
        {
            const BitmapTexture* bk = dynamic_cast<const BitmapTexture*>(camera.mDOF.bokehTexture.get());
            WriteU32(f, (anyTexture || (bk && bk->mImportanceMap)) ? 2 : 1);
        }
        WriteU32(f, (uint32_t)tlas.size());      WriteBytes(f, tlas.data(), tlas.size() * sizeof(RtbBvhNode));
        WriteU32(f, (uint32_t)objects.size());   WriteBytes(f, objects.data(), objects.size() * sizeof(RtbObject));
        WriteU32(f, (uint32_t)shapes.size());    WriteBytes(f, shapes.data(), shapes.size() * sizeof(RtbShape));
        WriteU32(f, (uint32_t)meshes.size());
        for (const Mesh& m : meshes)
        {
            WriteU32(f, (uint32_t)m.nodes.size());
            WriteU32(f, (uint32_t)(m.triangles.size() / 9));
            WriteU32(f, m.numVertices);
            WriteU32(f, (uint32_t)m.materialIds.size());
            WriteBytes(f, m.nodes.data(), m.nodes.size() * sizeof(RtbBvhNode));
            WriteBytes(f, m.triangles.data(), m.triangles.size() * 4);
            WriteBytes(f, m.indices.data(), m.indices.size() * 4);
            WriteBytes(f, m.shading.data(), m.shading.size() * 4);
            WriteBytes(f, m.materialIds.data(), m.materialIds.size() * 4);
        }
        WriteU32(f, (uint32_t)materials.size()); WriteBytes(f, materials.data(), materials.size() * sizeof(RtbMaterial));
        WriteU32(f, (uint32_t)lights.size());    WriteBytes(f, lights.data(), lights.size() * sizeof(RtbLight));
        WriteU32(f, (uint32_t)globalLights.size()); WriteBytes(f, globalLights.data(), globalLights.size() * 4);

        RtbCamera c;
        memset(&c, 0, sizeof(c));
        c.localToWorld = ToMat(camera.mLocalToWorld);
        c.aspectRatio = camera.mAspectRatio;
        c.tanHalfFoV = camera.mTanHalfFoV;
        c.dofEnable = camera.mDOF.enable;
        c.bokehShape = (uint32_t)camera.mDOF.bokehShape;
        c.aperture = camera.mDOF.aperture;
        c.focalPlaneDistance = camera.mDOF.focalPlaneDistance;
        c.barrelDistortionConstFactor = camera.barrelDistortionConstFactor;
        c.barrelDistortionVariableFactor = camera.barrelDistortionVariableFactor;
        WritePod(f, c);
        WritePod(f, camera.mFieldOfView);
        // texture-shaped bokeh: the importance map BitmapTexture::MakeSamplable built (its Distribution's cumulative table)
        const BitmapTexture* bokeh = dynamic_cast<const BitmapTexture*>(camera.mDOF.bokehTexture.get());
        const bool hasBokeh = bokeh && bokeh->mImportanceMap;
        if (anyTexture || hasBokeh)
        {
            WriteU32(f, (uint32_t)textures.size());
            for (const Texture& t : textures)
            {
                const uint32_t h[8] = { t.desc.type, t.desc.width, t.desc.height, t.desc.numOctaves, t.desc.texA, t.desc.texB, t.desc.texWeight, 0u };
                WriteBytes(f, h, sizeof(h)); WriteBytes(f, t.desc.colorA, 16); WriteBytes(f, t.desc.colorB, 16);
                WriteBytes(f, t.texels.data(), t.texels.size() * 4);
            }
            WriteU32(f, (uint32_t)materialTextures.size()); WriteBytes(f, materialTextures.data(), materialTextures.size() * sizeof(RtbMaterialTextures));
            WriteU32(f, (uint32_t)lightTextures.size()); WriteBytes(f, lightTextures.data(), lightTextures.size() * 4);
            if (hasBokeh)
            {
                WriteU32(f, bokeh->mBitmap->GetWidth()); WriteU32(f, bokeh->mBitmap->GetHeight());
                WriteBytes(f, bokeh->mImportanceMap->mCDF, ((size_t)bokeh->mImportanceMap->mSize + 1) * 4);
            }
        }
        fclose(f);
    }
};

