// rt_texture.h — host side of the reference's texture objects (Core/Textures/*.h) and of its Bitmap loader, kept as
// the plugin surface: same class names, same construction rules, same failure behaviour.  Nothing is evaluated on
// the host: Scene::Flatten turns these objects into the POD RtbTexture array that rtb_upload_scene copies to HBM.
//
//   Bitmap (BMP subset)          Core/Utils/Bitmap.cpp:197-333 (Init, Load), Core/Utils/BitmapBMP.cpp:46-139 (LoadBMP),
//                                Bitmap::GetPixel Core/Utils/Bitmap.cpp:335-518, Convert_sRGB_To_Linear Color/ColorHelpers.h:17-27
//   BitmapTexture                Core/Textures/BitmapTexture.cpp:17-21 (filter = Bilinear_SmoothStep)
//   CheckerboardTexture          Core/Textures/CheckerboardTexture.cpp:9-25
//   NoiseTexture                 Core/Textures/NoiseTexture.cpp:44-52
//   MixTexture                   Core/Textures/MixTexture.cpp:8-17
//
// Bitmap decodes into the float quadruples the reference's Bitmap::GetPixel returns for every texel (format decode +
// sRGB -> linear for gamma-space files), which is what the device samples.  DDS and OpenEXR containers are
// recognised and REFUSED with an error (the reference would load them; silently dropping the texture would render a
// different image); any other unreadable file yields "no texture" exactly like the reference (Bitmap::Load fails,
// helpers::LoadTexture returns nullptr, Demo/MeshLoader.cpp:31-76).
#pragma once

#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "rt_math.h"
#include "../../include/rtb200.h"

namespace rt {

using math::Vector4;

class Bitmap
{
public:
    uint32_t GetWidth() const { return mWidth; }
    uint32_t GetHeight() const { return mHeight; }
    const std::vector<float>& GetTexels() const { return mTexels; }     // width*height*4, texel (x,y) = GetPixel(x,y)
    const std::string& GetError() const { return mError; }              // non-empty: the file is valid but unsupported here

    // Bitmap::Load, Bitmap.cpp:298-333 (LoadBMP, then LoadDDS, then LoadEXR)
    bool Load(const char* path)
    {
        FILE* file = fopen(path, "rb");
        if (!file) return false;                         // "Failed to load source image from file"
        const bool ok = LoadBMP(file);
        if (!ok && mError.empty())
        {
            // the reference would now try its DDS and OpenEXR readers
            uint8_t magic[4] = { 0, 0, 0, 0 };
            fseek(file, 0, SEEK_SET);
            if (fread(magic, 1, 4, file) == 4)
            {
                if (!memcmp(magic, "DDS ", 4)) mError = std::string("DDS bitmaps are not supported by this build: '") + path + "'";
                if (magic[0] == 0x76 && magic[1] == 0x2F && magic[2] == 0x31 && magic[3] == 0x01) mError = std::string("OpenEXR bitmaps are not supported by this build: '") + path + "'";
            }
        }
        fclose(file);
        return ok;
    }

private:
    static float Unorm8(uint8_t b) { return static_cast<float>(b) * (1.0f / 255.0f); }     // every 8-bit path of GetPixel reduces to this (see header)
    // Convert_sRGB_To_Linear, ColorHelpers.h:17-27 (two fused multiply-adds and a multiply, per lane)
    static float SrgbToLinear(float c)
    {
        float r = std::fmaf(c, 0.305306011f, 0.682171111f);
        r = std::fmaf(c, r, 0.012522878f);
        return r * c;
    }

#pragma pack(push, 2)
    struct BitmapFileHeader { uint16_t bfType; uint32_t bfSize; uint16_t bfReserved1, bfReserved2; uint32_t bfOffBits; };
    struct BitmapInfoHeader
    {
        uint32_t biSize, biWidth, biHeight; uint16_t biPlanes, biBitCount; uint32_t biCompression, biSizeImage, biXPelsPerMeter, biYPelsPerMeter, biClrUsed, biClrImportant;
    };
#pragma pack(pop)

    // Bitmap::LoadBMP, BitmapBMP.cpp:46-139: 24-bit BGR, 8-bit palette (biClrUsed > 0) or 8-bit grey (biClrUsed == 0),
    // uncompressed, rows padded to 4 bytes, kept in FILE order (row 0 of the file is texel row 0), gamma space
    bool LoadBMP(FILE* file)
    {
        BitmapFileHeader fh;
        if (fread(&fh, sizeof(fh), 1, file) != 1) return false;
        if (fh.bfType != 0x4D42) return false;
        BitmapInfoHeader ih;
        if (fread(&ih, sizeof(ih), 1, file) != 1) return false;
        if (ih.biPlanes != 1 || ih.biCompression != 0) return false;
        enum { kBGR, kPalette, kGrey } format;
        uint32_t bytesPerPixel;
        if (ih.biBitCount == 24) { format = kBGR; bytesPerPixel = 3; }
        else if (ih.biBitCount == 8 && ih.biClrUsed > 0) { format = kPalette; bytesPerPixel = 1; }
        else if (ih.biBitCount == 8 && ih.biClrUsed == 0) { format = kGrey; bytesPerPixel = 1; }
        else return false;
        if (ih.biWidth == 0 || ih.biHeight == 0) return false;
        if (ih.biWidth > 65536u || ih.biHeight > 65536u) { mError = "bitmap larger than 65536 texels per side"; return false; }
        const uint32_t stride = (ih.biWidth * bytesPerPixel + 3u) & ~3u;
        std::vector<uint8_t> palette((size_t)ih.biClrUsed * 4u);
        if (!palette.empty() && fread(palette.data(), palette.size(), 1, file) != 1) return false;
        if (fseek(file, (long)fh.bfOffBits, SEEK_SET) != 0) return false;
        std::vector<uint8_t> data((size_t)stride * ih.biHeight);
        if (fread(data.data(), data.size(), 1, file) != 1) return false;

        mWidth = ih.biWidth; mHeight = ih.biHeight;
        mTexels.resize((size_t)mWidth * mHeight * 4);
        for (uint32_t y = 0; y < mHeight; ++y)
        {
            const uint8_t* row = data.data() + (size_t)stride * y;
            for (uint32_t x = 0; x < mWidth; ++x)
            {
                float c[4];
                if (format == kBGR)
                {
                    const uint8_t* s = row + 3u * (size_t)x;                 // Vector4_LoadBGR_UNorm: w lane = 0
                    c[0] = Unorm8(s[2]); c[1] = Unorm8(s[1]); c[2] = Unorm8(s[0]); c[3] = 0.0f;
                }
                else if (format == kPalette)
                {
                    const size_t index = row[x];                             // B8G8R8A8_UNorm_Palette: swizzle <2,1,0,3>
                    uint8_t s[4] = { 0, 0, 0, 0 };
                    if (4u * index + 4u <= palette.size()) memcpy(s, palette.data() + 4u * index, 4);
                    c[0] = Unorm8(s[2]); c[1] = Unorm8(s[1]); c[2] = Unorm8(s[0]); c[3] = Unorm8(s[3]);
                }
                else
                {
                    c[0] = c[1] = c[2] = c[3] = Unorm8(row[x]);              // R8_UNorm: splat
                }
                float* out = &mTexels[4 * ((size_t)y * mWidth + x)];
                for (int k = 0; k < 4; ++k) out[k] = SrgbToLinear(c[k]);     // initData.linearSpace = false
            }
        }
        return true;
    }

    uint32_t mWidth = 0, mHeight = 0;
    std::vector<float> mTexels;
    std::string mError;
};
using BitmapPtr = std::shared_ptr<Bitmap>;

// ---- ITexture ---------------------------------------------------------------------------------------------------

class ITexture
{
public:
    virtual ~ITexture() = default;
    virtual const char* GetName() const = 0;
    virtual uint32_t GetType() const = 0;
};
using TexturePtr = std::shared_ptr<ITexture>;

class BitmapTexture : public ITexture
{
public:
    explicit BitmapTexture(const BitmapPtr& bitmap) : mBitmap(bitmap) {}
    const char* GetName() const override { return "bitmap"; }
    uint32_t GetType() const override { return RTB_TEXTURE_BITMAP; }
    const BitmapPtr& GetBitmap() const { return mBitmap; }
private:
    BitmapPtr mBitmap;
};

class CheckerboardTexture : public ITexture
{
public:
    CheckerboardTexture(const Vector4& colorA, const Vector4& colorB) : mColorA(colorA), mColorB(colorB) {}
    const char* GetName() const override { return "checkerboard"; }
    uint32_t GetType() const override { return RTB_TEXTURE_CHECKERBOARD; }
    Vector4 mColorA, mColorB;
};

class NoiseTexture : public ITexture
{
public:
    NoiseTexture(const Vector4& colorA, const Vector4& colorB, uint32_t numOctaves) : mColorA(colorA), mColorB(colorB), mNumOctaves(numOctaves) {}
    const char* GetName() const override { return "noise"; }
    uint32_t GetType() const override { return RTB_TEXTURE_NOISE; }
    Vector4 mColorA, mColorB;
    uint32_t mNumOctaves;
};

class MixTexture : public ITexture
{
public:
    MixTexture(const TexturePtr& a, const TexturePtr& b, const TexturePtr& mask) : mTextureA(a), mTextureB(b), mTextureMask(mask) {}
    const char* GetName() const override { return "mix"; }
    uint32_t GetType() const override { return RTB_TEXTURE_MIX; }
    TexturePtr mTextureA, mTextureB, mTextureMask;
};

namespace helpers {

// LoadBitmapObject, Demo/MeshLoader.cpp:31-58: ".png" / ".jpg" are redirected to ".bmp", bitmaps are cached by full path.
// `error` is set when the file exists in a format the reference reads but this build does not (the scene load then fails).
inline BitmapPtr LoadBitmapObject(const std::string& baseDir, const std::string& path, std::string& error)
{
    if (path.empty()) return nullptr;
    std::string fullPath = baseDir + path;
    if (fullPath.length() >= 4 && (fullPath.rfind(".png") == fullPath.length() - 4 || fullPath.rfind(".jpg") == fullPath.length() - 4))
        fullPath.replace(fullPath.length() - 4, 4, ".bmp");
    static std::map<std::string, BitmapPtr> bitmapsList;
    BitmapPtr& bitmapPtr = bitmapsList[fullPath];
    if (!bitmapPtr)
    {
        bitmapPtr = std::make_shared<Bitmap>();
        if (!bitmapPtr->Load(fullPath.c_str()))
        {
            if (!bitmapPtr->GetError().empty()) error = bitmapPtr->GetError();
            return nullptr;
        }
    }
    if (!bitmapPtr->GetError().empty()) error = bitmapPtr->GetError();
    return bitmapPtr;
}

// helpers::LoadTexture, Demo/MeshLoader.cpp:60-76: nullptr when the bitmap cannot be loaded or is empty
inline TexturePtr LoadTexture(const std::string& baseDir, const std::string& path, std::string& error)
{
    BitmapPtr bitmap = LoadBitmapObject(baseDir, path, error);
    if (!bitmap) return nullptr;
    if (bitmap->GetWidth() > 0 && bitmap->GetHeight() > 0) return std::make_shared<BitmapTexture>(bitmap);
    return nullptr;
}

} // namespace helpers
} // namespace rt
