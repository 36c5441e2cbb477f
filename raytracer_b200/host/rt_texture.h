// rt_texture.h — host side of the reference's texture objects (Core/Textures/*.h) and of its Bitmap loader, kept as
// the plugin surface: same class names, same construction rules, same failure behaviour.  Nothing is evaluated on
// the host: Scene::Flatten turns these objects into the POD RtbTexture array that rtb_upload_scene copies to HBM.
//
//   Bitmap (BMP + DDS)           Core/Utils/Bitmap.cpp:197-333 (Init, Load), BitmapBMP.cpp:46-139, BitmapDDS.cpp:228-540, BlockCompression.cpp,
//                                Bitmap::GetPixel Core/Utils/Bitmap.cpp:335-518, Convert_sRGB_To_Linear Color/ColorHelpers.h:17-27
//   BitmapTexture                Core/Textures/BitmapTexture.cpp:17-21 (filter = Bilinear_SmoothStep)
//   CheckerboardTexture          Core/Textures/CheckerboardTexture.cpp:9-25
//   NoiseTexture                 Core/Textures/NoiseTexture.cpp:44-52
//   MixTexture                   Core/Textures/MixTexture.cpp:8-17
//
// Bitmap decodes into the float quadruples the reference's Bitmap::GetPixel returns for every texel (format decode +
// sRGB -> linear for gamma-space files), which is what the device samples.  BMP and DDS (Core/Utils/BitmapDDS.cpp:
// uncompressed 8 / 16 / 32-bit integer, half and float layouts, R11G11B10, R9G9B9E5, BC1 / BC4 / BC5 blocks) are read;
// OpenEXR files (which the reference reads through the third-party tinyexr) go through rt_exr.h: scan-line images with
// 3 or 4 half / float channels, uncompressed or RLE / ZIP / ZIPS / PIZ compressed; any unreadable file yields "no
// texture" exactly like the reference (Bitmap::Load fails, helpers::LoadTexture returns nullptr, MeshLoader.cpp:31-76).
#pragma once

#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <new>
#include <string>
#include <vector>

#include "rt_math.h"
#include "rt_exr.h"
#include "../../include/rtb200.h"

namespace rt {

using math::Vector4;

class Bitmap
{
public:
    // the subset of Bitmap::Format (Core/Utils/Bitmap.h:16-42) the BMP and DDS readers can produce
    enum class Format : uint8_t
    {
        Unknown = 0, R8_UNorm, R8G8_UNorm, B8G8R8_UNorm, B8G8R8A8_UNorm, R8G8B8A8_UNorm, B8G8R8A8_UNorm_Palette, B5G6R5_UNorm,
        R16_UNorm, R16G16_UNorm, R16G16B16A16_UNorm, R32_Float, R32G32_Float, R32G32B32_Float, R32G32B32A32_Float,
        R11G11B10_Float, R16_Half, R16G16_Half, R16G16B16_Half, R16G16B16A16_Half, R9G9B9E5_SharedExp, BC1, BC4, BC5,
    };

    uint32_t GetWidth() const { return mWidth; }
    uint32_t GetHeight() const { return mHeight; }
    const std::vector<float>& GetTexels() const { return mTexels; }     // width*height*4, texel (x,y) = GetPixel(x,y)
    const std::string& GetError() const { return mError; }              // non-empty: the file is valid but unsupported here

    // BitmapTexture::MakeSamplable (BitmapTexture.cpp:121-152) + Distribution::Initialize (Math/Distribution.cpp:27-84):
    // importance = Dot3(c_rgbIntensityWeights, GetPixel) per texel, accumulated in float, normalised by 1 / total.
    // Returns the cumulative distribution (width * height + 1 entries) that Distribution::SampleDiscrete searches.
    const std::vector<float>& GetImportanceCdf()
    {
        if (!mImportanceCdf.empty() || mTexels.empty()) return mImportanceCdf;
        const size_t n = (size_t)mWidth * mHeight;
        mImportanceCdf.resize(n + 1);
        float accumulated = 0.0f;
        mImportanceCdf[0] = 0.0f;
        for (size_t i = 0; i < n; ++i)
        {
            const float* c = &mTexels[4 * i];
            const float pdf = (0.21f * c[0] + 0.72f * c[1]) + (0.07f * c[2] + 0.0f);     // dpps 0x7F: (x + y) + (z + 0)
            accumulated += pdf;
            mImportanceCdf[i + 1] = accumulated;
        }
        const float cdfNormFactor = 1.0f / accumulated;
        for (size_t i = 0; i <= n; ++i) mImportanceCdf[i] *= cdfNormFactor;
        return mImportanceCdf;
    }

    // Bitmap::Load, Bitmap.cpp:298-333 (LoadBMP, then LoadDDS, then LoadEXR)
    bool Load(const char* path)
    {
        FILE* file = fopen(path, "rb");
        if (!file) return false;                         // "Failed to load source image from file"
        bool ok = LoadBMP(file);
        if (!ok && mError.empty())
        {
            fseek(file, 0, SEEK_SET);
            ok = LoadDDS(file, path);
        }
        if (!ok && mError.empty()) ok = LoadEXR(path);
        fclose(file);
        if (ok) Decode();
        return ok;
    }

private:
    // ---- raw storage as the reference keeps it (Bitmap::Init, Bitmap.cpp:197-254: data + a zeroed 64-byte margin) ----
    Format mFormat = Format::Unknown;
    bool mLinearSpace = true;
    uint32_t mStride = 0;
    std::vector<uint8_t> mData, mPalette;

    static uint32_t BitsPerPixel(Format f)              // Bitmap::BitsPerPixel, Bitmap.cpp:17-49
    {
        switch (f)
        {
        case Format::R8_UNorm: case Format::B8G8R8A8_UNorm_Palette: return 8;
        case Format::R8G8_UNorm: case Format::B5G6R5_UNorm: case Format::R16_UNorm: case Format::R16_Half: return 16;
        case Format::B8G8R8_UNorm: return 24;
        case Format::B8G8R8A8_UNorm: case Format::R8G8B8A8_UNorm: case Format::R16G16_UNorm: case Format::R32_Float:
        case Format::R11G11B10_Float: case Format::R16G16_Half: case Format::R9G9B9E5_SharedExp: return 32;
        case Format::R16G16B16_Half: return 48;
        case Format::R16G16B16A16_UNorm: case Format::R32G32_Float: case Format::R16G16B16A16_Half: return 64;
        case Format::R32G32B32_Float: return 96;
        case Format::R32G32B32A32_Float: return 128;
        case Format::BC1: case Format::BC4: return 4;
        case Format::BC5: return 8;
        default: return 0;
        }
    }

    bool Init(uint32_t width, uint32_t height, Format format, uint32_t stride, bool linearSpace)
    {
        const uint32_t minStride = (uint32_t)((uint64_t)width * BitsPerPixel(format) / 8);
        mStride = stride > minStride ? stride : minStride;
        const uint64_t dataSize = (uint64_t)height * mStride;
        if (dataSize == 0) return false;
        if (width > 65536u || height > 65536u) { mError = "bitmap larger than 65536 texels per side"; return false; }
        mData.assign((size_t)dataSize + 64, 0);
        mWidth = width; mHeight = height; mFormat = format; mLinearSpace = linearSpace;
        return true;
    }
    size_t DataSize() const { return mData.size() - 64; }

    static float Unorm8(uint32_t b) { return static_cast<float>(b) * (1.0f / 255.0f); }     // every 8-bit path of GetPixel reduces to this
    // Convert_sRGB_To_Linear, ColorHelpers.h:17-27 (two fused multiply-adds and a multiply, per lane)
    static float SrgbToLinear(float c)
    {
        float r = std::fmaf(c, 0.305306011f, 0.682171111f);
        r = std::fmaf(c, r, 0.012522878f);
        return r * c;
    }
    static float BitsToFloat(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
    static uint16_t Load16(const uint8_t* p) { uint16_t v; memcpy(&v, p, 2); return v; }
    static uint32_t Load32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
    static float LoadF32(const uint8_t* p) { float v; memcpy(&v, p, 4); return v; }

    // Half::ToFloat, software path (Math/HalfImpl.h:68-105; the oracle is not built with RT_USE_FP16C)
    static float HalfToFloat(uint16_t value)
    {
        uint32_t mantissa = static_cast<uint32_t>(value & 0x03FF);
        uint32_t exponent = (value & 0x7C00);
        if (exponent == 0x7C00) exponent = 0x8f;
        else if (exponent != 0) exponent = static_cast<uint32_t>((static_cast<int>(value) >> 10) & 0x1F);
        else if (mantissa != 0)
        {
            exponent = 1;
            do { exponent--; mantissa <<= 1; } while ((mantissa & 0x0400) == 0);
            mantissa &= 0x03FF;
        }
        else exponent = static_cast<uint32_t>(-112);
        return BitsToFloat(((static_cast<uint32_t>(value) & 0x8000) << 16) | ((exponent + 112) << 23) | (mantissa << 13));
    }

    // helper::DecodeBC_Grayscale, BlockCompression.cpp:72-100
    static float DecodeBCGrayscale(const uint8_t* blockData, uint32_t x, uint32_t y)
    {
        const float intColor0 = blockData[0], intColor1 = blockData[1];
        const float color0 = intColor0 / 255.0f, color1 = intColor1 / 255.0f;
        uint64_t code; memcpy(&code, blockData + 2, 8);
        const uint32_t index = (uint32_t)((code >> (uint64_t)(3 * (4 * y + x))) % 8);
        if (index == 0) return color0;
        if (index == 1) return color1;
        if (intColor0 > intColor1) return (color0 * (float)(8 - index) + color1 * (float)(index - 1)) / 7.0f;
        if (index == 6) return 0.0f;
        if (index == 7) return 1.0f;
        return (color0 * (float)(6 - index) + color1 * (float)(index - 1)) / 5.0f;
    }

    // Bitmap::GetPixel before the colour-space conversion, Bitmap.cpp:335-506
    void GetRawPixel(uint32_t x, uint32_t y, float c[4]) const
    {
        const uint8_t* row = mData.data() + (size_t)mStride * y;
        c[0] = c[1] = c[2] = c[3] = 0.0f;
        switch (mFormat)
        {
        case Format::R8_UNorm: c[0] = c[1] = c[2] = c[3] = Unorm8(row[x]); break;                      // FromInteger(v) * (1/255): splat
        case Format::R8G8_UNorm: { const uint8_t* s = row + 2u * (size_t)x; c[0] = Unorm8(s[0]); c[1] = Unorm8(s[1]); break; }
        case Format::B8G8R8_UNorm: { const uint8_t* s = row + 3u * (size_t)x; c[0] = Unorm8(s[2]); c[1] = Unorm8(s[1]); c[2] = Unorm8(s[0]); break; }
        case Format::B8G8R8A8_UNorm: { const uint8_t* s = row + 4u * (size_t)x; c[0] = Unorm8(s[2]); c[1] = Unorm8(s[1]); c[2] = Unorm8(s[0]); c[3] = Unorm8(s[3]); break; }
        case Format::R8G8B8A8_UNorm: { const uint8_t* s = row + 4u * (size_t)x; c[0] = Unorm8(s[0]); c[1] = Unorm8(s[1]); c[2] = Unorm8(s[2]); c[3] = Unorm8(s[3]); break; }
        case Format::B8G8R8A8_UNorm_Palette:
        {
            const size_t index = row[x];
            uint8_t s[4] = { 0, 0, 0, 0 };
            if (4u * index + 4u <= mPalette.size()) memcpy(s, mPalette.data() + 4u * index, 4);
            c[0] = Unorm8(s[2]); c[1] = Unorm8(s[1]); c[2] = Unorm8(s[0]); c[3] = Unorm8(s[3]);
            break;
        }
        case Format::B5G6R5_UNorm:
        {
            // Vector4_Load_B5G6R5_Norm: masked integer * (2^-k / max) = channel * (1 / max), rounded once
            const uint32_t v = Load16(row + 2u * (size_t)x);
            c[0] = (float)((v >> 11) & 0x1Fu) * (1.0f / 31.0f); c[1] = (float)((v >> 5) & 0x3Fu) * (1.0f / 63.0f); c[2] = (float)(v & 0x1Fu) * (1.0f / 31.0f);
            break;
        }
        case Format::R16_UNorm: c[0] = c[1] = c[2] = c[3] = (float)Load16(row + 2u * (size_t)x) * (1.0f / 65535.0f); break;
        case Format::R16G16_UNorm: { const uint8_t* s = row + 4u * (size_t)x; c[0] = (float)Load16(s) * (1.0f / 65535.0f); c[1] = (float)Load16(s + 2) * (1.0f / 65535.0f); break; }
        case Format::R16G16B16A16_UNorm: { const uint8_t* s = row + 8u * (size_t)x; for (int k = 0; k < 4; ++k) c[k] = (float)Load16(s + 2 * k) * (1.0f / 65535.0f); break; }
        case Format::R32_Float: c[0] = c[1] = c[2] = c[3] = LoadF32(row + 4u * (size_t)x); break;       // Vector4(float): splat
        case Format::R32G32_Float: { const uint8_t* s = row + 8u * (size_t)x; c[0] = LoadF32(s); c[1] = LoadF32(s + 4); break; }
        case Format::R32G32B32_Float: { const uint8_t* s = row + 12u * (size_t)x; c[0] = LoadF32(s); c[1] = LoadF32(s + 4); c[2] = LoadF32(s + 8); break; }
        case Format::R32G32B32A32_Float: { const uint8_t* s = row + 16u * (size_t)x; for (int k = 0; k < 4; ++k) c[k] = LoadF32(s + 4 * k); break; }
        case Format::R16_Half: c[0] = c[1] = c[2] = c[3] = HalfToFloat(Load16(row + 2u * (size_t)x)); break;
        case Format::R16G16_Half: { const uint8_t* s = row + 4u * (size_t)x; c[0] = HalfToFloat(Load16(s)); c[1] = HalfToFloat(Load16(s + 2)); break; }
        case Format::R16G16B16_Half: { const uint8_t* s = row + 6u * (size_t)x; for (int k = 0; k < 3; ++k) c[k] = HalfToFloat(Load16(s + 2 * k)); break; }      // Load_Half4 & mask<1,1,1,0>
        case Format::R16G16B16A16_Half: { const uint8_t* s = row + 8u * (size_t)x; for (int k = 0; k < 4; ++k) c[k] = HalfToFloat(Load16(s + 2 * k)); break; }
        case Format::R9G9B9E5_SharedExp:
        {
            // SharedExpFloat3::ToVector, Packed.h:126-131
            const uint32_t v = Load32(row + 4u * (size_t)x);
            const float scale = BitsToFloat(0x33800000u + ((v >> 27) << 23));
            c[0] = scale * (float)(v & 0x1FFu); c[1] = scale * (float)((v >> 9) & 0x1FFu); c[2] = scale * (float)((v >> 18) & 0x1FFu); c[3] = scale * 0.0f;
            break;
        }
        case Format::R11G11B10_Float:
        {
            // PackedFloat3::ToVector, Packed.h:160-167 (bit assembly; no Inf / NaN / denormal handling, like the reference)
            const uint32_t v = Load32(row + 4u * (size_t)x);
            const uint32_t xm = v & 0x3Fu, xe = (v >> 6) & 0x1Fu, ym = (v >> 11) & 0x3Fu, ye = (v >> 17) & 0x1Fu, zm = (v >> 22) & 0x1Fu, ze = (v >> 27) & 0x1Fu;
            c[0] = BitsToFloat(((xe + 112u) << 23) | (xm << 17)); c[1] = BitsToFloat(((ye + 112u) << 23) | (ym << 17)); c[2] = BitsToFloat(((ze + 112u) << 23) | (zm << 17));
            break;
        }
        case Format::BC1:
        {
            // DecodeBC1, BlockCompression.cpp:52-70: lerp of the masked 5:6:5 integers, then the per-channel scale
            const size_t blocksInRow = mWidth / 4u;
            const uint8_t* block = mData.data() + 8u * (blocksInRow * (y / 4u) + (x / 4u));
            const uint32_t v0 = Load32(block), v1 = Load32(block + 2), code = Load32(block + 4);
            const uint32_t index = (code >> (2u * (4u * (y % 4u) + (x % 4u)))) % 4u;
            const float weights[] = { 0.0f, 1.0f, 1.0f / 3.0f, 2.0f / 3.0f };
            const uint32_t masks[3] = { 0x1Fu << 11, 0x3Fu << 5, 0x1Fu };
            const float scale[3] = { 1.0f / 2048.0f / 31.0f, 1.0f / 32.0f / 63.0f, 1.0f / 31.0f };
            for (int k = 0; k < 3; ++k)
            {
                const float a = (float)(int32_t)(v0 & masks[k]), b = (float)(int32_t)(v1 & masks[k]);
                c[k] = std::fmaf(b - a, weights[index], a) * scale[k];
            }
            c[3] = std::fmaf(0.0f - 0.0f, weights[index], 0.0f) * 0.0f;
            break;
        }
        case Format::BC4:
        {
            const size_t blocksInRow = mWidth / 4u;
            const float v = DecodeBCGrayscale(mData.data() + 8u * (blocksInRow * (y / 4u) + (x / 4u)), x % 4u, y % 4u);
            c[0] = c[1] = c[2] = v; c[3] = 1.0f;
            break;
        }
        case Format::BC5:
        {
            const size_t blocksInRow = mWidth / 4u;
            const uint8_t* block = mData.data() + 16u * (blocksInRow * (y / 4u) + (x / 4u));
            const float red = DecodeBCGrayscale(block, x % 4u, y % 4u), green = DecodeBCGrayscale(block + 8, x % 4u, y % 4u);
            c[0] = green; c[1] = red; c[2] = 0.0f; c[3] = 1.0f;          // (green, red, 0, 1), BlockCompression.cpp:141
            break;
        }
        default: break;
        }
    }

    void Decode()
    {
        mTexels.resize((size_t)mWidth * mHeight * 4);
        for (uint32_t y = 0; y < mHeight; ++y)
            for (uint32_t x = 0; x < mWidth; ++x)
            {
                float c[4];
                GetRawPixel(x, y, c);
                float* out = &mTexels[4 * ((size_t)y * mWidth + x)];
                for (int k = 0; k < 4; ++k) out[k] = mLinearSpace ? c[k] : SrgbToLinear(c[k]);
            }
        mData.clear(); mData.shrink_to_fit(); mPalette.clear();
    }

    // Bitmap::LoadEXR, BitmapEXR.cpp:11-190: single-part scan-line files with 3 or 4 channels of one type (half or
    // float); the planes arrive in the file's channel order (B, G, R or A, B, G, R) and are stored reversed; linear space
    bool LoadEXR(const char* path)
    {
        exr::Image image;
        std::string error;
        try { if (!exr::Load(path, image, error)) { mError = error; return false; } }
        catch (const std::bad_alloc&) { return false; }
        const size_t n = image.channels.size();
        if (n != 3 && n != 4) return false;                                        // "Unsupported EXR format."
        const int32_t type = image.channels[0].pixelType;
        for (const exr::Channel& c : image.channels) if (c.pixelType != type) return false;      // "All channels must be of the same type"
        if (type != exr::FLOAT && type != exr::HALF) return false;                 // "Unsupported EXR format: %i"
        const Format format = type == exr::FLOAT ? (n == 3 ? Format::R32G32B32_Float : Format::R32G32B32A32_Float)
                                                 : (n == 3 ? Format::R16G16B16_Half : Format::R16G16B16A16_Half);
        if (!Init((uint32_t)image.width, (uint32_t)image.height, format, 0, true)) return false;
        const size_t elem = type == exr::FLOAT ? 4 : 2;
        for (size_t y = 0; y < (size_t)image.height; ++y)
            for (size_t x = 0; x < (size_t)image.width; ++x)
                for (size_t k = 0; k < n; ++k)
                    memcpy(&mData[(size_t)mStride * y + (x * n + k) * elem], &image.planes[n - 1 - k][(y * image.width + x) * elem], elem);
        return true;
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
        Format format;
        if (ih.biBitCount == 24) format = Format::B8G8R8_UNorm;
        else if (ih.biBitCount == 8 && ih.biClrUsed > 0) format = Format::B8G8R8A8_UNorm_Palette;
        else if (ih.biBitCount == 8 && ih.biClrUsed == 0) format = Format::R8_UNorm;
        else return false;
        if (ih.biWidth == 0 || ih.biHeight == 0 || ih.biWidth > 65536u || ih.biHeight > 65536u) return false;
        if (ih.biClrUsed > 256u) return false;      // an 8-bit index cannot address more; a crafted header must not size the allocation
        const uint32_t stride = (ih.biWidth * BitsPerPixel(format) / 8 + 3u) & ~3u;
        if (!Init(ih.biWidth, ih.biHeight, format, stride, false)) return false;
        mPalette.assign((size_t)ih.biClrUsed * 4u, 0);
        if (!mPalette.empty() && fread(mPalette.data(), mPalette.size(), 1, file) != 1) return false;
        if (fseek(file, (long)fh.bfOffBits, SEEK_SET) != 0) return false;
        if (fread(mData.data(), DataSize(), 1, file) != 1) return false;
        return true;
    }

    // Bitmap::LoadDDS, BitmapDDS.cpp:228-540: only the top mip level is read, formats as listed there
    bool LoadDDS(FILE* file, const char* path)
    {
        struct PixelFormat { uint32_t size, flags, fourCC, rgbBitCount, rBitMask, gBitMask, bBitMask, aBitMask; };
        struct Header { uint32_t magic, size, flags, height, width, pitchOrLinearSize, depth, mipMapCount, reserved1[11]; PixelFormat pf; uint32_t caps[4]; uint32_t reserved2; };
        struct HeaderDX10 { uint32_t dxgiFormat, resourceDimension, miscFlag, arraySize, miscFlags2; };
        static_assert(sizeof(Header) == 128, "DDS header");
        Header header;
        if (fread(&header, sizeof(header), 1, file) != 1) return false;
        if (header.magic != 0x20534444u) return false;
        if (header.width < 1 || header.height < 1 || header.width >= 65535u || header.height >= 65535u) return false;
        auto fourcc = [](char a, char b, char c, char d) { return (uint32_t)(uint8_t)a | ((uint32_t)(uint8_t)b << 8) | ((uint32_t)(uint8_t)c << 16) | ((uint32_t)(uint8_t)d << 24); };
        Format format = Format::Unknown;
        bool linear = true;
        const PixelFormat& pf = header.pf;
        if (pf.flags & 0x40u)                   // DDPF_RGB
        {
            if (pf.rgbBitCount == 32)
            {
                if (pf.rBitMask == 0x00FF0000u && pf.gBitMask == 0x0000FF00u && pf.bBitMask == 0x000000FFu && pf.aBitMask == 0xFF000000u) { format = Format::B8G8R8A8_UNorm; linear = false; }
                else if (pf.rBitMask == 0x000000FFu && pf.gBitMask == 0x0000FF00u && pf.bBitMask == 0x00FF0000u && pf.aBitMask == 0xFF000000u) { format = Format::R8G8B8A8_UNorm; linear = false; }
                else if (pf.rBitMask == 0xFFFFFFFFu && pf.gBitMask == 0 && pf.bBitMask == 0 && pf.aBitMask == 0) format = Format::R32_Float;
                else if (pf.rBitMask == 0xFFFFu && pf.gBitMask == 0xFFFF0000u && pf.bBitMask == 0 && pf.aBitMask == 0) format = Format::R16G16_UNorm;
            }
            else if (pf.rgbBitCount == 16)
            {
                if (pf.rBitMask == 0xf800u && pf.gBitMask == 0x07e0u && pf.bBitMask == 0x001fu && pf.aBitMask == 0) { format = Format::B5G6R5_UNorm; linear = false; }
            }
        }
        else if (pf.flags & 0x4u)               // DDPF_FOURCC
        {
            if (pf.fourCC == 111u) format = Format::R16_Half;                  // D3DFMT_R16F
            else if (pf.fourCC == 112u) format = Format::R16G16_Half;          // D3DFMT_G16R16F
            else if (pf.fourCC == 113u) format = Format::R16G16B16A16_Half;    // D3DFMT_A16B16G16R16F
            else if (pf.fourCC == 114u) format = Format::R32_Float;            // D3DFMT_R32F
            else if (pf.fourCC == 115u) format = Format::R32G32_Float;         // D3DFMT_G32R32F
            else if (pf.fourCC == 116u) format = Format::R32G32B32A32_Float;   // D3DFMT_A32B32G32R32F
            else if (pf.fourCC == 36u) format = Format::R16G16B16A16_UNorm;    // D3DFMT_A16B16G16R16
            else if (pf.fourCC == fourcc('D', 'X', 'T', '1')) format = Format::BC1;
            else if (pf.fourCC == fourcc('A', 'T', 'I', '1') || pf.fourCC == fourcc('B', 'C', '4', 'U') || pf.fourCC == fourcc('B', 'C', '4', 'S')) format = Format::BC4;
            else if (pf.fourCC == fourcc('A', 'T', 'I', '2') || pf.fourCC == fourcc('B', 'C', '5', 'U')) format = Format::BC5;
            else if (pf.fourCC == fourcc('D', 'X', '1', '0'))
            {
                HeaderDX10 dx10;
                if (fread(&dx10, sizeof(dx10), 1, file) != 1) return false;
                switch (dx10.dxgiFormat)
                {
                case 10: format = Format::R16G16B16A16_Half; break;
                case 34: format = Format::R16G16_Half; break;
                case 54: format = Format::R16_Half; break;
                default: break;
                }
                switch (dx10.dxgiFormat)         // (the reference's second if-chain starts here: a half format found above survives)
                {
                case 2: format = Format::R32G32B32A32_Float; break;
                case 6: format = Format::R32G32B32_Float; break;
                case 16: format = Format::R32G32_Float; break;
                case 41: format = Format::R32_Float; break;
                case 67: format = Format::R9G9B9E5_SharedExp; break;
                case 26: format = Format::R11G11B10_Float; break;
                case 87: format = Format::B8G8R8A8_UNorm; break;
                case 91: format = Format::B8G8R8A8_UNorm; linear = false; break;
                case 49: format = Format::R8G8_UNorm; break;
                case 61: format = Format::R8_UNorm; break;
                case 85: format = Format::B5G6R5_UNorm; linear = false; break;
                case 11: format = Format::R16G16B16A16_UNorm; break;
                case 35: format = Format::R16G16_UNorm; break;
                case 56: format = Format::R16_UNorm; break;
                case 71: format = Format::BC1; break;
                case 72: format = Format::BC1; linear = false; break;
                case 80: format = Format::BC4; break;
                case 83: format = Format::BC5; break;
                default: break;
                }
            }
        }
        else if (pf.flags & 0x20000u)           // DDPF_LUMINANCE
        {
            if (pf.rgbBitCount == 8)
            {
                if (pf.rBitMask == 0xFFu && pf.gBitMask == 0 && pf.bBitMask == 0 && pf.aBitMask == 0) { format = Format::R8_UNorm; linear = false; }
                else if (pf.rBitMask == 0xFFu && pf.gBitMask == 0 && pf.bBitMask == 0 && pf.aBitMask == 0xFF00u) { format = Format::R8G8_UNorm; linear = false; }
            }
            else if (pf.rgbBitCount == 16)
            {
                if (pf.rBitMask == 0xFFFFu && pf.gBitMask == 0 && pf.bBitMask == 0 && pf.aBitMask == 0) format = Format::R16_UNorm;
                else if (pf.rBitMask == 0xFFu && pf.gBitMask == 0 && pf.bBitMask == 0 && pf.aBitMask == 0xFF00u) { format = Format::R8G8_UNorm; linear = false; }
            }
        }
        if (format == Format::Unknown) return false;     // "Unsupported DDS format"
        if ((format == Format::BC1 || format == Format::BC4 || format == Format::BC5) && ((header.width | header.height) & 3u))
        {
            mError = std::string("block-compressed DDS whose size is not a multiple of 4 (the reference reads out of bounds there): '") + path + "'";
            return false;
        }
        if (!Init(header.width, header.height, format, 0, linear)) return false;
        if (fread(mData.data(), DataSize(), 1, file) != 1) return false;
        return true;
    }

    uint32_t mWidth = 0, mHeight = 0;
    std::vector<float> mTexels;
    std::vector<float> mImportanceCdf;
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
