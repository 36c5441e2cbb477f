// rt_exr.h — OpenEXR scan-line reader for Bitmap::LoadEXR (Core/Utils/BitmapEXR.cpp:11-190).
//
// The reference reads EXR files through the third-party tinyexr (External/tinyexr, v0.9.5-era single header) and then
// keeps only single-part scan-line images with 3 or 4 channels of one type (half or float).  This reader is written
// from the published OpenEXR file layout and delivers what that call sequence delivers - per-channel planes in the
// file's channel order - including tinyexr's behaviours a caller can observe:
//   * every one of the eight standard header attributes must be present, otherwise the load fails;
//   * a block whose stored size equals its raw size is taken as raw bytes (RLE / ZIP / ZIPS);
//   * with lineOrder = DECREASING_Y a compressed file's rows come out flipped (row = height - 1 - line,
//     tinyexr.h:9662-9690), an uncompressed file's rows are placed by block position instead (tinyexr.h:10087-10190);
// Compressions: NONE, RLE, ZIPS, ZIP (zlib), PIZ (wavelet + Huffman).  Tiled, multi-part and deep files: refused, as the
// reference refuses them.
#pragma once

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include <zlib.h>

namespace rt {
namespace exr {

enum PixelType : int32_t { UINT = 0, HALF = 1, FLOAT = 2 };
enum Compression : uint8_t { NONE = 0, RLE = 1, ZIPS = 2, ZIP = 3, PIZ = 4 };

struct Channel { std::string name; int32_t pixelType = 0; };

struct Image
{
    int32_t width = 0, height = 0;
    std::vector<Channel> channels;                 // file order (alphabetical in a conforming file)
    std::vector<std::vector<uint8_t>> planes;      // per channel: width * height elements of 2 (half) or 4 bytes
};

namespace detail {

inline size_t PixelSize(int32_t type) { return type == HALF ? 2u : 4u; }

// OpenEXR's byte predictor + even / odd interleave, undone (ImfZip.cpp / ImfRle.cpp: the last two steps of every
// byte-oriented compressor)
inline void Unpredict(std::vector<uint8_t>& buf, std::vector<uint8_t>& out)
{
    const size_t n = buf.size();
    for (size_t t = 1; t < n; ++t) buf[t] = (uint8_t)((int)buf[t - 1] + (int)buf[t] - 128);
    out.resize(n);
    const size_t half = (n + 1) / 2;
    for (size_t i = 0, s = 0; s < n; ++i)
    {
        out[s++] = buf[i];
        if (s < n) out[s++] = buf[half + i];
    }
}

inline bool UnRle(const uint8_t* src, size_t srcLen, std::vector<uint8_t>& dst)
{
    size_t o = 0;
    const size_t n = dst.size();
    for (size_t i = 0; i < srcLen;)
    {
        const int count = (int8_t)src[i++];
        if (count < 0)
        {
            const size_t len = (size_t)(-count);
            if (i + len > srcLen || o + len > n) return false;
            memcpy(&dst[o], src + i, len);
            i += len; o += len;
        }
        else
        {
            const size_t len = (size_t)count + 1;
            if (i >= srcLen || o + len > n) return false;
            memset(&dst[o], src[i++], len);
            o += len;
        }
    }
    return o == n;
}

// ---- PIZ (ImfPizCompressor.cpp, ImfHuf.cpp, ImfWav.cpp) --------------------------------------------------------------

constexpr int HUF_ENCBITS = 16, HUF_DECBITS = 14;
constexpr int HUF_ENCSIZE = (1 << HUF_ENCBITS) + 1, HUF_DECSIZE = 1 << HUF_DECBITS, HUF_DECMASK = HUF_DECSIZE - 1;
constexpr int SHORT_ZEROCODE_RUN = 59, LONG_ZEROCODE_RUN = 63, SHORTEST_LONG_RUN = 2 + LONG_ZEROCODE_RUN - SHORT_ZEROCODE_RUN;

struct HufDec { int len = 0, lit = 0; std::vector<int> p; };

struct BitReader
{
    const uint8_t* p; const uint8_t* end; uint64_t c = 0; int lc = 0;
    bool GetBits(int n, uint64_t& out)
    {
        while (lc < n) { if (p >= end) return false; c = (c << 8) | *p++; lc += 8; }
        lc -= n;
        out = (c >> lc) & ((1ull << n) - 1);
        return true;
    }
};

inline void HufCanonicalCodeTable(std::vector<uint64_t>& hcode)
{
    uint64_t n[59] = { 0 };
    for (int i = 0; i < HUF_ENCSIZE; ++i) n[hcode[i]] += 1;
    uint64_t c = 0;
    for (int i = 58; i > 0; --i) { const uint64_t nc = ((c + n[i]) >> 1); n[i] = c; c = nc; }
    for (int i = 0; i < HUF_ENCSIZE; ++i)
    {
        const int l = (int)hcode[i];
        if (l > 0) hcode[i] = (uint64_t)l | (n[l]++ << 6);
    }
}

inline bool HufUnpackEncTable(BitReader& br, int im, int iM, std::vector<uint64_t>& hcode)
{
    hcode.assign(HUF_ENCSIZE, 0);
    for (; im <= iM; im++)
    {
        uint64_t l;
        if (!br.GetBits(6, l)) return false;
        hcode[im] = l;
        if (l == (uint64_t)LONG_ZEROCODE_RUN)
        {
            uint64_t zerun;
            if (!br.GetBits(8, zerun)) return false;
            zerun += SHORTEST_LONG_RUN;
            if (im + (int)zerun > iM + 1) return false;
            while (zerun--) hcode[im++] = 0;
            im--;
        }
        else if (l >= (uint64_t)SHORT_ZEROCODE_RUN)
        {
            uint64_t zerun = l - SHORT_ZEROCODE_RUN + 2;
            if (im + (int)zerun > iM + 1) return false;
            while (zerun--) hcode[im++] = 0;
            im--;
        }
    }
    HufCanonicalCodeTable(hcode);
    return true;
}

inline bool HufBuildDecTable(const std::vector<uint64_t>& hcode, int im, int iM, std::vector<HufDec>& hdecod)
{
    hdecod.assign(HUF_DECSIZE, HufDec());
    for (; im <= iM; im++)
    {
        const uint64_t c = hcode[im] >> 6;
        const int l = (int)(hcode[im] & 63);
        if (c >> l) return false;
        if (l > HUF_DECBITS)
        {
            HufDec& pl = hdecod[c >> (l - HUF_DECBITS)];
            if (pl.len) return false;
            pl.lit++;
            pl.p.push_back(im);
        }
        else if (l)
        {
            const size_t first = (size_t)(c << (HUF_DECBITS - l));
            for (uint64_t i = 1ull << (HUF_DECBITS - l); i > 0; i--)
            {
                HufDec& pl = hdecod[first + (size_t)i - 1];
                if (pl.len || !pl.p.empty()) return false;
                pl.len = l;
                pl.lit = im;
            }
        }
    }
    return true;
}

inline bool HufDecode(const std::vector<uint64_t>& hcode, const std::vector<HufDec>& hdecod, const uint8_t* in, int ni /* bits */, int rlc, size_t no,
                      std::vector<uint16_t>& out)
{
    out.clear(); out.reserve(no);
    uint64_t c = 0;
    int lc = 0;
    const uint8_t* ie = in + (ni + 7) / 8;
    auto emit = [&](int sym) -> bool
    {
        if (sym == rlc)
        {
            if (lc < 8) { if (in >= ie) return false; c = (c << 8) | *in++; lc += 8; }
            lc -= 8;
            const int cs = (int)((c >> lc) & 0xFF);
            if (out.empty() || out.size() + (size_t)cs > no) return false;
            const uint16_t s = out.back();
            for (int k = 0; k < cs; ++k) out.push_back(s);
        }
        else
        {
            if (out.size() >= no) return false;
            out.push_back((uint16_t)sym);
        }
        return true;
    };
    while (in < ie)
    {
        c = (c << 8) | *in++; lc += 8;
        while (lc >= HUF_DECBITS)
        {
            const HufDec& pl = hdecod[(c >> (lc - HUF_DECBITS)) & HUF_DECMASK];
            if (pl.len)
            {
                lc -= pl.len;
                if (!emit(pl.lit)) return false;
            }
            else
            {
                if (pl.p.empty()) return false;
                int j = 0;
                for (; j < pl.lit; j++)
                {
                    const int l = (int)(hcode[pl.p[j]] & 63);
                    while (lc < l && in < ie) { c = (c << 8) | *in++; lc += 8; }
                    if (lc >= l && (hcode[pl.p[j]] >> 6) == ((c >> (lc - l)) & ((1ull << l) - 1)))
                    {
                        lc -= l;
                        if (!emit(pl.p[j])) return false;
                        break;
                    }
                }
                if (j == pl.lit) return false;
            }
        }
    }
    const int i = (8 - ni) & 7;
    c >>= i; lc -= i;
    while (lc > 0)
    {
        const HufDec& pl = hdecod[(c << (HUF_DECBITS - lc)) & HUF_DECMASK];
        if (!pl.len) return false;
        lc -= pl.len;
        if (lc < 0) return false;
        if (!emit(pl.lit)) return false;
    }
    return out.size() == no;
}

inline bool HufUncompress(const uint8_t* compressed, size_t nCompressed, size_t nRaw, std::vector<uint16_t>& raw)
{
    if (nCompressed == 0) { raw.assign(nRaw, 0); return nRaw == 0; }
    if (nCompressed < 20) return false;
    uint32_t im, iM, nBits;
    memcpy(&im, compressed, 4); memcpy(&iM, compressed + 4, 4); memcpy(&nBits, compressed + 12, 4);
    if (im >= (uint32_t)HUF_ENCSIZE || iM >= (uint32_t)HUF_ENCSIZE) return false;
    BitReader br{ compressed + 20, compressed + nCompressed };
    std::vector<uint64_t> freq;
    if (!HufUnpackEncTable(br, (int)im, (int)iM, freq)) return false;
    const uint8_t* ptr = br.p;          // the table is padded to whole bytes
    if ((size_t)(ptr - compressed) + (nBits + 7) / 8 > nCompressed) return false;
    std::vector<HufDec> hdec;
    if (!HufBuildDecTable(freq, (int)im, (int)iM, hdec)) return false;
    return HufDecode(freq, hdec, ptr, (int)nBits, (int)iM, nRaw, raw);
}

// inverse 2-D wavelet (ImfWav.cpp wav2Decode), 14-bit and 16-bit variants
inline void Wdec14(uint16_t l, uint16_t h, uint16_t& a, uint16_t& b)
{
    const int16_t ls = (int16_t)l, hs = (int16_t)h;
    const int hi = hs, ai = ls + (hi & 1) + (hi >> 1);
    const int16_t as = (int16_t)ai, bs = (int16_t)(ai - hi);
    a = (uint16_t)as; b = (uint16_t)bs;
}

inline void Wdec16(uint16_t l, uint16_t h, uint16_t& a, uint16_t& b)
{
    constexpr int A_OFFSET = 1 << 15, MOD_MASK = (1 << 16) - 1;
    const int m = l, d = h;
    const int bb = (m - (d >> 1)) & MOD_MASK;
    const int aa = (d + bb - A_OFFSET) & MOD_MASK;
    b = (uint16_t)bb; a = (uint16_t)aa;
}

inline void Wav2Decode(uint16_t* in, int nx, int ox, int ny, int oy, uint16_t mx)
{
    const bool w14 = mx < (1 << 14);
    const int n = nx > ny ? ny : nx;
    int p = 1, p2;
    while (p <= n) p <<= 1;
    p >>= 1; p2 = p; p >>= 1;
    while (p >= 1)
    {
        uint16_t* py = in;
        uint16_t* ey = in + oy * (ny - p2);
        const int oy1 = oy * p, oy2 = oy * p2, ox1 = ox * p, ox2 = ox * p2;
        uint16_t i00, i01, i10, i11;
        for (; py <= ey; py += oy2)
        {
            uint16_t* px = py;
            uint16_t* ex = py + ox * (nx - p2);
            for (; px <= ex; px += ox2)
            {
                uint16_t* p01 = px + ox1; uint16_t* p10 = px + oy1; uint16_t* p11 = p10 + ox1;
                if (w14)
                {
                    Wdec14(*px, *p10, i00, i10); Wdec14(*p01, *p11, i01, i11);
                    Wdec14(i00, i01, *px, *p01); Wdec14(i10, i11, *p10, *p11);
                }
                else
                {
                    Wdec16(*px, *p10, i00, i10); Wdec16(*p01, *p11, i01, i11);
                    Wdec16(i00, i01, *px, *p01); Wdec16(i10, i11, *p10, *p11);
                }
            }
            if (nx & p)
            {
                uint16_t* p10 = px + oy1;
                if (w14) Wdec14(*px, *p10, i00, *p10); else Wdec16(*px, *p10, i00, *p10);
                *px = i00;
            }
        }
        if (ny & p)
        {
            uint16_t* px = py;
            uint16_t* ex = py + ox * (nx - p2);
            for (; px <= ex; px += ox2)
            {
                uint16_t* p01 = px + ox1;
                if (w14) Wdec14(*px, *p01, i00, *p01); else Wdec16(*px, *p01, i00, *p01);
                *px = i00;
            }
        }
        p2 = p; p >>= 1;
    }
}

// one PIZ block -> raw scan-line bytes (channel-planar per line, like the other compressors' output)
inline bool UnPiz(const uint8_t* src, size_t srcLen, const std::vector<Channel>& channels, int width, int numLines, std::vector<uint8_t>& dst)
{
    constexpr int USHORT_RANGE = 1 << 16, BITMAP_SIZE = USHORT_RANGE >> 3;
    if (srcLen < 4) return false;
    std::vector<uint8_t> bitmap(BITMAP_SIZE, 0);
    uint16_t minNonZero, maxNonZero;
    memcpy(&minNonZero, src, 2); memcpy(&maxNonZero, src + 2, 2);
    const uint8_t* ptr = src + 4;
    if (maxNonZero >= BITMAP_SIZE) return false;
    if (minNonZero <= maxNonZero)
    {
        const size_t len = (size_t)(maxNonZero - minNonZero + 1);
        if ((size_t)(ptr - src) + len > srcLen) return false;
        memcpy(&bitmap[minNonZero], ptr, len);
        ptr += len;
    }
    // reverse look-up table
    std::vector<uint16_t> lut(USHORT_RANGE, 0);
    int k = 0;
    for (int i = 0; i < USHORT_RANGE; ++i) if (i == 0 || (bitmap[i >> 3] & (1 << (i & 7)))) lut[k++] = (uint16_t)i;
    const uint16_t maxValue = (uint16_t)(k - 1);

    if ((size_t)(ptr - src) + 4 > srcLen) return false;
    int32_t length;
    memcpy(&length, ptr, 4); ptr += 4;
    if (length < 0 || (size_t)(ptr - src) + (size_t)length > srcLen) return false;
    size_t total = 0;
    for (const Channel& c : channels) total += (size_t)width * numLines * (PixelSize(c.pixelType) / 2);
    std::vector<uint16_t> tmp;
    if (!HufUncompress(ptr, (size_t)length, total, tmp)) return false;

    // wavelet decode per channel (each 16-bit half of a 32-bit channel is its own plane)
    size_t start = 0;
    std::vector<size_t> chanStart;
    for (const Channel& c : channels)
    {
        const int size = (int)(PixelSize(c.pixelType) / 2);
        chanStart.push_back(start);
        for (int j = 0; j < size; ++j) Wav2Decode(&tmp[start + j], width, size, numLines, width * size, maxValue);
        start += (size_t)width * numLines * size;
    }
    for (uint16_t& v : tmp) v = lut[v];
    // rearrange: per line, per channel
    dst.resize(total * 2);
    uint8_t* out = dst.data();
    std::vector<size_t> cursor = chanStart;
    for (int y = 0; y < numLines; ++y)
        for (size_t c = 0; c < channels.size(); ++c)
        {
            const size_t n = (size_t)width * (PixelSize(channels[c].pixelType) / 2);
            memcpy(out, &tmp[cursor[c]], n * 2);
            out += n * 2; cursor[c] += n;
        }
    return true;
}

inline bool ReadString(const std::vector<uint8_t>& f, size_t& pos, std::string& out)
{
    out.clear();
    while (pos < f.size() && f[pos] != 0) out.push_back((char)f[pos++]);
    if (pos >= f.size()) return false;
    ++pos;
    return true;
}

} // namespace detail

// Returns false when the file is not a readable single-part scan-line EXR (`error` stays empty: the reference then
// reports a failed texture load), or when it is one that this reader knowingly does not handle (`error` says why).
inline bool Load(const char* path, Image& image, std::string& error)
{
    using namespace detail;
    FILE* fp = fopen(path, "rb");
    if (!fp) return false;
    std::vector<uint8_t> f;
    fseek(fp, 0, SEEK_END);
    const long size = ftell(fp);
    fseek(fp, 0, SEEK_SET);
    if (size > 0) { f.resize((size_t)size); if (fread(f.data(), 1, f.size(), fp) != f.size()) f.clear(); }
    fclose(fp);
    if (f.size() < 8 || f[0] != 0x76 || f[1] != 0x2f || f[2] != 0x31 || f[3] != 0x01) return false;
    if (f[4] != 2) return false;                                   // version
    const bool tiled = (f[5] & 0x2) != 0, nonImage = (f[5] & 0x8) != 0, multipart = (f[5] & 0x10) != 0;
    if (multipart || nonImage) return false;                       // "Multipart EXR are not supported"
    size_t pos = 8;

    bool hasChannels = false, hasCompression = false, hasDataWindow = false, hasDisplayWindow = false, hasLineOrder = false,
         hasAspect = false, hasCenter = false, hasWidth = false;
    uint8_t compression = 0, lineOrder = 0;
    int32_t dataWindow[4] = { 0, 0, 0, 0 };
    for (;;)
    {
        if (pos >= f.size()) return false;
        if (f[pos] == 0) { ++pos; break; }
        std::string name, type;
        if (!ReadString(f, pos, name) || !ReadString(f, pos, type) || pos + 4 > f.size()) return false;
        uint32_t len;
        memcpy(&len, &f[pos], 4); pos += 4;
        if (pos + len > f.size()) return false;
        const uint8_t* d = &f[pos];
        if (name == "channels")
        {
            size_t p = pos;
            while (p < pos + len && f[p] != 0)
            {
                Channel c;
                if (!ReadString(f, p, c.name) || p + 16 > pos + len) return false;
                memcpy(&c.pixelType, &f[p], 4);
                p += 16;                                            // pixel type, pLinear + 3 reserved, x / y sampling
                image.channels.push_back(c);
            }
            hasChannels = !image.channels.empty();
        }
        else if (name == "compression" && len >= 1) { compression = d[0]; hasCompression = true; }
        else if (name == "dataWindow" && len >= 16) { memcpy(dataWindow, d, 16); hasDataWindow = true; }
        else if (name == "displayWindow" && len >= 16) hasDisplayWindow = true;
        else if (name == "lineOrder" && len >= 1) { lineOrder = d[0]; hasLineOrder = true; }
        else if (name == "pixelAspectRatio" && len >= 4) hasAspect = true;
        else if (name == "screenWindowCenter" && len >= 8) hasCenter = true;
        else if (name == "screenWindowWidth" && len >= 4) hasWidth = true;
        pos += len;
    }
    if (!(hasChannels && hasCompression && hasDataWindow && hasDisplayWindow && hasLineOrder && hasAspect && hasCenter && hasWidth)) return false;
    if (tiled) return false;                                        // "Tiled EXR are not supported"
    if (compression > PIZ) return false;                            // PXR24 / B44 / DWA: tinyexr reports an error as well
    for (const Channel& c : image.channels) if (c.pixelType < UINT || c.pixelType > FLOAT) return false;

    const int64_t width = (int64_t)dataWindow[2] - dataWindow[0] + 1, height = (int64_t)dataWindow[3] - dataWindow[1] + 1;
    if (width <= 0 || height <= 0 || width > 65536 || height > 65536) return false;
    image.width = (int32_t)width; image.height = (int32_t)height;
    size_t lineBytes = 0;
    for (const Channel& c : image.channels) lineBytes += (size_t)width * PixelSize(c.pixelType);
    const int linesPerBlock = compression == ZIP ? 16 : (compression == PIZ ? 32 : 1);
    const size_t numBlocks = (size_t)((height + linesPerBlock - 1) / linesPerBlock);
    if (pos + numBlocks * 8 > f.size()) return false;
    // a corrupt data window must not turn into a multi-gigabyte allocation: no compressor here gets beyond ~1000 : 1
    if ((uint64_t)lineBytes * (uint64_t)height / 4096u > f.size()) return false;
    image.planes.resize(image.channels.size());
    for (size_t c = 0; c < image.channels.size(); ++c) image.planes[c].assign((size_t)width * height * PixelSize(image.channels[c].pixelType), 0);

    std::vector<uint8_t> raw, scratch;
    for (size_t b = 0; b < numBlocks; ++b)
    {
        uint64_t offset;
        memcpy(&offset, &f[pos + b * 8], 8);
        if (f.size() < 8 || offset > f.size() - 8) return false;          // (subtraction form: offset + 8 would wrap for offsets near 2^64)
        int32_t lineNo, dataLen;
        memcpy(&lineNo, &f[offset], 4); memcpy(&dataLen, &f[offset + 4], 4);
        if (dataLen < 0 || (uint64_t)dataLen > f.size() - 8 - offset) return false;
        const int64_t endLine = std::min<int64_t>((int64_t)lineNo + linesPerBlock, (int64_t)dataWindow[3] + 1);
        const int64_t numLines = endLine - lineNo;
        const int64_t firstLine = (int64_t)lineNo - dataWindow[1];
        if (numLines <= 0 || firstLine < 0 || firstLine + numLines > height) return false;
        const uint8_t* src = &f[offset + 8];
        const size_t rawSize = lineBytes * (size_t)numLines;
        raw.assign(rawSize, 0);
        if (compression == NONE || ((size_t)dataLen == rawSize && compression != PIZ))
        {
            if ((size_t)dataLen < rawSize) return false;
            memcpy(raw.data(), src, rawSize);
        }
        else if (compression == RLE)
        {
            scratch.assign(rawSize, 0);
            if (!UnRle(src, (size_t)dataLen, scratch)) return false;
            Unpredict(scratch, raw);
        }
        else if (compression == ZIP || compression == ZIPS)
        {
            scratch.assign(rawSize, 0);
            uLongf outLen = (uLongf)rawSize;
            if (uncompress(scratch.data(), &outLen, src, (uLong)dataLen) != Z_OK || outLen != rawSize) return false;
            Unpredict(scratch, raw);
        }
        else
        {
            if ((size_t)dataLen == rawSize) memcpy(raw.data(), src, rawSize);
            else if (!UnPiz(src, (size_t)dataLen, image.channels, (int)width, (int)numLines, raw) || raw.size() != rawSize) { error.clear(); return false; }
        }
        // scatter: per line, the channels one after the other
        const uint8_t* p = raw.data();
        for (int64_t v = 0; v < numLines; ++v)
        {
            // tinyexr places the lines of an uncompressed file by the block's position in the offset table and those of
            // a compressed file by the block's line number (tinyexr.h:10087-10190 vs 9607-10060)
            const int64_t at = (compression == NONE ? (int64_t)b : firstLine) + v;
            if (at >= height) return false;
            const int64_t row = lineOrder == 0 ? at : height - 1 - at;
            for (size_t c = 0; c < image.channels.size(); ++c)
            {
                const size_t n = (size_t)width * PixelSize(image.channels[c].pixelType);
                memcpy(&image.planes[c][(size_t)row * n], p, n);
                p += n;
            }
        }
    }
    return true;
}

} // namespace exr
} // namespace rt
