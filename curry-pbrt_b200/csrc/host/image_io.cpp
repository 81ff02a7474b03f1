// PNG in (texture load, src/texture/image/png.rs:6-27: 8-bit RGB assumed), OpenEXR in
// (src/texture/image/exr.rs:7-61) and PNG out (Film::write_image, src/film.rs:35-38 ->
// src/texture/image.rs:108-127).  zlib does the inflate/deflate; the container formats, scanline
// filters and EXR predictors are handled here.
#include "scene_host.h"

#include <zlib.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>

namespace cpbrt_host {

static uint32_t be32(const uint8_t* p) { return (uint32_t(p[0]) << 24) | (uint32_t(p[1]) << 16) | (uint32_t(p[2]) << 8) | p[3]; }

static int paeth(int a, int b, int c) {
    int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
    if (pa <= pb && pa <= pc) return a;
    return pb <= pc ? b : c;
}

bool read_png_rgb8(const std::string& path, uint32_t* w, uint32_t* h, std::vector<uint8_t>* rgb, std::string* err) {
    std::ifstream f(path, std::ios::binary);
    if (!f) { *err = "cannot open image '" + path + "'"; return false; }
    std::stringstream ss; ss << f.rdbuf();
    std::string d = ss.str();
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', '\r', '\n', 0x1a, '\n'};
    if (d.size() < 8 || memcmp(d.data(), sig, 8) != 0) { *err = "not a PNG: " + path; return false; }
    size_t pos = 8;
    uint32_t W = 0, H = 0; int depth = 0, ctype = 0, interlace = 0;
    std::string idat;
    while (pos + 12 <= d.size()) {
        const uint8_t* p = (const uint8_t*)d.data() + pos;
        uint32_t len = be32(p);
        std::string typ(d.data() + pos + 4, 4);
        if (pos + 12 + len > d.size()) { *err = "truncated PNG: " + path; return false; }
        const uint8_t* body = p + 8;
        if (typ == "IHDR" && len != 13) { *err = "bad PNG IHDR length: " + path; return false; }
        if (typ == "IHDR") { W = be32(body); H = be32(body + 4); depth = body[8]; ctype = body[9]; interlace = body[12]; }
        else if (typ == "IDAT") idat.append((const char*)body, len);
        else if (typ == "IEND") break;
        pos += 12 + len;
    }
    // The reference indexes the decoded buffer as packed RGB8 (png.rs:17-24); other layouts
    // would be misread there, so they are rejected here instead of silently reinterpreted.
    if (depth != 8 || ctype != 2 || interlace != 0) { *err = "only 8-bit non-interlaced RGB PNG is supported: " + path; return false; }
    const size_t bpp = 3, stride = size_t(W) * bpp;
    // a corrupted header must not turn into a multi-terabyte allocation: deflate expands at most ~1032x
    if (W == 0 || H == 0 || W > (1u << 16) || H > (1u << 16) || (stride + 1) * H > idat.size() * 1100 + 4096) {
        *err = "PNG header and data sizes do not fit together: " + path; return false;
    }
    std::vector<uint8_t> raw((stride + 1) * H);
    uLongf out_len = raw.size();
    if (uncompress(raw.data(), &out_len, (const Bytef*)idat.data(), idat.size()) != Z_OK || out_len != raw.size()) {
        *err = "PNG inflate failed: " + path; return false;
    }
    rgb->assign(stride * H, 0);
    for (uint32_t y = 0; y < H; ++y) {
        const uint8_t* in = raw.data() + y * (stride + 1);
        uint8_t* out = rgb->data() + y * stride;
        const uint8_t* up = y ? out - stride : nullptr;
        int ft = in[0];
        for (size_t i = 0; i < stride; ++i) {
            int a = i >= bpp ? out[i - bpp] : 0, b = up ? up[i] : 0, c = (up && i >= bpp) ? up[i - bpp] : 0;
            int x = in[1 + i], v;
            switch (ft) {
                case 0: v = x; break;
                case 1: v = x + a; break;
                case 2: v = x + b; break;
                case 3: v = x + ((a + b) >> 1); break;
                case 4: v = x + paeth(a, b, c); break;
                default: *err = "bad PNG filter: " + path; return false;
            }
            out[i] = (uint8_t)v;
        }
    }
    *w = W; *h = H;
    return true;
}

// ---- OpenEXR ------------------------------------------------------------------------------------
// ExrImageFileReader::read_file (src/texture/image/exr.rs:7-61) reads layer 0 of the file with the
// `exr` crate (0.7.4) and copies the channels named R, G and B -- f16, f32 or u32 samples, each
// converted to f32 -- into an RGB buffer in data-window row order; any other channel name panics
// there and is an error here.  This reader covers single-part scanline files, uncompressed or with
// the lossless RLE / ZIPS / ZIP / PIZ codecs (zlib is already linked for PNG).  Tiled, deep and
// multi-part files and the lossy codecs (PXR24, B44, DWA) are rejected by name.
static float half_to_float(uint16_t h) {
    const uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
    uint32_t exp = (h >> 10) & 0x1Fu, man = h & 0x3FFu, bits;
    if (exp == 0) {
        if (man == 0) bits = sign;
        else {   // subnormal half -> normal float
            int e = -1;
            do { ++e; man <<= 1; } while (!(man & 0x400u));
            bits = sign | ((uint32_t)(127 - 15 - e) << 23) | ((man & 0x3FFu) << 13);
        }
    } else if (exp == 31) bits = sign | 0x7F800000u | (man << 13);
    else bits = sign | ((exp + 127 - 15) << 23) | (man << 13);
    float f; memcpy(&f, &bits, 4);
    return f;
}

namespace {
struct ExrCursor {
    const uint8_t* p; size_t n, pos = 0; bool ok = true;
    bool need(size_t k) { if (k > n || pos > n - k) ok = false; return ok; }   // no wrap-around for hostile offsets
    uint8_t u8() { return need(1) ? p[pos++] : 0; }
    uint32_t u32() { if (!need(4)) return 0; uint32_t v; memcpy(&v, p + pos, 4); pos += 4; return v; }   // little-endian host
    uint64_t u64() { if (!need(8)) return 0; uint64_t v; memcpy(&v, p + pos, 8); pos += 8; return v; }
    std::string cstr() { std::string s; while (need(1) && p[pos]) s.push_back((char)p[pos++]); if (ok) ++pos; return s; }
};
struct ExrChannel { std::string name; uint32_t type; int rgb; };
}  // namespace

// undo the byte predictor and the even / odd split the RLE and ZIP codecs apply before compressing
static void exr_unfilter(std::vector<uint8_t>& t, std::vector<uint8_t>& out) {
    for (size_t i = 1; i < t.size(); ++i) t[i] = (uint8_t)(t[i - 1] + t[i] - 128);
    out.resize(t.size());
    const size_t half = (t.size() + 1) / 2;
    for (size_t i = 0; i < t.size(); ++i) out[i] = (i & 1) ? t[half + i / 2] : t[i / 2];
}

// ---- PIZ codec (OpenEXR's default for photographs; published format: ImfPizCompressor / ImfHuf / ImfWav) ----
// A block is: the range [min, max] of non-zero bytes of a 65536-bit "value occurs" bitmap and those
// bytes, then a canonical-Huffman stream of 16-bit symbols (with a run-length escape symbol).  The
// symbols are the 2-D Haar-like wavelet coefficients of every channel (per 16-bit half of a 32-bit
// sample), computed on values that were first mapped through the bitmap's rank table.
namespace {
struct PizBits {
    const uint8_t* in; const uint8_t* end; uint64_t c = 0; int lc = 0;
    void byte() { c = (c << 8) | (in < end ? *in : 0); ++in; lc += 8; }
    uint32_t get(int n) { while (lc < n) byte(); lc -= n; return (uint32_t)((c >> lc) & ((1u << n) - 1u)); }
};
constexpr int kHufEncSize = (1 << 16) + 1, kHufDecBits = 14, kHufDecSize = 1 << kHufDecBits;
struct HufDec { int len = 0; uint32_t lit = 0; std::vector<uint32_t> longs; };

bool huf_uncompress(const uint8_t* data, size_t n, std::vector<uint16_t>& out) {
    if (n == 0) return out.empty();
    if (n < 20) return false;
    auto u32 = [&](size_t o) { uint32_t v; memcpy(&v, data + o, 4); return v; };
    uint32_t im = u32(0); const uint32_t iM = u32(4), n_bits = u32(12);
    if (im >= (uint32_t)kHufEncSize || iM >= (uint32_t)kHufEncSize) return false;
    // code lengths, 6 bits each, runs of zeros packed (59..62: 2..5 zeros, 63: 6 + next 8 bits)
    std::vector<uint64_t> code(kHufEncSize, 0);
    PizBits tb{data + 20, data + n};
    for (uint32_t i = im; i <= iM; ++i) {
        const uint32_t l = tb.get(6);
        code[i] = l;
        if (l == 63) { uint32_t z = tb.get(8) + 6; if (i + z > iM + 1) return false; while (z--) code[i++] = 0; --i; }
        else if (l >= 59) { uint32_t z = l - 59 + 2; if (i + z > iM + 1) return false; while (z--) code[i++] = 0; --i; }
    }
    const uint8_t* bits = tb.in;
    if (bits > data + n) return false;
    // canonical codes: shorter codes have numerically larger prefixes
    uint64_t cnt[59] = {0};
    for (int i = 0; i < kHufEncSize; ++i) { if (code[i] > 58) return false; ++cnt[code[i]]; }
    uint64_t c = 0;
    for (int i = 58; i > 0; --i) { const uint64_t nc = (c + cnt[i]) >> 1; cnt[i] = c; c = nc; }
    for (int i = 0; i < kHufEncSize; ++i) { const uint64_t l = code[i]; if (l > 0) code[i] = l | (cnt[l]++ << 6); }
    // decoding table indexed by the next 14 bits
    std::vector<HufDec> dec(kHufDecSize);
    for (uint32_t i = im; i <= iM; ++i) {
        const uint64_t cd = code[i] >> 6; const int l = (int)(code[i] & 63);
        if (l == 0) continue;
        if (cd >> l) return false;
        if (l > kHufDecBits) { HufDec& d = dec[cd >> (l - kHufDecBits)]; if (d.len) return false; d.longs.push_back(i); }
        else { HufDec* d = &dec[cd << (kHufDecBits - l)]; for (uint64_t k = 1ull << (kHufDecBits - l); k > 0; --k, ++d) { if (d->len || !d->longs.empty()) return false; d->len = l; d->lit = i; } }
    }
    const size_t want = out.size(); size_t no = 0;
    PizBits b{bits, data + n};
    const uint8_t* ie = bits + (n_bits + 7) / 8;
    if (ie > data + n) return false;
    auto emit = [&](uint32_t sym) -> bool {
        if (sym == iM) {                      // run-length escape: repeat the previous symbol `next 8 bits` times
            if (b.lc < 8) b.byte();
            b.lc -= 8;
            uint32_t run = (uint32_t)((b.c >> b.lc) & 0xFF);
            if (no == 0 || no + run > want) return false;
            const uint16_t v = out[no - 1];
            while (run--) out[no++] = v;
            return true;
        }
        if (no >= want) return false;
        out[no++] = (uint16_t)sym;
        return true;
    };
    while (b.in < ie) {
        b.byte();
        while (b.lc >= kHufDecBits) {
            const HufDec& d = dec[(b.c >> (b.lc - kHufDecBits)) & (kHufDecSize - 1)];
            if (d.len) { b.lc -= d.len; if (!emit(d.lit)) return false; }
            else {
                if (d.longs.empty()) return false;
                size_t j = 0;
                for (; j < d.longs.size(); ++j) {
                    const int l = (int)(code[d.longs[j]] & 63);
                    while (b.lc < l && b.in < ie) b.byte();
                    if (b.lc >= l && (code[d.longs[j]] >> 6) == ((b.c >> (b.lc - l)) & ((1ull << l) - 1))) { b.lc -= l; if (!emit(d.longs[j])) return false; break; }
                }
                if (j == d.longs.size()) return false;
            }
        }
    }
    const int pad = (8 - (int)n_bits) & 7;     // the last byte's unused low bits
    b.c >>= pad; b.lc -= pad;
    while (b.lc > 0) {
        const HufDec& d = dec[(b.c << (kHufDecBits - b.lc)) & (kHufDecSize - 1)];
        if (!d.len) return false;
        b.lc -= d.len;
        if (!emit(d.lit)) return false;
    }
    return no == want;
}

inline void wdec14(uint16_t l, uint16_t h, uint16_t& a, uint16_t& b) {
    const int ls = (int16_t)l, hi = (int16_t)h;
    const int ai = ls + (hi & 1) + (hi >> 1);
    a = (uint16_t)(int16_t)ai; b = (uint16_t)(int16_t)(ai - hi);
}
inline void wdec16(uint16_t l, uint16_t h, uint16_t& a, uint16_t& b) {
    const int m = l, d = h;
    const int bb = (m - (d >> 1)) & 0xFFFF;
    const int aa = (d + bb - 0x8000) & 0xFFFF;
    b = (uint16_t)bb; a = (uint16_t)aa;
}
// inverse of the hierarchical 2-D wavelet transform over an nx x ny grid with strides ox, oy
void wav2_decode(uint16_t* in, int nx, int ox, int ny, int oy, uint16_t mx) {
    const bool w14 = mx < (1 << 14);
    const int n = nx > ny ? ny : nx;
    int p = 1, p2;
    while (p <= n) p <<= 1;
    p >>= 1; p2 = p; p >>= 1;
    while (p >= 1) {
        uint16_t* py = in;
        uint16_t* const ey = in + (ptrdiff_t)oy * (ny - p2);
        const ptrdiff_t oy1 = (ptrdiff_t)oy * p, oy2 = (ptrdiff_t)oy * p2, ox1 = (ptrdiff_t)ox * p, ox2 = (ptrdiff_t)ox * p2;
        uint16_t i00, i01, i10, i11;
        for (; py <= ey; py += oy2) {
            uint16_t* px = py;
            uint16_t* const ex = py + (ptrdiff_t)ox * (nx - p2);
            for (; px <= ex; px += ox2) {
                uint16_t *p01 = px + ox1, *p10 = px + oy1, *p11 = p10 + ox1;
                if (w14) { wdec14(*px, *p10, i00, i10); wdec14(*p01, *p11, i01, i11); wdec14(i00, i01, *px, *p01); wdec14(i10, i11, *p10, *p11); }
                else { wdec16(*px, *p10, i00, i10); wdec16(*p01, *p11, i01, i11); wdec16(i00, i01, *px, *p01); wdec16(i10, i11, *p10, *p11); }
            }
            if (nx & p) {      // odd column
                uint16_t* p10 = px + oy1;
                if (w14) wdec14(*px, *p10, i00, *p10); else wdec16(*px, *p10, i00, *p10);
                *px = i00;
            }
        }
        if (ny & p) {          // odd line
            uint16_t* px = py;
            uint16_t* const ex = py + (ptrdiff_t)ox * (nx - p2);
            for (; px <= ex; px += ox2) {
                uint16_t* p01 = px + ox1;
                if (w14) wdec14(*px, *p01, i00, *p01); else wdec16(*px, *p01, i00, *p01);
                *px = i00;
            }
        }
        p2 = p; p >>= 1;
    }
}

// One PIZ block -> the block's bytes in the uncompressed scanline layout.  words[c] = 16-bit words per sample of channel c.
bool piz_decode_block(const uint8_t* src, size_t size, const std::vector<int>& words, int nx, int ny, std::vector<uint8_t>& raw) {
    if (size < 4) return false;
    uint16_t min_nz, max_nz; memcpy(&min_nz, src, 2); memcpy(&max_nz, src + 2, 2);
    size_t pos = 4;
    std::vector<uint8_t> bitmap(8192, 0);
    if (min_nz <= max_nz) {
        if (max_nz >= 8192 || pos + (size_t)(max_nz - min_nz + 1) > size) return false;
        memcpy(bitmap.data() + min_nz, src + pos, (size_t)(max_nz - min_nz + 1));
        pos += (size_t)(max_nz - min_nz + 1);
    }
    std::vector<uint16_t> lut(65536, 0);
    int k = 0;
    for (int i = 0; i < 65536; ++i) if (i == 0 || (bitmap[i >> 3] & (1 << (i & 7)))) lut[k++] = (uint16_t)i;
    const uint16_t max_value = (uint16_t)(k - 1);
    if (pos + 4 > size) return false;
    int32_t length; memcpy(&length, src + pos, 4); pos += 4;
    if (length < 0 || pos + (size_t)length > size) return false;
    size_t total = 0;
    for (int w : words) total += (size_t)w * nx * ny;
    std::vector<uint16_t> tmp(total);
    if (!huf_uncompress(src + pos, (size_t)length, tmp)) return false;
    size_t off = 0;
    std::vector<size_t> start(words.size());
    for (size_t c = 0; c < words.size(); ++c) {
        start[c] = off;
        for (int j = 0; j < words[c]; ++j) wav2_decode(tmp.data() + off + j, nx, words[c], ny, nx * words[c], max_value);
        off += (size_t)words[c] * nx * ny;
    }
    for (auto& v : tmp) v = lut[v];
    raw.resize(total * 2);
    uint8_t* q = raw.data();
    for (int y = 0; y < ny; ++y)
        for (size_t c = 0; c < words.size(); ++c) {
            const size_t n = (size_t)nx * words[c];
            memcpy(q, tmp.data() + start[c] + (size_t)y * n, n * 2);
            q += n * 2;
        }
    return true;
}
}  // namespace

bool read_exr_rgb(const std::string& path, uint32_t* w, uint32_t* h, std::vector<float>* rgb, std::string* err) {
    std::ifstream f(path, std::ios::binary);
    if (!f) { *err = "cannot open image '" + path + "'"; return false; }
    std::stringstream ss; ss << f.rdbuf();
    const std::string d = ss.str();
    ExrCursor c{(const uint8_t*)d.data(), d.size()};
    if (c.u32() != 20000630u) { *err = "not an OpenEXR file: " + path; return false; }
    const uint32_t version = c.u32();
    if ((version & 0xFFu) != 2 || (version & (0x200u | 0x800u | 0x1000u))) { *err = "only single-part scanline OpenEXR files are supported: " + path; return false; }
    std::vector<ExrChannel> channels;
    int compression = -1; int32_t win[4] = {0, 0, -1, -1}; bool have_window = false;
    for (;;) {
        const std::string name = c.cstr();
        if (!c.ok) { *err = "truncated OpenEXR header: " + path; return false; }
        if (name.empty()) break;
        const std::string type = c.cstr();
        const uint32_t size = c.u32();
        if (!c.need(size)) { *err = "truncated OpenEXR header: " + path; return false; }
        ExrCursor a{c.p + c.pos, size};
        if (name == "channels" && type == "chlist") {
            for (;;) {
                ExrChannel ch; ch.name = a.cstr();
                if (!a.ok || ch.name.empty()) break;
                ch.type = a.u32(); a.u32();                 // pixel type; pLinear + 3 reserved bytes
                const uint32_t xs = a.u32(), ys = a.u32();
                if (!a.ok) break;
                if (xs != 1 || ys != 1) { *err = "sub-sampled OpenEXR channels are not supported: " + path; return false; }
                // the reference accepts exactly the names R, G, B and panics on anything else (exr.rs:35-46)
                ch.rgb = ch.name == "R" ? 0 : ch.name == "G" ? 1 : ch.name == "B" ? 2 : -1;
                if (ch.rgb < 0) { *err = "OpenEXR channel '" + ch.name + "' is neither R, G nor B: " + path; return false; }
                if (ch.type > 2) { *err = "unknown OpenEXR pixel type: " + path; return false; }
                channels.push_back(ch);
            }
        } else if (name == "compression" && size >= 1) compression = a.u8();
        else if (name == "dataWindow" && size >= 16) { for (int k = 0; k < 4; ++k) win[k] = (int32_t)a.u32(); have_window = true; }
        c.pos += size;
    }
    if (channels.empty() || compression < 0 || !have_window || win[2] < win[0] || win[3] < win[1]) { *err = "incomplete OpenEXR header: " + path; return false; }
    static const char* codec_names[] = {"NONE", "RLE", "ZIPS", "ZIP", "PIZ", "PXR24", "B44", "B44A", "DWAA", "DWAB"};
    if (compression > 4) {
        *err = std::string("OpenEXR compression ") + (compression < 10 ? codec_names[compression] : "?") + " is not supported (NONE, RLE, ZIPS, ZIP, PIZ are): " + path;
        return false;
    }
    const uint64_t W = (uint64_t)((int64_t)win[2] - win[0] + 1), H = (uint64_t)((int64_t)win[3] - win[1] + 1);
    if (W > (1u << 20) || H > (1u << 20)) { *err = "OpenEXR data window too large: " + path; return false; }
    const uint32_t lines_per_block = compression == 3 ? 16u : (compression == 4 ? 32u : 1u);
    const uint64_t n_blocks = (H + lines_per_block - 1) / lines_per_block;
    size_t line_bytes = 0;
    for (const auto& ch : channels) line_bytes += (ch.type == 1 ? 2 : 4) * (size_t)W;
    // as for PNG: the pixel data cannot be more than ~1100x the file (zlib) -- rejects corrupted windows
    if ((uint64_t)line_bytes * H > (uint64_t)d.size() * 1100 + 4096 || n_blocks * 8 > d.size()) { *err = "OpenEXR data window does not fit the file: " + path; return false; }
    std::vector<uint64_t> offsets(n_blocks);
    for (auto& o : offsets) o = c.u64();
    if (!c.ok) { *err = "truncated OpenEXR offset table: " + path; return false; }
    rgb->assign((size_t)W * H * 3, 0.f);
    std::vector<uint8_t> tmp, raw;
    for (uint64_t b = 0; b < n_blocks; ++b) {
        if (offsets[b] >= c.n) { *err = "bad OpenEXR chunk offset: " + path; return false; }
        ExrCursor k{c.p, c.n, (size_t)offsets[b]};
        const int32_t y0 = (int32_t)k.u32();
        const uint32_t size = k.u32();
        if (!k.ok || !k.need(size) || y0 < win[1] || y0 > win[3]) { *err = "bad OpenEXR chunk: " + path; return false; }
        const uint64_t row0 = (uint64_t)((int64_t)y0 - win[1]);
        const uint64_t rows = std::min<uint64_t>(lines_per_block, H - row0);
        const size_t want = line_bytes * rows;
        const uint8_t* src = k.p + k.pos;
        if (compression == 0 || size == want) {      // a block that does not shrink is stored raw
            if (size != want) { *err = "bad OpenEXR chunk size: " + path; return false; }
            raw.assign(src, src + want);
        } else if (compression == 1) {               // RLE: n < 0 -> -n literal bytes, n >= 0 -> next byte n + 1 times
            tmp.clear();
            size_t i = 0;
            while (i < size) {
                const int n = (int8_t)src[i++];
                if (n < 0) { if (i + (size_t)(-n) > size) break; tmp.insert(tmp.end(), src + i, src + i + (size_t)(-n)); i += (size_t)(-n); }
                else { if (i >= size) break; tmp.insert(tmp.end(), (size_t)n + 1, src[i++]); }
            }
            if (tmp.size() != want) { *err = "OpenEXR RLE block has the wrong length: " + path; return false; }
            exr_unfilter(tmp, raw);
        } else if (compression == 4) {
            std::vector<int> words;
            for (const auto& ch : channels) words.push_back(ch.type == 1 ? 1 : 2);
            if (!piz_decode_block(src, size, words, (int)W, (int)rows, raw) || raw.size() != want) { *err = "OpenEXR PIZ block could not be decoded: " + path; return false; }
        } else {                                     // ZIPS / ZIP: zlib stream
            tmp.assign(want, 0);
            uLongf out_len = want;
            if (uncompress(tmp.data(), &out_len, src, size) != Z_OK || out_len != want) { *err = "OpenEXR inflate failed: " + path; return false; }
            exr_unfilter(tmp, raw);
        }
        // within a block: scanline by scanline, inside a scanline channel by channel (file order = alphabetical)
        const uint8_t* q = raw.data();
        for (uint64_t r = 0; r < rows; ++r) {
            for (const auto& ch : channels) {
                float* out = rgb->data() + ((row0 + r) * W) * 3 + ch.rgb;
                for (uint64_t x = 0; x < W; ++x) {
                    float v;
                    if (ch.type == 1) { uint16_t hv; memcpy(&hv, q, 2); q += 2; v = half_to_float(hv); }
                    else if (ch.type == 2) { memcpy(&v, q, 4); q += 4; }
                    else { uint32_t u; memcpy(&u, q, 4); q += 4; v = (float)u; }
                    out[x * 3] = v;
                }
            }
        }
    }
    *w = (uint32_t)W; *h = (uint32_t)H;
    return true;
}

static void put_be32(std::string& s, uint32_t v) {
    char b[4] = {char(v >> 24), char(v >> 16), char(v >> 8), char(v)};
    s.append(b, 4);
}
static void put_chunk(std::string& out, const char* typ, const std::string& body) {
    put_be32(out, (uint32_t)body.size());
    std::string tb(typ, 4); tb += body;
    out += tb;
    put_be32(out, (uint32_t)crc32(0, (const Bytef*)tb.data(), (uInt)tb.size()));
}

bool write_png_rgb8(const std::string& path, const uint8_t* rgb, uint32_t w, uint32_t h, std::string* err) {
    std::string raw;
    raw.reserve((size_t(w) * 3 + 1) * h);
    for (uint32_t y = 0; y < h; ++y) { raw.push_back(0); raw.append((const char*)rgb + size_t(y) * w * 3, size_t(w) * 3); }
    uLongf clen = compressBound(raw.size());
    std::string comp(clen, 0);
    if (compress2((Bytef*)&comp[0], &clen, (const Bytef*)raw.data(), raw.size(), 6) != Z_OK) { *err = "deflate failed"; return false; }
    comp.resize(clen);
    std::string out("\x89PNG\r\n\x1a\n", 8);
    std::string ihdr; put_be32(ihdr, w); put_be32(ihdr, h);
    ihdr.push_back(8); ihdr.push_back(2); ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0);
    put_chunk(out, "IHDR", ihdr);
    put_chunk(out, "IDAT", comp);
    put_chunk(out, "IEND", std::string());
    FILE* fp = fopen(path.c_str(), "wb");
    if (!fp) { *err = "cannot create '" + path + "'"; return false; }
    bool ok = fwrite(out.data(), 1, out.size(), fp) == out.size();
    fclose(fp);
    if (!ok) *err = "short write to '" + path + "'";
    return ok;
}

// gamma_correct (src/math/mod.rs:51-57)
static float gamma_correct(float f) {
    if (f <= 0.0031308f) return 12.92f * f;
    return 1.055f * std::pow(f, 1.f / 2.4f) - 0.055f;
}

// ImageTexture::into_file quantisation (image.rs:124): clamp(v*255+0.5, 0, 255) as u8.
// Rust's `as u8` truncates toward zero and maps NaN to 0; the reference clamp() leaves NaN
// untouched (both comparisons false), so NaN -> 0.
void film_to_rgb8(const float* film, size_t n_floats, uint8_t* out) {
    for (size_t i = 0; i < n_floats; ++i) {
        float v = gamma_correct(film[i]) * 255.f + 0.5f;
        if (v < 0.f) v = 0.f; else if (v > 255.f) v = 255.f;
        out[i] = (v != v) ? 0 : (uint8_t)v;
    }
}

}  // namespace cpbrt_host
