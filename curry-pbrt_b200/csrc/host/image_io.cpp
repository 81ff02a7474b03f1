// PNG in (texture load, src/texture/image/png.rs:6-27: 8-bit RGB assumed) and PNG out
// (Film::write_image, src/film.rs:35-38 -> src/texture/image.rs:108-127).  zlib does the
// inflate/deflate; the container format and scanline filters are handled here.
#include "scene_host.h"

#include <zlib.h>

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
        if (typ == "IHDR") { W = be32(body); H = be32(body + 4); depth = body[8]; ctype = body[9]; interlace = body[12]; }
        else if (typ == "IDAT") idat.append((const char*)body, len);
        else if (typ == "IEND") break;
        pos += 12 + len;
    }
    // The reference indexes the decoded buffer as packed RGB8 (png.rs:17-24); other layouts
    // would be misread there, so they are rejected here instead of silently reinterpreted.
    if (depth != 8 || ctype != 2 || interlace != 0) { *err = "only 8-bit non-interlaced RGB PNG is supported: " + path; return false; }
    const size_t bpp = 3, stride = size_t(W) * bpp;
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
