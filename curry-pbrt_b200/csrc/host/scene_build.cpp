// Scene construction with the reference's semantics: the top-level factories of
// render_from_file (src/render.rs:63-78), SceneParseStack::parse (src/scene.rs:41-174), the
// camera clipper (src/camera/perspective.rs:54-78, src/camera/mod.rs:58-63), material / light /
// texture / shape property parsing (src/material/mod.rs:28-154, src/light/mod.rs:43-74,
// src/scene/texture_map.rs:17-51, src/geometry/shape/mod.rs:198-216), flattened into the
// index-based arrays of cpbrt_scene_desc.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <fstream>
#include <map>
#include <memory>

#include "scene_host.h"
#include "../bvh_build.h"      // parallel_for

namespace cpbrt_host {

namespace {

#include "cie_tables.inc"

[[noreturn]] void fail(const std::string& m) { throw ParseError(m); }

// ------------------------------------------------------------------ transforms (transform.rs)
Xf xf_from_matrix(const M4& m) {   // From<Matrix4> (transform.rs:131-135)
    Xf r; r.m = m;
    if (!inverse(m, &r.m_inv)) fail("singular matrix in transform directive");
    return r;
}
Xf xf_translate(V3 d) {
    return {M4::rows(1, 0, 0, d.x, 0, 1, 0, d.y, 0, 0, 1, d.z, 0, 0, 0, 1),
            M4::rows(1, 0, 0, -d.x, 0, 1, 0, -d.y, 0, 0, 1, -d.z, 0, 0, 0, 1)};
}
Xf xf_scale(V3 s) {
    return {M4::rows(s.x, 0, 0, 0, 0, s.y, 0, 0, 0, 0, s.z, 0, 0, 0, 0, 1),
            M4::rows(1.f / s.x, 0, 0, 0, 0, 1.f / s.y, 0, 0, 0, 0, 1.f / s.z, 0, 0, 0, 0, 1)};
}
Xf xf_rotate(float angle, V3 axis) {   // transform.rs:38-62 (transpose of pbrt's matrix, Q17)
    V3 a = normalize(axis);
    float radians = angle * (3.14159265358979323846f / 180.0f);   // f32::to_radians
    float s = std::sin(radians), c = std::cos(radians);
    M4 m = M4::rows(
        a.x * a.x + (1.f - a.x * a.x) * c, a.x * a.y * (1.f - c) + a.z * s, a.x * a.z * (1.f - c) - a.y * s, 0.f,
        a.x * a.y * (1.f - c) - a.z * s, a.y * a.y + (1.f - a.y * a.y) * c, a.y * a.z * (1.f - c) + a.x * s, 0.f,
        a.x * a.z * (1.f - c) + a.y * s, a.y * a.z * (1.f - c) - a.x * s, a.z * a.z + (1.f - a.z * a.z) * c, 0.f,
        0.f, 0.f, 0.f, 1.f);
    return {m, m.transpose()};
}
Xf xf_look_at(V3 pos, V3 look, V3 up) {   // transform.rs:28-37
    V3 dir = normalize(look - pos);
    V3 right = normalize(cross(normalize(up), dir));
    V3 new_up = cross(dir, right);
    M4 m_inv = M4::rows(right.x, new_up.x, dir.x, pos.x, right.y, new_up.y, dir.y, pos.y,
                        right.z, new_up.z, dir.z, pos.z, 0.f, 0.f, 0.f, 1.f);
    Xf r; r.m_inv = m_inv;
    if (!inverse(m_inv, &r.m)) fail("LookAt: singular matrix");
    return r;
}
Xf xf_perspective(float fov, float near, float far) {   // transform.rs:103-124
    float inv_tan = 1.f / std::tan(fov * (3.14159265358979323846f / 180.0f) / 2.f);
    float t = far / (far - near);
    return xf_from_matrix(M4::rows(inv_tan, 0, 0, 0, 0, inv_tan, 0, 0, 0, 0, t, -t * near, 0, 0, 1, 0));
}

V3 take_v3(PropertySet& ps) {
    Basics b; std::vector<float> f;
    if (!ps.take_front(3, &b) || !basics_get_floats(b, &f)) fail("transform directive: expected 3 numbers");
    return {f[0], f[1], f[2]};
}

// Transform::parse_from_segment (transform.rs:171-195)
bool parse_transform_directive(const Segment& seg, Xf* out) {
    if (seg.is_block) return false;
    PropertySet ps = seg.object;
    const std::string& t = seg.type;
    if (t == "ConcatTransform" || t == "Transform") {
        Basics b; std::vector<float> f;
        if (!ps.take_front(16, &b) || !basics_get_floats(b, &f)) fail(t + ": expected 16 numbers");
        M4 m;   // Matrix4::from_vec fills column-major
        for (int c = 0; c < 4; ++c) for (int r = 0; r < 4; ++r) m.m[r][c] = f[c * 4 + r];
        *out = xf_from_matrix(m);
        return true;
    }
    if (t == "Translate") { *out = xf_translate(take_v3(ps)); return true; }
    if (t == "Rotate") {
        Basics b; float angle;
        if (!ps.take_front(1, &b) || !basics_get_float(b, &angle)) fail("Rotate: expected an angle");
        V3 axis = take_v3(ps);
        *out = xf_rotate(angle, axis);
        return true;
    }
    if (t == "Scale") { *out = xf_scale(take_v3(ps)); return true; }
    if (t == "LookAt") {
        V3 pos = take_v3(ps), look = take_v3(ps), up = take_v3(ps);
        *out = xf_look_at(pos, look, up);
        return true;
    }
    return false;
}

// ------------------------------------------------------------------ typed property access
bool get_float_prop(const PropertySet& ps, const char* name, float* out) {
    const Property* p = ps.find(name);
    if (!p) return false;
    if (!basics_get_float(p->values, out)) fail(std::string("property '") + name + "': expected a number");
    return true;
}
bool get_usize_prop(const PropertySet& ps, const char* name, uint64_t* out) {
    const Property* p = ps.find(name);
    if (!p) return false;
    int32_t i;
    if (!basics_get_integer(p->values, &i)) fail(std::string("property '") + name + "': expected an integer");
    *out = (uint64_t)(int64_t)i;   // `as usize`
    return true;
}
bool get_string_prop(const PropertySet& ps, const char* name, std::string* out, std::string* file = nullptr) {
    const Property* p = ps.find(name);
    if (!p) return false;
    const Basic* b = basics_get_string(p->values);
    if (!b) fail(std::string("property '") + name + "': expected a string");
    *out = b->str();
    if (file) *file = b->file();
    return true;
}
// RGBSpectrum::parse_from_property (rgb_spectrum.rs:189-209)
void spectrum_from_property(const Property& p, float rgb[3]) {
    if (p.type_name == "rgb") {
        std::vector<float> f;
        if (!basics_get_floats(p.values, &f) || f.size() != 3) fail("rgb property needs 3 numbers");
        rgb[0] = f[0]; rgb[1] = f[1]; rgb[2] = f[2];
    } else if (p.type_name == "spectrum") {
        std::vector<float> f;
        if (!basics_get_floats(p.values, &f) || f.size() % 2) fail("spectrum property needs (lambda, value) pairs");
        std::vector<std::pair<float, float>> s;
        for (size_t i = 0; i < f.size(); i += 2) s.push_back({f[i], f[i + 1]});
        if (s.empty()) fail("empty spectrum");
        spectrum_from_sampled(std::move(s), rgb);
    } else {
        rgb[0] = rgb[1] = rgb[2] = 1.f;   // parse_default
    }
}
void get_spectrum_default(const PropertySet& ps, const char* name, float rgb[3]) {   // get_default::<Spectrum>
    const Property* p = ps.find(name);
    if (p) spectrum_from_property(*p, rgb);
    else rgb[0] = rgb[1] = rgb[2] = 1.f;
}

// ------------------------------------------------------------------ builder state
struct HShape {          // Arc<dyn Shape>
    uint32_t kind, data;
    int32_t xform;       // innermost transform record, chain through .outer
};
struct HPrimitive { HShape shape; int32_t material, light; };

struct Stack {           // SceneParseStack (scene.rs:31-39) -- copied on Attribute/Object entry
    int32_t material = -1;
    bool has_transform = false;
    Xf transform;
    bool has_area_light = false;
    float area_L[3] = {0, 0, 0};
    std::map<std::string, std::pair<int, int32_t>> textures;   // name -> (channels, texture id)
    std::map<std::string, int32_t> named_material;
    bool has_object_name = false;
    std::string object_name;
};

struct CameraInfo {
    Xf raster_to_camera;
    float res_x, res_y;
    bool has_world = false;
    Xf stored;           // TransformCamera.transform (already inverted: m = camera->world)
};

struct Builder {
    SceneData* sd;
    CameraInfo cam;
    std::map<std::string, std::vector<HPrimitive>> objects;
    std::map<std::string, std::pair<int, int32_t>> image_cache;   // path|channels|gamma -> texture id

    int32_t add_transform(const Xf& t, int32_t outer = -1) {
        cpbrt_transform r;
        for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { r.m[i * 4 + j] = t.m.m[i][j]; r.m_inv[i * 4 + j] = t.m_inv.m[i][j]; }
        r.outer = outer; r._pad[0] = r._pad[1] = r._pad[2] = 0;
        sd->transforms.push_back(r);
        return (int32_t)sd->transforms.size() - 1;
    }
    static Xf xf_of(const cpbrt_transform& r) {
        Xf t;
        for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { t.m.m[i][j] = r.m[i * 4 + j]; t.m_inv.m[i][j] = r.m_inv[i * 4 + j]; }
        return t;
    }
    // TransformShape::new(shape, t) on a shape that may already be wrapped: the new wrapper is
    // the outermost link.  Chains are copied because inner records are shared between shapes.
    int32_t wrap_chain(int32_t inner, int32_t outer_rec, std::map<int32_t, int32_t>& memo) {
        if (inner < 0) return outer_rec;
        auto it = memo.find(inner);
        if (it != memo.end()) return it->second;
        cpbrt_transform r = sd->transforms[inner];
        int32_t rest = wrap_chain(r.outer, outer_rec, memo);
        r.outer = rest;
        sd->transforms.push_back(r);
        int32_t id = (int32_t)sd->transforms.size() - 1;
        memo[inner] = id;
        return id;
    }

    // ---- textures --------------------------------------------------------------------------
    int32_t const_texture(int channels, const float v[3]) {
        cpbrt_texture t{};
        t.kind = CPBRT_TEX_CONSTANT; t.channels = (uint32_t)channels;
        t.value[0] = v[0]; t.value[1] = channels == 3 ? v[1] : 0.f; t.value[2] = channels == 3 ? v[2] : 0.f;
        sd->textures.push_back(t);
        return (int32_t)sd->textures.size() - 1;
    }
    int32_t const_spectrum(float v) { float a[3] = {v, v, v}; return const_texture(3, a); }
    int32_t const_float(float v) { float a[3] = {v, 0, 0}; return const_texture(1, a); }

    // ImageTexture::<T>::from_file (image.rs:71-86) (+ apply_inverse_gamma_correct for
    // `spectrum` textures declared through a Texture directive, texture_map.rs:41-49)
    int32_t image_texture(const std::string& path, int channels, bool inverse_gamma) {
        std::string key = path + "|" + std::to_string(channels) + "|" + (inverse_gamma ? "g" : "l");
        auto it = image_cache.find(key);
        if (it != image_cache.end()) return it->second.second;
        size_t dot = path.find_last_of('.');
        std::string ext = dot == std::string::npos ? "" : path.substr(dot + 1);
        if (ext != "png" && ext != "exr") fail("unknown image extension: " + path);   // image.rs:73-77
        uint32_t w, h; std::vector<uint8_t> rgb; std::vector<float> frgb; std::string err;
        if (ext == "png") {
            if (!read_png_rgb8(path, &w, &h, &rgb, &err)) fail(err);
            frgb.resize(rgb.size());
            for (size_t i = 0; i < rgb.size(); ++i) frgb[i] = rgb[i] / 255.f;   // to_rgb (png.rs:25-27)
        } else if (!read_exr_rgb(path, &w, &h, &frgb, &err)) fail(err);
        cpbrt_texture t{};
        t.kind = CPBRT_TEX_IMAGE; t.channels = (uint32_t)channels; t.width = w; t.height = h;
        t.offset = sd->texels.size();
        auto inv_gamma = [](float f) {   // inverse_gamma_correct (math/mod.rs:59-65), 1.05 sic (Q19)
            if (f <= 0.04045f) return f * 1.f / 12.92f;
            return std::pow((f + 0.055f) / 1.05f, 2.4f);
        };
        for (size_t i = 0; i < size_t(w) * h; ++i) {
            const float s[3] = {frgb[i * 3], frgb[i * 3 + 1], frgb[i * 3 + 2]};
            if (channels == 3) {
                for (int c = 0; c < 3; ++c) sd->texels.push_back(inverse_gamma ? inv_gamma(s[c]) : s[c]);
            } else {
                float y = 0.212671f * s[0] + 0.715160f * s[1] + 0.072169f * s[2];   // from_rgb_spectrum -> y()
                sd->texels.push_back(inverse_gamma ? inv_gamma(y) : y);
            }
        }
        sd->textures.push_back(t);
        int32_t id = (int32_t)sd->textures.size() - 1;
        image_cache[key] = {channels, id};
        return id;
    }

    // get_texture::<T> (texture/mod.rs:90-99) ; -1 when the property is absent
    int32_t get_texture(const PropertySet& ps, const char* name, int channels, const Stack& st) {
        const Property* p = ps.find(name);
        if (!p) return -1;
        if (p->type_name == "texture") {
            const Basic* b = basics_get_string(p->values);
            if (!b) fail("texture property needs a name");
            auto it = st.textures.find(b->str());
            if (it == st.textures.end() || it->second.first != channels) fail("texture '" + b->str() + "' not declared with a matching type");
            return it->second.second;
        }
        if (channels == 3) { float rgb[3]; spectrum_from_property(*p, rgb); return const_texture(3, rgb); }
        float f;
        if (!basics_get_float(p->values, &f)) fail(std::string("property '") + name + "': expected a number");
        return const_float(f);
    }
    int32_t tex_or_spectrum(const PropertySet& ps, const char* name, const Stack& st, float dflt) {
        int32_t t = get_texture(ps, name, 3, st);
        return t >= 0 ? t : const_spectrum(dflt);
    }
    int32_t tex_or_float(const PropertySet& ps, const char* name, const Stack& st, float dflt) {
        int32_t t = get_texture(ps, name, 1, st);
        return t >= 0 ? t : const_float(dflt);
    }

    // parse_material_with_type (material/mod.rs:52-154)
    int32_t parse_material(const std::string& type, const PropertySet& ps, const Stack& st) {
        cpbrt_material m{};
        for (int i = 0; i < CPBRT_MAT_MAX_TEX; ++i) m.tex[i] = -1;
        m.sub[0] = m.sub[1] = -1;
        if (type == "matte") {
            m.type = CPBRT_MAT_MATTE;
            m.tex[0] = tex_or_spectrum(ps, "Kd", st, 0.5f);
            m.tex[1] = tex_or_float(ps, "sigma", st, 0.f);
        } else if (type == "glass") {
            m.type = CPBRT_MAT_GLASS;
            m.tex[0] = tex_or_spectrum(ps, "Kr", st, 0.5f);
            m.tex[1] = tex_or_spectrum(ps, "Kt", st, 1.f);
            m.tex[2] = tex_or_float(ps, "index", st, 1.5f);
        } else if (type == "mirror") {
            m.type = CPBRT_MAT_MIRROR;
            m.tex[0] = tex_or_spectrum(ps, "Kr", st, 1.f);
        } else if (type == "plastic") {
            m.type = CPBRT_MAT_PLASTIC;
            m.tex[0] = tex_or_spectrum(ps, "Kd", st, 0.25f);
            m.tex[1] = tex_or_spectrum(ps, "Ks", st, 0.25f);
            m.tex[2] = tex_or_float(ps, "roughness", st, 0.1f);
        } else if (type == "uber") {
            m.type = CPBRT_MAT_UBER;
            m.tex[0] = tex_or_spectrum(ps, "Kd", st, 0.25f);
            m.tex[1] = tex_or_spectrum(ps, "Ks", st, 0.25f);
            m.tex[2] = tex_or_spectrum(ps, "Kr", st, 0.f);
            m.tex[3] = tex_or_spectrum(ps, "Kt", st, 0.f);
            m.tex[4] = tex_or_float(ps, "roughness", st, 0.1f);
            m.tex[5] = get_texture(ps, "uroughness", 1, st);
            m.tex[6] = get_texture(ps, "uroughness", 1, st);   // sic: vroughness reads "uroughness" (mod.rs:105)
            int32_t eta = get_texture(ps, "eta", 1, st);
            if (eta < 0) eta = tex_or_float(ps, "index", st, 1.5f);
            m.tex[7] = eta;
            m.tex[8] = tex_or_spectrum(ps, "opacity", st, 1.f);
        } else if (type == "translucent") {
            m.type = CPBRT_MAT_TRANSLUCENT;
            m.tex[0] = tex_or_spectrum(ps, "Kd", st, 0.25f);
            m.tex[1] = tex_or_spectrum(ps, "Ks", st, 0.25f);
            m.tex[2] = tex_or_float(ps, "roughness", st, 0.1f);
            m.tex[3] = tex_or_spectrum(ps, "reflect", st, 0.5f);
            m.tex[4] = tex_or_spectrum(ps, "transmit", st, 0.5f);
        } else if (type == "mix") {
            m.type = CPBRT_MAT_MIX;
            m.tex[0] = tex_or_spectrum(ps, "amount", st, 0.f);
            std::string n1, n2;
            if (!get_string_prop(ps, "namedmaterial1", &n1) || !get_string_prop(ps, "namedmaterial2", &n2)) fail("mix: namedmaterial1/2 required");
            auto i1 = st.named_material.find(n1), i2 = st.named_material.find(n2);
            if (i1 == st.named_material.end() || i2 == st.named_material.end()) fail("mix: unknown named material");
            m.sub[0] = i1->second; m.sub[1] = i2->second;
        } else {
            fail("unknown material type '" + type + "'");
        }
        sd->materials.push_back(m);
        return (int32_t)sd->materials.size() - 1;
    }

    // parse_shape (shape/mod.rs:198-216) -> bare shapes
    std::vector<HShape> parse_shape(const PropertySet& ps) {
        const char* name = ps.get_name();
        if (!name) fail("Shape needs a type name");
        std::vector<HShape> r;
        std::string n = name;
        if (n == "sphere") {
            float radius = 1.f;
            get_float_prop(ps, "radius", &radius);
            sd->sphere_radius.push_back(radius);
            r.push_back({CPBRT_SHAPE_SPHERE, (uint32_t)sd->sphere_radius.size() - 1, -1});
        } else if (n == "trianglemesh") {
            const Property* pi = ps.find("indices");
            const Property* pp = ps.find("P");
            if (!pi || !pp) fail("trianglemesh needs 'indices' and 'P'");
            std::vector<float> pf;
            if (!basics_get_floats(pp->values, &pf) || pf.size() % 3) fail("trianglemesh: bad 'P'");
            size_t nv = pf.size() / 3;
            std::vector<uint64_t> idx; idx.reserve(pi->values.size());
            for (auto& b : pi->values) {
                if (b.kind != Basic::Int) fail("trianglemesh: 'indices' must be integers");
                idx.push_back((uint64_t)(int64_t)b.i);
            }
            add_mesh(idx, pf.data(), nv, &r);
        } else if (n == "plymesh") {
            std::string fn, file;
            if (!get_string_prop(ps, "filename", &fn, &file)) fail("plymesh needs 'filename'");
            load_ply(join_path(dir_of(file), fn), &r);
        } else {
            fail("unknown shape '" + n + "'");
        }
        return r;
    }
    // create_triangles (triangle.rs:30-36): one Triangle per index triple; uvs default
    void add_mesh(const std::vector<uint64_t>& idx, const float* p, size_t nv, std::vector<HShape>* out) {
        const size_t nt = idx.size() / 3, t0 = sd->triangles.size(), s0 = out->size();
        sd->triangles.resize(t0 + nt);
        out->resize(s0 + nt);
        std::atomic<bool> bad(false);
        cpbrt::parallel_for(nt, 1 << 15, [&](uint64_t lo, uint64_t hi) {
            for (uint64_t t = lo; t < hi; ++t) {
                cpbrt_triangle tri;
                for (int k = 0; k < 3; ++k) {
                    const uint64_t v = idx[t * 3 + k];
                    if (v >= nv) { bad.store(true); return; }
                    tri.p[k][0] = p[v * 3]; tri.p[k][1] = p[v * 3 + 1]; tri.p[k][2] = p[v * 3 + 2];
                }
                tri.uv[0][0] = 0.f; tri.uv[0][1] = 0.f; tri.uv[1][0] = 1.f; tri.uv[1][1] = 0.f; tri.uv[2][0] = 1.f; tri.uv[2][1] = 1.f;
                sd->triangles[t0 + t] = tri;
                (*out)[s0 + t] = HShape{CPBRT_SHAPE_TRIANGLE, (uint32_t)(t0 + t), -1};
            }
        });
        if (bad.load()) fail("trianglemesh: vertex index out of range");
    }
    void load_ply(const std::string& path, std::vector<HShape>* out) {
        std::ifstream f(path, std::ios::binary);
        if (!f) fail("cannot open ply '" + path + "'");
        std::string line;
        auto getl = [&]() { if (!std::getline(f, line)) fail("ply: truncated header"); };   // a failed read leaves `line` as it was: the property loop would never end
        getl(); if (line != "ply") fail("ply: bad magic");
        getl(); if (line != "format binary_little_endian 1.0") fail("ply: only binary_little_endian 1.0");
        getl(); if (line.compare(0, 14, "element vertex") != 0) fail("ply: expected element vertex");
        size_t nv = std::stoull(line.substr(15));
        const char* need[8] = {"x", "y", "z", "nx", "ny", "nz", "u", "v"};
        bool seen[8] = {};
        for (;;) {
            getl();
            if (line.compare(0, 8, "property") != 0) break;
            if (line.compare(9, 5, "float") != 0) fail("ply: vertex properties must be float");
            std::string prop = line.substr(15);
            for (int i = 0; i < 8; ++i) if (prop == need[i]) seen[i] = true;
        }
        for (int i = 0; i < 8; ++i) if (!seen[i]) fail(std::string("ply: required property ") + need[i] + " missing");
        if (line.compare(0, 12, "element face") != 0) fail("ply: expected element face");
        size_t nf = std::stoull(line.substr(13));
        getl(); if (line != "property list uint8 int vertex_indices") fail("ply: bad face property");
        getl(); if (line != "end_header") fail("ply: expected end_header");
        {   // the counts come from the file: they must fit what is left of it before anything is allocated
            const std::streampos here = f.tellg();
            f.seekg(0, std::ios::end);
            const uint64_t left = (uint64_t)(f.tellg() - here);
            f.seekg(here);
            if (nv > left / 32 || nf > left / 13 || nv * 32 + nf * 13 > left) fail("ply: truncated file");
        }
        // one read for the whole body (a 10-million-triangle mesh is ~300 MB: record-by-record stream reads took seconds)
        std::vector<char> body(nv * 32 + nf * 13);
        f.read(body.data(), (std::streamsize)body.size());
        if (!f) fail("ply: truncated file");
        std::vector<float> verts(nv * 3);
        cpbrt::parallel_for(nv, 1 << 16, [&](uint64_t lo, uint64_t hi) {
            for (uint64_t i = lo; i < hi; ++i) memcpy(&verts[i * 3], &body[i * 32], 12);     // x y z; normals and uvs are not kept (plymesh.rs:82-131)
        });
        std::vector<uint64_t> idx(nf * 3);
        const char* fp = body.data() + nv * 32;
        std::atomic<bool> not_tri(false);
        cpbrt::parallel_for(nf, 1 << 16, [&](uint64_t lo, uint64_t hi) {
            for (uint64_t i = lo; i < hi; ++i) {
                if ((uint8_t)fp[i * 13] != 3) { not_tri.store(true); return; }
                int32_t v[3]; memcpy(v, fp + i * 13 + 1, 12);
                for (int k = 0; k < 3; ++k) idx[i * 3 + k] = (uint64_t)(int64_t)v[k];
            }
        });
        if (not_tri.load()) fail("ply: only triangles");
        add_mesh(idx, verts.data(), nv, out);
    }

    // ---- bounds / clipper ------------------------------------------------------------------
    B3 base_bound(const HShape& s) const {
        if (s.kind == CPBRT_SHAPE_SPHERE) {   // sphere.rs:22-27
            float r = sd->sphere_radius[s.data];
            return bounds_of_points({-r, -r, -r}, {r, r, r});
        }
        const cpbrt_triangle& t = sd->triangles[s.data];   // triangle.rs:114-119
        B3 b = bounds_of_points({t.p[0][0], t.p[0][1], t.p[0][2]}, {t.p[1][0], t.p[1][1], t.p[1][2]});
        return bounds_or_point(b, {t.p[2][0], t.p[2][1], t.p[2][2]});
    }
    // TransformCamera::clip + PerspectiveCamera::clip (camera/mod.rs:58-63, perspective.rs:54-78)
    bool clip(const HPrimitive& p) const {
        const int r = clip_status(p);
        if (r == 2) fail("clipper: point at w == 0 (the reference panics here)");
        return r == 1;
    }
    // 0 = kept, 1 = clipped, 2 = a corner projects to w == 0.  Const and allocation-free after its first call on a
    // thread, so the primitives of a big mesh are tested in parallel (parse, "Shape").
    int clip_status(const HPrimitive& p) const {
        if (p.material < 0) return 0;
        B3 b = base_bound(p.shape);
        // chain links from innermost outwards; the outermost link is the TransformShape the
        // clipper composes with world->camera (shape_apply, shape/mod.rs:86-91)
        static thread_local std::vector<int32_t> chain;
        chain.clear();
        for (int32_t x = p.shape.xform; x >= 0; x = sd->transforms[x].outer) chain.push_back(x);
        if (cam.has_world) {
            Xf w2c = cam.stored.inverse();
            Xf outer = chain.empty() ? Xf::identity() : xf_of(sd->transforms[chain.back()]);
            Xf composed = outer.apply(w2c);
            for (size_t i = 0; i + 1 < chain.size(); ++i) b = bounds_apply(b, xf_of(sd->transforms[chain[i]]).m);
            b = bounds_apply(b, composed.m);
        } else {
            for (size_t i = 0; i < chain.size(); ++i) b = bounds_apply(b, xf_of(sd->transforms[chain[i]]).m);
        }
        const M4& c2r = cam.raster_to_camera.m_inv;
        for (int corner = 0; corner < 8; ++corner) {
            V3 pt = (corner % 2 == 0) ? b.mn : b.mx;   // Index<usize> for Bounds: index % 2 (bounds.rs:162-169)
            bool ok = true;
            V3 r = apply_point(c2r, pt, &ok);
            if (!ok) return 2;
            bool clipped = r.z < 0.f || r.x >= cam.res_x || r.x < 0.f || r.y >= cam.res_y || r.y < 0.f;
            if (!clipped) return 0;
        }
        return 1;
    }
    void add_primitive(const HPrimitive& p, bool use_clipper) {
        if (use_clipper && clip(p)) { ++sd->n_clipped; return; }
        cpbrt_primitive q{};
        q.shape.kind = p.shape.kind; q.shape.data = p.shape.data; q.shape.xform = p.shape.xform; q.shape._pad = 0;
        q.material = p.material; q.light = p.light;
        sd->primitives.push_back(q);
    }

    // ---- SceneParseStack::parse (scene.rs:41-174) ------------------------------------------
    void parse(Stack& st, const Segment& seg) {
        if (seg.is_block) {
            if (seg.type == "Attribute") {
                Stack inner = st;
                for (auto& c : seg.children) parse(inner, c);
            } else if (seg.type == "Object") {
                Stack inner = st;
                if (!seg.has_name) fail("ObjectBegin needs a name");
                inner.has_object_name = true; inner.object_name = seg.name;
                for (auto& c : seg.children) parse(inner, c);
            } else {
                fail("unexpected block '" + seg.type + "' inside World");
            }
            return;
        }
        const PropertySet& ps = seg.object;
        const std::string& t = seg.type;
        if (t == "MakeNamedMaterial") {
            std::string mtype;
            if (!get_string_prop(ps, "type", &mtype)) fail("MakeNamedMaterial needs 'string type'");
            const char* name = ps.get_name();
            if (!name) fail("MakeNamedMaterial needs a name");
            int32_t id = parse_material(mtype, ps, st);
            st.named_material[name] = id;
        } else if (t == "Material") {
            const char* name = ps.get_name();
            if (!name) fail("Material needs a type name");
            st.material = parse_material(name, ps, st);
        } else if (t == "Shape") {
            std::vector<HShape> shapes = parse_shape(ps);
            if (st.has_transform) {   // shapes_apply (shape/mod.rs:93-99)
                int32_t rec = add_transform(st.transform);
                for (auto& s : shapes) s.xform = rec;
            }
            // a mesh of a million triangles: the clipper (two matrix products and sixteen point transforms per
            // primitive) runs on all host threads, the primitives are then pushed in order as before
            std::vector<uint8_t> clip_flag;
            if (shapes.size() >= 65536 && !st.has_area_light && !st.has_object_name && st.material >= 0) {
                clip_flag.resize(shapes.size());
                cpbrt::parallel_for(shapes.size(), 1 << 14, [&](uint64_t lo, uint64_t hi) {
                    for (uint64_t i = lo; i < hi; ++i) clip_flag[i] = (uint8_t)clip_status(HPrimitive{shapes[i], st.material, -1});
                });
                for (uint8_t f : clip_flag) if (f == 2) fail("clipper: point at w == 0 (the reference panics here)");
                sd->primitives.reserve(sd->primitives.size() + shapes.size());
            }
            for (size_t si = 0; si < shapes.size(); ++si) {
                auto& s = shapes[si];
                HPrimitive p{s, -1, -1};
                if (!clip_flag.empty()) {
                    if (clip_flag[si]) { ++sd->n_clipped; continue; }
                    p.material = st.material;
                    add_primitive(p, false);
                    continue;
                }
                if (st.has_area_light) {
                    cpbrt_light l{};
                    l.type = CPBRT_LIGHT_AREA;
                    l.shape.kind = s.kind; l.shape.data = s.data; l.shape.xform = s.xform;
                    l.L[0] = st.area_L[0]; l.L[1] = st.area_L[1]; l.L[2] = st.area_L[2];
                    l.tex = -1;
                    sd->lights.push_back(l);
                    p.light = (int32_t)sd->lights.size() - 1;
                } else {
                    if (st.material < 0) fail("Shape without a current Material");
                    p.material = st.material;
                }
                if (st.has_object_name) objects[st.object_name].push_back(p);
                else add_primitive(p, true);
            }
        } else if (t == "ObjectInstance") {
            const char* name = ps.get_name();
            if (!name) fail("ObjectInstance needs a name");
            auto it = objects.find(name);
            if (it != objects.end()) {
                std::vector<HPrimitive> prims = it->second;
                if (st.has_transform) {   // primitives_apply (primitive.rs:115-130): wrap, do not compose
                    int32_t rec = add_transform(st.transform);
                    std::map<int32_t, int32_t> memo;
                    for (auto& p : prims) p.shape.xform = wrap_chain(p.shape.xform, rec, memo);
                }
                for (auto& p : prims) add_primitive(p, true);
            }
        } else if (t == "LightSource") {
            const char* name = ps.get_name();
            if (!name) fail("LightSource needs a type name");
            std::string n = name;
            cpbrt_light l{};
            l.tex = -1; l.shape.xform = -1;
            if (n == "point") {   // light/mod.rs:45-48, point.rs:10-26
                l.type = CPBRT_LIGHT_POINT;
                get_spectrum_default(ps, "I", l.L);
                V3 p = {0.f, 0.f, 0.f};
                if (st.has_transform) p = apply_point(st.transform.m, p);
                l.v[0] = p.x; l.v[1] = p.y; l.v[2] = p.z;
            } else if (n == "distant") {   // light/mod.rs:49-57, distant.rs:10-26
                l.type = CPBRT_LIGHT_DISTANT;
                get_spectrum_default(ps, "L", l.L);
                V3 w = {0.f, 0.f, -1.f};
                const Property* from = ps.find("from");
                if (from) {
                    const Property* to = ps.find("to");
                    std::vector<float> a, b;
                    if (!to || !basics_get_floats(from->values, &a) || !basics_get_floats(to->values, &b) || a.size() < 3 || b.size() < 3)
                        fail("distant light: 'from' needs a matching 'to'");
                    w = V3{b[0], b[1], b[2]} - V3{a[0], a[1], a[2]};
                }
                w = normalize(w);
                if (st.has_transform) w = apply_vector(st.transform.m, w);
                l.v[0] = w.x; l.v[1] = w.y; l.v[2] = w.z;
            } else if (n == "infinite") {   // light/mod.rs:58-61
                l.type = CPBRT_LIGHT_INFINITE;
                const Property* mp = ps.find("mapname");
                if (mp) {
                    if (mp->type_name != "string") fail("infinite light: mapname must be a string");
                    const Basic* b = basics_get_string(mp->values);
                    if (!b) fail("infinite light: mapname must be a string");
                    l.tex = image_texture(join_path(dir_of(b->file()), b->str()), 3, false);
                } else {
                    l.tex = const_spectrum(1.f);
                }
            } else {
                fail("unknown light '" + n + "'");
            }
            sd->lights.push_back(l);
        } else if (t == "AreaLightSource") {
            const char* name = ps.get_name();
            if (!name || std::string(name) != "diffuse") fail("only AreaLightSource \"diffuse\" exists");
            st.has_area_light = true;
            get_spectrum_default(ps, "L", st.area_L);
        } else if (t == "Texture") {   // texture_map.rs:17-51
            std::string fn, file;
            if (!get_string_prop(ps, "filename", &fn, &file)) fail("Texture needs 'string filename'");
            PropertySet copy = ps;
            Basics b1, b2;
            if (!copy.take_front(1, &b1) || !copy.take_front(1, &b2) || !basics_get_string(b1) || !basics_get_string(b2))
                fail("Texture needs a name and a type");
            std::string name = b1[0].str(), ttype = b2[0].str();
            std::string path = join_path(dir_of(file), fn);
            if (ttype == "float") st.textures[name] = {1, image_texture(path, 1, false)};
            else if (ttype == "spectrum") st.textures[name] = {3, image_texture(path, 3, true)};
            else fail("Texture type must be float or spectrum");
        } else if (t == "Transform") {
            Xf x;
            if (!parse_transform_directive(seg, &x)) fail("bad Transform");
            st.transform = x; st.has_transform = true;
        } else {
            Xf x;
            if (parse_transform_directive(seg, &x)) {
                st.transform = st.has_transform ? st.transform.apply(x) : x;   // scene.rs:158-165
                st.has_transform = true;
            } else {
                fprintf(stderr, "cpbrt: Parsing scene error: unknown directive '%s' ignored\n", t.c_str());
            }
        }
    }
};

const Segment* find_object(const std::vector<Segment>& segs, const char* type) {
    for (auto& s : segs) if (!s.is_block && s.type == type) return &s;
    return nullptr;
}

}  // namespace

void spectrum_from_sampled(std::vector<std::pair<float, float>> sampled, float rgb[3]) {
    std::stable_sort(sampled.begin(), sampled.end(), [](const std::pair<float, float>& a, const std::pair<float, float>& b) { return a.first < b.first; });
    float xyz[3] = {0.f, 0.f, 0.f};
    const size_t n = sampled.size();
    for (int i = 0; i < kCieSamples; ++i) {
        float l = (float)(kCieFirstLambda + i), val;
        if (l <= sampled.front().first) val = sampled.front().second;
        else if (l >= sampled.back().first) val = sampled.back().second;
        else {
            // find_interval (rgb_spectrum.rs:223-237)
            size_t first = 0, len = n;
            while (len > 0) {
                size_t half = len / 2, middle = first + half;
                if (sampled[middle].first <= l) { first = middle + 1; len -= half + 1; }
                else len = half;
            }
            size_t off = first - 1;
            if (off > n - 2) off = n - 2;
            float t = (l - sampled[off].first) / (sampled[off + 1].first - sampled[off].first);
            val = sampled[off].second * t + sampled[off + 1].second * (1.f - t);   // reversed lerp (math/mod.rs:13)
        }
        for (int c = 0; c < 3; ++c) xyz[c] += kCieXYZ[i][c] * val;
    }
    float scale = ((float)(kCieFirstLambda + kCieSamples - 1) - (float)kCieFirstLambda) / (kCieYIntegral * (float)kCieSamples);
    for (int c = 0; c < 3; ++c) xyz[c] *= scale;
    rgb[0] = 3.240479f * xyz[0] - 1.537150f * xyz[1] - 0.498535f * xyz[2];
    rgb[1] = -0.969256f * xyz[0] + 1.875991f * xyz[1] + 0.041556f * xyz[2];
    rgb[2] = 0.055648f * xyz[0] - 0.204043f * xyz[1] + 1.057311f * xyz[2];
}

void SceneData::refresh_desc() {
    desc.abi_version = CPBRT_ABI_VERSION;
    desc.n_primitives = primitives.size(); desc.primitives = primitives.data();
    desc.n_triangles = triangles.size();   desc.triangles = triangles.data();
    desc.n_spheres = sphere_radius.size(); desc.sphere_radius = sphere_radius.data();
    desc.n_transforms = (uint32_t)transforms.size(); desc.transforms = transforms.data();
    desc.n_materials = (uint32_t)materials.size();   desc.materials = materials.data();
    desc.n_textures = (uint32_t)textures.size();     desc.textures = textures.data();
    desc.n_texels = texels.size();         desc.texels = texels.data();
    desc.n_lights = (uint32_t)lights.size();         desc.lights = lights.data();
    desc.sampler.permutation = permutation.data();
}

void build_scene_from_segments(const std::vector<Segment>& segments, uint64_t perm_seed, SceneData* sd) {
    Builder b;
    b.sd = sd;
    memset(&sd->desc, 0, sizeof sd->desc);

    // parse_find_eat::<Transform>: the FIRST top-level transform directive only (Q23)
    bool has_cam_xf = false; Xf cam_xf;
    for (auto& s : segments) if (parse_transform_directive(s, &cam_xf)) { has_cam_xf = true; break; }

    const Segment* cam = find_object(segments, "Camera");
    if (!cam) fail("no Camera directive");
    const char* cam_name = cam->object.get_name();
    if (!cam_name || std::string(cam_name) != "perspective") fail("only Camera \"perspective\" exists");
    float fov = 90.f;
    get_float_prop(cam->object, "fov", &fov);

    const Segment* smp = find_object(segments, "Sampler");
    if (!smp) fail("no Sampler directive");
    const char* smp_name = smp->object.get_name();
    if (!smp_name || std::string(smp_name) != "halton") fail("only Sampler \"halton\" exists");
    uint64_t spp = 0;
    if (!get_usize_prop(smp->object, "pixelsamples", &spp)) fail("Sampler needs 'integer pixelsamples'");

    const Segment* film = find_object(segments, "Film");
    if (!film) fail("no Film directive");
    const char* film_name = film->object.get_name();
    if (!film_name || std::string(film_name) != "image") fail("only Film \"image\" exists");
    uint64_t xres = 640, yres = 480;
    get_usize_prop(film->object, "xresolution", &xres);
    get_usize_prop(film->object, "yresolution", &yres);
    sd->film_filename = "curry-pbrt.png";
    get_string_prop(film->object, "filename", &sd->film_filename);

    const Segment* integ = find_object(segments, "Integrator");
    if (!integ) fail("no Integrator directive");
    const char* integ_name = integ->object.get_name();
    uint64_t max_depth = 5;
    get_usize_prop(integ->object, "maxdepth", &max_depth);
    if (integ_name && std::string(integ_name) == "directlighting") sd->desc.integrator = CPBRT_INTEGRATOR_DIRECT;
    else if (integ_name && std::string(integ_name) == "path") sd->desc.integrator = CPBRT_INTEGRATOR_PATH;
    else fail("unknown integrator");
    sd->desc.max_depth = (uint32_t)max_depth;
    sd->desc.film_width = (uint32_t)xres; sd->desc.film_height = (uint32_t)yres;

    // PerspectiveCamera::new (perspective.rs:11-46)
    {
        const float near = 1e-2f, far = 1000.f;
        float rx = (float)xres, ry = (float)yres;
        float aspect = rx / ry;
        float mnx, mny, mxx, mxy;
        if (aspect > 1.f) { mnx = rmin(-aspect, aspect); mxx = rmax(-aspect, aspect); mny = -1.f; mxy = 1.f; }
        else { mnx = -1.f; mxx = 1.f; float a = -1.f / aspect, c = 1.f / aspect; mny = rmin(a, c); mxy = rmax(a, c); }
        float dx = mxx - mnx, dy = mxy - mny;
        Xf s2r = xf_translate({-mnx, -mxy, 0.f}).apply(xf_scale({1.f / dx, -(1.f / dy), 1.f})).apply(xf_scale({rx, ry, 1.f}));
        Xf c2s = xf_perspective(fov, near, far);
        Xf c2r = c2s.apply(s2r);
        b.cam.raster_to_camera = c2r.inverse();
        b.cam.res_x = rx; b.cam.res_y = ry;
    }
    cpbrt_camera& C = sd->desc.camera;
    memset(&C, 0, sizeof C);
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) {
        C.raster_to_camera[i * 4 + j] = b.cam.raster_to_camera.m.m[i][j];
        C.camera_to_raster[i * 4 + j] = b.cam.raster_to_camera.m_inv.m[i][j];
        C.camera_to_world[i * 4 + j] = (i == j) ? 1.f : 0.f;
    }
    float lens_radius;
    if (get_float_prop(cam->object, "lensradius", &lens_radius)) {   // with_lens (lens.rs:40-51)
        C.has_lens = 1; C.lens_radius = lens_radius; C.focal_distance = 1e6f;
        get_float_prop(cam->object, "focaldistance", &C.focal_distance);
    }
    if (has_cam_xf) {   // camera_apply -> TransformCamera::new stores transform.inverse() (camera/mod.rs:49-55,74-79)
        b.cam.has_world = true;
        b.cam.stored = cam_xf.inverse();
        C.has_camera_to_world = 1;
        for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) C.camera_to_world[i * 4 + j] = b.cam.stored.m.m[i][j];
    }

    halton_sampler_params((uint32_t)spp, (uint32_t)xres, (uint32_t)yres, &sd->desc.sampler);
    sd->permutation.resize(CPBRT_PERM_TABLE_LEN);
    make_permutation(perm_seed, sd->permutation.data());

    // SceneBuilder::build_with_clipper over the World block (scene.rs:181-199)
    const Segment* world = nullptr;
    for (auto& s : segments) if (s.is_block && s.type == "World") { world = &s; break; }
    if (!world) fail("no WorldBegin block");
    Stack st;
    for (auto& c : world->children) b.parse(st, c);

    sd->refresh_desc();
}

}  // namespace cpbrt_host
