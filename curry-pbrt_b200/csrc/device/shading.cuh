// Shading-side device functions: textures, BxDFs / BSDF, shape sampling, lights.
//   textures   src/texture/constant.rs, src/texture/image.rs:92-101
//   BSDF       src/material/bxdf/mod.rs:71-214, lambertian.rs, oren_nayar.rs, specular.rs,
//              src/math/distribution.rs:16-53, src/material/{matte,mirror,glass}.rs
//   shapes     src/geometry/shape/mod.rs:17-63, triangle.rs:120-126, sphere.rs:54-105,
//              transform.rs:60-70
//   lights     src/light/{mod,area,point,distant}.rs
#pragma once
#include "halton.cuh"
#include "shapes.cuh"

namespace cpbrt {

// ---------------------------------------------------------------------------------- textures
CP_DEV Spec tex_eval(const DevScene& sc, int32_t id, P2 uv) {
    const DevTexture& t = sc.textures[id];
    if (t.kind == 0) return spec3(t.value[0], t.value[1], t.value[2]);
    // ImageTexture::evaluate: v flipped, nearest texel, saturating cast then min (Q20)
    float u = uv.x, v = 1.f - uv.y;
    uint32_t ix = min(f32_to_u32_sat(u * (float)t.width), t.width - 1);
    uint32_t iy = min(f32_to_u32_sat(v * (float)t.height), t.height - 1);
    const float* px = sc.texels + t.offset + ((uint64_t)ix + (uint64_t)iy * t.width) * t.channels;
    if (t.channels == 3) return spec3(__ldg(px), __ldg(px + 1), __ldg(px + 2));
    return spec3(__ldg(px), 0.f, 0.f);
}

// ---------------------------------------------------------------------------------- BxDFs
// Distribution1D::from + sample_remap (distribution.rs:16-53) for a table of LEN entries
template <int LEN>
CP_DEV void distribution_sample_remap(const float* f, int len, float u, int* idx, float* pdf_out, float* remap) {
    float cdf[LEN], pdf[LEN];
    float n = (float)len, last = 0.f;
#pragma unroll
    for (int i = 0; i < LEN; ++i) if (i < len) { cdf[i] = last + f[i] / n; last = cdf[i]; }
    float f_sum = last;
#pragma unroll
    for (int i = 0; i < LEN; ++i) if (i < len) { cdf[i] /= f_sum; pdf[i] = f[i] / (n * f_sum); }
    *idx = len - 1; *pdf_out = 0.f; *remap = 1.f;
    bool found = false;
#pragma unroll
    for (int i = 0; i < LEN; ++i) if (i < len) {
        if (!found && u <= cdf[i]) { *idx = i; *pdf_out = pdf[i]; *remap = (cdf[i] - u) / pdf[i]; found = true; }
        if (!found && i == len - 1) *pdf_out = pdf[i];
    }
}

CP_DEV float fresnel_dielectric(float eta_i_in, float eta_t_in, float cos_i) {   // specular.rs:24-44
    float ei = cos_i > 0.f ? eta_i_in : eta_t_in, et = cos_i > 0.f ? eta_t_in : eta_i_in;
    cos_i = fabsf(cos_i);
    float sin_i = sqrtf(1.f - cos_i * cos_i);
    float sin_t = sin_i * ei / et;
    if (sin_t >= 1.f) return 1.f;
    float cos_t = sqrtf(1.f - sin_t * sin_t);
    float t0 = et * cos_i, t1 = ei * cos_t;
    float r_perl = (t0 - t1) / (t0 + t1);
    t0 = ei * cos_i; t1 = et * cos_t;
    float r_perp = (t0 - t1) / (t0 + t1);
    return (r_perl * r_perl + r_perp * r_perp) / 2.f;
}

enum LobeKind : int { LOBE_LAMBERT_R = 0, LOBE_OREN_NAYAR = 1 };
enum DeltaKind : int { DELTA_SPEC_R_NOOP = 0, DELTA_SPEC_R_DIELECTRIC = 1, DELTA_SPEC_T = 2 };

struct Lobe { int kind; Spec r; float a, b; };
struct DeltaLobe { int kind; Spec r; float eta_a, eta_b; };

CP_DEV Spec lobe_f(const Lobe& l, V3 wo, V3 wi) {
    if (l.kind == LOBE_LAMBERT_R) return l.r * kInvPi;   // lambertian.rs:14-18
    // OrenNayar::f (oren_nayar.rs:18-30)
    float sin_alpha, tan_beta;
    if (cos_theta(wi) < cos_theta(wo)) { sin_alpha = sin_theta(wi); tan_beta = tan_theta(wo); }
    else { sin_alpha = sin_theta(wo); tan_beta = tan_theta(wi); }
    return l.r * ((l.a + l.b * rmax(0.f, cos_delta_phi(wi, wo)) * sin_alpha * tan_beta) / kPi);
}
// BxDF::sample_f default for a reflective lobe (bxdf/mod.rs:20-37)
CP_DEV void lobe_sample_f(const Lobe& l, V3 wo, P2 u, V3* wi, Spec* f, float* pdf) {
    V3 w = cosine_sample_hemisphere(u, pdf);
    if (wo.z < 0.f) w.z *= -1.f;
    *wi = w;
    *f = lobe_f(l, wo, w);
}
// DeltaBxDF::sample_f (specular.rs:56-66, 87-101)
CP_DEV bool delta_sample_f(const DeltaLobe& l, V3 wo, V3* wi, Spec* s) {
    if (l.kind == DELTA_SPEC_T) {
        float ei = cos_theta(wo) > 0.f ? l.eta_a : l.eta_b, et = cos_theta(wo) > 0.f ? l.eta_b : l.eta_a;
        float eta = ei / et;
        float sg = wo.z != wo.z ? wo.z : copysignf(1.f, wo.z);   // f32::signum
        if (!refract(wo, normalize(v3(0.f, 0.f, sg)), eta, wi)) return false;
        float ft = 1.f - fresnel_dielectric(l.eta_a, l.eta_b, cos_theta(*wi));
        if (ft == 0.f) return false;
        *s = l.r * ft / fabsf(cos_theta(*wi));
        return true;
    }
    *wi = v3(-wo.x, -wo.y, wo.z);
    float c = cos_theta(*wi);
    float fr = l.kind == DELTA_SPEC_R_DIELECTRIC ? fresnel_dielectric(l.eta_a, l.eta_b, c) : 1.f;
    if (fr == 0.f) return false;
    *s = l.r * fr / fabsf(c);
    return true;
}

template <int NL, int ND>
struct BSDF {
    static constexpr bool kNoDelta = ND == 0;
    V3 n, sn, snx, sny;
    Lobe lobes[NL > 0 ? NL : 1]; int n_lobes;
    DeltaLobe deltas[ND > 0 ? ND : 1]; int n_deltas;

    CP_DEV void init(V3 normal) {   // BSDF::new(n, n) (bxdf/mod.rs:83-97)
        sn = normal;
        coordinate_system(sn, &snx, &sny);
        n = normalize(v3(dot(normal, snx), dot(normal, sny), dot(normal, sn)));
        n_lobes = 0; n_deltas = 0;
    }
    CP_DEV void init_frame(const float4* __restrict__ f) { load_frame(f); n_lobes = 0; n_deltas = 0; }
    // the same frame read from DevScene::prim_frame (host-evaluated, see scene_dev.cuh)
    CP_DEV void load_frame(const float4* __restrict__ f) {
        const float4 f0 = __ldg(f), f1 = __ldg(f + 1), f2 = __ldg(f + 2);
        sn = v3(f0.x, f0.y, f0.z); snx = v3(f0.w, f1.x, f1.y); sny = v3(f1.z, f1.w, f2.x); n = v3(f2.y, f2.z, f2.w);
    }
    CP_DEV V3 local_to_world(V3 w) const {
        return normalize(v3((snx.x * w.x + sny.x * w.y) + sn.x * w.z, (snx.y * w.x + sny.y * w.y) + sn.y * w.z, (snx.z * w.x + sny.z * w.y) + sn.z * w.z));
    }
    CP_DEV V3 world_to_local(V3 w) const { return normalize(v3(dot(w, snx), dot(w, sny), dot(w, sn))); }
    // wo in the local frame: every sampling / evaluation routine below needs it, computed once per hit (set_wo)
    V3 wo_local;
    CP_DEV void set_wo(V3 wo) { wo_local = world_to_local(wo); }

    // sample_no_delta_f (bxdf/mod.rs:136-159)
    CP_DEV bool sample_no_delta_f(V3 wo, P2 u, V3* wi, Spec* f, float* pdf) const {
        if (NL == 0 || n_lobes == 0) { *wi = v3(0.f, 0.f, 0.f); *pdf = 0.f; return false; }
        float fl = u.x * (float)n_lobes;                      // sample_usize_remap (math/mod.rs:81-86)
        float tr = truncf(fl), remap = fl - tr;
        uint32_t index = min(f32_to_u32_sat(tr), (uint32_t)n_lobes - 1);
        P2 u2; u2.x = remap; u2.y = u.y;
        V3 wo_l = wo_local, wi_l; float p;
        lobe_sample_f(lobes[NL == 1 ? 0 : index], wo_l, u2, &wi_l, f, &p);
        *wi = local_to_world(wi_l);
        *pdf = p / (float)n_lobes;
        return true;
    }
    // sample_delta_f (bxdf/mod.rs:160-175)
    CP_DEV bool sample_delta_f(V3 wo, float u, V3* wi, Spec* f, float* pdf) const {
        if (ND == 0 || n_deltas == 0) { *wi = v3(0.f, 0.f, 0.f); *pdf = 0.f; return false; }
        V3 wo_l = wo_local;
        V3 w[ND > 0 ? ND : 1]; Spec s[ND > 0 ? ND : 1]; float lm[ND > 0 ? ND : 1]; int cnt = 0;
#pragma unroll
        for (int i = 0; i < ND; ++i) if (i < n_deltas) { V3 wi_l; Spec si; if (delta_sample_f(deltas[i], wo_l, &wi_l, &si)) { w[cnt] = wi_l; s[cnt] = si; lm[cnt] = lum(si); ++cnt; } }
        if (cnt == 0) { *wi = v3(0.f, 0.f, 0.f); *pdf = 0.f; return false; }   // the reference panics on an empty table
        int i; float p, remap;
        distribution_sample_remap<(ND > 0 ? ND : 1)>(lm, cnt, u, &i, &p, &remap);
        V3 wsel = w[0]; Spec ssel = s[0];
#pragma unroll
        for (int k = 1; k < ND; ++k) if (k == i) { wsel = w[k]; ssel = s[k]; }
        *wi = local_to_world(wsel); *f = ssel; *pdf = p;
        return true;
    }
    // no_delta_f_pdf (bxdf/mod.rs:176-198); all supported non-delta lobes are reflective
    CP_DEV bool no_delta_f_pdf(V3 wo_w, V3 wi_w, Spec* f, float* pdf) const {
        if (NL == 0 || n_lobes == 0) { *pdf = 0.f; return false; }
        V3 wo = wo_local, wi = world_to_local(wi_w);
        bool reflect = dot(wo, n) * dot(wi, n) > 0.f;
        if (!reflect) { *pdf = 0.f / 0.f; return false; }   // empty transmit list: f = None
        Spec acc = spec(0.f); float p = 0.f;
#pragma unroll
        for (int i = 0; i < NL; ++i) if (i < n_lobes) { acc = acc + lobe_f(lobes[i], wo, wi); p += wi.z * kInvPi; }   // BxDF::pdf default (:38-40)
        *f = acc; *pdf = p / (float)n_lobes;
        return true;
    }
    // sample_f (bxdf/mod.rs:199-214)
    CP_DEV bool sample_f(const DevSampler& S, SamplerState& st, V3 wo, V3* wi, Spec* f, float* pdf, bool* is_delta) const {
        float w[2]; w[0] = (float)n_lobes; w[1] = (float)n_deltas;
        int i; float p, remap;
        bool has; float ip;
#if CPBRT_HALTON_PAIR
        if (NL > 0 && n_lobes > 0) {
            // the second dimension is used whenever the non-delta group is chosen: both are drawn
            // together (halton.cuh) and the second is handed back if a delta lobe was picked
            const P2 uu = sampler_get_2d(S, st);
            distribution_sample_remap<2>(w, 2, uu.x, &i, &p, &remap);
            if (i == 0) { P2 u; u.x = remap; u.y = uu.y; has = sample_no_delta_f(wo, u, wi, f, &ip); *is_delta = false; }
            else { st.dim -= 1; has = sample_delta_f(wo, remap, wi, f, &ip); *is_delta = true; }
            *pdf = p * ip;
            return has;
        }
#endif
        distribution_sample_remap<2>(w, 2, sampler_get_1d(S, st), &i, &p, &remap);
        if (i == 0) { P2 u; u.x = remap; u.y = sampler_get_1d(S, st); has = sample_no_delta_f(wo, u, wi, f, &ip); *is_delta = false; }
        else { has = sample_delta_f(wo, remap, wi, f, &ip); *is_delta = true; }
        *pdf = p * ip;
        return has;
    }
};

// ---------------------------------------------------------------------------------- shape sampling
CP_DEV TriVerts light_tri(const DevLight& l) {
    TriVerts t;
    t.p0 = v3(l.tri[0], l.tri[1], l.tri[2]); t.p1 = v3(l.tri[3], l.tri[4], l.tri[5]); t.p2 = v3(l.tri[6], l.tri[7], l.tri[8]);
    return t;
}
// Shape::sample for the light's bare shape (triangle.rs:120-126, sphere.rs:54-65)
CP_DEV void bare_sample(const DevScene& sc, const DevLight& l, SamplerState& st, ShapePoint* sp, float* pdf) {
    if (l.shape_kind == 1) {
        V3 d = uniform_sample_sphere_q21(sampler_get_2d(sc.sampler, st));
        V3 p = d * l.radius;
        sp->n = normalize(d);
        p = p * (l.radius / length(p));
        sp->p = p; sp->p_error = gamma_n(5) * vabs(p); sp->uv = sphere_uv(l.radius, p);
    } else {
        TriVerts t = light_tri(l);
        P2 b = uniform_sample_triangle(sampler_get_2d(sc.sampler, st));
        float b0 = b.x, b1 = b.y, b2 = 1.f - b.x - b.y;
        sp->p = tri_point(t, b0, b1, b2); sp->n = v3(l.tri_n[0], l.tri_n[1], l.tri_n[2]); sp->uv = tri_uv(nullptr, b0, b1, b2);
        sp->p_error = gamma_n(6) * tri_abs_sum(t, b0, b1, b2);
    }
    *pdf = 1.f / l.area;
}
// default_sample_by_point (shape/mod.rs:24-41)
CP_DEV void bare_default_sample_by_point(const DevScene& sc, const DevLight& l, V3 point, SamplerState& st, ShapePoint* sp, float* pdf) {
    float p0;
    bare_sample(sc, l, st, sp, &p0);
    V3 wi = sp->p - point;
    if (length_sq(wi) == 0.f) { *pdf = 0.f; return; }
    float p = p0 * length_sq(wi) / -dot(normalize(wi), sp->n);
    *pdf = (p != p || isinf(p)) ? 0.f : p;
}
CP_DEV void bare_sample_by_point(const DevScene& sc, const DevLight& l, V3 point, SamplerState& st, ShapePoint* sp, float* pdf) {
    if (l.shape_kind != 1) { bare_default_sample_by_point(sc, l, point, st, sp, pdf); return; }
    // Sphere::sample_by_point (sphere.rs:66-95): cone sampling from outside
    const float radius = l.radius;
    float distance_2 = length_sq(point), radius_2 = radius * radius;
    if (distance_2 <= radius_2) { bare_default_sample_by_point(sc, l, point, st, sp, pdf); return; }
    float distance = sqrtf(distance_2);
    V3 z = point / distance, x, y;
    coordinate_system(z, &x, &y);
    P2 u = sampler_get_2d(sc.sampler, st);
    float sin_theta_max_2 = radius_2 / distance_2;
    float cos_theta_max = sqrtf(rmax(1.f - sin_theta_max_2, 0.f));
    float cos_t = (1.f - u.x) + u.x * cos_theta_max;
    float sin_t = sqrtf(rmax(1.f - cos_t * cos_t, 0.f));
    float phi = u.y * 2.f * kPi;
    float ds = distance * cos_t - sqrtf(rmax(radius_2 - distance_2 * sin_t * sin_t, 0.f));
    float cos_alpha = (distance_2 + radius_2 - ds * ds) / (2.f * distance * radius);
    float sin_alpha = sqrtf(rmax(1.f - cos_alpha * cos_alpha, 0.f));
    V3 d = cos_alpha * z + sin_alpha * cosf(phi) * x + sin_alpha * sinf(phi) * y;
    sp->n = normalize(d);
    V3 p = d * radius;
    sp->p = p; sp->p_error = gamma_n(5) * vabs(p); sp->uv = sphere_uv(radius, p);
    *pdf = 1.f / (2.f * kPi * (1.f - cos_theta_max));
}
CP_DEV float bare_default_by_point_pdf(const DevLight& l, V3 point, const ShapePoint& sp) {   // shape/mod.rs:42-52
    V3 d = point - sp.p;
    float d2 = length_sq(d), dist = sqrtf(d2);
    float pdf = d2 / (fabsf(dot(d / dist, sp.n)) * l.area);
    return (pdf != pdf || isinf(pdf)) ? 0.f : pdf;
}
CP_DEV float bare_by_point_pdf(const DevLight& l, V3 point, const ShapePoint& sp) {
    if (l.shape_kind != 1) return bare_default_by_point_pdf(l, point, sp);
    float distance_2 = length_sq(point);   // sphere.rs:96-105
    if (distance_2 < l.radius * l.radius) return bare_default_by_point_pdf(l, point, sp);
    float sin_theta_max_2 = l.radius * l.radius / distance_2;
    float cos_theta_max = sqrtf(1.f - sin_theta_max_2);
    return 1.f / (2.f * kPi * (1.f - cos_theta_max));
}
// TransformShape::{sample_by_point, by_point_pdf} (transform.rs:60-70)
CP_DEV void shape_sample_by_point(const DevScene& sc, const DevLight& l, V3 point, SamplerState& st, ShapePoint* sp, float* pdf) {
    int32_t chain[8]; int n = 0;
    for (int32_t x = l.shape_xform; x >= 0 && n < 8; x = sc.transforms[x].outer) chain[n++] = x;
    for (int i = n - 1; i >= 0; --i) point = apply_point(sc.transforms[chain[i]].m_inv, point);
    bare_sample_by_point(sc, l, point, st, sp, pdf);
    for (int i = 0; i < n; ++i) *sp = shape_point_apply(*sp, sc.transforms[chain[i]].m, sc.transforms[chain[i]].m_inv);
}
CP_DEV float shape_by_point_pdf(const DevScene& sc, const DevLight& l, V3 point, ShapePoint sp) {
    int32_t chain[8]; int n = 0;
    for (int32_t x = l.shape_xform; x >= 0 && n < 8; x = sc.transforms[x].outer) chain[n++] = x;
    for (int i = n - 1; i >= 0; --i) {
        const DevTransform& t = sc.transforms[chain[i]];
        point = apply_point(t.m_inv, point);
        sp = shape_point_apply(sp, t.m_inv, t.m);
    }
    return bare_by_point_pdf(l, point, sp);
}

// ---------------------------------------------------------------------------------- lights
struct LightSample { V3 wi; bool has_li; Spec li; float pdf; Ray vis; };

CP_DEV Ray visibility_new(const ShapePoint& from, const ShapePoint& to) {   // light/mod.rs:103-107, ray.rs:30-36
    V3 f = point_offset_by_error(from, to.p - from.p);
    V3 t = point_offset_by_error(to, f - to.p);
    Ray r; r.o = f; r.d = t - f; r.t_max = 1.f - 0.00001f;
    return r;
}
CP_DEV bool light_is_delta(const DevLight& l) { return l.type == 1 || l.type == 2; }

// ---- InfiniteAreaLight (src/light/infinite_area.rs) -------------------------------------------
CP_DEV P2 spherical_to_normalize_phi_theta(V3 w) {   // math/mod.rs:135-141
    const float p = atan2f(w.y, w.x);
    P2 r;
    r.x = (p < 0.f ? p + 2.f * kPi : p) / 2.f * kInvPi;
    r.y = acosf(rclamp(w.z, -1.f, 1.f)) * kInvPi;
    return r;
}
CP_DEV V3 normalize_phi_theta_to_spherical(P2 pt) {   // math/mod.rs:143-151
    const float theta = pt.y * kPi, phi = pt.x * kPi * 2.f;
    const float ct = cosf(theta), st = sinf(theta), cp = cosf(phi), sp = sinf(phi);
    return v3(st * cp, st * sp, ct);
}
// Distribution1D::sample_remap (distribution.rs:48-56): first i with u <= cdf[i].  The reference
// scans linearly; the cdf is non-decreasing, so a binary search lands on the same entry (and on
// the same "nothing matched" fallback when the table is all NaN).
CP_DEV void dist_sample_remap(const float* __restrict__ pdf, const float* __restrict__ cdf, uint32_t len, float u, uint32_t* idx, float* p, float* remap) {
    uint32_t lo = 0, hi = len;
    while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (u <= __ldg(&cdf[mid])) hi = mid; else lo = mid + 1; }
    if (lo < len) { const float c = __ldg(&cdf[lo]), q = __ldg(&pdf[lo]); *idx = lo; *p = q; *remap = (c - u) / q; }
    else { *idx = len - 1; *p = __ldg(&pdf[len - 1]); *remap = 1.f; }
}
CP_DEV Spec infinite_le(const DevScene& sc, const DevLight& l, V3 d, P2* uv_out) {   // le_out_scene (:35-39)
    P2 uv = spherical_to_normalize_phi_theta(normalize(d));
    uv.y = 1.f - uv.y;
    if (uv_out) *uv_out = uv;
    return tex_eval(sc, l.tex, uv);
}
CP_DEV float infinite_out_scene_pdf(const DevScene& sc, const DevLight& l, P2 uv) {   // out_scene_pdf (:40-51) + Distribution2D::continuous_pdf (distribution.rs:119-126)
    const float nx = (float)l.env_h, ny = (float)l.env_w;   // n = (row distribution, first row)
    const uint32_t r = f32_to_u32_sat(rmin(uv.x * nx, nx - 1.f)), c = f32_to_u32_sat(rmin(uv.y * ny, ny - 1.f));
    const float* t = sc.env + l.env_off;
    const float pdf = (__ldg(&t[r]) * __ldg(&t[2ull * l.env_h + (uint64_t)r * l.env_w + c])) * nx * ny;
    const float st = sinf(uv.y * kPi);
    return st != 0.f ? pdf / (2.f * kPi * kPi * st) : 0.f;
}
// le_out_scene of every light summed into `L` with weight beta (path.rs:26-31, direct_light.rs:36-38):
// only infinite lights answer with Some(..), the others add zero
CP_DEV Spec add_out_scene(const DevScene& sc, V3 d, Spec L, Spec beta, bool weighted) {
    for (uint32_t k = 0; k < sc.n_lights; ++k) {
        const DevLight& l = sc.lights[k];
        const Spec le = l.type == 3 ? infinite_le(sc, l, d, nullptr) : spec(0.f);
        L = weighted ? L + beta * le : L + le;
    }
    return L;
}

CP_DEV LightSample light_sample_li(const DevScene& sc, const DevLight& l, const ShapePoint& point, SamplerState& st) {
    LightSample r; r.has_li = false; r.pdf = 0.f; r.wi = v3(0.f, 0.f, 0.f); r.li = spec(0.f);
    if (l.type == 0) {   // AreaLight::sample_li (area.rs:30-51)
        ShapePoint lp; float pdf;
        shape_sample_by_point(sc, l, point.p, st, &lp, &pdf);
        r.vis = visibility_new(point, lp);
        if (length_sq(lp.p - point.p) == 0.f || pdf == 0.f) return r;
        r.wi = normalize(lp.p - point.p); r.has_li = true; r.li = spec3(l.L[0], l.L[1], l.L[2]); r.pdf = pdf;
    } else if (l.type == 1) {   // DeltaLight adaptor (light/mod.rs:90-98) + PointLight (point.rs:33-38)
        sampler_get_2d(sc.sampler, st);
        V3 lp = v3(l.v[0], l.v[1], l.v[2]);
        V3 wi = lp - point.p;
        r.li = spec3(l.L[0], l.L[1], l.L[2]) / length_sq(wi);
        ShapePoint to; to.p = lp; to.n = normalize(-wi); to.uv.x = 0.f; to.uv.y = 0.f; to.p_error = v3(0.f, 0.f, 0.f);
        r.wi = normalize(wi); r.has_li = true; r.pdf = 1.f; r.vis = visibility_new(point, to);
    } else if (l.type == 3) {   // InfiniteAreaLight::sample_li (infinite_area.rs:52-72)
        const P2 u = sampler_get_2d(sc.sampler, st);
        const float* t = sc.env + l.env_off;
        const uint32_t w = l.env_w, h = l.env_h;
        uint32_t row, col; float rp, cp, xr, yr;
        dist_sample_remap(t, t + h, h, u.x, &row, &rp, &xr);                         // sample_continuous (distribution.rs:57-60)
        const float row_pdf = rp * (float)h, x = ((float)row + xr) / (float)h;
        dist_sample_remap(t + 2ull * h + (uint64_t)row * w, t + 2ull * h + (uint64_t)h * w + (uint64_t)row * w, w, u.y, &col, &cp, &yr);
        const float col_pdf = cp * (float)w, y = ((float)col + yr) / (float)w;
        float pdf = row_pdf * col_pdf;
        P2 uv; uv.x = x; uv.y = 1.f - y;
        const V3 wi = normalize_phi_theta_to_spherical(uv);
        const float sth = sinf(uv.y * kPi);
        pdf = sth != 0.f ? pdf / (2.f * kPi * kPi * sth) : 0.f;
        r.wi = wi; r.has_li = true; r.li = tex_eval(sc, l.tex, uv); r.pdf = pdf;
        r.vis.o = point_offset_by_error(point, wi); r.vis.d = wi; r.vis.t_max = kFloatMax;
    } else {   // DistantLight (distant.rs:28-34), VisibilityTester::new_od
        sampler_get_2d(sc.sampler, st);
        V3 w = v3(l.v[0], l.v[1], l.v[2]);
        r.wi = -w; r.has_li = true; r.li = spec3(l.L[0], l.L[1], l.L[2]); r.pdf = 1.f;
        r.vis.o = point_offset_by_error(point, -w); r.vis.d = -w; r.vis.t_max = kFloatMax;
    }
    return r;
}

}  // namespace cpbrt
