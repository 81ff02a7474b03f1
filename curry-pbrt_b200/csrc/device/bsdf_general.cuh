// General BSDF for the materials that build a variable lobe list: plastic, uber, translucent and
// mix (src/material/{plastic,uber,translucent,mix}.rs), with the Trowbridge-Reitz microfacet
// lobes (src/material/bxdf/microfacet.rs:16-267), Lambertian transmission
// (lambertian.rs:20-36) and the ScaledBxDF / ScaledDeltaBxDF wrappers that BSDF::mix adds
// (bxdf/mod.rs:49-69, 218-269, 294-300).
//
// Matte, mirror and glass keep their fixed-shape BSDF<NL, ND> (shading.cuh) and their own
// shading kernels; everything else goes through GBSDF and k_shade<SHADE_GENERAL>.  The interface
// (init, no_delta_f_pdf, sample_no_delta_f, sample_f, n_lobes) is the one k_shade expects.
// Lobes are addressed by a run-time index, so a GBSDF lives in local memory: this is the
// coverage path, not the tuned one.
#pragma once
#include "shading.cuh"

namespace cpbrt {

static constexpr int kMaxLobes = 8;      // uber: 1 + 1, translucent: 4, doubled once by a mix of two such
static constexpr int kMaxScales = 4;     // nesting depth of mix materials

// more trig helpers of src/math/mod.rs:152-200
CP_DEV float cos_2_theta(V3 w) { return w.z * w.z; }
CP_DEV float tan_2_theta(V3 w) { return sin_2_theta(w) / cos_2_theta(w); }
CP_DEV float cos_phi(V3 w) { const float s = sin_theta(w); return s == 0.f ? 1.f : rclamp(w.x / s, -1.f, 1.f); }
CP_DEV float sin_phi(V3 w) { const float s = sin_theta(w); return s == 0.f ? 0.f : rclamp(w.y / s, -1.f, 1.f); }
CP_DEV float cos_2_phi(V3 w) { const float c = cos_phi(w); return c * c; }
CP_DEV float sin_2_phi(V3 w) { const float s = sin_phi(w); return s * s; }

// TrowbridgeReitzDistribution (microfacet.rs:16-121)
struct TrowbridgeReitz {
    float ax, ay;
    CP_DEV float d(V3 wh) const {
        const float t2 = tan_2_theta(wh);
        if (t2 != t2 || isinf(t2)) return 0.f;
        const float c2 = cos_2_theta(wh), c4 = c2 * c2;
        const float e = (cos_2_phi(wh) / (ax * ax) + sin_2_phi(wh) / (ay * ay)) * t2;
        return 1.f / (kPi * ax * ay * c4 * (1.f + e) * (1.f + e));
    }
    CP_DEV float lambda(V3 w) const {
        const float at = fabsf(tan_theta(w));
        if (at != at || isinf(at)) return 0.f;
        const float alpha = sqrtf(cos_2_phi(w) * ax * ax + sin_2_phi(w) * ay * ay);
        const float a = alpha * at;
        return (-1.f + sqrtf(1.f + a * a)) / 2.f;
    }
    CP_DEV float g(V3 wo, V3 wi) const { return 1.f / (1.f + lambda(wo) + lambda(wi)); }
    CP_DEV float g1(V3 w) const { return 1.f / (1.f + lambda(w)); }
    // sample_wh (microfacet.rs:40-95): visible-normal sampling in the stretched frame
    CP_DEV V3 sample_wh(V3 wo, P2 u, float* pdf) const {
        const bool flip = wo.z < 0.f;
        const V3 wi = flip ? -wo : wo;
        const V3 ws = v3(ax * wi.x, ay * wi.y, wi.z);
        const float cti = cos_theta(ws);
        float slope_x, slope_y;
        if (cti > 0.9999f) {
            const float r = sqrtf(u.x / (1.f - u.x));
            const float phi = 6.28318530718f * u.y;
            slope_x = r * cosf(phi); slope_y = r * sinf(phi);
        } else {
            const float st = sqrtf(rmax(0.f, 1.f - cti * cti));
            const float tt = st / cti;
            const float a = 1.f / tt;
            const float G1 = 2.f / (1.f + sqrtf(1.f + 1.f / (a * a)));
            const float A = 2.f * u.x / G1 - 1.f;
            float tmp = 1.f / (A * A - 1.f);
            if (tmp > 1e10f) tmp = 1e10f;
            const float B = tt;
            const float D = sqrtf(rmax(B * B * tmp * tmp - (A * A - B * B) * tmp, 0.f));
            const float sx1 = B * tmp - D, sx2 = B * tmp + D;
            slope_x = (A < 0.f || sx2 > 1.f / tt) ? sx1 : sx2;
            float S, u2;
            if (u.y > 0.5f) { S = 1.f; u2 = 2.f * (u.y - 0.5f); } else { S = -1.f; u2 = 2.f * (0.5f - u.y); }
            const float z = (u2 * (u2 * (u2 * 0.27385f - 0.73369f) + 0.46341f)) / (u2 * (u2 * (u2 * 0.093073f + 0.309420f) - 1.f) + 0.597999f);
            slope_y = S * z * sqrtf(1.f + slope_x * slope_x);
        }
        const float rot = cos_phi(ws) * slope_x - sin_phi(ws) * slope_y;
        slope_y = sin_phi(ws) * slope_x + cos_phi(ws) * slope_y;
        slope_x = rot;
        slope_x *= ax; slope_y *= ay;
        V3 wh = normalize(v3(-slope_x, -slope_y, 1.f));
        if (flip) wh = -wh;
        *pdf = d(wh) * g1(wo) * fabsf(dot(wo, wh)) / fabsf(cos_theta(wo));
        return wh;
    }
};
// roughness_to_alpha (microfacet.rs:28-33)
CP_DEV float roughness_to_alpha(float rough) {
    rough = rmax(rough, 1e-3f);
    const float x = logf(rough), x2 = x * x;
    return 1.62142f + 0.819955f * x + 0.1734f * x2 + 0.0171201f * x * x2 + 0.000640711f * x2 * x2;
}

enum GLobeKind : int { G_LAMBERT_R = 0, G_LAMBERT_T = 1, G_OREN_NAYAR = 2, G_MICRO_R = 3, G_MICRO_T = 4 };

// a non-delta BxDF, optionally inside ScaledBxDF layers
struct GLobe {
    int kind;
    Spec r;
    float a, b;                      // Oren-Nayar
    TrowbridgeReitz dist;            // microfacet
    float fr_eta_i, fr_eta_t;        // FresnelDielectric of a microfacet lobe
    float eta_a, eta_b;              // microfacet transmission
    Spec scales[kMaxScales]; int n_scales;

    CP_DEV bool transmit() const { return kind == G_LAMBERT_T || kind == G_MICRO_T; }

    CP_DEV bool f_inner(V3 wo, V3 wi, Spec* out) const {
        if (kind == G_LAMBERT_R || kind == G_LAMBERT_T) { *out = r * kInvPi; return true; }   // lambertian.rs:14-36
        if (kind == G_OREN_NAYAR) {   // oren_nayar.rs:18-30
            float sin_alpha, tan_beta;
            if (cos_theta(wi) < cos_theta(wo)) { sin_alpha = sin_theta(wi); tan_beta = tan_theta(wo); }
            else { sin_alpha = sin_theta(wo); tan_beta = tan_theta(wi); }
            *out = r * ((a + b * rmax(0.f, cos_delta_phi(wi, wo)) * sin_alpha * tan_beta) / kPi);
            return true;
        }
        if (kind == G_MICRO_R) {   // MicrofacetReflection::f (microfacet.rs:143-165)
            const float co = fabsf(cos_theta(wo)), ci = fabsf(cos_theta(wi));
            const V3 wi_n = normalize(wi), wo_n = normalize(wo);
            V3 wh = wi_n + wo_n;
            if ((wh.x == 0.f && wh.y == 0.f && wh.z == 0.f) || co == 0.f || ci == 0.f) return false;
            wh = normalize(wh);
            const float fr = fresnel_dielectric(fr_eta_i, fr_eta_t, dot(wi_n, wh.z < 0.f ? -wh : wh));
            *out = r * dist.d(wh) * dist.g(wo_n, wi_n) * fr / (4.f * co * ci);
            return true;
        }
        // MicrofacetTransmission::f (microfacet.rs:206-242)
        if (wo.z * wi.z > 0.f) return false;
        const float co = cos_theta(wo), ci = cos_theta(wi);
        if (ci == 0.f || co == 0.f) return false;
        const float eta = cos_theta(wo) > 0.f ? eta_b / eta_a : eta_a / eta_b;
        V3 wh = normalize(wo + wi * eta);
        if (wh.z < 0.f) wh = -wh;
        const float sqrt_denom = dot(wo, wh) + eta * dot(wi, wh);
        const float fr = fresnel_dielectric(fr_eta_i, fr_eta_t, dot(wo, wh));
        const float factor = 1.f / eta;
        const float v = fabsf(dist.d(wh) * dist.g(wo, wi) * eta * eta * fabsf(dot(wi, wh)) * fabsf(dot(wo, wh)) * factor * factor /
                              (ci * co * sqrt_denom * sqrt_denom));
        *out = (spec(1.f) - spec(fr)) * r * v;
        return true;
    }
    CP_DEV Spec apply_scales(Spec v) const {
        for (int i = 0; i < n_scales; ++i) v = v * scales[i];
        return v;
    }
    CP_DEV bool f(V3 wo, V3 wi, Spec* out) const {
        if (!f_inner(wo, wi, out)) return false;
        *out = apply_scales(*out);
        return true;
    }
    // BxDF::sample_f default (bxdf/mod.rs:20-37) or the microfacet overrides (microfacet.rs:166-179, 246-266)
    CP_DEV bool sample_f(V3 wo, P2 u, V3* wi, Spec* fo, float* pdf_out) const {
        bool has_f;
        *wi = v3(0.f, 0.f, 0.f); *pdf_out = 0.f;
        if (kind == G_MICRO_R) {
            if (wo.z == 0.f) return false;
            float pdf; const V3 wh = dist.sample_wh(wo, u, &pdf);
            if (dot(wo, wh) < 0.f) return false;
            const V3 w = -wo + 2.f * dot(wo, wh) * wh;
            if (w.z * wo.z < 0.f) return false;
            *wi = w;
            has_f = f_inner(wo, w, fo);
            *pdf_out = pdf / (4.f * dot(wo, wh));
        } else if (kind == G_MICRO_T) {
            if (wo.z == 0.f) return false;
            float pdf; const V3 wh = dist.sample_wh(wo, u, &pdf);
            if (dot(wo, wh) < 0.f) return false;
            float eta_i, eta_o;
            if (cos_theta(wo) > 0.f) { eta_i = eta_a / eta_b; eta_o = eta_b / eta_a; } else { eta_i = eta_b / eta_a; eta_o = eta_a / eta_b; }
            V3 w;
            if (!refract(wo, normalize(wh), eta_i, &w)) return false;
            const float sqrt_denom = dot(wo, wh) + eta_o * dot(w, wh);
            const float dwh_dwi = fabsf(eta_o * eta_o * dot(w, wh)) / (sqrt_denom * sqrt_denom);
            *wi = w;
            has_f = f_inner(wo, w, fo);
            *pdf_out = pdf * dwh_dwi;
        } else {
            float pdf; V3 w = cosine_sample_hemisphere(u, &pdf);
            if (!transmit()) { if (wo.z < 0.f) w.z *= -1.f; } else { if (wo.z > 0.f) w.z *= -1.f; }
            *wi = w;
            has_f = f_inner(wo, w, fo);
            *pdf_out = pdf;
        }
        if (has_f) *fo = apply_scales(*fo);
        return has_f;
    }
};

struct GDelta {   // specular.rs:45-102 inside ScaledDeltaBxDF layers
    DeltaLobe lobe;
    Spec scales[kMaxScales]; int n_scales;
    CP_DEV bool sample_f(V3 wo, V3* wi, Spec* s) const {
        if (!delta_sample_f(lobe, wo, wi, s)) return false;
        for (int i = 0; i < n_scales; ++i) *s = *s * scales[i];
        return true;
    }
};

struct GBSDF {   // BSDF (bxdf/mod.rs:71-270)
    static constexpr bool kNoDelta = false;
    V3 n, sn, snx, sny;
    GLobe lobes[kMaxLobes]; int n_lobes;
    GDelta deltas[kMaxLobes]; int n_deltas;
    bool overflow;

    CP_DEV void init(V3 normal) {
        sn = normal;
        coordinate_system(sn, &snx, &sny);
        n = normalize(v3(dot(normal, snx), dot(normal, sny), dot(normal, sn)));
        n_lobes = 0; n_deltas = 0; overflow = false;
    }
    CP_DEV void init_frame(const float4* __restrict__ f) { load_frame(f); n_lobes = 0; n_deltas = 0; overflow = false; }
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

    CP_DEV GLobe* add_lobe(int kind, Spec r) {
        if (n_lobes == kMaxLobes) { overflow = true; return &lobes[kMaxLobes - 1]; }
        GLobe* l = &lobes[n_lobes++];
        l->kind = kind; l->r = r; l->a = 0.f; l->b = 0.f; l->dist.ax = 1.f; l->dist.ay = 1.f;
        l->fr_eta_i = 1.f; l->fr_eta_t = 1.f; l->eta_a = 1.f; l->eta_b = 1.f; l->n_scales = 0;
        return l;
    }
    CP_DEV void add_microfacet_r(Spec ks, float ru, float rv, float eta) {
        GLobe* l = add_lobe(G_MICRO_R, ks);
        l->dist.ax = ru; l->dist.ay = rv; l->fr_eta_i = 1.f; l->fr_eta_t = eta;
    }
    CP_DEV void add_delta(int kind, Spec r, float ea, float eb) {
        if (n_deltas == kMaxLobes) { overflow = true; return; }
        GDelta* d = &deltas[n_deltas++];
        d->lobe.kind = kind; d->lobe.r = r; d->lobe.eta_a = ea; d->lobe.eta_b = eb; d->n_scales = 0;
    }

    // sample_no_delta_f + choose_no_delta_f (bxdf/mod.rs:136-159)
    CP_DEV bool sample_no_delta_f(V3 wo, P2 u, V3* wi, Spec* f, float* pdf) const {
        if (n_lobes == 0) { *wi = v3(0.f, 0.f, 0.f); *pdf = 0.f; return false; }
        const float fl = u.x * (float)n_lobes;                // sample_usize_remap (math/mod.rs:81-86)
        const float tr = truncf(fl), remap = fl - tr;
        const uint32_t index = min(f32_to_u32_sat(tr), (uint32_t)n_lobes - 1);
        P2 u2; u2.x = remap; u2.y = u.y;
        const V3 wo_l = wo_local; V3 wi_l; float p;
        const bool has = lobes[index].sample_f(wo_l, u2, &wi_l, f, &p);
        *wi = local_to_world(wi_l);
        *pdf = p / (float)n_lobes;
        return has;
    }
    // sample_delta_f (bxdf/mod.rs:160-175)
    CP_DEV bool sample_delta_f(V3 wo, float u, V3* wi, Spec* f, float* pdf) const {
        *wi = v3(0.f, 0.f, 0.f); *pdf = 0.f;
        if (n_deltas == 0) return false;
        const V3 wo_l = wo_local;
        V3 w[kMaxLobes]; Spec s[kMaxLobes]; float lm[kMaxLobes]; int cnt = 0;
        for (int i = 0; i < n_deltas; ++i) { if (deltas[i].sample_f(wo_l, &w[cnt], &s[cnt])) { lm[cnt] = lum(s[cnt]); ++cnt; } }
        if (cnt == 0) return false;   // the reference panics on an empty table
        int i; float p, remap;
        distribution_sample_remap<kMaxLobes>(lm, cnt, u, &i, &p, &remap);
        *wi = local_to_world(w[i]); *f = s[i]; *pdf = p;
        return true;
    }
    // no_delta_f_pdf (bxdf/mod.rs:176-198): reflective lobes when wo, wi lie on the same side of n,
    // transmissive ones otherwise; pdf = mean of the default pdfs over that subset
    CP_DEV bool no_delta_f_pdf(V3 wo_w, V3 wi_w, Spec* f, float* pdf) const {
        if (n_lobes == 0) { *pdf = 0.f; return false; }
        const V3 wo = wo_local, wi = world_to_local(wi_w);
        const bool reflect = dot(wo, n) * dot(wi, n) > 0.f;
        bool has = false; Spec acc = spec(0.f); float p = 0.f; int len = 0;
        for (int i = 0; i < n_lobes; ++i) {
            if (lobes[i].transmit() == reflect) continue;
            ++len;
            Spec tf;
            if (lobes[i].f(wo, wi, &tf)) { acc = (has ? acc : spec(0.f)) + tf; has = true; p += wi.z * kInvPi; }   // BxDF::pdf default (:38-40)
        }
        *f = acc; *pdf = p / (float)len;
        return has;
    }
    // sample_f (bxdf/mod.rs:199-214)
    CP_DEV bool sample_f(const DevSampler& S, SamplerState& st, V3 wo, V3* wi, Spec* f, float* pdf, bool* is_delta) const {
        float w[2]; w[0] = (float)n_lobes; w[1] = (float)n_deltas;
        int i; float p, remap;
        distribution_sample_remap<2>(w, 2, sampler_get_1d(S, st), &i, &p, &remap);
        bool has; float ip;
        if (i == 0) { P2 u; u.x = remap; u.y = sampler_get_1d(S, st); has = sample_no_delta_f(wo, u, wi, f, &ip); *is_delta = false; }
        else { has = sample_delta_f(wo, remap, wi, f, &ip); *is_delta = true; }
        *pdf = p * ip;
        return has;
    }
};

// One leaf material (anything but mix) appended to `b`.
CP_DEV void append_material(const DevScene& sc, const DevMaterial& m, P2 uv, GBSDF& b) {
    auto S = [&](int slot) { return tex_eval(sc, m.tex[slot], uv); };
    auto F = [&](int slot) { return tex_eval(sc, m.tex[slot], uv).r; };
    switch (m.type) {
        case 0: {   // matte.rs:17-28
            const Spec kd = S(0);
            const float sigma = rclamp(F(1), 0.f, 90.f);
            if (sigma == 0.f) b.add_lobe(G_LAMBERT_R, kd);
            else {
                GLobe* l = b.add_lobe(G_OREN_NAYAR, kd);
                const float s = sigma * (kPi / 180.f), s2 = s * s;
                l->a = 1.f - (s2 / (2.f * (s + 0.33f)));
                l->b = 0.45f * s2 / (s2 + 0.09f);
            }
            break;
        }
        case 1: b.add_delta(DELTA_SPEC_R_NOOP, S(0), 1.f, 1.f); break;   // mirror.rs:15-23
        case 2: {   // glass.rs:17-29
            const Spec r = S(0), t = S(1);
            const float eta = F(2);
            b.add_delta(DELTA_SPEC_R_DIELECTRIC, r, 1.f, eta);
            b.add_delta(DELTA_SPEC_T, t, 1.f, eta);
            break;
        }
        case 3: {   // plastic.rs:10-31
            const Spec kd = S(0);
            if (!is_black(kd)) b.add_lobe(G_LAMBERT_R, kd);
            const Spec ks = S(1);
            if (!is_black(ks)) { const float rough = roughness_to_alpha(F(2)); b.add_microfacet_r(ks, rough, rough, 1.5f); }
            break;
        }
        case 4: {   // uber.rs:16-62
            const float eta = F(7);
            const Spec opacity = S(8);
            const Spec t = spec(1.f) - opacity;
            if (!is_black(t)) b.add_delta(DELTA_SPEC_T, t, 1.f, 1.f);
            const Spec kd = opacity * S(0);
            if (!is_black(kd)) b.add_lobe(G_LAMBERT_R, kd);
            const Spec ks = opacity * S(1);
            if (!is_black(ks)) {
                float ru = m.tex[5] >= 0 ? F(5) : F(4);
                float rv = m.tex[6] >= 0 ? F(6) : ru;
                ru = roughness_to_alpha(ru); rv = roughness_to_alpha(rv);
                b.add_microfacet_r(ks, ru, rv, eta);
            }
            const Spec kr = opacity * S(2);
            if (!is_black(kr)) b.add_delta(DELTA_SPEC_R_DIELECTRIC, kr, 1.f, eta);
            const Spec kt = opacity * S(3);
            if (!is_black(kt)) b.add_delta(DELTA_SPEC_T, kt, 1.f, eta);
            break;
        }
        case 5: {   // translucent.rs:13-57
            const float eta = 1.5f;
            const Spec r = S(3), t = S(4);
            if (is_black(r) && is_black(t)) break;
            const Spec kd = S(0);
            if (!is_black(kd)) {
                if (!is_black(r)) b.add_lobe(G_LAMBERT_R, r * kd);
                if (!is_black(t)) b.add_lobe(G_LAMBERT_T, t * kd);
            }
            const Spec ks = S(1);
            if (!is_black(ks)) {
                const float rough = roughness_to_alpha(F(2));
                if (!is_black(r)) b.add_microfacet_r(r * ks, rough, rough, eta);
                if (!is_black(t)) {
                    GLobe* l = b.add_lobe(G_MICRO_T, t * ks);
                    l->dist.ax = rough; l->dist.ay = rough; l->eta_a = 1.f; l->eta_b = eta; l->fr_eta_i = 1.f; l->fr_eta_t = eta;
                }
            }
            break;
        }
        default: break;
    }
}

// Material::compute_scattering_functions for any material, mix trees included (mix.rs:11-16,
// BSDF::mix bxdf/mod.rs:218-269).  mix(A, B, s): A's lobes get the scale s appended, B's lobes
// follow with 1 - s; a black s keeps only B, a black 1 - s only A.  The tree is walked with an
// explicit stack; lobes of A and B land in the same arrays, in the order BSDF::mix produces.
CP_DEV void build_general_bsdf(const DevScene& sc, uint32_t mat, const ShapePoint& sp, GBSDF& b, const float4* frame = nullptr) {
    if (frame) b.init_frame(frame); else b.init(sp.n);
    // frame: material, phase (0 = enter, 1 = left done, 2 = right done), lobe / delta counts at entry and after left
    struct Frame { uint32_t mat; int phase; int l0, d0, l1, d1; };
    Frame st[kMaxScales + 1];
    int depth = 0;
    st[0].mat = mat; st[0].phase = 0; st[0].l0 = 0; st[0].d0 = 0; st[0].l1 = 0; st[0].d1 = 0;
    while (depth >= 0) {
        Frame& fr = st[depth];
        const DevMaterial& m = sc.materials[fr.mat];
        if (m.type != 6) { append_material(sc, m, sp.uv, b); --depth; continue; }
        if (fr.phase == 0) {
            fr.phase = 1; fr.l0 = b.n_lobes; fr.d0 = b.n_deltas;
            if (depth == kMaxScales) { b.overflow = true; --depth; continue; }
            Frame& c = st[++depth]; c.mat = (uint32_t)m.sub[0]; c.phase = 0;
        } else if (fr.phase == 1) {
            fr.phase = 2; fr.l1 = b.n_lobes; fr.d1 = b.n_deltas;
            Frame& c = st[++depth]; c.mat = (uint32_t)m.sub[1]; c.phase = 0;
        } else {
            const Spec s1 = tex_eval(sc, m.tex[0], sp.uv), s2 = spec(1.f) - s1;
            if (is_black(s1)) {          // *self = rhs
                const int nl = b.n_lobes - fr.l1, nd = b.n_deltas - fr.d1;
                for (int i = 0; i < nl; ++i) b.lobes[fr.l0 + i] = b.lobes[fr.l1 + i];
                for (int i = 0; i < nd; ++i) b.deltas[fr.d0 + i] = b.deltas[fr.d1 + i];
                b.n_lobes = fr.l0 + nl; b.n_deltas = fr.d0 + nd;
            } else if (is_black(s2)) {   // rhs dropped
                b.n_lobes = fr.l1; b.n_deltas = fr.d1;
            } else {
                for (int i = fr.l0; i < b.n_lobes; ++i) { GLobe& l = b.lobes[i]; if (l.n_scales < kMaxScales) l.scales[l.n_scales++] = i < fr.l1 ? s1 : s2; else b.overflow = true; }
                for (int i = fr.d0; i < b.n_deltas; ++i) { GDelta& d = b.deltas[i]; if (d.n_scales < kMaxScales) d.scales[d.n_scales++] = i < fr.d1 ? s1 : s2; else b.overflow = true; }
            }
            --depth;
        }
    }
}

}  // namespace cpbrt
