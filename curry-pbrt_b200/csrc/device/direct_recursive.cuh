// DirectLightIntegrator::li_depth with its recursion over delta lobes
// (src/integrator/direct_light.rs:12-41):
//
//     l = uniform_sample_one_light(...)
//     if depth + 1 < max_depth:
//         for (wi, s) in bsdf.sample_all_delta_f(wo):          # bxdf/mod.rs:127-135
//             l += li_depth(ray(sp, wi), depth + 1) * s * |n . wi|
//
// A glass surface has two delta lobes, so this is a TREE of rays per camera sample, and the
// Halton dimensions are handed out in depth-first order: what a sibling draws depends on how
// much the subtree before it consumed.  That does not fit queues of independent rays, so this
// one configuration (directlighting, maxdepth >= 2, a material with delta lobes) is rendered by
// one thread per camera sample walking its tree with an explicit stack, tracing with the
// reference-order traversals of bvh_trace.cuh.  Everything else uses the wavefront stages.
#pragma once
#include "bsdf_general.cuh"
#include "bvh_trace.cuh"

namespace cpbrt {

static constexpr int kMaxDirectDepth = 12;

struct RayCounts { uint32_t closest, any; };

// uniform_sample_one_light (integrator/mod.rs:13-97), evaluated on the spot
CP_DEV Spec uniform_sample_one_light_inline(const DevScene& sc, const ShapePoint& sp, const GBSDF& bsdf, V3 wo, SamplerState& st,
                                            RayCounts* rc, uint32_t* overflow) {
    Spec l = spec(0.f);
    if (sc.n_lights != 0) {
        const uint32_t li = sampler_get_usize(sc.sampler, st, sc.n_lights);
        const DevLight& light = sc.lights[li];
        const V3 n = sp.n;
        const LightSample ls = light_sample_li(sc, light, sp, st);
        if (ls.has_li && ls.pdf != 0.f) {
            Spec f; float f_pdf;
            if (bsdf.no_delta_f_pdf(wo, ls.wi, &f, &f_pdf) && f_pdf != 0.f) {
                ++rc->any;
                if (!bvh_any<false>(sc, ls.vis, nullptr, overflow)) {
                    Spec ld;
                    if (light_is_delta(light)) ld = ls.li * f * fabsf(dot(n, ls.wi)) / ls.pdf;
                    else ld = ls.li * f * fabsf(dot(n, ls.wi)) * power_heuristic(ls.pdf, f_pdf) / ls.pdf;
                    l = l + ld;
                }
            }
        }
        if (!light_is_delta(light)) {
            const P2 u = sampler_get_2d(sc.sampler, st);
            V3 wi; Spec f; float f_pdf;
            if (bsdf.sample_no_delta_f(wo, u, &wi, &f, &f_pdf) && f_pdf != 0.f) {
                Ray ray; ray.o = point_offset_by_error(sp, wi); ray.d = wi; ray.t_max = kFloatMax;
                Spec ld = spec(0.f);
                ++rc->closest;
                float t;
                const uint32_t slot = bvh_closest<false>(sc, ray, &t, nullptr, overflow);
                if (slot != 0xFFFFFFFFu) {
                    const PrimRecord pr = load_prim(sc, slot);
                    if (prim_source(pr) == (kSourceLight | li)) {
                        const ShapePoint hsp = prim_shape_point(sc, slot, pr, ray, t);
                        const float li_pdf = shape_by_point_pdf(sc, light, sp.p, hsp);
                        if (li_pdf != 0.f) ld = spec3(light.L[0], light.L[1], light.L[2]) * f * fabsf(dot(n, wi)) * power_heuristic(f_pdf, li_pdf) / f_pdf;
                    }
                } else if (light.type == 3) {
                    P2 uv; const Spec le = infinite_le(sc, light, wi, &uv);
                    const float le_pdf = infinite_out_scene_pdf(sc, light, uv);
                    if (le_pdf != 0.f) ld = le * f * fabsf(dot(n, wi)) * power_heuristic(f_pdf, le_pdf) / f_pdf;
                }
                l = l + ld;
            }
        }
    }
    return l * (float)sc.n_lights;
}

struct DirectFrame {
    Spec l;                         // this call's `l` so far
    int next, n;                    // delta-lobe children still to visit
    float4 child_o[kMaxLobes];      // {ray origin, |n . wi|}
    float4 child_d[kMaxLobes];      // {ray direction, -}
    Spec child_s[kMaxLobes];        // the lobe's value
};

// One li_depth call up to (not including) its recursive calls: fills `fr`.
CP_DEV void direct_visit(const DevScene& sc, const Ray& ray, uint32_t depth, SamplerState& st, DirectFrame& fr, RayCounts* rc,
                         uint32_t* overflow, uint32_t* bsdf_overflow) {
    fr.l = spec(0.f); fr.next = 0; fr.n = 0;
    ++rc->closest;
    float t;
    const uint32_t slot = bvh_closest<false>(sc, ray, &t, nullptr, overflow);
    if (slot == 0xFFFFFFFFu) { fr.l = add_out_scene(sc, ray.d, fr.l, spec(1.f), false); return; }
    const PrimRecord pr = load_prim(sc, slot);
    const uint32_t src = prim_source(pr);
    if (src & kSourceLight) { const DevLight& l = sc.lights[src & ~kSourceLight]; fr.l = fr.l + spec3(l.L[0], l.L[1], l.L[2]); return; }
    Ray full = ray; full.t_max = kFloatMax;
    const ShapePoint sp = prim_shape_point(sc, slot, pr, full, t);
    GBSDF bsdf;
    build_general_bsdf(sc, src, sp, bsdf);
    if (bsdf.overflow) *bsdf_overflow = 1;
    const V3 wo = -normalize(ray.d);
    bsdf.set_wo(wo);
    fr.l = fr.l + uniform_sample_one_light_inline(sc, sp, bsdf, wo, st, rc, overflow);
    if (depth + 1 < sc.max_depth) {
        const V3 wo_l = bsdf.wo_local;
        for (int i = 0; i < bsdf.n_deltas; ++i) {      // sample_all_delta_f
            V3 wi_l; Spec s;
            if (!bsdf.deltas[i].sample_f(wo_l, &wi_l, &s)) continue;
            const V3 wi = bsdf.local_to_world(wi_l);
            fr.child_o[fr.n] = make_float4(0.f, 0.f, 0.f, fabsf(dot(sp.n, wi)));
            const V3 o = point_offset_by_error(sp, wi);
            fr.child_o[fr.n].x = o.x; fr.child_o[fr.n].y = o.y; fr.child_o[fr.n].z = o.z;
            fr.child_d[fr.n] = make_float4(wi.x, wi.y, wi.z, 0.f);
            fr.child_s[fr.n] = s;
            ++fr.n;
        }
    }
}

}  // namespace cpbrt
