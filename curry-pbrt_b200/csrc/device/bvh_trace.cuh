// BVH traversal on the device with the reference's visiting order, so that the chosen
// primitive is the one the recursive CPU code returns, ties included:
//   closest hit  BVHAggregate::intersect_through_bound (src/aggregate/bvh.rs:151-191)
//     - a Parent's own bound is tested on entry against the CURRENT t_max
//     - the child on the far side of the split axis is visited first (bvh.rs:173-186)
//     - leaves test every primitive, no box test, hits accepted at t <= t_max (later wins)
//   any hit      intersect_predicate_through_bound (bvh.rs:125-150): left, then right
//   box test     Bounds3f::intersect_predicate_cached (src/geometry/bounds.rs:306-323),
//                including its widened [t_min, t_max] interval (Q6)
// The recursion becomes an explicit stack.  A deferred (second) child's box test is split:
// everything that does not depend on t_max is decided when the parent is read, its t_min is
// kept in the stack entry and `t_min < t_max` is re-evaluated with the t_max current at pop
// time -- exactly what the recursive code sees when it reaches that child.
#pragma once
#include "shapes.cuh"

namespace cpbrt {

struct RayBoxSetup {   // RayIntersectCache (bounds.rs:326-368)
    V3 o, inv_d;
    bool pos[3];       // is_positive_d: d >= 0
};
CP_DEV RayBoxSetup ray_box_setup(const Ray& r) {
    RayBoxSetup s;
    s.o = r.o;
    s.inv_d = v3(1.f / r.d.x, 1.f / r.d.y, 1.f / r.d.z);
    s.pos[0] = r.d.x >= 0.f; s.pos[1] = r.d.y >= 0.f; s.pos[2] = r.d.z >= 0.f;
    return s;
}

// The t_max-independent part of intersect_predicate_cached; on success *t_min_out holds the
// value compared against ray.t_max at the end of the reference function.
CP_DEV bool box_test_static(const RayBoxSetup& s, float mnx, float mny, float mnz, float mxx, float mxy, float mxz, float* t_min_out) {
    const float widen = 1.f + 2.f * gamma_n(3);
    float t_min = ((s.pos[0] ? mnx : mxx) - s.o.x) * s.inv_d.x;
    float t_max = ((s.pos[0] ? mxx : mnx) - s.o.x) * s.inv_d.x * widen;
    {
        float lo = ((s.pos[1] ? mny : mxy) - s.o.y) * s.inv_d.y;
        float hi = ((s.pos[1] ? mxy : mny) - s.o.y) * s.inv_d.y * widen;
        if (t_min > hi || lo > t_max) return false;
        t_min = rmin(t_min, lo); t_max = rmax(t_max, hi);
    }
    {
        float lo = ((s.pos[2] ? mnz : mxz) - s.o.z) * s.inv_d.z;
        float hi = ((s.pos[2] ? mxz : mnz) - s.o.z) * s.inv_d.z * widen;
        if (t_min > hi || lo > t_max) return false;
        t_min = rmin(t_min, lo); t_max = rmax(t_max, hi);
    }
    *t_min_out = t_min;
    return t_max > 0.f;
}

static constexpr int kStackSize = 96;
static constexpr uint32_t kLeafBit = 0x80000000u;

struct TraceStats { uint32_t parent_visits, prim_tests; };

// Closest hit.  Returns the BVH slot (or 0xFFFFFFFF) and leaves the hit distance in *t_hit.
template <bool COUNT>
__device__ uint32_t bvh_closest(const DevScene& sc, Ray ray, float* t_hit, TraceStats* st, uint32_t* overflow) {
    uint32_t best = 0xFFFFFFFFu;
    if (sc.n_prims == 0) return best;
    const RayBoxSetup bs = ray_box_setup(ray);
    const TriRaySetup ts = tri_ray_setup(ray.d);
    uint2 stack[kStackSize];
    int sp = 0;

    auto leaf = [&](uint32_t first, uint32_t cnt) {
        for (uint32_t s = first; s < first + cnt; ++s) {
            if (COUNT) ++st->prim_tests;
            PrimRecord pr = load_prim(sc, s);
            float t;
            if (prim_intersect_t(sc, pr, ray, ts, &t)) { ray.t_max = t; best = s; }
        }
    };

    uint32_t cur;
    if (sc.root_cnt > 0) { leaf(sc.root_ref, sc.root_cnt); *t_hit = ray.t_max; return best; }
    {
        if (COUNT) ++st->parent_visits;
        float tmin;
        if (!(box_test_static(bs, sc.root_box[0], sc.root_box[1], sc.root_box[2], sc.root_box[3], sc.root_box[4], sc.root_box[5], &tmin) && tmin < ray.t_max)) {
            *t_hit = ray.t_max; return best;
        }
        cur = sc.root_ref;
    }
    for (;;) {
        const float4* np = sc.nodes + 4ull * cur;
        const float4 q0 = __ldg(np), q1 = __ldg(np + 1), q2 = __ldg(np + 2), q3 = __ldg(np + 3);
        const uint32_t lref = __float_as_uint(q3.x), rref = __float_as_uint(q3.y);
        const uint32_t lw = __float_as_uint(q3.z), rcnt = __float_as_uint(q3.w);
        const uint32_t lcnt = lw & 0x3FFFFFFFu, axis = lw >> 30;
        const bool positive = axis == 0 ? bs.pos[0] : (axis == 1 ? bs.pos[1] : bs.pos[2]);
        // reference: positive -> visit `right` first, then `left`; otherwise left first
        const bool first_is_right = positive;
        const uint32_t f_ref = first_is_right ? rref : lref, f_cnt = first_is_right ? rcnt : lcnt;
        const uint32_t s_ref = first_is_right ? lref : rref, s_cnt = first_is_right ? lcnt : rcnt;

        // deferred child
        if (s_cnt > 0) {
            if (sp < kStackSize) stack[sp++] = make_uint2(s_ref | kLeafBit, s_cnt); else *overflow = 1;
        } else {
            if (COUNT) ++st->parent_visits;
            float tmin; bool ok;
            if (first_is_right) ok = box_test_static(bs, q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, &tmin);
            else ok = box_test_static(bs, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, &tmin);
            if (ok) { if (sp < kStackSize) stack[sp++] = make_uint2(s_ref, __float_as_uint(tmin)); else *overflow = 1; }
        }
        // first child
        bool descend = false;
        if (f_cnt > 0) {
            leaf(f_ref, f_cnt);
        } else {
            if (COUNT) ++st->parent_visits;
            float tmin; bool ok;
            if (first_is_right) ok = box_test_static(bs, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, &tmin);
            else ok = box_test_static(bs, q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, &tmin);
            if (ok && tmin < ray.t_max) { cur = f_ref; descend = true; }
        }
        while (!descend) {
            if (sp == 0) { *t_hit = ray.t_max; return best; }
            const uint2 e = stack[--sp];
            if (e.x & kLeafBit) leaf(e.x & ~kLeafBit, e.y);
            else if (__uint_as_float(e.y) < ray.t_max) { cur = e.x; descend = true; }
        }
    }
}

// Any hit (shadow rays): left then right, first accepted primitive ends the search.
template <bool COUNT>
__device__ bool bvh_any(const DevScene& sc, const Ray& ray, TraceStats* st, uint32_t* overflow) {
    if (sc.n_prims == 0) return false;
    const RayBoxSetup bs = ray_box_setup(ray);
    const TriRaySetup ts = tri_ray_setup(ray.d);
    uint2 stack[kStackSize];
    int sp = 0;

    auto leaf = [&](uint32_t first, uint32_t cnt) -> bool {
        for (uint32_t s = first; s < first + cnt; ++s) {
            if (COUNT) ++st->prim_tests;
            PrimRecord pr = load_prim(sc, s);
            float t;
            if (prim_intersect_t(sc, pr, ray, ts, &t)) return true;
        }
        return false;
    };
    if (sc.root_cnt > 0) return leaf(sc.root_ref, sc.root_cnt);
    uint32_t cur;
    {
        if (COUNT) ++st->parent_visits;
        float tmin;
        if (!(box_test_static(bs, sc.root_box[0], sc.root_box[1], sc.root_box[2], sc.root_box[3], sc.root_box[4], sc.root_box[5], &tmin) && tmin < ray.t_max)) return false;
        cur = sc.root_ref;
    }
    for (;;) {
        const float4* np = sc.nodes + 4ull * cur;
        const float4 q0 = __ldg(np), q1 = __ldg(np + 1), q2 = __ldg(np + 2), q3 = __ldg(np + 3);
        const uint32_t lref = __float_as_uint(q3.x), rref = __float_as_uint(q3.y);
        const uint32_t lcnt = __float_as_uint(q3.z) & 0x3FFFFFFFu, rcnt = __float_as_uint(q3.w);
        // right child: decided now (t_max never changes for a predicate ray), visited after left
        if (rcnt > 0) {
            if (sp < kStackSize) stack[sp++] = make_uint2(rref | kLeafBit, rcnt); else *overflow = 1;
        } else {
            if (COUNT) ++st->parent_visits;
            float tmin;
            if (box_test_static(bs, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, &tmin) && tmin < ray.t_max) {
                if (sp < kStackSize) stack[sp++] = make_uint2(rref, 0u); else *overflow = 1;
            }
        }
        bool descend = false;
        if (lcnt > 0) {
            if (leaf(lref, lcnt)) return true;
        } else {
            if (COUNT) ++st->parent_visits;
            float tmin;
            if (box_test_static(bs, q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, &tmin) && tmin < ray.t_max) { cur = lref; descend = true; }
        }
        while (!descend) {
            if (sp == 0) return false;
            const uint2 e = stack[--sp];
            if (e.x & kLeafBit) { if (leaf(e.x & ~kLeafBit, e.y)) return true; }
            else { cur = e.x; descend = true; }
        }
    }
}

}  // namespace cpbrt
