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

static constexpr int kStackSize = 64;
static constexpr uint32_t kLeafBit = 0x80000000u;
static constexpr uint32_t kDone = 0xFFFFFFFFu;      // "no more work": has the leaf bit set on purpose

struct TraceStats { uint32_t parent_visits, prim_tests; };

// Work items of the traversal ("tasks"), packed in a uint2:
//   Parent task  x = node index,             y = bits(t_min of that node's own box test)
//   Leaf task    x = kLeafBit | first slot,  y = primitive count
// The loop keeps exactly ONE place in the code where Parent nodes are expanded and ONE where
// primitives are tested, so the lanes of a warp that are at different depths of different
// subtrees still execute the same instructions (SIMT "while-while" form).
struct TaskStack {
    uint2 e[kStackSize];
    int n = 0;
    CP_DEV void push(uint2 t, uint32_t* overflow) { if (n < kStackSize) e[n++] = t; else *overflow = 1; }
    CP_DEV uint2 pop() { return n ? e[--n] : make_uint2(kDone, 0u); }
};

struct NodeChildren { uint32_t ref[2], cnt[2]; bool ok[2]; float tmin[2]; uint32_t axis; };

// Reads one Parent node (64 B) and runs the t_max-independent part of both children's box tests.
template <bool COUNT>
CP_DEV NodeChildren expand_node(const DevScene& sc, uint32_t node, const RayBoxSetup& bs, TraceStats* st) {
    const float4* np = sc.nodes + 4ull * node;
    const float4 q0 = __ldg(np), q1 = __ldg(np + 1), q2 = __ldg(np + 2), q3 = __ldg(np + 3);
    NodeChildren c;
    c.ref[0] = __float_as_uint(q3.x); c.ref[1] = __float_as_uint(q3.y);
    const uint32_t lw = __float_as_uint(q3.z);
    c.cnt[0] = lw & 0x3FFFFFFFu; c.cnt[1] = __float_as_uint(q3.w); c.axis = lw >> 30;
    c.tmin[0] = 0.f; c.tmin[1] = 0.f;
    c.ok[0] = c.cnt[0] > 0 || box_test_static(bs, q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, &c.tmin[0]);
    c.ok[1] = c.cnt[1] > 0 || box_test_static(bs, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, &c.tmin[1]);
    if (COUNT) st->parent_visits += (c.cnt[0] == 0) + (c.cnt[1] == 0);
    return c;
}
CP_DEV uint2 child_task(const NodeChildren& c, int k) {
    return c.cnt[k] > 0 ? make_uint2(c.ref[k] | kLeafBit, c.cnt[k]) : make_uint2(c.ref[k], __float_as_uint(c.tmin[k]));
}

// Closest hit.  Returns the BVH slot (or 0xFFFFFFFF) and leaves the hit distance in *t_hit.
template <bool COUNT>
__device__ uint32_t bvh_closest(const DevScene& sc, Ray ray, float* t_hit, TraceStats* st, uint32_t* overflow) {
    uint32_t best = 0xFFFFFFFFu;
    *t_hit = ray.t_max;
    if (sc.n_prims == 0) return best;
    const RayBoxSetup bs = ray_box_setup(ray);
    const TriRaySetup ts = tri_ray_setup(ray.d);
    TaskStack stack;
    uint2 task;
    if (sc.root_cnt > 0) task = make_uint2(sc.root_ref | kLeafBit, sc.root_cnt);
    else {
        if (COUNT) ++st->parent_visits;
        float tmin;
        const bool ok = box_test_static(bs, sc.root_box[0], sc.root_box[1], sc.root_box[2], sc.root_box[3], sc.root_box[4], sc.root_box[5], &tmin);
        task = ok ? make_uint2(sc.root_ref, __float_as_uint(tmin)) : make_uint2(kDone, 0u);
    }
    for (;;) {
        // ---- Parent tasks: descend until the next leaf (or until nothing is left) ----------
        while (!(task.x & kLeafBit)) {
            // the part of the box test that depends on t_max, with the t_max current NOW
            if (!(__uint_as_float(task.y) < ray.t_max)) { task = stack.pop(); continue; }
            const NodeChildren c = expand_node<COUNT>(sc, task.x, bs, st);
            // reference order: positive direction along the split axis -> right child first
            const bool positive = c.axis == 0 ? bs.pos[0] : (c.axis == 1 ? bs.pos[1] : bs.pos[2]);
            const int first = positive ? 1 : 0, second = first ^ 1;
            if (c.ok[second]) stack.push(child_task(c, second), overflow);
            task = c.ok[first] ? child_task(c, first) : stack.pop();
        }
        if (task.x == kDone) break;
        // ---- Leaf task: every primitive, no box test, later hit wins ties -------------------
        const uint32_t first = task.x & ~kLeafBit, end = first + task.y;
        for (uint32_t s = first; s < end; ++s) {
            if (COUNT) ++st->prim_tests;
            const PrimRecord pr = load_prim(sc, s);
            float t;
            if (prim_intersect_t(sc, pr, ray, ts, &t)) { ray.t_max = t; best = s; }
        }
        task = stack.pop();
    }
    *t_hit = ray.t_max;
    return best;
}

// Any hit (shadow rays): left then right, first accepted primitive ends the search.  t_max
// never changes, so a Parent child whose full box test fails is dropped when its parent is read.
template <bool COUNT>
__device__ bool bvh_any(const DevScene& sc, const Ray& ray, TraceStats* st, uint32_t* overflow) {
    if (sc.n_prims == 0) return false;
    const RayBoxSetup bs = ray_box_setup(ray);
    const TriRaySetup ts = tri_ray_setup(ray.d);
    TaskStack stack;
    uint2 task;
    if (sc.root_cnt > 0) task = make_uint2(sc.root_ref | kLeafBit, sc.root_cnt);
    else {
        if (COUNT) ++st->parent_visits;
        float tmin;
        const bool ok = box_test_static(bs, sc.root_box[0], sc.root_box[1], sc.root_box[2], sc.root_box[3], sc.root_box[4], sc.root_box[5], &tmin);
        task = (ok && tmin < ray.t_max) ? make_uint2(sc.root_ref, 0u) : make_uint2(kDone, 0u);
    }
    for (;;) {
        while (!(task.x & kLeafBit)) {
            const NodeChildren c = expand_node<COUNT>(sc, task.x, bs, st);
            const bool ok0 = c.ok[0] && (c.cnt[0] > 0 || c.tmin[0] < ray.t_max);
            const bool ok1 = c.ok[1] && (c.cnt[1] > 0 || c.tmin[1] < ray.t_max);
            if (ok1) stack.push(child_task(c, 1), overflow);
            task = ok0 ? child_task(c, 0) : stack.pop();
        }
        if (task.x == kDone) return false;
        const uint32_t first = task.x & ~kLeafBit, end = first + task.y;
        for (uint32_t s = first; s < end; ++s) {
            if (COUNT) ++st->prim_tests;
            const PrimRecord pr = load_prim(sc, s);
            float t;
            if (prim_intersect_t(sc, pr, ray, ts, &t)) return true;
        }
        task = stack.pop();
    }
}

}  // namespace cpbrt
