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
    bool any_neg_zero; // some d component is -0.0: the reference then pairs d >= 0 (true) with 1/d = -inf
};
CP_DEV RayBoxSetup ray_box_setup(const Ray& r) {
    RayBoxSetup s;
    s.o = r.o;
    s.inv_d = v3(1.f / r.d.x, 1.f / r.d.y, 1.f / r.d.z);
    s.pos[0] = r.d.x >= 0.f; s.pos[1] = r.d.y >= 0.f; s.pos[2] = r.d.z >= 0.f;
    s.any_neg_zero = (r.d.x == 0.f && signbit(r.d.x)) || (r.d.y == 0.f && signbit(r.d.y)) || (r.d.z == 0.f && signbit(r.d.z));
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
// rays that the fast search could not settle by itself: replays in the reference's own order (a rival within the margin),
// and in the flat search the hand-overs to the reference traversal
__device__ unsigned long long g_trace_slow_paths;

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

// Tuning constants of the persistent-warp kernels (128 threads per block), per traversal kind and
// warp schedule; measured with tools/variant_sweep.py.
//   blocks  resident blocks per SM the wavefront kernel is compiled for (register budget
//           65536 / (128 * blocks): 8 -> 64 registers, 6 -> 80)
//   refill  idle lanes a warp waits for before it writes results back and fetches new rays
static constexpr int kTraceBlock = 128;
#ifndef CPBRT_BLOCKS_CLOSEST_LOOP
#define CPBRT_BLOCKS_CLOSEST_LOOP 7
#endif
#ifndef CPBRT_BLOCKS_CLOSEST_VOTE
#define CPBRT_BLOCKS_CLOSEST_VOTE 6
#endif
#ifndef CPBRT_BLOCKS_SHADOW_LOOP
#define CPBRT_BLOCKS_SHADOW_LOOP 7
#endif
#ifndef CPBRT_BLOCKS_SHADOW_VOTE
#define CPBRT_BLOCKS_SHADOW_VOTE 8
#endif
#ifndef CPBRT_REFILL_CLOSEST_LOOP
#define CPBRT_REFILL_CLOSEST_LOOP 16
#endif
#ifndef CPBRT_REFILL_SHADOW_LOOP
#define CPBRT_REFILL_SHADOW_LOOP 20
#endif
#ifndef CPBRT_REFILL_CLOSEST_VOTE
#define CPBRT_REFILL_CLOSEST_VOTE 12
#endif
#ifndef CPBRT_REFILL_SHADOW_VOTE
#define CPBRT_REFILL_SHADOW_VOTE 12
#endif

struct NodeChildren { uint32_t ref[2], cnt[2]; bool ok[2]; float tmin[2]; uint32_t axis; };

// Reads one Parent node (64 B) and runs the t_max-independent part of both children's box tests.
template <bool COUNT>
CP_DEV NodeChildren expand_node(const DevScene& sc, uint32_t node, const RayBoxSetup& bs, TraceStats* st) {
    CP_CHECK(node < sc.n_nodes, 301);
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

// ------------------------------------------------------------------------------------------
// Persistent-warp driver: the traversal above, restructured so that a lane whose ray is
// finished pulls the next ray from a global counter instead of idling until the slowest ray
// of its warp is done (ray costs are heavy-tailed: a ray grazing the terrain pierces thousands
// of boxes, its warp neighbours a few dozen).  One outer iteration =
//     refill idle lanes -> every lane descends to its next leaf -> every lane tests one leaf
// so lanes only ever wait for one descent, not for a whole ray.
//   fetch(i, &ray, &user)            load work item i
//   finish(done, user, ray, slot, t) called by ALL lanes whenever any lane finished a ray
//                                    (`done` marks the lanes that did) -- may use warp votes
// Visiting order, box arithmetic and primitive tests are exactly those of bvh_closest /
// bvh_any; only the scheduling differs.
template <bool ANY, bool COUNT, class User, class Fetch, class Finish>
__device__ void trace_persistent(const DevScene& sc, uint32_t n_items, uint32_t* work_counter, TraceStats* st, uint32_t* overflow,
                                 Fetch fetch, Finish finish) {
    const int lane = threadIdx.x & 31;
    Ray ray; ray.o = v3(0.f, 0.f, 0.f); ray.d = v3(0.f, 0.f, 1.f); ray.t_max = 0.f;
    RayBoxSetup bs = ray_box_setup(ray);
    TriRaySetup ts = tri_ray_setup(ray.d);
    TaskStack stack;
    uint2 task = make_uint2(kDone, 0u);
    uint32_t best = 0xFFFFFFFFu;
    User user = User();
    bool have_ray = false, exhausted = false;

    for (;;) {
        // ---- refill -------------------------------------------------------------------------
        const unsigned idle = __ballot_sync(0xFFFFFFFFu, !have_ray);
        if (idle && !exhausted) {
            const int leader = __ffs(idle) - 1;
            uint32_t base = 0;
            if (lane == leader) base = atomicAdd(work_counter, (uint32_t)__popc(idle));
            base = __shfl_sync(0xFFFFFFFFu, base, leader);
            if (base >= n_items) exhausted = true;   // warp-uniform
            if (!have_ray) {
                const uint32_t i = base + __popc(idle & ((1u << lane) - 1u));
                if (i < n_items) {
                    fetch(i, &ray, &user);
                    have_ray = true; best = 0xFFFFFFFFu; stack.n = 0;
                    bs = ray_box_setup(ray); ts = tri_ray_setup(ray.d);
                    if (sc.n_prims == 0) task = make_uint2(kDone, 0u);
                    else if (sc.root_cnt > 0) task = make_uint2(sc.root_ref | kLeafBit, sc.root_cnt);
                    else {
                        if (COUNT) ++st->parent_visits;
                        float tmin;
                        bool ok = box_test_static(bs, sc.root_box[0], sc.root_box[1], sc.root_box[2], sc.root_box[3], sc.root_box[4], sc.root_box[5], &tmin);
                        if (ANY) ok = ok && tmin < ray.t_max;
                        task = ok ? make_uint2(sc.root_ref, __float_as_uint(tmin)) : make_uint2(kDone, 0u);
                    }
                }
            }
        }
        if (__all_sync(0xFFFFFFFFu, !have_ray)) break;

        // ---- Parent tasks: descend to the next leaf -----------------------------------------
        while (have_ray && !(task.x & kLeafBit)) {
            if (!ANY && !(__uint_as_float(task.y) < ray.t_max)) { task = stack.pop(); continue; }
            const NodeChildren c = expand_node<COUNT>(sc, task.x, bs, st);
            if (ANY) {
                const bool ok0 = c.ok[0] && (c.cnt[0] > 0 || c.tmin[0] < ray.t_max);
                const bool ok1 = c.ok[1] && (c.cnt[1] > 0 || c.tmin[1] < ray.t_max);
                if (ok1) stack.push(child_task(c, 1), overflow);
                task = ok0 ? child_task(c, 0) : stack.pop();
            } else {
                const bool positive = c.axis == 0 ? bs.pos[0] : (c.axis == 1 ? bs.pos[1] : bs.pos[2]);
                const int first = positive ? 1 : 0, second = first ^ 1;
                if (c.ok[second]) stack.push(child_task(c, second), overflow);
                task = c.ok[first] ? child_task(c, first) : stack.pop();
            }
        }
        // ---- Leaf task ----------------------------------------------------------------------
        bool finished = false;
        if (have_ray) {
            if (task.x != kDone) {
                const uint32_t first = task.x & ~kLeafBit, end = first + task.y;
                bool stop = false;
                for (uint32_t s = first; s < end && !stop; ++s) {
                    if (COUNT) ++st->prim_tests;
                    const PrimRecord pr = load_prim(sc, s);
                    float t;
                    if (prim_intersect_t(sc, pr, ray, ts, &t)) { best = s; if (ANY) stop = true; else ray.t_max = t; }
                }
                task = stop ? make_uint2(kDone, 0u) : stack.pop();
            }
            finished = task.x == kDone;
        }
        // ---- results -------------------------------------------------------------------------
        if (__any_sync(0xFFFFFFFFu, finished)) {
            finish(finished, user, ray, best, ray.t_max);
            if (finished) have_ray = false;
        }
    }
}

// ==========================================================================================
// Two-phase closest hit: same answer as the reference-order traversal above, far less work.
//
// The reference visits the FAR child first and its box test is loose (Q5, Q6), so a ray
// touches every box near its whole length -- tens of thousands of dependent fetches for a ray
// grazing a terrain.  What the reference returns, though, is easy to characterise: process the
// reachable hit primitives in reference order, accept P when t_scaled_P <= t_max * det_P
// (i.e. t_P <= t_max up to a few ulp), and return the last accepted one.  Removing any other
// primitive from that sequence, or starting with a smaller t_max that is still safely above
// the final t, does not change the last accepted primitive.  Hence:
//   phase 1 (search)  near-first traversal with a tight slab test; finds the smallest t and
//                     whether a rival hit lies within `margin` of it.  A primitive only
//                     counts if every ancestor passes the reference's own box predicate, so
//                     nothing the reference cannot reach is ever reported.
//   phase 2 (replay)  only when a rival exists (ray through a shared edge / vertex, coincident
//                     surfaces): reference-order traversal, reference box and primitive tests,
//                     started at t_max = t_best + margin and restricted to boxes that overlap
//                     [t_best - margin, t_best + margin]; this reproduces the reference's
//                     tie-breaking exactly.
// `margin` covers the rounding of the acceptance test and of the slab distances (relative
// 2e-5 plus 2e-5 * max|bound - o| / max|d|); it only decides how often phase 2 runs.
struct TightSetup { float abs_m, t_init; };

struct ChildTests {
    uint32_t ref[2], cnt[2]; uint32_t axis;
    bool ref_ok[2]; float ref_tmin[2];     // reference predicate (t_max-independent part) and its t_min
    float entry[2], exit[2];               // tight slab interval (exit widened by 1 + 2*gamma(3))
};

CP_DEV void slab_tests(const RayBoxSetup& s, float mnx, float mny, float mnz, float mxx, float mxy, float mxz,
                       bool* ref_ok, float* ref_tmin, float* entry, float* exit) {
    const float widen = 1.f + 2.f * gamma_n(3);
    const float lo0 = ((s.pos[0] ? mnx : mxx) - s.o.x) * s.inv_d.x, hi0 = ((s.pos[0] ? mxx : mnx) - s.o.x) * s.inv_d.x * widen;
    const float lo1 = ((s.pos[1] ? mny : mxy) - s.o.y) * s.inv_d.y, hi1 = ((s.pos[1] ? mxy : mny) - s.o.y) * s.inv_d.y * widen;
    const float lo2 = ((s.pos[2] ? mnz : mxz) - s.o.z) * s.inv_d.z, hi2 = ((s.pos[2] ? mxz : mnz) - s.o.z) * s.inv_d.z * widen;
    // reference predicate, same sequence of comparisons as box_test_static
    bool ok = true;
    float t_min = lo0, t_max = hi0;
    if (t_min > hi1 || lo1 > t_max) ok = false;
    t_min = rmin(t_min, lo1); t_max = rmax(t_max, hi1);
    if (t_min > hi2 || lo2 > t_max) ok = false;
    t_min = rmin(t_min, lo2); t_max = rmax(t_max, hi2);
    *ref_ok = ok && t_max > 0.f;
    *ref_tmin = t_min;
    // tight interval; fmaxf / fminf drop NaNs, which only makes the interval more permissive.
    // With a -0.0 direction component the reference's near/far choice (d >= 0) disagrees with
    // the sign of 1/d and its slab comes out inverted; the tight test orders each slab itself.
    if (s.any_neg_zero) {
        *entry = fmaxf(fmaxf(fminf(lo0, hi0), fminf(lo1, hi1)), fminf(lo2, hi2));
        *exit = fminf(fminf(fmaxf(lo0, hi0), fmaxf(lo1, hi1)), fmaxf(lo2, hi2));
    } else {
        *entry = fmaxf(fmaxf(lo0, lo1), lo2);
        *exit = fminf(fminf(hi0, hi1), hi2);
    }
}

CP_DEV ChildTests expand_node2(const DevScene& sc, uint32_t node, const RayBoxSetup& bs) {
    CP_CHECK(node < sc.n_nodes, 301);
    const float4* np = sc.nodes + 4ull * node;
    const float4 q0 = __ldg(np), q1 = __ldg(np + 1), q2 = __ldg(np + 2), q3 = __ldg(np + 3);
    ChildTests c;
    c.ref[0] = __float_as_uint(q3.x); c.ref[1] = __float_as_uint(q3.y);
    const uint32_t lw = __float_as_uint(q3.z);
    c.cnt[0] = lw & 0x3FFFFFFFu; c.cnt[1] = __float_as_uint(q3.w); c.axis = lw >> 30;
    slab_tests(bs, q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, &c.ref_ok[0], &c.ref_tmin[0], &c.entry[0], &c.exit[0]);
    slab_tests(bs, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, &c.ref_ok[1], &c.ref_tmin[1], &c.entry[1], &c.exit[1]);
    return c;
}

CP_DEV float hit_margin(float t, float abs_m) { return 2e-5f * fabsf(t) + abs_m; }
CP_DEV float slab_slack(float a, float b, float abs_m) { return 1e-5f * fmaxf(fabsf(a), fabsf(b)) + abs_m; }

// Scale of the numbers entering the slab / triangle arithmetic for this ray: 2e-5 * max|bound - o| / max|d|
CP_DEV float ray_abs_margin(const DevScene& sc, const Ray& ray) {
    float m = 0.f;
    m = fmaxf(m, fmaxf(fabsf(sc.root_box[0] - ray.o.x), fabsf(sc.root_box[3] - ray.o.x)));
    m = fmaxf(m, fmaxf(fabsf(sc.root_box[1] - ray.o.y), fabsf(sc.root_box[4] - ray.o.y)));
    m = fmaxf(m, fmaxf(fabsf(sc.root_box[2] - ray.o.z), fabsf(sc.root_box[5] - ray.o.z)));
    const float dmax = fmaxf(fmaxf(fabsf(ray.d.x), fabsf(ray.d.y)), fabsf(ray.d.z));
    const float r = 2e-5f * m / dmax;
    return (r == r && r < 3.0e38f) ? r : 3.0e38f;   // degenerate rays: every hit is a "rival" -> replay decides
}

enum TraceMode : int { MODE_SEARCH = 0, MODE_REPLAY = 1 };

// Persistent-warp closest hit (two-phase), loop-scheduled: every lane descends to its next leaf,
// then every lane tests that leaf.  The better schedule for small trees and coherent rays (a few
// node steps per leaf, little to wait for); see closest_persistent below for the other one.
// COUNT: node steps (64-byte Parent fetches) and primitive tests (48-byte records) of THIS traversal are
// added to *st -- the kernel's own work, for the roofline (the reference-order counters live in bvh_closest).
template <class User, bool COUNT = false, class Fetch, class Finish>
__device__ void closest_persistent_loop(const DevScene& sc, uint32_t n_items, uint32_t* work_counter, uint32_t* overflow, Fetch fetch, Finish finish, TraceStats* st = nullptr) {
    const int lane = threadIdx.x & 31;
    Ray ray; ray.o = v3(0.f, 0.f, 0.f); ray.d = v3(0.f, 0.f, 1.f); ray.t_max = 0.f;
    RayBoxSetup bs = ray_box_setup(ray);
    TriRaySetup ts = tri_ray_setup(ray.d);
    TaskStack stack;
    uint2 task = make_uint2(kDone, 0u);
    User user = User();
    bool have_ray = false, exhausted = false;
    int mode = MODE_SEARCH;
    float t_init = 0.f, abs_m = 0.f;
    // search state: best / second-best accepted distance; replay state: ray.t_max and `best`
    uint32_t best = 0xFFFFFFFFu;
    float t_best = 0.f, t_second = 0.f, t_cut = 0.f, t_lo = 0.f;

    auto root_task = [&](bool replay) -> uint2 {
        if (sc.n_prims == 0) return make_uint2(kDone, 0u);
        if (sc.root_cnt > 0) return make_uint2(sc.root_ref | kLeafBit, sc.root_cnt);
        bool rok; float rtmin, en, ex;
        slab_tests(bs, sc.root_box[0], sc.root_box[1], sc.root_box[2], sc.root_box[3], sc.root_box[4], sc.root_box[5], &rok, &rtmin, &en, &ex);
        if (!rok) return make_uint2(kDone, 0u);
        if (replay) return make_uint2(sc.root_ref, __float_as_uint(rtmin));          // t_min < t_max checked when the task is taken
        if (!(rtmin < t_init)) return make_uint2(kDone, 0u);
        return make_uint2(sc.root_ref, __float_as_uint(en));
    };

    for (;;) {
        // ---- refill -------------------------------------------------------------------------
        const unsigned idle = __ballot_sync(0xFFFFFFFFu, !have_ray);
        // refill in groups: one atomic per >= CPBRT_REFILL_* idle lanes (or when the warp is empty)
        if (!exhausted && (__popc(idle) >= CPBRT_REFILL_CLOSEST_LOOP || idle == 0xFFFFFFFFu)) {
            const int leader = __ffs(idle) - 1;
            uint32_t base = 0;
            if (lane == leader) base = atomicAdd(work_counter, (uint32_t)__popc(idle));
            base = __shfl_sync(0xFFFFFFFFu, base, leader);
            if (base >= n_items) exhausted = true;
            if (!have_ray) {
                const uint32_t i = base + __popc(idle & ((1u << lane) - 1u));
                if (i < n_items) {
                    fetch(i, &ray, &user);
                    have_ray = true; stack.n = 0; mode = MODE_SEARCH;
                    bs = ray_box_setup(ray); ts = tri_ray_setup(ray.d);
                    t_init = ray.t_max; abs_m = ray_abs_margin(sc, ray);
                    best = 0xFFFFFFFFu; t_best = t_init; t_second = kFloatMax; t_cut = t_init;
                    task = root_task(false);
                }
            }
        }
        if (__all_sync(0xFFFFFFFFu, !have_ray)) break;

        // ---- Parent tasks ---------------------------------------------------------------------
        while (have_ray && !(task.x & kLeafBit)) {
            const float key = __uint_as_float(task.y);
            if (mode == MODE_SEARCH ? (key > t_cut) : !(key < ray.t_max)) { task = stack.pop(); continue; }
            const ChildTests c = expand_node2(sc, task.x, bs);
            if (COUNT) ++st->parent_visits;
            if (mode == MODE_SEARCH) {
                // near-first; a child is entered if the reference could reach it (its own
                // predicate, against the initial t_max) and the tight interval meets [0, t_cut]
                bool go[2]; float key2[2];
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    const float sl = slab_slack(c.entry[k], c.exit[k], abs_m);
                    const bool tight = !(c.entry[k] > c.exit[k] + sl) && !(c.exit[k] + sl < 0.f) && !(c.entry[k] - sl > t_cut);
                    // a Leaf has no box in the reference (always entered there); its stored
                    // box only prunes primitives that cannot be hit within [0, t_cut]
                    go[k] = tight && (c.cnt[k] > 0 || (c.ref_ok[k] && c.ref_tmin[k] < t_init));
                    key2[k] = c.entry[k] - sl;
                }
                // both candidate tasks are formed with straight-line selects (no indexed access
                // into the small per-child arrays, which would push them to local memory)
                const uint2 t0 = c.cnt[0] > 0 ? make_uint2(c.ref[0] | kLeafBit, c.cnt[0]) : make_uint2(c.ref[0], __float_as_uint(key2[0]));
                const uint2 t1 = c.cnt[1] > 0 ? make_uint2(c.ref[1] | kLeafBit, c.cnt[1]) : make_uint2(c.ref[1], __float_as_uint(key2[1]));
                const bool near0 = go[0] && (!go[1] || key2[0] <= key2[1]);
                const bool go_near = near0 ? go[0] : go[1], go_far = near0 ? go[1] : go[0];
                const uint2 t_near = near0 ? t0 : t1, t_far = near0 ? t1 : t0;
                if (go_far) stack.push(t_far, overflow);
                task = go_near ? t_near : stack.pop();
            } else {
                // reference order and reference box rule, skipping boxes that cannot hold a
                // hit inside the replay window
                bool go[2];
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    const float sl = slab_slack(c.entry[k], c.exit[k], abs_m);
                    const bool window = !(c.exit[k] + sl < t_lo) && !(c.entry[k] - sl > t_cut);
                    go[k] = window && (c.cnt[k] > 0 || c.ref_ok[k]);
                }
                const bool positive = c.axis == 0 ? bs.pos[0] : (c.axis == 1 ? bs.pos[1] : bs.pos[2]);
                const uint2 t0 = c.cnt[0] > 0 ? make_uint2(c.ref[0] | kLeafBit, c.cnt[0]) : make_uint2(c.ref[0], __float_as_uint(c.ref_tmin[0]));
                const uint2 t1 = c.cnt[1] > 0 ? make_uint2(c.ref[1] | kLeafBit, c.cnt[1]) : make_uint2(c.ref[1], __float_as_uint(c.ref_tmin[1]));
                // positive direction along the split axis -> right child (1) first
                const bool go_first = positive ? go[1] : go[0], go_second = positive ? go[0] : go[1];
                const uint2 t_first = positive ? t1 : t0, t_second = positive ? t0 : t1;
                if (go_second) stack.push(t_second, overflow);
                task = go_first ? t_first : stack.pop();
            }
        }
        // ---- Leaf task --------------------------------------------------------------------------
        bool finished = false;
        if (have_ray) {
            if (task.x != kDone) {
                const uint32_t first = task.x & ~kLeafBit, end = first + task.y;
                for (uint32_t s = first; s < end; ++s) {
                    const PrimRecord pr = load_prim(sc, s);
                    if (COUNT) ++st->prim_tests;
                    float t;
                    if (mode == MODE_SEARCH) {
                        Ray probe = ray; probe.t_max = fminf(t_init, t_cut);
                        if (prim_intersect_t(sc, pr, probe, ts, &t)) {
                            if (best == 0xFFFFFFFFu || t < t_best) { if (best != 0xFFFFFFFFu) t_second = fminf(t_second, t_best); t_best = t; best = s; }
                            else t_second = fminf(t_second, t);
                            t_cut = t_best + hit_margin(t_best, abs_m);
                        }
                    } else {
                        if (prim_intersect_t(sc, pr, ray, ts, &t)) { ray.t_max = t; best = s; }
                    }
                }
                task = stack.pop();
            }
            if (task.x == kDone) {
                if (mode == MODE_SEARCH && best != 0xFFFFFFFFu && !(t_second > t_cut)) {
                    // a rival within the margin: let the reference's own order decide
                    mode = MODE_REPLAY; atomicAdd(&g_trace_slow_paths, 1ull);
                    t_lo = t_best - hit_margin(t_best, abs_m);
                    ray.t_max = fminf(t_init, t_cut);
                    best = 0xFFFFFFFFu; stack.n = 0;
                    task = root_task(true);
                    // (t_cut stays t_best + margin: upper edge of the replay window)
                } else {
                    if (mode == MODE_SEARCH) ray.t_max = best != 0xFFFFFFFFu ? t_best : t_init;
                    finished = true;
                }
            }
        }
        // ---- results ----------------------------------------------------------------------------
        if (__any_sync(0xFFFFFFFFu, finished)) {
            finish(finished, user, ray, best, ray.t_max);
            if (finished) have_ray = false;
        }
    }
}

// Persistent-warp any hit, loop-scheduled.  The result is a pure existence test against a constant t_max, so
// the visiting order is free: nearer child first, reference predicate (required for parity)
// plus the tight interval (prunes what the reference's loose test lets through).
template <class User, bool COUNT = false, class Fetch, class Finish>
__device__ void any_persistent_loop(const DevScene& sc, uint32_t n_items, uint32_t* work_counter, uint32_t* overflow, Fetch fetch, Finish finish, TraceStats* st = nullptr) {
    const int lane = threadIdx.x & 31;
    Ray ray; ray.o = v3(0.f, 0.f, 0.f); ray.d = v3(0.f, 0.f, 1.f); ray.t_max = 0.f;
    RayBoxSetup bs = ray_box_setup(ray);
    TriRaySetup ts = tri_ray_setup(ray.d);
    TaskStack stack;
    uint2 task = make_uint2(kDone, 0u);
    User user = User();
    bool have_ray = false, exhausted = false, hit = false;
    float abs_m = 0.f;
    for (;;) {
        const unsigned idle = __ballot_sync(0xFFFFFFFFu, !have_ray);
        // refill in groups: one atomic per >= CPBRT_REFILL_* idle lanes (or when the warp is empty)
        if (!exhausted && (__popc(idle) >= CPBRT_REFILL_SHADOW_LOOP || idle == 0xFFFFFFFFu)) {
            const int leader = __ffs(idle) - 1;
            uint32_t base = 0;
            if (lane == leader) base = atomicAdd(work_counter, (uint32_t)__popc(idle));
            base = __shfl_sync(0xFFFFFFFFu, base, leader);
            if (base >= n_items) exhausted = true;
            if (!have_ray) {
                const uint32_t i = base + __popc(idle & ((1u << lane) - 1u));
                if (i < n_items) {
                    fetch(i, &ray, &user);
                    have_ray = true; stack.n = 0; hit = false;
                    bs = ray_box_setup(ray); ts = tri_ray_setup(ray.d);
                    abs_m = ray_abs_margin(sc, ray);
                    if (sc.n_prims == 0) task = make_uint2(kDone, 0u);
                    else if (sc.root_cnt > 0) task = make_uint2(sc.root_ref | kLeafBit, sc.root_cnt);
                    else {
                        bool rok; float rtmin, en, ex;
                        slab_tests(bs, sc.root_box[0], sc.root_box[1], sc.root_box[2], sc.root_box[3], sc.root_box[4], sc.root_box[5], &rok, &rtmin, &en, &ex);
                        task = (rok && rtmin < ray.t_max) ? make_uint2(sc.root_ref, 0u) : make_uint2(kDone, 0u);
                    }
                }
            }
        }
        if (__all_sync(0xFFFFFFFFu, !have_ray)) break;
        while (have_ray && !(task.x & kLeafBit)) {
            const ChildTests c = expand_node2(sc, task.x, bs);
            if (COUNT) ++st->parent_visits;
            bool go[2]; float key[2];
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const float sl = slab_slack(c.entry[k], c.exit[k], abs_m);
                const bool tight = !(c.entry[k] > c.exit[k] + sl) && !(c.exit[k] + sl < 0.f) && !(c.entry[k] - sl > ray.t_max);
                go[k] = tight && (c.cnt[k] > 0 || (c.ref_ok[k] && c.ref_tmin[k] < ray.t_max));
                key[k] = c.entry[k];
            }
            const uint2 t0 = c.cnt[0] > 0 ? make_uint2(c.ref[0] | kLeafBit, c.cnt[0]) : make_uint2(c.ref[0], 0u);
            const uint2 t1 = c.cnt[1] > 0 ? make_uint2(c.ref[1] | kLeafBit, c.cnt[1]) : make_uint2(c.ref[1], 0u);
            const bool near0 = go[0] && (!go[1] || key[0] <= key[1]);
            const bool go_near = near0 ? go[0] : go[1], go_far = near0 ? go[1] : go[0];
            const uint2 t_near = near0 ? t0 : t1, t_far = near0 ? t1 : t0;
            if (go_far) stack.push(t_far, overflow);
            task = go_near ? t_near : stack.pop();
        }
        bool finished = false;
        if (have_ray) {
            if (task.x != kDone) {
                const uint32_t first = task.x & ~kLeafBit, end = first + task.y;
                for (uint32_t s = first; s < end && !hit; ++s) {
                    const PrimRecord pr = load_prim(sc, s);
                    if (COUNT) ++st->prim_tests;
                    float t;
                    if (prim_intersect_t(sc, pr, ray, ts, &t)) hit = true;
                }
                task = hit ? make_uint2(kDone, 0u) : stack.pop();
            }
            finished = task.x == kDone;
        }
        if (__any_sync(0xFFFFFFFFu, finished)) {
            finish(finished, user, ray, hit ? 0u : 0xFFFFFFFFu, ray.t_max);
            if (finished) have_ray = false;
        }
    }
}

// Persistent-warp closest hit (two-phase), vote-scheduled.  fetch / finish as in trace_persistent.
//
// A ray alternates between node steps (expand one Parent) and primitive steps (test one primitive
// of a leaf).  Running "all node steps until every lane has a leaf, then the leaves" leaves most
// of the warp idle on incoherent rays: the lanes reach their next leaf after very different
// numbers of node steps (5 of 32 lanes active on random rays through a million triangles).  Here
// every iteration the warp VOTES: the kind of step more lanes are waiting for is executed, for
// exactly those lanes, and the others simply wait for their kind's turn.  Finished rays are
// written back, and their lanes refilled, in groups of CPBRT_REFILL_* lanes.
//   lane state   EMPTY (have_ray = false) | ACTIVE (task is a Parent or a leaf cursor) | DONE (task = kDone)
//   leaf cursor  x = kLeafBit | next slot, y = primitives left
template <class User, bool COUNT = false, class Fetch, class Finish>
__device__ void closest_persistent(const DevScene& sc, uint32_t n_items, uint32_t* work_counter, uint32_t* overflow, Fetch fetch, Finish finish, TraceStats* st = nullptr) {
    const int lane = threadIdx.x & 31;
    Ray ray; ray.o = v3(0.f, 0.f, 0.f); ray.d = v3(0.f, 0.f, 1.f); ray.t_max = 0.f;
    RayBoxSetup bs = ray_box_setup(ray);
    TriRaySetup ts = tri_ray_setup(ray.d);
    TaskStack stack;
    uint2 task = make_uint2(kDone, 0u);
    User user = User();
    bool have_ray = false, exhausted = false;
    int mode = MODE_SEARCH;
    float t_init = 0.f, abs_m = 0.f;
    // search state: best / second-best accepted distance; replay state: ray.t_max and `best`
    uint32_t best = 0xFFFFFFFFu;
    float t_best = 0.f, t_second = 0.f, t_cut = 0.f, t_lo = 0.f;

    auto root_task = [&](bool replay) -> uint2 {
        if (sc.n_prims == 0) return make_uint2(kDone, 0u);
        if (sc.root_cnt > 0) return make_uint2(sc.root_ref | kLeafBit, sc.root_cnt);
        bool rok; float rtmin, en, ex;
        slab_tests(bs, sc.root_box[0], sc.root_box[1], sc.root_box[2], sc.root_box[3], sc.root_box[4], sc.root_box[5], &rok, &rtmin, &en, &ex);
        if (!rok) return make_uint2(kDone, 0u);
        if (replay) return make_uint2(sc.root_ref, __float_as_uint(rtmin));          // t_min < t_max checked when the task is taken
        if (!(rtmin < t_init)) return make_uint2(kDone, 0u);
        return make_uint2(sc.root_ref, __float_as_uint(en));
    };

    for (;;) {
        // ---- a ray whose search ended with a rival inside the margin starts its replay ----------
        if (have_ray && task.x == kDone && mode == MODE_SEARCH && best != 0xFFFFFFFFu && !(t_second > t_cut)) {
            mode = MODE_REPLAY; atomicAdd(&g_trace_slow_paths, 1ull);
            t_lo = t_best - hit_margin(t_best, abs_m);
            ray.t_max = fminf(t_init, t_cut);      // t_cut stays t_best + margin: upper edge of the replay window
            best = 0xFFFFFFFFu; stack.n = 0;
            task = root_task(true);
        }
        const bool want_node = have_ray && !(task.x & kLeafBit);
        const bool want_prim = have_ray && (task.x & kLeafBit) && task.x != kDone;
        const unsigned m_node = __ballot_sync(0xFFFFFFFFu, want_node), m_prim = __ballot_sync(0xFFFFFFFFu, want_prim);
        const unsigned m_active = m_node | m_prim;
        // ---- write back finished rays and refill, in groups ---------------------------------------
        if (__popc(~m_active) >= CPBRT_REFILL_CLOSEST_VOTE || m_active == 0u) {
            const bool finished = have_ray && task.x == kDone;
            if (__any_sync(0xFFFFFFFFu, finished)) {
                if (finished && mode == MODE_SEARCH) ray.t_max = best != 0xFFFFFFFFu ? t_best : t_init;
                finish(finished, user, ray, best, ray.t_max);
                if (finished) have_ray = false;
            }
            if (!exhausted) {
                const unsigned idle = ~m_active;
                const int leader = __ffs(idle) - 1;
                uint32_t base = 0;
                if (lane == leader) base = atomicAdd(work_counter, (uint32_t)__popc(idle));
                base = __shfl_sync(0xFFFFFFFFu, base, leader);
                if (base >= n_items) exhausted = true;
                if (!have_ray) {
                    const uint32_t i = base + __popc(idle & ((1u << lane) - 1u));
                    if (i < n_items) {
                        fetch(i, &ray, &user);
                        have_ray = true; stack.n = 0; mode = MODE_SEARCH;
                        bs = ray_box_setup(ray); ts = tri_ray_setup(ray.d);
                        t_init = ray.t_max; abs_m = ray_abs_margin(sc, ray);
                        best = 0xFFFFFFFFu; t_best = t_init; t_second = kFloatMax; t_cut = t_init;
                        task = root_task(false);
                    }
                }
                continue;      // vote again with the new rays
            }
            if (m_active == 0u) break;      // nothing left anywhere in this warp
        }

        if (__popc(m_node) >= __popc(m_prim)) {
            // ---- node step ------------------------------------------------------------------------
            if (want_node) {
                const float key = __uint_as_float(task.y);
                if (mode == MODE_SEARCH ? (key > t_cut) : !(key < ray.t_max)) { task = stack.pop(); }
                else {
                    const ChildTests c = expand_node2(sc, task.x, bs);
            if (COUNT) ++st->parent_visits;
                    if (mode == MODE_SEARCH) {
                        // near-first; a child is entered if the reference could reach it (its own
                        // predicate, against the initial t_max) and the tight interval meets [0, t_cut]
                        bool go[2]; float key2[2];
#pragma unroll
                        for (int k = 0; k < 2; ++k) {
                            const float sl = slab_slack(c.entry[k], c.exit[k], abs_m);
                            const bool tight = !(c.entry[k] > c.exit[k] + sl) && !(c.exit[k] + sl < 0.f) && !(c.entry[k] - sl > t_cut);
                            // a Leaf has no box in the reference (always entered there); its stored
                            // box only prunes primitives that cannot be hit within [0, t_cut]
                            go[k] = tight && (c.cnt[k] > 0 || (c.ref_ok[k] && c.ref_tmin[k] < t_init));
                            key2[k] = c.entry[k] - sl;
                        }
                        // both candidate tasks are formed with straight-line selects (no indexed access
                        // into the small per-child arrays, which would push them to local memory)
                        const uint2 t0 = c.cnt[0] > 0 ? make_uint2(c.ref[0] | kLeafBit, c.cnt[0]) : make_uint2(c.ref[0], __float_as_uint(key2[0]));
                        const uint2 t1 = c.cnt[1] > 0 ? make_uint2(c.ref[1] | kLeafBit, c.cnt[1]) : make_uint2(c.ref[1], __float_as_uint(key2[1]));
                        const bool near0 = go[0] && (!go[1] || key2[0] <= key2[1]);
                        const bool go_near = near0 ? go[0] : go[1], go_far = near0 ? go[1] : go[0];
                        const uint2 t_near = near0 ? t0 : t1, t_far = near0 ? t1 : t0;
                        if (go_far) stack.push(t_far, overflow);
                        task = go_near ? t_near : stack.pop();
                    } else {
                        // reference order and reference box rule, skipping boxes that cannot hold a
                        // hit inside the replay window
                        bool go[2];
#pragma unroll
                        for (int k = 0; k < 2; ++k) {
                            const float sl = slab_slack(c.entry[k], c.exit[k], abs_m);
                            const bool window = !(c.exit[k] + sl < t_lo) && !(c.entry[k] - sl > t_cut);
                            go[k] = window && (c.cnt[k] > 0 || c.ref_ok[k]);
                        }
                        const bool positive = c.axis == 0 ? bs.pos[0] : (c.axis == 1 ? bs.pos[1] : bs.pos[2]);
                        const uint2 t0 = c.cnt[0] > 0 ? make_uint2(c.ref[0] | kLeafBit, c.cnt[0]) : make_uint2(c.ref[0], __float_as_uint(c.ref_tmin[0]));
                        const uint2 t1 = c.cnt[1] > 0 ? make_uint2(c.ref[1] | kLeafBit, c.cnt[1]) : make_uint2(c.ref[1], __float_as_uint(c.ref_tmin[1]));
                        // positive direction along the split axis -> right child (1) first
                        const bool go_first = positive ? go[1] : go[0], go_second = positive ? go[0] : go[1];
                        const uint2 t_first = positive ? t1 : t0, t_second_task = positive ? t0 : t1;
                        if (go_second) stack.push(t_second_task, overflow);
                        task = go_first ? t_first : stack.pop();
                    }
                }
            }
        } else {
            // ---- primitive step: one primitive of the lane's leaf ----------------------------------
            if (want_prim) {
                const uint32_t s = task.x & ~kLeafBit;
                const PrimRecord pr = load_prim(sc, s);
                    if (COUNT) ++st->prim_tests;
                float t;
                if (mode == MODE_SEARCH) {
                    Ray probe = ray; probe.t_max = fminf(t_init, t_cut);
                    if (prim_intersect_t(sc, pr, probe, ts, &t)) {
                        if (best == 0xFFFFFFFFu || t < t_best) { if (best != 0xFFFFFFFFu) t_second = fminf(t_second, t_best); t_best = t; best = s; }
                        else t_second = fminf(t_second, t);
                        t_cut = t_best + hit_margin(t_best, abs_m);
                    }
                } else {
                    if (prim_intersect_t(sc, pr, ray, ts, &t)) { ray.t_max = t; best = s; }
                }
                task.x += 1u; task.y -= 1u;
                if (task.y == 0u) task = stack.pop();
            }
        }
    }
}

// Persistent-warp any hit, vote-scheduled like closest_persistent.  The result is a pure
// existence test against a constant t_max, so the visiting order is free: nearer child first,
// reference predicate (required for parity) plus the tight interval (prunes what the reference's
// loose test lets through).
template <class User, bool COUNT = false, class Fetch, class Finish>
__device__ void any_persistent(const DevScene& sc, uint32_t n_items, uint32_t* work_counter, uint32_t* overflow, Fetch fetch, Finish finish, TraceStats* st = nullptr) {
    const int lane = threadIdx.x & 31;
    Ray ray; ray.o = v3(0.f, 0.f, 0.f); ray.d = v3(0.f, 0.f, 1.f); ray.t_max = 0.f;
    RayBoxSetup bs = ray_box_setup(ray);
    TriRaySetup ts = tri_ray_setup(ray.d);
    TaskStack stack;
    uint2 task = make_uint2(kDone, 0u);
    User user = User();
    bool have_ray = false, exhausted = false, hit = false;
    float abs_m = 0.f;
    for (;;) {
        const bool want_node = have_ray && !(task.x & kLeafBit);
        const bool want_prim = have_ray && (task.x & kLeafBit) && task.x != kDone;
        const unsigned m_node = __ballot_sync(0xFFFFFFFFu, want_node), m_prim = __ballot_sync(0xFFFFFFFFu, want_prim);
        const unsigned m_active = m_node | m_prim;
        if (__popc(~m_active) >= CPBRT_REFILL_SHADOW_VOTE || m_active == 0u) {
            const bool finished = have_ray && task.x == kDone;
            if (__any_sync(0xFFFFFFFFu, finished)) {
                finish(finished, user, ray, hit ? 0u : 0xFFFFFFFFu, ray.t_max);
                if (finished) have_ray = false;
            }
            if (!exhausted) {
                const unsigned idle = ~m_active;
                const int leader = __ffs(idle) - 1;
                uint32_t base = 0;
                if (lane == leader) base = atomicAdd(work_counter, (uint32_t)__popc(idle));
                base = __shfl_sync(0xFFFFFFFFu, base, leader);
                if (base >= n_items) exhausted = true;
                if (!have_ray) {
                    const uint32_t i = base + __popc(idle & ((1u << lane) - 1u));
                    if (i < n_items) {
                        fetch(i, &ray, &user);
                        have_ray = true; stack.n = 0; hit = false;
                        bs = ray_box_setup(ray); ts = tri_ray_setup(ray.d);
                        abs_m = ray_abs_margin(sc, ray);
                        if (sc.n_prims == 0) task = make_uint2(kDone, 0u);
                        else if (sc.root_cnt > 0) task = make_uint2(sc.root_ref | kLeafBit, sc.root_cnt);
                        else {
                            bool rok; float rtmin, en, ex;
                            slab_tests(bs, sc.root_box[0], sc.root_box[1], sc.root_box[2], sc.root_box[3], sc.root_box[4], sc.root_box[5], &rok, &rtmin, &en, &ex);
                            task = (rok && rtmin < ray.t_max) ? make_uint2(sc.root_ref, 0u) : make_uint2(kDone, 0u);
                        }
                    }
                }
                continue;
            }
            if (m_active == 0u) break;
        }
        if (__popc(m_node) >= __popc(m_prim)) {
            if (want_node) {
                const ChildTests c = expand_node2(sc, task.x, bs);
            if (COUNT) ++st->parent_visits;
                bool go[2]; float key[2];
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    const float sl = slab_slack(c.entry[k], c.exit[k], abs_m);
                    const bool tight = !(c.entry[k] > c.exit[k] + sl) && !(c.exit[k] + sl < 0.f) && !(c.entry[k] - sl > ray.t_max);
                    go[k] = tight && (c.cnt[k] > 0 || (c.ref_ok[k] && c.ref_tmin[k] < ray.t_max));
                    key[k] = c.entry[k];
                }
                const uint2 t0 = c.cnt[0] > 0 ? make_uint2(c.ref[0] | kLeafBit, c.cnt[0]) : make_uint2(c.ref[0], 0u);
                const uint2 t1 = c.cnt[1] > 0 ? make_uint2(c.ref[1] | kLeafBit, c.cnt[1]) : make_uint2(c.ref[1], 0u);
                const bool near0 = go[0] && (!go[1] || key[0] <= key[1]);
                const bool go_near = near0 ? go[0] : go[1], go_far = near0 ? go[1] : go[0];
                const uint2 t_near = near0 ? t0 : t1, t_far = near0 ? t1 : t0;
                if (go_far) stack.push(t_far, overflow);
                task = go_near ? t_near : stack.pop();
            }
        } else {
            if (want_prim) {
                const PrimRecord pr = load_prim(sc, task.x & ~kLeafBit);
                if (COUNT) ++st->prim_tests;
                float t;
                if (prim_intersect_t(sc, pr, ray, ts, &t)) { hit = true; task = make_uint2(kDone, 0u); }
                else {
                    task.x += 1u; task.y -= 1u;
                    if (task.y == 0u) task = stack.pop();
                }
            }
        }
    }
}

}  // namespace cpbrt
