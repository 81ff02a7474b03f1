// Closest hit / any hit for TINY scenes (at most kFlatMaxPrims primitives: the Cornell box has 33)
// without walking the tree.
//
// On a tree this small the warp spends its time diverging: 14 of 32 lanes active in the
// loop-scheduled walk (ncu, profiles/r01_s3_ncu_cornell.txt), ~140 warp instructions per ray.  What
// the reference returns does not depend on HOW the candidates are found (bvh_trace.cuh, "two-phase"):
// it is the nearest accepted primitive, unless a rival lies within `margin` of it, in which case the
// reference's own visiting order decides.  So here every lane tests its ray against EVERY primitive,
// in the same order as its 31 neighbours:
//   phase 1  the edge-function part of the watertight test (triangle.rs:196-222) for all
//            untransformed triangles, straight from shared memory, no divergence at all; the lanes
//            keep a bit mask of the triangles whose edge test passes (the ray's LINE meets them).
//            Spheres / transformed shapes ("specials") are pre-selected by their world bounds.
//   phase 2  the exact reference test (same arithmetic, same bits) for the few marked candidates,
//            tracking the best and second-best distance exactly like the search phase of
//            closest_persistent.
//   check    the reference only reaches a primitive through Parent boxes that pass its own loose
//            predicate (bounds.rs:306-323).  That predicate is MONOTONE in the box: every slab
//            bound is a monotone f32 function of the box corner, so a box that contains another
//            passes whenever the inner one does (finite ray, no zero direction component).  The
//            primitive's own bound is inside every ancestor's, hence "own bound passes, and its
//            t_min is below the hit distance" proves that the reference reaches and accepts the hit.
//   slow     anything else -- a rival within the margin, an own-bound test that fails, a ray with
//            a zero or non-finite direction component -- is handed to the reference-order
//            traversal (bvh_closest / bvh_any) in the same thread.  ~1e-4 of the rays.
// Same hits, same distances (bits) as the tree walks; tests/test_gpu_parity.py runs all schedules
// against the oracle and against each other.
#pragma once
#include "bvh_trace.cuh"

namespace cpbrt {

static constexpr int kFlatMaxPrims = 64;
// resident blocks per SM of the flat kernels (launch bounds = register budget); the batch entry points use CPBRT_BLOCKS_FLAT
#ifndef CPBRT_BLOCKS_FLAT
#define CPBRT_BLOCKS_FLAT 8
#endif
// The render's own kernels: 5 / 6 blocks (87 / 77 registers, no spills) instead of 8 (64 registers, ~100 B of
// spills).  Alone the kernels are only 3-4 % faster that way, but two pass lanes run at once (cpbrt_render) and a
// grid that does not claim the whole register file lets the other lane's shading blocks move in next to it:
// Cornell box 841 -> 888 Msamples/s together with 6 matte blocks (profiles/r02_block_count_sweep.txt).
#ifndef CPBRT_BLOCKS_FLAT_CLOSEST
#define CPBRT_BLOCKS_FLAT_CLOSEST 5
#endif
#ifndef CPBRT_BLOCKS_FLAT_SHADOW
#define CPBRT_BLOCKS_FLAT_SHADOW 6
#endif

#ifndef CPBRT_FLAT_UNROLL
// Items per trip of the edge phase.  Alone the trace kernels are 1-3 % faster unrolled by 4, but the frame -- two pass
// lanes, this kernel sharing the SMs and their instruction caches with the 300 KB matte kernel -- is 1.7 % faster
// with the rolled loop (888 -> 903 Msamples/s, repeatable to 0.1 %; 2 and 8 are worse: profiles/r02_block_count_sweep.txt).
#define CPBRT_FLAT_UNROLL 1
#endif
static constexpr int kFlatUnroll = CPBRT_FLAT_UNROLL;
// Tables in shared memory (copied from DevScene::flat at kernel start), float4 units:
//   pairs  [c * pstride + 3 * k + 0..2]        item k of phase 1 for a ray whose dominant axis is c (components
//                                              rotated like permute() of shapes.cuh): two triangles that share an
//                                              edge, as four vertices {S0.xyz S1.x}{S1.yz Ua.xy}{Ua.z Ub.xyz}
//                                              (S0 S1 shared, Ua / Ub the third vertex of triangle a / b); a
//                                              triangle without a partner has Ub = Ua and no b bit in item_mask
//   tris   tri_off + [c * stride + 3 * s + 0..2]  triangle of BVH slot s in its own vertex order, rotated; the
//                                              exact test of phase 2 (zeros for specials)
//   boxes  box_off + [2 * s + 0..1]            world bounds of slot s: {min.xyz, max.x}{max.yz, -, -}
//   slots  tab_off (bytes)                     BVH slot of candidate bit 2 * k (triangle a) and 2 * k + 1 (b)
// Why pairs: the edge function of an ordered vertex pair, E(u, v) = u.x * v.y - u.y * v.x, satisfies
// E(v, u) == -E(u, v) bit for bit, and the reference's rejection `(some e < 0) && (some e > 0)`
// (triangle.rs:213-222) does not change when all three values are negated or reordered.  The two
// triangles of a quad therefore need four vertex transforms and five edge functions instead of six and six.
extern __shared__ float4 flat_smem[];

struct FlatLayout { uint32_t tri_off, box_off, tab_off, total; };
__host__ __device__ inline FlatLayout flat_layout(uint32_t n, uint32_t stride, uint32_t pstride) {
    FlatLayout L; L.tri_off = 3u * pstride; L.box_off = L.tri_off + 3u * stride; L.tab_off = L.box_off + 2u * n; L.total = L.tab_off + 4u;
    return L;
}
CP_DEV void flat_load(const DevScene& sc) {
    const uint32_t total = flat_layout(sc.flat_n, sc.flat_stride, sc.flat_pstride).total;
    for (uint32_t i = threadIdx.x; i < total; i += blockDim.x) flat_smem[i] = __ldg(sc.flat + i);
    __syncthreads();
}
static inline size_t flat_smem_bytes(uint32_t n, uint32_t stride, uint32_t pstride) { return (size_t)flat_layout(n, stride, pstride).total * sizeof(float4); }

struct FlatTri { V3 p0, p1, p2; };
CP_DEV FlatTri flat_tri_rel(uint32_t base, uint32_t s, V3 op) {   // rotated vertices minus the rotated origin == permute(p - o)
    const float4 a = flat_smem[base + 3u * s], b = flat_smem[base + 3u * s + 1u], c = flat_smem[base + 3u * s + 2u];
    FlatTri t;
    t.p0 = v3(a.x - op.x, a.y - op.y, a.z - op.z);
    t.p1 = v3(a.w - op.x, b.x - op.y, b.y - op.z);
    t.p2 = v3(b.z - op.x, b.w - op.y, c.x - op.z);
    return t;
}
CP_DEV uint32_t flat_slot_of(const DevScene& sc, uint32_t bit) {
    return reinterpret_cast<const uint8_t*>(flat_smem + flat_layout(sc.flat_n, sc.flat_stride, sc.flat_pstride).tab_off)[bit];
}

// phase 1: bit 2k / 2k+1 set <=> the edge functions of triangle a / b of item k do not reject the ray (triangle.rs:213-222)
CP_DEV uint64_t flat_edge_candidates(const DevScene& sc, uint32_t kz, V3 op, float sx, float sy) {
    uint32_t m[2] = {0u, 0u};
    const uint32_t base = kz * sc.flat_pstride;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const uint32_t lo = 16u * h;
        if (lo < sc.flat_items) {
            const uint32_t cnt = sc.flat_items - lo < 16u ? sc.flat_items - lo : 16u;
            uint32_t bit = 1u, mm = 0u;
#pragma unroll kFlatUnroll
            for (uint32_t k = 0; k < cnt; ++k) {
                const float4 a = flat_smem[base + 3u * (lo + k)], b = flat_smem[base + 3u * (lo + k) + 1u], c = flat_smem[base + 3u * (lo + k) + 2u];
                // sheared x, y of the four vertices (the z components do not enter the edge functions)
                const float s0z = a.z - op.z, s1z = b.y - op.z, uaz = c.x - op.z, ubz = c.w - op.z;
                const float s0x = (a.x - op.x) + sx * s0z, s0y = (a.y - op.y) + sy * s0z;
                const float s1x = (a.w - op.x) + sx * s1z, s1y = (b.x - op.y) + sy * s1z;
                const float uax = (b.z - op.x) + sx * uaz, uay = (b.w - op.y) + sy * uaz;
                const float ubx = (c.y - op.x) + sx * ubz, uby = (c.z - op.y) + sy * ubz;
                const float es = s0x * s1y - s0y * s1x;        // E(S0, S1)
                const float a1 = s1x * uay - s1y * uax;        // E(S1, Ua)
                const float a2 = uax * s0y - uay * s0x;        // E(Ua, S0)
                const float b1 = s0x * uby - s0y * ubx;        // E(S0, Ub)
                const float b2 = ubx * s1y - uby * s1x;        // E(Ub, S1)
                // (any e < 0) && (any e > 0); fminf / fmaxf skip NaNs exactly like the comparisons do
                const float mna = fminf(fminf(es, a1), a2), mxa = fmaxf(fmaxf(es, a1), a2);
                const float mnb = fminf(fminf(-es, b1), b2), mxb = fmaxf(fmaxf(-es, b1), b2);
                if (!(mna < 0.f && mxa > 0.f)) mm |= bit;
                if (!(mnb < 0.f && mxb > 0.f)) mm |= bit + bit;
                bit <<= 2;
            }
            m[h] = mm;
        }
    }
    return (((uint64_t)m[1] << 32) | m[0]) & sc.flat_item_mask;
}

CP_DEV bool flat_ray_ordinary(const Ray& r) {   // finite origin, finite direction without zero components
    const float big = kFloatMax;
    return fabsf(r.d.x) > 0.f && fabsf(r.d.y) > 0.f && fabsf(r.d.z) > 0.f && fabsf(r.d.x) <= big && fabsf(r.d.y) <= big && fabsf(r.d.z) <= big &&
           fabsf(r.o.x) <= big && fabsf(r.o.y) <= big && fabsf(r.o.z) <= big;
}

// specials whose world bounds the ray can touch within [0, t_max] (tight slab test with slack: never drops
// a primitive the exact test could accept)
CP_DEV uint64_t flat_special_candidates(const DevScene& sc, const RayBoxSetup& bs, float abs_m, float t_max) {
    uint64_t out = 0, todo = sc.flat_special_mask;      // warp-uniform
    const uint32_t box = flat_layout(sc.flat_n, sc.flat_stride, sc.flat_pstride).box_off;
    while (todo) {
        const int s = __ffsll((long long)todo) - 1; todo &= todo - 1;
        const float4 b0 = flat_smem[box + 2u * s], b1 = flat_smem[box + 2u * s + 1u];
        bool rok; float rtmin, en, ex;
        slab_tests(bs, b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, &rok, &rtmin, &en, &ex);
        const float sl = slab_slack(en, ex, abs_m);
        if (!(en > ex + sl) && !(ex + sl < 0.f) && !(en - sl > t_max)) out |= 1ull << s;
    }
    return out;
}

// the reference's own box predicate on the primitive's world bounds; `limit` = what its t_min must stay below
CP_DEV bool flat_reachable(const DevScene& sc, const RayBoxSetup& bs, uint32_t s, float limit) {
    const uint32_t box = flat_layout(sc.flat_n, sc.flat_stride, sc.flat_pstride).box_off;
    const float4 b0 = flat_smem[box + 2u * s], b1 = flat_smem[box + 2u * s + 1u];
    float tmin;
    return box_test_static(bs, b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, &tmin) && tmin < limit;
}

// Closest hit of one ray per lane.  Returns the BVH slot (0xFFFFFFFF = miss), *t_hit = its distance (ray.t_max on a miss).
CP_DEV uint32_t flat_closest(const DevScene& sc, const Ray& ray, bool active, float* t_hit, uint32_t* overflow) {
    const float t_init = ray.t_max;
    const bool ordinary = flat_ray_ordinary(ray);
    const bool fast = active && ordinary;
    const TriRaySetup ts = tri_ray_setup(ray.d);
    const RayBoxSetup bs = ray_box_setup(ray);
    const V3 op = permute(ray.o, ts);
    const float abs_m = ray_abs_margin(sc, ray);
    const uint32_t base = flat_layout(sc.flat_n, sc.flat_stride, sc.flat_pstride).tri_off + (uint32_t)ts.kz * sc.flat_stride;
    uint64_t ct = flat_edge_candidates(sc, (uint32_t)ts.kz, op, ts.sx, ts.sy);
    uint64_t cs = sc.flat_special_mask ? flat_special_candidates(sc, bs, abs_m, t_init) : 0ull;
    if (!fast) { ct = 0; cs = 0; }
    uint32_t best = 0xFFFFFFFFu;
    float t_best = t_init, t_second = kFloatMax, t_cut = t_init;
    auto accept = [&](uint32_t s, float t) {
        if (best == 0xFFFFFFFFu || t < t_best) { if (best != 0xFFFFFFFFu) t_second = fminf(t_second, t_best); t_best = t; best = s; }
        else t_second = fminf(t_second, t);
        t_cut = t_best + hit_margin(t_best, abs_m);
    };
    // phase 2: exact tests
    while (ct) {
        const uint32_t s = flat_slot_of(sc, (uint32_t)(__ffsll((long long)ct) - 1)); ct &= ct - 1;
        const FlatTri tr = flat_tri_rel(base, s, op);
        float t;
        if (tri_intersect_permuted(tr.p0, tr.p1, tr.p2, fminf(t_init, t_cut), ts, &t, nullptr)) accept(s, t);
    }
    while (cs) {
        const uint32_t s = (uint32_t)(__ffsll((long long)cs) - 1); cs &= cs - 1;
        Ray probe = ray; probe.t_max = fminf(t_init, t_cut);
        float t;
        if (prim_intersect_t(sc, load_prim(sc, s), probe, ts, &t)) accept(s, t);
    }
    bool slow = active && !ordinary;
    if (fast && best != 0xFFFFFFFFu) slow = !(t_second > t_cut) || !flat_reachable(sc, bs, best, t_best);
    float t = best != 0xFFFFFFFFu ? t_best : t_init;
    if (slow) { best = bvh_closest<false>(sc, ray, &t, nullptr, overflow); atomicAdd(&g_trace_slow_paths, 1ull); }      // the reference's own order decides
    *t_hit = t;
    return best;
}

// Any hit: is some primitive that the reference can reach accepted within ray.t_max?
CP_DEV bool flat_any(const DevScene& sc, const Ray& ray, bool active, uint32_t* overflow) {
    const bool ordinary = flat_ray_ordinary(ray);
    const bool fast = active && ordinary;
    const TriRaySetup ts = tri_ray_setup(ray.d);
    const RayBoxSetup bs = ray_box_setup(ray);
    const V3 op = permute(ray.o, ts);
    const uint32_t base = flat_layout(sc.flat_n, sc.flat_stride, sc.flat_pstride).tri_off + (uint32_t)ts.kz * sc.flat_stride;
    uint64_t ct = flat_edge_candidates(sc, (uint32_t)ts.kz, op, ts.sx, ts.sy);
    uint64_t cs = sc.flat_special_mask ? flat_special_candidates(sc, bs, ray_abs_margin(sc, ray), ray.t_max) : 0ull;
    if (!fast) { ct = 0; cs = 0; }
    bool hit = false, unsure = active && !ordinary;
    while (ct) {
        const uint32_t s = flat_slot_of(sc, (uint32_t)(__ffsll((long long)ct) - 1)); ct &= ct - 1;
        const FlatTri tr = flat_tri_rel(base, s, op);
        float t;
        if (tri_intersect_permuted(tr.p0, tr.p1, tr.p2, ray.t_max, ts, &t, nullptr)) {
            if (flat_reachable(sc, bs, s, ray.t_max)) { hit = true; ct = 0; cs = 0; } else unsure = true;
        }
    }
    while (cs) {
        const uint32_t s = (uint32_t)(__ffsll((long long)cs) - 1); cs &= cs - 1;
        float t;
        if (prim_intersect_t(sc, load_prim(sc, s), ray, ts, &t)) {
            if (flat_reachable(sc, bs, s, ray.t_max)) { hit = true; cs = 0; } else unsure = true;
        }
    }
    if (!hit && unsure) { hit = bvh_any<false>(sc, ray, nullptr, overflow); atomicAdd(&g_trace_slow_paths, 1ull); }
    return hit;
}

// Drivers with the fetch / finish interface of the persistent-warp traversals: one queue item per
// thread and trip, whole warps (finish may use warp votes).
template <class User, class Fetch, class Finish>
__device__ void closest_flat(const DevScene& sc, uint32_t n_items, uint32_t* overflow, Fetch fetch, Finish finish) {
    flat_load(sc);
    const uint32_t n_round = (n_items + 31u) & ~31u;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += gridDim.x * blockDim.x) {
        const bool active = i < n_items;
        Ray ray; ray.o = v3(0.f, 0.f, 0.f); ray.d = v3(0.25f, 0.5f, 1.f); ray.t_max = 0.f;
        User user = User();
        if (active) fetch(i, &ray, &user);
        float t;
        const uint32_t slot = flat_closest(sc, ray, active, &t, overflow);
        ray.t_max = t;
        finish(active, user, ray, slot, t);
    }
}
template <class User, class Fetch, class Finish>
__device__ void any_flat(const DevScene& sc, uint32_t n_items, uint32_t* overflow, Fetch fetch, Finish finish) {
    flat_load(sc);
    const uint32_t n_round = (n_items + 31u) & ~31u;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += gridDim.x * blockDim.x) {
        const bool active = i < n_items;
        Ray ray; ray.o = v3(0.f, 0.f, 0.f); ray.d = v3(0.25f, 0.5f, 1.f); ray.t_max = 0.f;
        User user = User();
        if (active) fetch(i, &ray, &user);
        const bool hit = flat_any(sc, ray, active, overflow);
        finish(active, user, ray, hit ? 0u : 0xFFFFFFFFu, ray.t_max);
    }
}

}  // namespace cpbrt
