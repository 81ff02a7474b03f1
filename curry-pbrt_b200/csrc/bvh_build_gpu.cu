// Top levels of the exact full-sweep SAH build (BVHAggregate::create_sah_node,
// src/aggregate/bvh.rs:24-124) on the GPU.
//
// The reference sorts a node's primitives along each axis (stable, by bound centre), sweeps
// prefix / suffix bounds, takes the first minimum of the SAH cost over (axis, position), permutes
// the primitives into that axis' order and recurses.  On the host the top of the tree is the slow
// part: one node, three 10-million-element sorts, nothing to run in parallel.  Here the tree is
// built level by level for every segment of at least `min_len` primitives:
//     keys   (segment rank << 32 | order-preserving centre bits), one stable radix sort per axis
//            -- stability makes ties fall back to the current slot order, like the reference's
//            sort_by_cached_key on the node's current order
//     sweep  prefix and suffix bound unions as scans-by-key over the sorted order, SAH cost per
//            split position with the reference's f32 expression
//     pick   64-bit atomicMin of (ordered cost, axis, position): first minimum in the reference's
//            (axis, position) visiting order; a NaN in the very first cost sticks, as it does there
//     split  gather into the winning axis' order; children of >= min_len primitives stay on the
//            device for the next level, smaller ones are handed to the host builder
// Sorting, scanning and compaction are CUB device primitives; everything specific to the
// algorithm is the kernels below.  The host finishes the pending segments in parallel
// (finish_sah_bvh) and produces the same primitive order and topology as the pure host build
// (tests/test_oracle.py, tests/test_gpu_parity.py).
#include <cuda_runtime.h>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <chrono>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "bvh_build.h"

namespace cpbrt {

namespace {

struct Box { float mn[3], mx[3]; };
static_assert(sizeof(Box) == sizeof(BuildBounds), "layout");

__device__ __forceinline__ float dmin(float l, float r) { return r < l ? r : l; }   // src/math/mod.rs:36-49
__device__ __forceinline__ float dmax(float l, float r) { return r > l ? r : l; }
struct Unite {
    __device__ __forceinline__ Box operator()(const Box& a, const Box& b) const {
        Box r;
#pragma unroll
        for (int k = 0; k < 3; ++k) { r.mn[k] = dmin(a.mn[k], b.mn[k]); r.mx[k] = dmax(a.mx[k], b.mx[k]); }
        return r;
    }
};
__device__ __forceinline__ float surface_area(const Box& b) {   // bounds.rs:255-258
    const float dx = b.mx[0] - b.mn[0], dy = b.mx[1] - b.mn[1], dz = b.mx[2] - b.mn[2];
    return 2.f * (dx * dy + dx * dz + dy * dz);
}
// OrderedFloat order (NaN greatest, -0 == +0) as an unsigned integer
__device__ __forceinline__ uint32_t ordered_bits(float f) {
    if (f != f) return 0xFFFFFFFFu;
    if (f == 0.f) f = 0.f;
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float from_ordered_bits(uint32_t k) {
    if (k == 0xFFFFFFFFu) return __uint_as_float(0x7FC00000u);
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k);
}

// per-segment record of the current level
struct Segment {
    uint32_t slot0;      // first global BVH slot
    uint32_t rank0;      // first position in the working arrays
    uint32_t len;
    int32_t parent;      // top node that owns this segment (-1: root)
    uint32_t side;       // which child of the parent
};

__global__ void k_make_keys(const Box* __restrict__ pb, const uint32_t* __restrict__ seg, uint32_t m, int axis,
                            unsigned long long* __restrict__ keys, uint32_t* __restrict__ vals) {
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < m; r += gridDim.x * blockDim.x) {
        const Box b = pb[r];
        const float c = b.mn[axis] + (b.mx[axis] - b.mn[axis]) / 2.f;   // bounds.rs:70-72
        keys[r] = ((unsigned long long)seg[r] << 32) | ordered_bits(c);
        vals[r] = r;
    }
}
__global__ void k_gather_boxes(const Box* __restrict__ pb, const uint32_t* __restrict__ idx, uint32_t m, Box* __restrict__ fwd,
                               Box* __restrict__ rev, const uint32_t* __restrict__ seg, uint32_t* __restrict__ seg_rev) {
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < m; r += gridDim.x * blockDim.x) {
        const Box b = pb[idx[r]];
        fwd[r] = b; rev[m - 1 - r] = b;
        seg_rev[m - 1 - r] = seg[r];
    }
}
// SAH cost of every split position of every segment along `axis`, folded into the segment's best
// candidate: key = ordered(cost) << 32 | axis << 30 | position.  Consecutive lanes hold consecutive
// sorted positions, so a warp sees a few contiguous runs of equal segments: one segmented
// shuffle reduction, then one atomicMin per run.
__global__ void k_costs(const Box* __restrict__ prefix, const Box* __restrict__ suffix_rev, const uint32_t* __restrict__ seg,
                        const Segment* __restrict__ segs, uint32_t m, int axis, unsigned long long* __restrict__ best,
                        uint32_t* __restrict__ first_is_nan) {
    const int lane = threadIdx.x & 31;
    const uint32_t m_round = (m + 31u) & ~31u;
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < m_round; r += gridDim.x * blockDim.x) {
        uint32_t s = 0xFFFFFFFFu;
        unsigned long long v = ~0ull;
        if (r < m) {
            s = seg[r];
            const Segment sg = segs[s];
            const uint32_t i = r - sg.rank0;
            if (i + 1 < sg.len) {
                const float full_sa = surface_area(prefix[sg.rank0 + sg.len - 1]);
                const float sa_f = surface_area(prefix[r]);
                const float sa_b = surface_area(suffix_rev[m - 1 - (r + 1)]);
                const float cost = 0.125f + ((float)(i + 1) * sa_f + (float)(sg.len - 1 - i) * sa_b) / full_sa;   // bvh.rs:70-74
                if (axis == 0 && i == 0 && cost != cost) first_is_nan[s] = 1u;
                v = ((unsigned long long)ordered_bits(cost) << 32) | ((unsigned long long)axis << 30) | i;
            }
        }
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const unsigned long long o = __shfl_down_sync(0xFFFFFFFFu, v, off);
            const uint32_t os = __shfl_down_sync(0xFFFFFFFFu, s, off);
            if (lane + off < 32 && os == s && o < v) v = o;
        }
        const uint32_t prev = __shfl_up_sync(0xFFFFFFFFu, s, 1);
        if (s != 0xFFFFFFFFu && (lane == 0 || prev != s) && v != ~0ull) atomicMin(&best[s], v);
    }
}

struct Decision { uint32_t axis, split, leaf; };

// One thread per segment: leaf or split, node creation, children for the next level / the host.
__global__ void k_decide(const Segment* __restrict__ segs, uint32_t n_segs, const unsigned long long* __restrict__ best,
                         const uint32_t* __restrict__ first_is_nan, const Box* __restrict__ prefix_axis0, uint32_t min_len,
                         Decision* __restrict__ dec, TopNode* __restrict__ nodes, uint32_t* __restrict__ n_nodes, uint32_t node_cap,
                         PendingSegment* __restrict__ pending, int32_t* __restrict__ pending_parent, uint32_t* __restrict__ pending_side,
                         uint32_t* __restrict__ n_pending, uint32_t pending_cap, uint32_t* __restrict__ child_count, int32_t* __restrict__ root_ref,
                         uint32_t* __restrict__ overflow) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_segs) return;
    const Segment sg = segs[s];
    unsigned long long key = best[s];
    float min_cost = from_ordered_bits((uint32_t)(key >> 32));
    uint32_t axis = (uint32_t)((key >> 30) & 3u), split = (uint32_t)(key & 0x3FFFFFFFu);
    if (first_is_nan[s]) { min_cost = __uint_as_float(0x7FC00000u); axis = 0; split = 0; }   // the seed never gets replaced (bvh.rs:77-88)
    Decision d; d.axis = axis; d.split = split; d.leaf = (min_cost > (float)sg.len) ? 1u : 0u;   // bvh.rs:92
    dec[s] = d;
    auto add_pending = [&](uint32_t a, uint32_t b, uint32_t force_leaf, int32_t parent, uint32_t side) {
        const uint32_t p = atomicAdd(n_pending, 1u);
        if (p >= pending_cap) { *overflow = 1u; return; }
        pending[p].start = a; pending[p].end = b; pending[p].force_leaf = force_leaf;
        pending_parent[p] = parent; pending_side[p] = side;
    };
    if (d.leaf) {
        add_pending(sg.slot0, sg.slot0 + sg.len, 1u, sg.parent, sg.side);
        child_count[s] = 0;
        return;
    }
    const uint32_t id = atomicAdd(n_nodes, 1u);
    if (id >= node_cap) { *overflow = 1u; child_count[s] = 0; return; }
    TopNode nd;
    nd.axis = (uint8_t)axis;
    const Box full = prefix_axis0[sg.rank0 + sg.len - 1];   // union of the whole segment (any axis' last prefix)
    for (int k = 0; k < 3; ++k) { nd.bound.mn[k] = full.mn[k]; nd.bound.mx[k] = full.mx[k]; }
    nd.child[0] = nd.child[1] = 0x7FFFFFFF;
    nodes[id] = nd;
    if (sg.parent < 0) *root_ref = (int32_t)id; else nodes[sg.parent].child[sg.side] = (int32_t)id;
    const uint32_t llen = split + 1, rlen = sg.len - llen;
    uint32_t cont = 0;
    if (llen >= min_len) cont |= 1u; else add_pending(sg.slot0, sg.slot0 + llen, 0u, (int32_t)id, 0u);
    if (rlen >= min_len) cont |= 2u; else add_pending(sg.slot0 + llen, sg.slot0 + sg.len, 0u, (int32_t)id, 1u);
    child_count[s] = (cont & 1u) + ((cont >> 1) & 1u);
    // remember who owns the continuing children: encoded in dec (leaf == 0), node id via a side table
    dec[s].leaf = 0u | (id << 1);     // bit 0 = leaf flag, upper bits = node id of this segment
}

// Permute every segment into its winning order, write the finished part out and flag what continues.
__global__ void k_split(const Box* __restrict__ pb, const uint32_t* __restrict__ order, const uint32_t* __restrict__ seg,
                        const Segment* __restrict__ segs, const Decision* __restrict__ dec, const uint32_t* __restrict__ idx0,
                        const uint32_t* __restrict__ idx1, const uint32_t* __restrict__ idx2, uint32_t m, uint32_t min_len,
                        Box* __restrict__ pb_new, uint32_t* __restrict__ order_new, uint32_t* __restrict__ keep,
                        Box* __restrict__ pb_final, uint32_t* __restrict__ order_final) {
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < m; r += gridDim.x * blockDim.x) {
        const uint32_t s = seg[r];
        const Segment sg = segs[s];
        const Decision d = dec[s];
        const bool leaf = d.leaf & 1u;
        uint32_t src = r;
        if (!leaf) src = d.axis == 0 ? idx0[r] : (d.axis == 1 ? idx1[r] : idx2[r]);
        const Box b = pb[src]; const uint32_t o = order[src];
        pb_new[r] = b; order_new[r] = o;
        const uint32_t i = r - sg.rank0;
        bool stays = false;
        if (!leaf) {
            const uint32_t llen = d.split + 1;
            stays = i < llen ? (llen >= min_len) : (sg.len - llen >= min_len);
        }
        keep[r] = stays ? 1u : 0u;
        if (!stays) { pb_final[sg.slot0 + i] = b; order_final[sg.slot0 + i] = o; }
    }
}

// Next level's segment table (one thread per old segment, children in slot order).
__global__ void k_next_segments(const Segment* __restrict__ segs, uint32_t n_segs, const Decision* __restrict__ dec,
                                const uint32_t* __restrict__ child_offset, const uint32_t* __restrict__ keep_offset, uint32_t min_len,
                                Segment* __restrict__ next) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_segs) return;
    const Decision d = dec[s];
    if (d.leaf & 1u) return;
    const Segment sg = segs[s];
    const int32_t id = (int32_t)(d.leaf >> 1);
    const uint32_t llen = d.split + 1, rlen = sg.len - llen;
    uint32_t c = child_offset[s];
    if (llen >= min_len) { Segment n; n.slot0 = sg.slot0; n.rank0 = keep_offset[sg.rank0]; n.len = llen; n.parent = id; n.side = 0; next[c++] = n; }
    if (rlen >= min_len) { Segment n; n.slot0 = sg.slot0 + llen; n.rank0 = keep_offset[sg.rank0 + llen]; n.len = rlen; n.parent = id; n.side = 1; next[c++] = n; }
}
__global__ void k_compact(const Box* __restrict__ pb_new, const uint32_t* __restrict__ order_new, const uint32_t* __restrict__ keep,
                          const uint32_t* __restrict__ keep_offset, uint32_t m, Box* __restrict__ pb_out, uint32_t* __restrict__ order_out) {
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < m; r += gridDim.x * blockDim.x) {
        if (!keep[r]) continue;
        const uint32_t q = keep_offset[r];
        pb_out[q] = pb_new[r]; order_out[q] = order_new[r];
    }
}
__global__ void k_segment_ids(const Segment* __restrict__ segs, uint32_t n_segs, uint32_t* __restrict__ seg) {
    // one block per segment would be wasteful for few huge segments: every thread finds its segment by binary search
    const uint32_t total = segs[n_segs - 1].rank0 + segs[n_segs - 1].len;
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < total; r += gridDim.x * blockDim.x) {
        uint32_t lo = 0, hi = n_segs;
        while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (segs[mid].rank0 <= r) lo = mid; else hi = mid; }
        seg[r] = lo;
    }
}


// ------------------------------------------------------------------------------------------------
// Small segments: everything below `min_len` primitives, one block per segment, the whole subtree in
// shared memory.  Same algorithm as the host builder's Builder::build (bvh_build.cpp) node by node:
//     keys    (ordered centre bits << 32 | local index) per axis; ranks by counting (keys are unique, so the
//             rank IS the stable sorted position -- ties fall back to the current order like the reference's
//             stable sort_by_cached_key)
//     sweep   one thread per axis: suffix areas from the back, then prefix box + cost from the front with the
//             reference's f32 expression (bvh.rs:70-74)
//     pick    one thread walks the 3 x (len - 1) costs in the reference's (axis, position) order: the first
//             cost seeds the minimum (even a NaN), afterwards only a strictly smaller one replaces it
//     split   gather boxes / order into the winning axis' order, push the right child, continue with the left
// Nodes go to a block-private range of the global node array (reserved with one atomicAdd: at most 2 len - 1);
// children are written as indices into that array.  A subtree's nodes are contiguous, which is also what the
// traversal wants.  The segment's root index goes to seg_root[k].
static constexpr int kSmallMax = 256;          // segment length the kernel handles (shared-memory arrays)
static constexpr int kSmallThreads = 128;
struct SmallNode { uint32_t leaf_axis; uint32_t a, b; float mn[3], mx[3]; };   // leaf_axis: bit 0 = leaf, bits 8.. = axis
static_assert(sizeof(SmallNode) == 36, "layout");

__global__ void __launch_bounds__(kSmallThreads) k_build_small(Box* __restrict__ pb_final, uint32_t* __restrict__ order_final,
                                                               const PendingSegment* __restrict__ pending, uint32_t n_pending,
                                                               SmallNode* __restrict__ nodes, uint32_t node_cap, uint32_t* __restrict__ n_nodes,
                                                               uint32_t* __restrict__ seg_root, uint32_t* __restrict__ overflow) {
    __shared__ Box box[kSmallMax], tmp_box[kSmallMax];
    __shared__ uint32_t ord[kSmallMax], tmp_ord[kSmallMax];
    __shared__ unsigned long long keys[kSmallMax];
    __shared__ uint16_t idx[3][kSmallMax];
    __shared__ float back_sa[3][kSmallMax], cost[3][kSmallMax];
    __shared__ uint32_t stk_start[kSmallMax], stk_len[kSmallMax], stk_link[kSmallMax];   // link: node index << 1 | side, 0xFFFFFFFF = segment root
    __shared__ uint32_t sh_base, sh_used, sh_sp, sh_decision[3];   // decision: leaf flag, axis, split
    const int tid = threadIdx.x;
    for (uint32_t k = blockIdx.x; k < n_pending; k += gridDim.x) {
        const PendingSegment seg = pending[k];
        const uint32_t len0 = seg.end - seg.start;
        if (seg.force_leaf != 0u && len0 > 0) {      // the SAH said "leaf" upstairs: one node, whatever the length
            if (tid == 0) {
                const uint32_t at = atomicAdd(n_nodes, 1u);
                if (at + 1u > node_cap) { *overflow = 1u; seg_root[k] = 0xFFFFFFFFu; }
                else {
                    SmallNode nd; nd.leaf_axis = 1u; nd.a = seg.start; nd.b = seg.end;
                    for (int c = 0; c < 3; ++c) { nd.mn[c] = 0.f; nd.mx[c] = 0.f; }
                    nodes[at] = nd; seg_root[k] = at;
                }
            }
            continue;
        }
        if (len0 > (uint32_t)kSmallMax || len0 == 0) { if (tid == 0) seg_root[k] = 0xFFFFFFFFu; continue; }   // left to the host builder
        __syncthreads();
        if (tid == 0) {
            sh_base = atomicAdd(n_nodes, 2u * len0 - 1u);
            sh_used = 0; sh_sp = 0;
            if (sh_base + 2u * len0 - 1u > node_cap) { *overflow = 1u; sh_base = 0xFFFFFFFFu; }
        }
        for (uint32_t i = tid; i < len0; i += kSmallThreads) { box[i] = pb_final[seg.start + i]; ord[i] = order_final[seg.start + i]; }
        __syncthreads();
        if (sh_base == 0xFFFFFFFFu) { if (tid == 0) seg_root[k] = 0xFFFFFFFFu; continue; }
        const uint32_t base = sh_base;
        // explicit recursion: (start, len, link) items
        uint32_t start = 0, len = len0, link = 0xFFFFFFFFu;
        bool have = true;
        while (have) {
            // ---- node [start, start + len) -------------------------------------------------------
            bool leaf = seg.force_leaf != 0u || len <= 1u;
            uint32_t axis = 0, split = 0;
            Box full = box[start];
            if (!leaf) {
                for (int a = 0; a < 3; ++a) {
                    for (uint32_t i = tid; i < len; i += kSmallThreads) {
                        const Box b = box[start + i];
                        const float c = b.mn[a] + (b.mx[a] - b.mn[a]) / 2.f;   // bounds.rs:70-72
                        keys[i] = ((unsigned long long)ordered_bits(c) << 32) | i;
                    }
                    __syncthreads();
                    for (uint32_t i = tid; i < len; i += kSmallThreads) {
                        const unsigned long long ki = keys[i];
                        uint32_t rank = 0;
                        for (uint32_t j = 0; j < len; ++j) rank += keys[j] < ki ? 1u : 0u;
                        idx[a][rank] = (uint16_t)i;
                    }
                    __syncthreads();
                }
                if (tid < 3) {      // one axis each: suffix areas, then prefix boxes and costs (sweep_axis of bvh_build.cpp)
                    const int a = tid;
                    Box acc = box[start + idx[a][len - 1]];
                    back_sa[a][len - 2] = surface_area(acc);
                    for (uint32_t i = len - 2; i-- > 0;) { acc = Unite()(acc, box[start + idx[a][i + 1]]); back_sa[a][i] = surface_area(acc); }
                    // the union of the whole node (same on every axis: min / max are exact) -- the host unites in slot order
                    Box all = box[start];
                    for (uint32_t i = 1; i < len; ++i) all = Unite()(all, box[start + i]);
                    const float full_sa = surface_area(all);
                    acc = box[start + idx[a][0]];
                    for (uint32_t i = 0; i + 1 < len; ++i) {
                        if (i) acc = Unite()(acc, box[start + idx[a][i]]);
                        cost[a][i] = 0.125f + ((float)(i + 1) * surface_area(acc) + (float)(len - 1 - i) * back_sa[a][i]) / full_sa;   // bvh.rs:70-74
                    }
                }
                __syncthreads();
                if (tid == 0) {     // bvh.rs:77-92
                    bool seeded = false; float min_cost = 0.f; uint32_t min_axis = 0, min_i = 0;
                    for (uint32_t a = 0; a < 3; ++a)
                        for (uint32_t i = 0; i + 1 < len; ++i) {
                            const float c = cost[a][i];
                            if (!seeded) { seeded = true; min_cost = c; min_axis = a; min_i = i; }
                            else if (c < min_cost) { min_cost = c; min_axis = a; min_i = i; }
                        }
                    sh_decision[0] = (min_cost > (float)len) ? 1u : 0u; sh_decision[1] = min_axis; sh_decision[2] = min_i;
                }
                __syncthreads();
                leaf = sh_decision[0] != 0u; axis = sh_decision[1]; split = sh_decision[2];
                for (uint32_t i = 1; i < len; ++i) full = Unite()(full, box[start + i]);     // every thread: the Parent's bound
            }
            // ---- emit the node -------------------------------------------------------------------
            __syncthreads();
            const uint32_t id = sh_used;           // node index within the block's range
            __syncthreads();                       // every thread has read it before thread 0 advances it
            if (tid == 0) {
                sh_used = id + 1;
                SmallNode nd;
                if (leaf) { nd.leaf_axis = 1u; nd.a = seg.start + start; nd.b = seg.start + start + len; for (int c = 0; c < 3; ++c) { nd.mn[c] = 0.f; nd.mx[c] = 0.f; } }
                else { nd.leaf_axis = axis << 8; nd.a = 0xFFFFFFFFu; nd.b = 0xFFFFFFFFu; for (int c = 0; c < 3; ++c) { nd.mn[c] = full.mn[c]; nd.mx[c] = full.mx[c]; } }
                nodes[base + id] = nd;
                if (link == 0xFFFFFFFFu) seg_root[k] = base + id;
                else if (link & 1u) nodes[base + (link >> 1)].b = base + id;
                else nodes[base + (link >> 1)].a = base + id;
            }
            if (!leaf) {
                // bvh.rs:100-112: slot j receives the primitive that was at local index idx[axis][j]
                for (uint32_t j = tid; j < len; j += kSmallThreads) { const uint32_t src = start + idx[axis][j]; tmp_box[j] = box[src]; tmp_ord[j] = ord[src]; }
                __syncthreads();
                for (uint32_t j = tid; j < len; j += kSmallThreads) { box[start + j] = tmp_box[j]; ord[start + j] = tmp_ord[j]; }
                if (tid == 0) {     // right child waits on the stack, the left one is next
                    const uint32_t sp = sh_sp;
                    stk_start[sp] = start + split + 1; stk_len[sp] = len - split - 1; stk_link[sp] = (id << 1) | 1u;
                    sh_sp = sp + 1;
                }
                __syncthreads();
                len = split + 1; link = id << 1;      // start unchanged
            } else {
                __syncthreads();
                if (sh_sp == 0) have = false;
                else {
                    const uint32_t sp = sh_sp - 1;
                    start = stk_start[sp]; len = stk_len[sp]; link = stk_link[sp];
                    __syncthreads();
                    if (tid == 0) sh_sp = sp;
                }
            }
            __syncthreads();
        }
        for (uint32_t i = tid; i < len0; i += kSmallThreads) { order_final[seg.start + i] = ord[i]; pb_final[seg.start + i] = box[i]; }
    }
}


// ------------------------------------------------------------------------------------------------
// Device-resident build (big scenes): bounds in, finished device layout out, nothing but the primitive order
// and a handful of counters ever crosses PCIe.
//
// Primitive::bound() per uploaded primitive, the same f32 operations as primitive_bounds() in bvh_build.cpp
// (triangle.rs:114-119, sphere.rs:22-27, TransformShape::bound transform.rs:52-54 with Bounds3f::apply
// bounds.rs:292-300: the eight corners through the matrix, homogeneous divide when w != 0).
struct DevXform { float m[16]; float m_inv[16]; int32_t outer; int32_t pad[3]; };      // == DevTransform of scene_dev.cuh
__global__ void k_prim_bounds(const cpbrt_primitive* __restrict__ prims, uint32_t n, const cpbrt_triangle* __restrict__ tris,
                              const float* __restrict__ radius, const DevXform* __restrict__ xf, Box* __restrict__ out, uint32_t* __restrict__ any_nan) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const cpbrt_shape sh = prims[i].shape;
        Box b;
        if (sh.kind == CPBRT_SHAPE_SPHERE) {
            const float r = radius[sh.data];
            for (int k = 0; k < 3; ++k) { b.mn[k] = dmin(-r, r); b.mx[k] = dmax(-r, r); }
        } else {
            const cpbrt_triangle& t = tris[sh.data];
            for (int k = 0; k < 3; ++k) {
                b.mn[k] = dmin(dmin(t.p[0][k], t.p[1][k]), t.p[2][k]);
                b.mx[k] = dmax(dmax(t.p[0][k], t.p[1][k]), t.p[2][k]);
            }
        }
        for (int32_t x = sh.xform; x >= 0; x = xf[x].outer) {
            const float* m = xf[x].m;
            Box r;
            for (int c = 0; c < 8; ++c) {
                const float p[3] = {(c & 1) ? b.mx[0] : b.mn[0], (c & 2) ? b.mx[1] : b.mn[1], (c & 4) ? b.mx[2] : b.mn[2]};
                float h[4], q[3];
                for (int j = 0; j < 4; ++j) h[j] = ((m[j * 4] * p[0] + m[j * 4 + 1] * p[1]) + m[j * 4 + 2] * p[2]) + m[j * 4 + 3] * 1.f;
                if (h[3] != 0.f) { q[0] = h[0] / h[3]; q[1] = h[1] / h[3]; q[2] = h[2] / h[3]; } else { q[0] = h[0]; q[1] = h[1]; q[2] = h[2]; }
                for (int k = 0; k < 3; ++k) {
                    if (c == 0) { r.mn[k] = q[k]; r.mx[k] = q[k]; }
                    else { r.mn[k] = dmin(r.mn[k], q[k]); r.mx[k] = dmax(r.mx[k], q[k]); }
                }
            }
            b = r;
        }
        bool nan = false;
        for (int k = 0; k < 3; ++k) nan = nan || b.mn[k] != b.mn[k] || b.mx[k] != b.mx[k];
        if (nan) *any_nan = 1u;      // the order of unions would matter: host build only
        out[i] = b;
    }
}
__global__ void k_iota(uint32_t* __restrict__ v, uint32_t n) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) v[i] = i;
}

// Device layout (scene_dev.cuh): Parent records {L.min.xyz L.max.x}{L.max.yz R.min.xy}{R.min.z R.max.xyz}{lref rref lcnt|axis<<30 rcnt}.
// Parent numbering: the level-synchronous ("top") Parents keep their ids 0 .. n_top-1, the Parents of the small
// subtrees follow in the order of the small-node array (a subtree's nodes are contiguous there).
// Child references while the tree is still in pieces: >= 0 a top node, <= -2 the small node -(c + 2).
__global__ void k_small_parent_flags(const SmallNode* __restrict__ small, uint32_t n_small, uint32_t* __restrict__ flag) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_small; i += gridDim.x * blockDim.x) flag[i] = (small[i].leaf_axis & 1u) ? 0u : 1u;
}
__global__ void k_link_pending(const int32_t* __restrict__ pending_parent, const uint32_t* __restrict__ pending_side, const uint32_t* __restrict__ seg_root,
                               uint32_t n_pending, TopNode* __restrict__ top) {
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n_pending; k += gridDim.x * blockDim.x)
        if (pending_parent[k] >= 0) top[pending_parent[k]].child[pending_side[k]] = -(int32_t)seg_root[k] - 2;
}
struct ChildRec { float mn[3], mx[3]; uint32_t ref, cnt; };
__device__ __forceinline__ ChildRec small_child(const SmallNode* __restrict__ small, const uint32_t* __restrict__ pidx, const Box* __restrict__ pb, uint32_t n_top, uint32_t s) {
    const SmallNode sn = small[s];
    ChildRec c;
    if (sn.leaf_axis & 1u) {       // Leaf: union of its primitives' bounds (only the two-phase traversal looks at it)
        Box u = pb[sn.a];
        for (uint32_t sl = sn.a + 1; sl < sn.b; ++sl) u = Unite()(u, pb[sl]);
        for (int k = 0; k < 3; ++k) { c.mn[k] = u.mn[k]; c.mx[k] = u.mx[k]; }
        c.ref = sn.a; c.cnt = sn.b - sn.a;
    } else {
        for (int k = 0; k < 3; ++k) { c.mn[k] = sn.mn[k]; c.mx[k] = sn.mx[k]; }
        c.ref = n_top + pidx[s]; c.cnt = 0;
    }
    return c;
}
__device__ __forceinline__ void write_parent(float4* __restrict__ out, uint32_t id, const ChildRec& l, const ChildRec& r, uint32_t axis) {
    out[4ull * id] = make_float4(l.mn[0], l.mn[1], l.mn[2], l.mx[0]);
    out[4ull * id + 1] = make_float4(l.mx[1], l.mx[2], r.mn[0], r.mn[1]);
    out[4ull * id + 2] = make_float4(r.mn[2], r.mx[0], r.mx[1], r.mx[2]);
    out[4ull * id + 3] = make_float4(__uint_as_float(l.ref), __uint_as_float(r.ref), __uint_as_float(l.cnt | (axis << 30)), __uint_as_float(r.cnt));
}
__global__ void k_layout_nodes(const TopNode* __restrict__ top, uint32_t n_top, const SmallNode* __restrict__ small, uint32_t n_small,
                               const uint32_t* __restrict__ pidx, const Box* __restrict__ pb, float4* __restrict__ out) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_top + n_small; i += gridDim.x * blockDim.x) {
        if (i < n_top) {
            const TopNode t = top[i];
            ChildRec c[2];
            for (int k = 0; k < 2; ++k) {
                if (t.child[k] >= 0) {
                    const TopNode ch = top[t.child[k]];
                    for (int j = 0; j < 3; ++j) { c[k].mn[j] = ch.bound.mn[j]; c[k].mx[j] = ch.bound.mx[j]; }
                    c[k].ref = (uint32_t)t.child[k]; c[k].cnt = 0;
                } else c[k] = small_child(small, pidx, pb, n_top, (uint32_t)(-(t.child[k] + 2)));
            }
            write_parent(out, i, c[0], c[1], t.axis);
        } else {
            const uint32_t s = i - n_top;
            const SmallNode sn = small[s];
            if (sn.leaf_axis & 1u) continue;
            const ChildRec l = small_child(small, pidx, pb, n_top, sn.a), r = small_child(small, pidx, pb, n_top, sn.b);
            write_parent(out, n_top + pidx[s], l, r, sn.leaf_axis >> 8);
        }
    }
}

// Primitive records, shading frames and uv triples in BVH order (scene_dev.cuh; the host versions are in
// cpbrt_api.cu: host_tri_frame etc. -- same f32 operations, IEEE division and square root on both sides).
__global__ void k_prim_records(const cpbrt_primitive* __restrict__ prims, const cpbrt_triangle* __restrict__ tris, const float* __restrict__ radius,
                               const uint32_t* __restrict__ order, uint32_t n, int any_uv, float4* __restrict__ rec, float4* __restrict__ frames,
                               float* __restrict__ uvs) {
    for (uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x; slot < n; slot += gridDim.x * blockDim.x) {
        const uint32_t up = order[slot];
        const cpbrt_primitive p = prims[up];
        uint32_t meta = ((uint32_t)(p.shape.xform + 1)) << 2;
        float4 q0 = make_float4(0, 0, 0, 0), q1 = q0, q2 = q0, f0 = q0, f1 = q0, f2 = q0;
        if (p.shape.kind == CPBRT_SHAPE_SPHERE) { meta |= 1u; q0.x = radius[p.shape.data]; }
        else {
            const cpbrt_triangle t = tris[p.shape.data];
            q0 = make_float4(t.p[0][0], t.p[0][1], t.p[0][2], t.p[1][0]);
            q1 = make_float4(t.p[1][1], t.p[1][2], t.p[2][0], t.p[2][1]);
            q2.x = t.p[2][2];
            if (any_uv) { meta |= 2u; for (int k = 0; k < 6; ++k) uvs[6ull * slot + k] = (&t.uv[0][0])[k]; }
            // Triangle::normal (triangle.rs:78-82) and the frame BSDF::new derives from it (bxdf/mod.rs:83-97)
            const float a[3] = {t.p[0][0] - t.p[2][0], t.p[0][1] - t.p[2][1], t.p[0][2] - t.p[2][2]};
            const float b[3] = {t.p[1][0] - t.p[2][0], t.p[1][1] - t.p[2][1], t.p[1][2] - t.p[2][2]};
            float c[3] = {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
            for (int pass = 0; pass < 2; ++pass) {
                const float len = sqrtf((c[0] * c[0] + c[1] * c[1]) + c[2] * c[2]);
                c[0] = c[0] / len; c[1] = c[1] / len; c[2] = c[2] / len;
            }
            float x[3], y[3];
            if (fabsf(c[0]) > fabsf(c[1])) { const float l = sqrtf(c[0] * c[0] + c[2] * c[2]); x[0] = -c[2] / l; x[1] = 0.f / l; x[2] = c[0] / l; }
            else { const float l = sqrtf(c[1] * c[1] + c[2] * c[2]); x[0] = 0.f / l; x[1] = c[2] / l; x[2] = -c[1] / l; }
            y[0] = c[1] * x[2] - c[2] * x[1]; y[1] = c[2] * x[0] - c[0] * x[2]; y[2] = c[0] * x[1] - c[1] * x[0];
            float nl[3] = {(c[0] * x[0] + c[1] * x[1]) + c[2] * x[2], (c[0] * y[0] + c[1] * y[1]) + c[2] * y[2], (c[0] * c[0] + c[1] * c[1]) + c[2] * c[2]};
            const float len = sqrtf((nl[0] * nl[0] + nl[1] * nl[1]) + nl[2] * nl[2]);
            nl[0] = nl[0] / len; nl[1] = nl[1] / len; nl[2] = nl[2] / len;
            f0 = make_float4(c[0], c[1], c[2], x[0]); f1 = make_float4(x[1], x[2], y[0], y[1]); f2 = make_float4(y[2], nl[0], nl[1], nl[2]);
        }
        q2.y = __uint_as_float(up); q2.z = __uint_as_float(meta);
        q2.w = __uint_as_float(p.material >= 0 ? (uint32_t)p.material : (0x80000000u | (uint32_t)p.light));
        rec[3ull * slot] = q0; rec[3ull * slot + 1] = q1; rec[3ull * slot + 2] = q2;
        frames[3ull * slot] = f0; frames[3ull * slot + 1] = f1; frames[3ull * slot + 2] = f2;
    }
}

// Device scratch of one build: one slab (a single cudaMalloc -- dozens of separate multi-hundred-megabyte
// allocations cost anything between 15 and 500 ms on the same box) handed out in 256-byte aligned pieces;
// whatever does not fit falls back to its own allocation.
struct DevMem {
    std::vector<void*> all;
    char* slab = nullptr; size_t slab_bytes = 0, used = 0;
    ~DevMem() { for (void* p : all) cudaFree(p); cudaFree(slab); }
    bool reserve(size_t bytes) {
        if (slab) return true;
        if (cudaMalloc((void**)&slab, bytes) != cudaSuccess) { cudaGetLastError(); slab = nullptr; return false; }
        slab_bytes = bytes; used = 0;
        return true;
    }
    template <class T> T* get(size_t n) {
        const size_t bytes = (std::max<size_t>(n, 1) * sizeof(T) + 255) & ~(size_t)255;
        if (slab && used + bytes <= slab_bytes) { T* p = (T*)(slab + used); used += bytes; return p; }
        void* p = nullptr;
        if (cudaMalloc(&p, bytes) != cudaSuccess) return nullptr;
        all.push_back(p);
        return (T*)p;
    }
};

#define CK(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) { *err = std::string(#x) + ": " + cudaGetErrorString(e__); return false; } } while (0)

// ------------------------------------------------------------------------------------------------
// One build's device working set.  run() performs the level-synchronous phase and the small-segment phase and
// leaves everything on the device; the two front ends below either download the pieces for the host stitcher
// (gpu_sah_top_phase: node dumps, tests) or finish the traversal layout on the device (gpu_sah_build_resident).
struct DeviceSah {
    DevMem M;
    cudaStream_t stream = nullptr;
    uint32_t n = 0, min_len = 2;
    int max_levels = 48;
    bool timing = false;
    double t_mark = 0.0;
    // results (device)
    Box* pb_final = nullptr; uint32_t* order_final = nullptr;
    TopNode* nodes = nullptr; PendingSegment* pending = nullptr; int32_t* pending_parent = nullptr; uint32_t* pending_side = nullptr;
    uint32_t* counters = nullptr;      // [0] nodes, [1] pending, [2] overflow, [3] root_ref (as int), [4] small nodes, [5] small overflow
    SmallNode* small = nullptr; uint32_t* seg_root = nullptr;
    Box* pb_in = nullptr;              // primitive bounds in upload order: filled by the caller before run()
    // results (host)
    uint32_t nn = 0, np = 0, n_small = 0;
    int levels = 0; int32_t root = 0x7FFFFFFF;
    std::vector<Segment> left;         // segments still active when the level budget ran out
    bool small_ran = false;

    static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
    void lap(const char* what) { if (timing) { cudaStreamSynchronize(stream); const double t = now(); fprintf(stderr, "[gpu_sah] %-28s %.3f s\n", what, t - t_mark); t_mark = t; } }
    bool init(uint32_t n_, uint32_t min_len_, int max_levels_, cudaStream_t st, std::string* err, size_t extra_bytes = 0) {
        n = n_; min_len = std::max(2u, min_len_); max_levels = max_levels_; stream = st;
        timing = getenv("CPBRT_TIMING") != nullptr; t_mark = now();
        if (n < 2 || n < min_len || n >= 0x3FFFFFFFu) { *err = "nothing to do on the device"; return false; }
        // 9 box arrays, 12 index arrays, 2 key arrays, CUB scratch and the per-segment tables; extra_bytes: what the
        // front end keeps next to them (the uploaded scene arrays of the resident build)
        M.reserve((size_t)n * (9 * sizeof(Box) + 12 * 4 + 2 * 8 + 8) + extra_bytes + (64u << 20));
        pb_in = M.get<Box>(n);
        if (!pb_in) { *err = "device memory for the BVH build"; return false; }
        return true;
    }
    bool run(bool small_phase, std::string* err);
};

bool DeviceSah::run(bool small_phase, std::string* err) {
    const uint32_t max_segs = n / min_len + 2;
    const uint32_t node_cap = 2 * max_segs * (uint32_t)std::max(1, max_levels) + 16, pending_cap = node_cap + 16;
    // fwd, rev and prefix are one piece: after the level-synchronous phase the small-subtree nodes live there
    // (2 n nodes of 36 bytes = 3 n boxes of 24 bytes)
    static_assert(2 * sizeof(SmallNode) == 3 * sizeof(Box), "the small nodes reuse three box arrays");
    Box* scan3 = M.get<Box>(3ull * n + 32);
    Box *pb = pb_in, *pb_new = M.get<Box>(n), *pb_alt = M.get<Box>(n), *fwd = scan3, *rev = scan3 ? scan3 + n : nullptr, *prefix = scan3 ? scan3 + 2ull * n : nullptr,
        *suffix_rev = M.get<Box>(n), *prefix0 = M.get<Box>(n);
    pb_final = M.get<Box>(n);
    uint32_t *order = M.get<uint32_t>(n), *order_new = M.get<uint32_t>(n), *order_alt = M.get<uint32_t>(n), *seg = M.get<uint32_t>(n),
             *seg_rev = M.get<uint32_t>(n), *vals = M.get<uint32_t>(n), *keep = M.get<uint32_t>(n + 1), *keep_off = M.get<uint32_t>(n + 1);
    order_final = M.get<uint32_t>(n);
    uint32_t* idx[3] = {M.get<uint32_t>(n), M.get<uint32_t>(n), M.get<uint32_t>(n)};
    unsigned long long *keys = M.get<unsigned long long>(n), *keys_out = M.get<unsigned long long>(n), *best = M.get<unsigned long long>(max_segs);
    Segment *segs = M.get<Segment>(max_segs), *segs_next = M.get<Segment>(max_segs);
    Decision* dec = M.get<Decision>(max_segs);
    uint32_t *first_nan = M.get<uint32_t>(max_segs), *child_count = M.get<uint32_t>(max_segs + 1), *child_off = M.get<uint32_t>(max_segs + 1);
    nodes = M.get<TopNode>(node_cap);
    pending = M.get<PendingSegment>(pending_cap);
    pending_parent = M.get<int32_t>(pending_cap); pending_side = M.get<uint32_t>(pending_cap);
    counters = M.get<uint32_t>(8);
    if (!pb || !pb_new || !pb_alt || !fwd || !rev || !prefix || !suffix_rev || !prefix0 || !pb_final || !order || !order_new || !order_alt || !order_final || !seg ||
        !seg_rev || !vals || !keep || !keep_off || !idx[0] || !idx[1] || !idx[2] || !keys || !keys_out || !best || !segs || !segs_next || !dec || !first_nan ||
        !child_count || !child_off || !nodes || !pending || !pending_parent || !pending_side || !counters) { *err = "device memory for the BVH build"; return false; }

    // CUB scratch: size it once for the largest problem
    size_t tmp_bytes = 0, need = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, need, keys, keys_out, vals, idx[0], (int)n, 0, 64, stream); tmp_bytes = std::max(tmp_bytes, need);
    cub::DeviceScan::InclusiveScanByKey(nullptr, need, seg, fwd, prefix, Unite(), (int)n, cub::Equality(), stream); tmp_bytes = std::max(tmp_bytes, need);
    cub::DeviceScan::ExclusiveSum(nullptr, need, keep, keep_off, (int)n + 1, stream); tmp_bytes = std::max(tmp_bytes, need);
    void* tmp = M.get<char>(tmp_bytes);
    if (!tmp) { *err = "device memory for the BVH build"; return false; }

    const int T = 256;
    auto grid = [&](uint32_t items) { return (int)std::min<uint32_t>((items + T - 1) / T, 148u * 16u); };
    k_iota<<<grid(n), T, 0, stream>>>(order, n);
    CK(cudaMemsetAsync(counters, 0, 8 * 4, stream));
    {
        int32_t none = 0x7FFFFFFF;
        CK(cudaMemcpyAsync(counters + 3, &none, 4, cudaMemcpyHostToDevice, stream));
        Segment root_seg; root_seg.slot0 = 0; root_seg.rank0 = 0; root_seg.len = n; root_seg.parent = -1; root_seg.side = 0;
        CK(cudaMemcpyAsync(segs, &root_seg, sizeof root_seg, cudaMemcpyHostToDevice, stream));
        CK(cudaStreamSynchronize(stream));      // the two sources above are locals
    }
    lap("allocations + setup");
    uint32_t m = n, n_segs = 1;
    int level = 0;
    for (; level < max_levels && n_segs > 0; ++level) {
        k_segment_ids<<<grid(m), T, 0, stream>>>(segs, n_segs, seg);
        CK(cudaMemsetAsync(best, 0xFF, (size_t)n_segs * 8, stream));
        CK(cudaMemsetAsync(first_nan, 0, (size_t)n_segs * 4, stream));
        int seg_bits = 1; while ((1u << seg_bits) < n_segs) ++seg_bits;
        for (int axis = 0; axis < 3; ++axis) {
            k_make_keys<<<grid(m), T, 0, stream>>>(pb, seg, m, axis, keys, vals);
            need = tmp_bytes;
            CK(cub::DeviceRadixSort::SortPairs(tmp, need, keys, keys_out, vals, idx[axis], (int)m, 0, 32 + seg_bits, stream));
            k_gather_boxes<<<grid(m), T, 0, stream>>>(pb, idx[axis], m, fwd, rev, seg, seg_rev);
            Box* pre = axis == 0 ? prefix0 : prefix;
            need = tmp_bytes;
            CK(cub::DeviceScan::InclusiveScanByKey(tmp, need, seg, fwd, pre, Unite(), (int)m, cub::Equality(), stream));
            need = tmp_bytes;
            CK(cub::DeviceScan::InclusiveScanByKey(tmp, need, seg_rev, rev, suffix_rev, Unite(), (int)m, cub::Equality(), stream));
            k_costs<<<grid(m), T, 0, stream>>>(pre, suffix_rev, seg, segs, m, axis, best, first_nan);
        }
        k_decide<<<(n_segs + 127) / 128, 128, 0, stream>>>(segs, n_segs, best, first_nan, prefix0, min_len, dec, nodes, counters + 0, node_cap, pending,
                                                           pending_parent, pending_side, counters + 1, pending_cap, child_count, (int32_t*)(counters + 3),
                                                           counters + 2);
        k_split<<<grid(m), T, 0, stream>>>(pb, order, seg, segs, dec, idx[0], idx[1], idx[2], m, min_len, pb_new, order_new, keep, pb_final, order_final);
        need = tmp_bytes;
        CK(cudaMemsetAsync(keep + m, 0, 4, stream));
        CK(cub::DeviceScan::ExclusiveSum(tmp, need, keep, keep_off, (int)m + 1, stream));
        CK(cudaMemsetAsync(child_count + n_segs, 0, 4, stream));
        need = tmp_bytes;
        CK(cub::DeviceScan::ExclusiveSum(tmp, need, child_count, child_off, (int)n_segs + 1, stream));
        k_next_segments<<<(n_segs + 127) / 128, 128, 0, stream>>>(segs, n_segs, dec, child_off, keep_off, min_len, segs_next);
        k_compact<<<grid(m), T, 0, stream>>>(pb_new, order_new, keep, keep_off, m, pb_alt, order_alt);
        uint32_t m_next = 0, s_next = 0, h_counters[4];
        CK(cudaMemcpyAsync(&m_next, keep_off + m, 4, cudaMemcpyDeviceToHost, stream));
        CK(cudaMemcpyAsync(&s_next, child_off + n_segs, 4, cudaMemcpyDeviceToHost, stream));
        CK(cudaMemcpyAsync(h_counters, counters, 16, cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
        if (h_counters[2]) { *err = "BVH top-phase tables overflowed"; return false; }
        std::swap(pb, pb_alt); std::swap(order, order_alt); std::swap(segs, segs_next);
        m = m_next; n_segs = s_next;
    }
    levels = level;
    lap("level-synchronous phase");
    // segments still active when the level budget ran out go to the host as they are
    left.assign(n_segs, Segment());
    if (n_segs) {
        CK(cudaMemcpyAsync(left.data(), segs, (size_t)n_segs * sizeof(Segment), cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
        // their primitives are in the working arrays: move them to their final slots
        for (const Segment& sg : left) {
            CK(cudaMemcpyAsync(pb_final + sg.slot0, pb + sg.rank0, (size_t)sg.len * sizeof(Box), cudaMemcpyDeviceToDevice, stream));
            CK(cudaMemcpyAsync(order_final + sg.slot0, order + sg.rank0, (size_t)sg.len * 4, cudaMemcpyDeviceToDevice, stream));
        }
    }
    uint32_t h_counters[4];
    CK(cudaMemcpyAsync(h_counters, counters, 16, cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    nn = h_counters[0]; np = h_counters[1]; root = (int32_t)h_counters[3];
    // ---- small phase: every pending segment of up to kSmallMax primitives, one block each ---------------------
    small_ran = false; n_small = 0;
    if (small_phase && np > 0) {
        const uint32_t small_cap = 2u * n;
        small = reinterpret_cast<SmallNode*>(scan3);
        seg_root = M.get<uint32_t>(np);
        if (!small || !seg_root) { *err = "device memory for the BVH build"; return false; }
        CK(cudaMemsetAsync(counters + 4, 0, 8, stream));
        // a block reserves 2 len - 1 nodes and may use fewer (multi-primitive Leaves): what stays unused reads as a Leaf
        CK(cudaMemsetAsync(small, 0xFF, (size_t)small_cap * sizeof(SmallNode), stream));
        k_build_small<<<(int)std::min<uint32_t>(np, 148u * 32u), kSmallThreads, 0, stream>>>(pb_final, order_final, pending, np, small, small_cap, counters + 4, seg_root, counters + 5);
        uint32_t used[2];
        CK(cudaMemcpyAsync(used, counters + 4, 8, cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
        lap("small-segment kernel");
        if (used[1]) { *err = "BVH small-phase node table overflowed"; return false; }
        n_small = used[0]; small_ran = true;
    }
    return true;
}

}  // namespace

bool gpu_sah_top_phase(const std::vector<BuildBounds>& bounds, uint32_t min_len, int max_levels, cudaStream_t stream, TopPhase* out, std::string* err) {
    const uint32_t n = (uint32_t)bounds.size();
    out->pb.clear(); out->order.clear(); out->nodes.clear(); out->pending.clear(); out->pending_parent.clear(); out->pending_side.clear();
    out->small_nodes.clear(); out->pending_root.clear();
    out->root = 0x7FFFFFFF; out->levels = 0;
    for (const BuildBounds& b : bounds)
        for (int k = 0; k < 3; ++k) if (b.mn[k] != b.mn[k] || b.mx[k] != b.mx[k]) { *err = "NaN bounds: the order of unions matters, host build only"; return false; }
    DeviceSah D;
    if (!D.init(n, min_len, max_levels, stream, err)) return false;
    CK(cudaMemcpyAsync(D.pb_in, bounds.data(), (size_t)n * sizeof(Box), cudaMemcpyHostToDevice, stream));
    const char* e = getenv("CPBRT_BVH_SMALL");
    if (!D.run(!e || atoi(e) != 0, err)) return false;
    const uint32_t nn = D.nn, np = D.np;
    bool host_work = !D.left.empty() || !D.small_ran;
    if (D.small_ran) {
        out->small_nodes.resize(D.n_small); out->pending_root.resize(np);
        static_assert(sizeof(SmallNode) == sizeof(SmallSubtreeNode), "layout");
        if (D.n_small) CK(cudaMemcpyAsync(out->small_nodes.data(), D.small, (size_t)D.n_small * sizeof(SmallNode), cudaMemcpyDeviceToHost, stream));
        CK(cudaMemcpyAsync(out->pending_root.data(), D.seg_root, (size_t)np * 4, cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
        for (uint32_t k = 0; k < np && !host_work; ++k) host_work = out->pending_root[k] == 0xFFFFFFFFu;
    }
    out->pb.resize(host_work ? n : 0); out->order.resize(n); out->nodes.resize(nn); out->pending.resize(np); out->pending_parent.resize(np); out->pending_side.resize(np);
    if (host_work) CK(cudaMemcpyAsync(out->pb.data(), D.pb_final, (size_t)n * sizeof(Box), cudaMemcpyDeviceToHost, stream));
    CK(cudaMemcpyAsync(out->order.data(), D.order_final, (size_t)n * 4, cudaMemcpyDeviceToHost, stream));
    if (nn) CK(cudaMemcpyAsync(out->nodes.data(), D.nodes, (size_t)nn * sizeof(TopNode), cudaMemcpyDeviceToHost, stream));
    if (np) {
        CK(cudaMemcpyAsync(out->pending.data(), D.pending, (size_t)np * sizeof(PendingSegment), cudaMemcpyDeviceToHost, stream));
        CK(cudaMemcpyAsync(out->pending_parent.data(), D.pending_parent, (size_t)np * 4, cudaMemcpyDeviceToHost, stream));
        CK(cudaMemcpyAsync(out->pending_side.data(), D.pending_side, (size_t)np * 4, cudaMemcpyDeviceToHost, stream));
    }
    CK(cudaStreamSynchronize(stream));
    out->root = D.root;
    for (const Segment& sg : D.left) {
        PendingSegment p; p.start = sg.slot0; p.end = sg.slot0 + sg.len; p.force_leaf = 0;
        out->pending.push_back(p); out->pending_parent.push_back(sg.parent); out->pending_side.push_back(sg.side);
    }
    out->levels = D.levels;
    D.lap("downloads");
    return true;
}

// The whole build on the device: bounds, SAH tree, traversal layout, primitive records.  Only the primitive order
// (4 bytes per primitive) comes back.  Returns false (with *err) whenever the device cannot finish by itself -- NaN
// bounds, a subtree the block kernel does not take, a root that is a Leaf -- and the caller falls back to the
// hybrid path, which produces the same tree.
bool gpu_sah_build_resident(const ResidentBuildInput& in, uint32_t min_len, int max_levels, cudaStream_t stream, ResidentBuildOutput* out, std::string* err) {
    const uint32_t n = in.n_prims;
    DeviceSah D;
    if (!D.init(n, min_len, max_levels, stream, err, (size_t)n * sizeof(cpbrt_primitive) + (size_t)in.n_tris * sizeof(cpbrt_triangle) + (size_t)in.n_spheres * 4)) return false;
    const int T = 256;
    auto grid = [&](uint32_t items) { return (int)std::min<uint32_t>((items + T - 1) / T, 148u * 16u); };
    cpbrt_primitive* d_prims_in = D.M.get<cpbrt_primitive>(n);
    cpbrt_triangle* d_tris = D.M.get<cpbrt_triangle>(in.n_tris);
    float* d_radius = D.M.get<float>(in.n_spheres);
    uint32_t* d_flag = D.M.get<uint32_t>(1);
    if (!d_prims_in || !d_tris || !d_radius || !d_flag) { *err = "device memory for the BVH build"; return false; }
    CK(cudaMemcpyAsync(d_prims_in, in.prims, (size_t)n * sizeof(cpbrt_primitive), cudaMemcpyHostToDevice, stream));
    if (in.n_tris) CK(cudaMemcpyAsync(d_tris, in.tris, (size_t)in.n_tris * sizeof(cpbrt_triangle), cudaMemcpyHostToDevice, stream));
    if (in.n_spheres) CK(cudaMemcpyAsync(d_radius, in.radius, (size_t)in.n_spheres * 4, cudaMemcpyHostToDevice, stream));
    out->h2d_bytes = (uint64_t)n * sizeof(cpbrt_primitive) + (uint64_t)in.n_tris * sizeof(cpbrt_triangle) + (uint64_t)in.n_spheres * 4;
    CK(cudaMemsetAsync(d_flag, 0, 4, stream));
    k_prim_bounds<<<grid(n), T, 0, stream>>>(d_prims_in, n, d_tris, d_radius, (const DevXform*)in.d_transforms, D.pb_in, d_flag);
    uint32_t any_nan = 0;
    CK(cudaMemcpyAsync(&any_nan, d_flag, 4, cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    D.lap("scene upload + bounds");
    if (any_nan) { *err = "NaN bounds: the order of unions matters, host build only"; return false; }
    if (!D.run(true, err)) return false;
    if (!D.left.empty()) { *err = "level budget exhausted"; return false; }
    if (D.root != 0 || D.nn == 0) { *err = "the root is a Leaf"; return false; }       // the traversal wants the root Parent at index 0
    const uint32_t np = D.np, n_top = D.nn, n_small = D.n_small;
    if (!D.small_ran && np) { *err = "small phase did not run"; return false; }
    // every pending segment must have been taken by the block kernel
    {
        std::vector<uint32_t> roots(np);
        if (np) CK(cudaMemcpyAsync(roots.data(), D.seg_root, (size_t)np * 4, cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
        for (uint32_t k = 0; k < np; ++k) if (roots[k] == 0xFFFFFFFFu) { *err = "a subtree was left to the host builder"; return false; }
    }
    // Parent numbers of the small subtrees
    uint32_t* pflag = D.M.get<uint32_t>(n_small + 1);
    uint32_t* pidx = D.M.get<uint32_t>(n_small + 1);
    if (!pflag || !pidx) { *err = "device memory for the BVH build"; return false; }
    size_t need = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, need, pflag, pidx, (int)n_small + 1, stream);
    void* tmp = D.M.get<char>(need);
    if (!tmp) { *err = "device memory for the BVH build"; return false; }
    CK(cudaMemsetAsync(pflag + n_small, 0, 4, stream));
    if (n_small) k_small_parent_flags<<<grid(n_small), T, 0, stream>>>(D.small, n_small, pflag);
    CK(cub::DeviceScan::ExclusiveSum(tmp, need, pflag, pidx, (int)n_small + 1, stream));
    uint32_t n_small_parents = 0;
    CK(cudaMemcpyAsync(&n_small_parents, pidx + n_small, 4, cudaMemcpyDeviceToHost, stream));
    TopNode root_node;
    CK(cudaMemcpyAsync(&root_node, D.nodes, sizeof root_node, cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    const uint32_t n_nodes = n_top + n_small_parents;
    // outputs: one allocation {Parents | records | frames | uv}, owned by the caller from here on
    const size_t off_rec = (size_t)n_nodes * 64, off_frames = off_rec + (size_t)n * 48, off_uv = off_frames + (size_t)n * 48;
    const size_t out_bytes = off_uv + (in.any_uv ? (size_t)n * 24 : 0);
    char* o_base = nullptr;
    if (cudaMalloc((void**)&o_base, out_bytes) != cudaSuccess) { cudaGetLastError(); *err = "device memory for the BVH layout"; return false; }
    void *o_nodes = o_base, *o_rec = o_base + off_rec, *o_frames = o_base + off_frames, *o_uv = in.any_uv ? o_base + off_uv : nullptr;
    auto drop = [&]() { cudaFree(o_base); };
    if (np) k_link_pending<<<grid(np), T, 0, stream>>>(D.pending_parent, D.pending_side, D.seg_root, np, D.nodes);
    k_layout_nodes<<<grid(n_top + n_small), T, 0, stream>>>(D.nodes, n_top, D.small, n_small, pidx, D.pb_final, (float4*)o_nodes);
    k_prim_records<<<grid(n), T, 0, stream>>>(d_prims_in, d_tris, d_radius, D.order_final, n, in.any_uv ? 1 : 0, (float4*)o_rec, (float4*)o_frames, (float*)o_uv);
    out->order.resize(n);
    cudaError_t e = cudaMemcpyAsync(out->order.data(), D.order_final, (size_t)n * 4, cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) { drop(); *err = std::string("BVH layout: ") + cudaGetErrorString(e); return false; }
    out->d_base = o_base; out->bytes = out_bytes;
    out->d_nodes = o_nodes; out->d_records = o_rec; out->d_frames = o_frames; out->d_uv = o_uv;
    out->n_nodes = n_nodes; out->levels = D.levels;
    for (int k = 0; k < 3; ++k) { out->root_box[k] = root_node.bound.mn[k]; out->root_box[3 + k] = root_node.bound.mx[k]; }
    D.lap("layout + records + order");
    return true;
}

}  // namespace cpbrt
