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

struct DevMem {
    std::vector<void*> all;
    ~DevMem() { for (void* p : all) cudaFree(p); }
    template <class T> T* get(size_t n) { void* p = nullptr; if (cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)) != cudaSuccess) return nullptr; all.push_back(p); return (T*)p; }
};

#define CK(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) { *err = std::string(#x) + ": " + cudaGetErrorString(e__); return false; } } while (0)

}  // namespace

bool gpu_sah_top_phase(const std::vector<BuildBounds>& bounds, uint32_t min_len, int max_levels, cudaStream_t stream, TopPhase* out, std::string* err) {
    const uint32_t n = (uint32_t)bounds.size();
    out->pb.clear(); out->order.clear(); out->nodes.clear(); out->pending.clear(); out->pending_parent.clear(); out->pending_side.clear();
    out->root = 0x7FFFFFFF; out->levels = 0;
    if (n < 2 || n < min_len || n >= 0x3FFFFFFFu) { *err = "nothing to do on the device"; return false; }
    if (min_len < 2) min_len = 2;
    for (const BuildBounds& b : bounds)
        for (int k = 0; k < 3; ++k) if (b.mn[k] != b.mn[k] || b.mx[k] != b.mx[k]) { *err = "NaN bounds: the order of unions matters, host build only"; return false; }

    DevMem M;
    const uint32_t max_segs = n / min_len + 2;
    const uint32_t node_cap = 2 * max_segs * (uint32_t)std::max(1, max_levels) + 16, pending_cap = node_cap + 16;
    Box *pb = M.get<Box>(n), *pb_new = M.get<Box>(n), *pb_alt = M.get<Box>(n), *fwd = M.get<Box>(n), *rev = M.get<Box>(n), *prefix = M.get<Box>(n),
        *suffix_rev = M.get<Box>(n), *prefix0 = M.get<Box>(n), *pb_final = M.get<Box>(n);
    uint32_t *order = M.get<uint32_t>(n), *order_new = M.get<uint32_t>(n), *order_alt = M.get<uint32_t>(n), *order_final = M.get<uint32_t>(n), *seg = M.get<uint32_t>(n),
             *seg_rev = M.get<uint32_t>(n), *vals = M.get<uint32_t>(n), *keep = M.get<uint32_t>(n + 1), *keep_off = M.get<uint32_t>(n + 1);
    uint32_t* idx[3] = {M.get<uint32_t>(n), M.get<uint32_t>(n), M.get<uint32_t>(n)};
    unsigned long long *keys = M.get<unsigned long long>(n), *keys_out = M.get<unsigned long long>(n), *best = M.get<unsigned long long>(max_segs);
    Segment *segs = M.get<Segment>(max_segs), *segs_next = M.get<Segment>(max_segs);
    Decision* dec = M.get<Decision>(max_segs);
    uint32_t *first_nan = M.get<uint32_t>(max_segs), *child_count = M.get<uint32_t>(max_segs + 1), *child_off = M.get<uint32_t>(max_segs + 1);
    TopNode* nodes = M.get<TopNode>(node_cap);
    PendingSegment* pending = M.get<PendingSegment>(pending_cap);
    int32_t* pending_parent = M.get<int32_t>(pending_cap); uint32_t* pending_side = M.get<uint32_t>(pending_cap);
    uint32_t* counters = M.get<uint32_t>(8);   // [0] nodes, [1] pending, [2] overflow, [3] root_ref (as int)
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

    CK(cudaMemcpyAsync(pb, bounds.data(), (size_t)n * sizeof(Box), cudaMemcpyHostToDevice, stream));
    {
        std::vector<uint32_t> iota(n);
        for (uint32_t i = 0; i < n; ++i) iota[i] = i;
        CK(cudaMemcpyAsync(order, iota.data(), (size_t)n * 4, cudaMemcpyHostToDevice, stream));
        CK(cudaStreamSynchronize(stream));
    }
    CK(cudaMemsetAsync(counters, 0, 8 * 4, stream));
    {
        int32_t none = 0x7FFFFFFF;
        CK(cudaMemcpyAsync(counters + 3, &none, 4, cudaMemcpyHostToDevice, stream));
        Segment root; root.slot0 = 0; root.rank0 = 0; root.len = n; root.parent = -1; root.side = 0;
        CK(cudaMemcpyAsync(segs, &root, sizeof root, cudaMemcpyHostToDevice, stream));
    }
    uint32_t m = n, n_segs = 1;
    const int T = 256;
    auto grid = [&](uint32_t items) { return (int)std::min<uint32_t>((items + T - 1) / T, 148u * 16u); };
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
    // segments still active when the level budget ran out go to the host as they are
    std::vector<Segment> left(n_segs);
    if (n_segs) {
        CK(cudaMemcpyAsync(left.data(), segs, (size_t)n_segs * sizeof(Segment), cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
        // their primitives are in the working arrays: move them to their final slots
        std::vector<Box> hb(m); std::vector<uint32_t> ho(m);
        CK(cudaMemcpyAsync(hb.data(), pb, (size_t)m * sizeof(Box), cudaMemcpyDeviceToHost, stream));
        CK(cudaMemcpyAsync(ho.data(), order, (size_t)m * 4, cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
        for (const Segment& sg : left) {
            CK(cudaMemcpyAsync(pb_final + sg.slot0, hb.data() + sg.rank0, (size_t)sg.len * sizeof(Box), cudaMemcpyHostToDevice, stream));
            CK(cudaMemcpyAsync(order_final + sg.slot0, ho.data() + sg.rank0, (size_t)sg.len * 4, cudaMemcpyHostToDevice, stream));
        }
        CK(cudaStreamSynchronize(stream));
    }
    uint32_t h_counters[4];
    CK(cudaMemcpyAsync(h_counters, counters, 16, cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    const uint32_t nn = h_counters[0], np = h_counters[1];
    out->pb.resize(n); out->order.resize(n); out->nodes.resize(nn); out->pending.resize(np); out->pending_parent.resize(np); out->pending_side.resize(np);
    CK(cudaMemcpyAsync(out->pb.data(), pb_final, (size_t)n * sizeof(Box), cudaMemcpyDeviceToHost, stream));
    CK(cudaMemcpyAsync(out->order.data(), order_final, (size_t)n * 4, cudaMemcpyDeviceToHost, stream));
    if (nn) CK(cudaMemcpyAsync(out->nodes.data(), nodes, (size_t)nn * sizeof(TopNode), cudaMemcpyDeviceToHost, stream));
    if (np) {
        CK(cudaMemcpyAsync(out->pending.data(), pending, (size_t)np * sizeof(PendingSegment), cudaMemcpyDeviceToHost, stream));
        CK(cudaMemcpyAsync(out->pending_parent.data(), pending_parent, (size_t)np * 4, cudaMemcpyDeviceToHost, stream));
        CK(cudaMemcpyAsync(out->pending_side.data(), pending_side, (size_t)np * 4, cudaMemcpyDeviceToHost, stream));
    }
    CK(cudaStreamSynchronize(stream));
    out->root = (int32_t)h_counters[3];
    for (const Segment& sg : left) {
        PendingSegment p; p.start = sg.slot0; p.end = sg.slot0 + sg.len; p.force_leaf = 0;
        out->pending.push_back(p); out->pending_parent.push_back(sg.parent); out->pending_side.push_back(sg.side);
    }
    out->levels = level;
    return true;
}

}  // namespace cpbrt
