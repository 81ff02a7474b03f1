// See bvh_build.h.  Reference: src/aggregate/bvh.rs:24-124, src/geometry/bounds.rs:70-72,255-258.
// Compiled with -ffp-contract=off: the SAH costs are compared as f32 values and must round
// exactly like the reference's.
#include "bvh_build.h"

#include <algorithm>
#include <atomic>
#include <cstring>
#include <functional>
#include <thread>

namespace cpbrt {

namespace {

inline float rmin(float l, float r) { return r < l ? r : l; }   // src/math/mod.rs:36-49
inline float rmax(float l, float r) { return r > l ? r : l; }

inline void unite(BuildBounds& a, const BuildBounds& b) {
    for (int k = 0; k < 3; ++k) { a.mn[k] = rmin(a.mn[k], b.mn[k]); a.mx[k] = rmax(a.mx[k], b.mx[k]); }
}
inline float surface_area(const BuildBounds& b) {   // bounds.rs:255-258
    float dx = b.mx[0] - b.mn[0], dy = b.mx[1] - b.mn[1], dz = b.mx[2] - b.mn[2];
    return 2.f * (dx * dy + dx * dz + dy * dz);
}
inline float center(const BuildBounds& b, int axis) {   // bounds.rs:70-72: min + (max - min) / 2
    return b.mn[axis] + (b.mx[axis] - b.mn[axis]) / 2.f;
}
// OrderedFloat order (NaN greatest and equal to itself, -0 == +0) as an unsigned integer
inline uint32_t sort_key(float f) {
    if (f != f) return 0xFFFFFFFFu;
    if (f == 0.f) f = 0.f;
    uint32_t u; memcpy(&u, &f, 4);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

struct Builder {
    std::vector<BuildBounds> pb;      // bounds of the primitive currently in each slot
    std::vector<uint32_t> order;
    std::vector<BuildNode> arena;
    std::atomic<uint32_t> n_nodes{0};
    std::atomic<int> spare_threads{0};

    uint32_t alloc() { return n_nodes.fetch_add(1); }

    uint32_t make_leaf(uint32_t start, uint32_t end) {
        uint32_t id = alloc();
        BuildNode& n = arena[id];
        n.leaf = true; n.axis = 0; n.a = start; n.b = end; memset(&n.bound, 0, sizeof n.bound);
        return id;
    }

    // Per-thread scratch: millions of small nodes would otherwise spend their time in malloc.  A
    // node is done with its scratch before it recurses, and a forked subtree runs on its own thread.
    struct Scratch { std::vector<uint64_t> keyed; std::vector<float> back_sa; std::vector<uint32_t> idx[3], o; std::vector<float> cost[3]; std::vector<BuildBounds> b; };
    static Scratch& scratch() { static thread_local Scratch s; return s; }

    // one axis of the sweep: sorted local order and the cost of every split position
    void sweep_axis(uint32_t start, uint32_t len, int axis, float full_sa, std::vector<uint32_t>& idx, std::vector<float>& cost) const {
        Scratch& S = scratch();
        std::vector<uint64_t>& keyed = S.keyed;
        keyed.resize(len);
        for (uint32_t i = 0; i < len; ++i) keyed[i] = ((uint64_t)sort_key(center(pb[start + i], axis)) << 32) | i;
        std::sort(keyed.begin(), keyed.end());   // ties fall back to the local index == stable sort
        idx.resize(len);
        for (uint32_t i = 0; i < len; ++i) idx[i] = (uint32_t)keyed[i];
        // suffix areas first, then a running prefix box
        std::vector<float>& back_sa = S.back_sa;
        back_sa.resize(len - 1);
        BuildBounds acc = pb[start + idx[len - 1]];
        back_sa[len - 2] = surface_area(acc);
        for (uint32_t i = len - 2; i-- > 0;) { unite(acc, pb[start + idx[i + 1]]); back_sa[i] = surface_area(acc); }
        cost.resize(len - 1);
        acc = pb[start + idx[0]];
        for (uint32_t i = 0; i + 1 < len; ++i) {
            if (i) unite(acc, pb[start + idx[i]]);
            cost[i] = 0.125f + ((float)(i + 1) * surface_area(acc) + (float)(len - 1 - i) * back_sa[i]) / full_sa;
        }
    }

    uint32_t build(uint32_t start, uint32_t end) {
        const uint32_t len = end - start;
        if (len <= 1) return make_leaf(start, end);
        BuildBounds full = pb[start];
        for (uint32_t i = start + 1; i < end; ++i) unite(full, pb[i]);
        const float full_sa = surface_area(full);

        // big nodes sweep their axes on separate threads and need their own vectors; small ones reuse the thread's scratch
        const bool big = len >= (1u << 16);
        std::vector<uint32_t> own_idx[3];
        std::vector<float> own_cost[3];
        Scratch& S = scratch();
        std::vector<uint32_t>* idx = big ? own_idx : S.idx;
        std::vector<float>* cost = big ? own_cost : S.cost;
        if (big && spare_threads.fetch_sub(2) >= 2) {
            std::thread t1([&] { sweep_axis(start, len, 1, full_sa, idx[1], cost[1]); });
            std::thread t2([&] { sweep_axis(start, len, 2, full_sa, idx[2], cost[2]); });
            sweep_axis(start, len, 0, full_sa, idx[0], cost[0]);
            t1.join(); t2.join();
            spare_threads.fetch_add(2);
        } else {
            if (big) spare_threads.fetch_add(2);
            for (int a = 0; a < 3; ++a) sweep_axis(start, len, a, full_sa, idx[a], cost[a]);
        }
        // the reference's running minimum over (axis, i): the first cost seeds it (even a NaN),
        // afterwards only a strictly smaller cost replaces it (bvh.rs:77-88)
        bool have = false; float min_cost = 0.f; int min_axis = 0; uint32_t min_i = 0;
        for (int a = 0; a < 3; ++a)
            for (uint32_t i = 0; i + 1 < len; ++i) {
                const float c = cost[a][i];
                if (!have) { have = true; min_cost = c; min_axis = a; min_i = i; }
                else if (c < min_cost) { min_cost = c; min_axis = a; min_i = i; }
            }
        if (min_cost > (float)len) return make_leaf(start, end);   // bvh.rs:92-99

        {   // bvh.rs:100-112: slot j receives the primitive that was at local index best[j]
            const std::vector<uint32_t>& best = idx[min_axis];
            std::vector<uint32_t>& o = S.o; std::vector<BuildBounds>& b = S.b;
            o.resize(len); b.resize(len);
            for (uint32_t j = 0; j < len; ++j) { o[j] = order[start + best[j]]; b[j] = pb[start + best[j]]; }
            memcpy(&order[start], o.data(), len * sizeof(uint32_t));
            memcpy(&pb[start], b.data(), len * sizeof(BuildBounds));
            if (big) { std::vector<uint32_t>().swap(o); std::vector<BuildBounds>().swap(b); }   // do not keep 10M-entry scratch alive
        }
        if (big) for (int a = 0; a < 3; ++a) { std::vector<uint32_t>().swap(own_idx[a]); std::vector<float>().swap(own_cost[a]); }

        const uint32_t mid = start + min_i + 1;
        uint32_t left, right;
        const bool fork = (mid - start) >= 4096 && (end - mid) >= 4096;
        if (fork && spare_threads.fetch_sub(1) >= 1) {
            std::thread t([&] { left = build(start, mid); });
            right = build(mid, end);
            t.join();
            spare_threads.fetch_add(1);
        } else {
            if (fork) spare_threads.fetch_add(1);
            left = build(start, mid);
            right = build(mid, end);
        }
        const uint32_t id = alloc();
        BuildNode& n = arena[id];
        n.leaf = false; n.axis = (uint8_t)min_axis; n.a = left; n.b = right; n.bound = full;
        return id;
    }
};

}  // namespace

static void apply_point(const float* m, const float p[3], float out[3]) {   // point.rs:10-14
    float r[4];
    for (int i = 0; i < 4; ++i) r[i] = ((m[i * 4] * p[0] + m[i * 4 + 1] * p[1]) + m[i * 4 + 2] * p[2]) + m[i * 4 + 3] * 1.f;
    if (r[3] != 0.f) { out[0] = r[0] / r[3]; out[1] = r[1] / r[3]; out[2] = r[2] / r[3]; }
    else { out[0] = r[0]; out[1] = r[1]; out[2] = r[2]; }
}

void parallel_for(uint64_t n, uint64_t min_chunk, const std::function<void(uint64_t, uint64_t)>& body) {
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    const uint64_t chunks = std::max<uint64_t>(1, std::min<uint64_t>(hw, n / std::max<uint64_t>(1, min_chunk)));
    if (chunks <= 1) { body(0, n); return; }
    std::vector<std::thread> pool;
    for (uint64_t c = 1; c < chunks; ++c) pool.emplace_back([&, c] { body(n * c / chunks, n * (c + 1) / chunks); });
    body(0, n / chunks);
    for (auto& t : pool) t.join();
}

void primitive_bounds(const cpbrt_primitive* prims, uint64_t n, const cpbrt_triangle* triangles, const float* sphere_radius,
                      const cpbrt_transform* transforms, std::vector<BuildBounds>* out) {
    out->resize(n);
    parallel_for(n, 1 << 16, [&](uint64_t lo, uint64_t hi) {
    for (uint64_t i = lo; i < hi; ++i) {
        const cpbrt_shape& sh = prims[i].shape;
        BuildBounds b;
        if (sh.kind == CPBRT_SHAPE_SPHERE) {
            const float r = sphere_radius[sh.data];
            for (int k = 0; k < 3; ++k) { b.mn[k] = rmin(-r, r); b.mx[k] = rmax(-r, r); }
        } else {
            const cpbrt_triangle& t = triangles[sh.data];   // Bounds3f::new(p0, p1) | p2
            for (int k = 0; k < 3; ++k) {
                b.mn[k] = rmin(rmin(t.p[0][k], t.p[1][k]), t.p[2][k]);
                b.mx[k] = rmax(rmax(t.p[0][k], t.p[1][k]), t.p[2][k]);
            }
        }
        for (int32_t x = sh.xform; x >= 0; x = transforms[x].outer) {
            const float* m = transforms[x].m;
            BuildBounds r;
            for (int c = 0; c < 8; ++c) {
                const float p[3] = {(c & 1) ? b.mx[0] : b.mn[0], (c & 2) ? b.mx[1] : b.mn[1], (c & 4) ? b.mx[2] : b.mn[2]};
                float q[3];
                apply_point(m, p, q);
                for (int k = 0; k < 3; ++k) {
                    if (c == 0) { r.mn[k] = q[k]; r.mx[k] = q[k]; }
                    else { r.mn[k] = rmin(r.mn[k], q[k]); r.mx[k] = rmax(r.mx[k], q[k]); }
                }
            }
            b = r;
        }
        (*out)[i] = b;
    }
    });
}

bool export_bvh(const BvhBuildResult& R, uint32_t* prim_order_out, cpbrt_bvh_dump* dump) {
    if (prim_order_out) memcpy(prim_order_out, R.order.data(), R.order.size() * sizeof(uint32_t));
    if (!dump) return true;
    const uint64_t nn = R.post_order.size();
    dump->n_nodes = nn;
    dump->root = nn ? R.ref_index[R.root] : 0;
    if (dump->capacity < nn) return false;
    for (uint64_t k = 0; k < nn; ++k) {
        const BuildNode& nd = R.arena[R.post_order[k]];
        if (dump->is_leaf) dump->is_leaf[k] = nd.leaf;
        if (dump->axis) dump->axis[k] = nd.axis;
        if (dump->a) dump->a[k] = nd.leaf ? nd.a : R.ref_index[nd.a];
        if (dump->b) dump->b[k] = nd.leaf ? nd.b : R.ref_index[nd.b];
        if (dump->bound) for (int c = 0; c < 3; ++c) { dump->bound[k * 6 + c] = nd.leaf ? 0.f : nd.bound.mn[c]; dump->bound[k * 6 + 3 + c] = nd.leaf ? 0.f : nd.bound.mx[c]; }
    }
    return true;
}

static void number_post_order(BvhBuildResult* out);

void finish_sah_bvh(TopPhase&& top, unsigned n_threads, BvhBuildResult* out) {
    const uint32_t n = (uint32_t)top.order.size();
    out->order.clear(); out->arena.clear(); out->post_order.clear(); out->ref_index.clear(); out->root = 0;
    if (n == 0) return;
    if (n_threads == 0) n_threads = std::max(1u, std::thread::hardware_concurrency());
    Builder b;
    b.pb.swap(top.pb);
    b.order.swap(top.order);
    b.arena.resize(2ull * n + top.nodes.size());
    // arena ids of the device-built Parents first, then the host subtrees
    const uint32_t n_top = (uint32_t)top.nodes.size();
    const uint32_t n_small = (uint32_t)top.small_nodes.size();
    b.n_nodes = n_top + n_small;
    const uint32_t np = (uint32_t)top.pending.size();
    std::vector<uint32_t> pending_id(np, 0);
    // subtrees the device already built: copied behind the top nodes, indices shifted
    parallel_for(n_small, 1 << 16, [&](uint64_t lo, uint64_t hi) {
        for (uint64_t i = lo; i < hi; ++i) {
            const SmallSubtreeNode& sn = top.small_nodes[i];
            BuildNode& nd = b.arena[n_top + i];
            nd.leaf = (sn.leaf_axis & 1u) != 0; nd.axis = (uint8_t)(sn.leaf_axis >> 8);
            nd.a = nd.leaf ? sn.a : n_top + sn.a; nd.b = nd.leaf ? sn.b : n_top + sn.b;
            for (int k = 0; k < 3; ++k) { nd.bound.mn[k] = sn.mn[k]; nd.bound.mx[k] = sn.mx[k]; }
        }
    });
    auto device_built = [&](uint32_t k) { return k < top.pending_root.size() && top.pending_root[k] != 0xFFFFFFFFu; };
    // many segments: one worker per thread pulls segments; few segments: let build() fork inside them
    const unsigned workers = std::max(1u, std::min<unsigned>(n_threads, np));
    b.spare_threads = (int)n_threads - (int)workers;
    std::atomic<uint32_t> next{0};
    auto work = [&]() {
        for (;;) {
            const uint32_t k = next.fetch_add(1);
            if (k >= np) break;
            const PendingSegment& p = top.pending[k];
            if (device_built(k)) pending_id[k] = n_top + top.pending_root[k];
            else pending_id[k] = p.force_leaf ? b.make_leaf(p.start, p.end) : b.build(p.start, p.end);
        }
    };
    {
        std::vector<std::thread> pool;
        for (unsigned t = 1; t < workers; ++t) pool.emplace_back(work);
        work();
        for (auto& t : pool) t.join();
    }
    for (uint32_t i = 0; i < n_top; ++i) {
        BuildNode& nd = b.arena[i];
        nd.leaf = false; nd.axis = top.nodes[i].axis; nd.bound = top.nodes[i].bound;
        nd.a = (uint32_t)top.nodes[i].child[0]; nd.b = (uint32_t)top.nodes[i].child[1];
    }
    uint32_t root = top.root == 0x7FFFFFFF ? 0u : (uint32_t)top.root;
    for (uint32_t k = 0; k < np; ++k) {
        if (top.pending_parent[k] < 0) root = pending_id[k];
        else if (top.pending_side[k] == 0) b.arena[top.pending_parent[k]].a = pending_id[k];
        else b.arena[top.pending_parent[k]].b = pending_id[k];
    }
    out->root = root;
    out->host_built_segments = 0;
    for (uint32_t k = 0; k < np; ++k) if (!device_built(k) && !top.pending[k].force_leaf) ++out->host_built_segments;
    b.arena.resize(b.n_nodes.load());
    out->order.swap(b.order);
    out->arena.swap(b.arena);
    number_post_order(out);
}

void build_sah_bvh(const std::vector<BuildBounds>& bounds, unsigned n_threads, BvhBuildResult* out) {
    const uint32_t n = (uint32_t)bounds.size();
    out->order.clear(); out->arena.clear(); out->post_order.clear(); out->ref_index.clear(); out->root = 0;
    if (n == 0) return;
    Builder b;
    b.pb = bounds;
    b.order.resize(n);
    for (uint32_t i = 0; i < n; ++i) b.order[i] = i;
    b.arena.resize(2ull * n);
    if (n_threads == 0) n_threads = std::max(1u, std::thread::hardware_concurrency());
    b.spare_threads = (int)n_threads - 1;
    out->root = b.build(0, n);
    b.arena.resize(b.n_nodes.load());
    out->order.swap(b.order);
    out->arena.swap(b.arena);
    number_post_order(out);
}

// reference numbering: post-order, left subtree first (bvh.rs:113-123)
static void number_post_order(BvhBuildResult* out) {
    const uint32_t nn = (uint32_t)out->arena.size();
    out->post_order.reserve(nn);
    out->ref_index.assign(nn, 0);
    std::vector<std::pair<uint32_t, int>> stack;
    stack.push_back({out->root, 0});
    while (!stack.empty()) {
        auto& top = stack.back();
        const BuildNode& node = out->arena[top.first];
        if (node.leaf || top.second == 2) {
            out->ref_index[top.first] = (uint32_t)out->post_order.size();
            out->post_order.push_back(top.first);
            stack.pop_back();
        } else {
            const uint32_t child = top.second == 0 ? node.a : node.b;
            ++top.second;
            stack.push_back({child, 0});
        }
    }
}

}  // namespace cpbrt
