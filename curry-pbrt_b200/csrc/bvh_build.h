// Exact full-sweep SAH build of the reference (BVHAggregate::create_sah_node,
// src/aggregate/bvh.rs:24-124) -- same split decisions, same primitive order, same node
// numbering -- organised for large inputs: keys are sorted as packed integers, the three
// axes of a big node are swept by separate threads, and independent subtrees are built in
// parallel.
#pragma once
#include <cstdint>
#include <functional>
#include <vector>

#include "../../include/cpbrt.h"

namespace cpbrt {

struct BuildBounds { float mn[3], mx[3]; };

struct BuildNode {      // BVHNode (bvh.rs:13-21)
    bool leaf;
    uint8_t axis;
    uint32_t a, b;      // Leaf: slot range [a, b) ; Parent: arena indices of left / right
    BuildBounds bound;  // Parent: union of its primitives' bounds
};

struct BvhBuildResult {
    std::vector<uint32_t> order;      // BVH slot -> uploaded primitive index
    std::vector<BuildNode> arena;     // nodes in allocation order
    uint32_t root = 0;                // arena index
    // the reference's numbering: nodes are pushed after their children (post-order), root last
    std::vector<uint32_t> post_order; // arena index of reference node k
    std::vector<uint32_t> ref_index;  // reference node number of arena node i
    uint32_t host_built_segments = 0; // hybrid build: pending segments the HOST builder had to finish (0: the whole SAH build ran on the device)
};

// body(lo, hi) over [0, n) split across the host threads (chunks of at least min_chunk items)
void parallel_for(uint64_t n, uint64_t min_chunk, const std::function<void(uint64_t, uint64_t)>& body);

// Primitive::bound() of every uploaded primitive (triangle.rs:114-119, sphere.rs:22-27,
// TransformShape::bound transform.rs:52-54 with Bounds3f::apply bounds.rs:292-300).
void primitive_bounds(const cpbrt_primitive* prims, uint64_t n, const cpbrt_triangle* triangles, const float* sphere_radius,
                      const cpbrt_transform* transforms, std::vector<BuildBounds>* out);

// Fills a caller-allocated cpbrt_bvh_dump / order array in the reference's node numbering.
// Returns false when dump->capacity is too small.
bool export_bvh(const BvhBuildResult& r, uint32_t* prim_order_out, cpbrt_bvh_dump* dump);

// bounds[i] = Primitive::bound() of uploaded primitive i.  n_threads == 0 -> hardware.
void build_sah_bvh(const std::vector<BuildBounds>& bounds, unsigned n_threads, BvhBuildResult* out);

// ---- hybrid build: the top of the tree on the GPU (bvh_build_gpu.cu), the rest on the host ------
struct TopNode {            // a Parent created by the device phase
    uint8_t axis;
    BuildBounds bound;
    int32_t child[2];       // TopNode index once known; pending segments attach through pending_parent / pending_side
};
struct PendingSegment { uint32_t start, end, force_leaf; };   // slot range below the level-synchronous phase (force_leaf: the SAH said "leaf")
// a node of a subtree the device built for a pending segment (bvh_build_gpu.cu, k_build_small)
struct SmallSubtreeNode { uint32_t leaf_axis; uint32_t a, b; float mn[3], mx[3]; };   // bit 0 = leaf, bits 8.. = axis; Leaf: slot range, Parent: indices into small_nodes
struct TopPhase {
    std::vector<BuildBounds> pb;          // bound of the primitive now in each slot
    std::vector<uint32_t> order;          // slot -> uploaded primitive index, as permuted by the device levels
    std::vector<TopNode> nodes;
    std::vector<PendingSegment> pending;
    std::vector<int32_t> pending_parent;  // TopNode that owns pending segment k (-1: it is the root)
    std::vector<uint32_t> pending_side;
    int32_t root = 0x7FFFFFFF;            // TopNode index of the root (0x7FFFFFFF: the root is a pending segment)
    int levels = 0;
    // pending segments already finished on the device: their subtrees' nodes, and per pending segment the index of
    // its root in small_nodes (0xFFFFFFFF / beyond the end: not built there, the host builder does it)
    std::vector<SmallSubtreeNode> small_nodes;
    std::vector<uint32_t> pending_root;
};
// The whole build on the device (bvh_build_gpu.cu, gpu_sah_build_resident): host scene arrays in, device traversal
// layout out (scene_dev.cuh: 64-byte Parent records, 48-byte primitive records, shading frames, uv triples).
struct ResidentBuildInput {
    const cpbrt_primitive* prims; uint32_t n_prims;
    const cpbrt_triangle* tris; uint32_t n_tris;
    const float* radius; uint32_t n_spheres;
    const void* d_transforms;         // DevTransform array already on the device
    bool any_uv;
};
struct ResidentBuildOutput {
    void* d_base = nullptr; size_t bytes = 0;        // ONE cudaMalloc'ed block, owned by the caller; the four arrays below are pieces of it
    void *d_nodes = nullptr, *d_records = nullptr, *d_frames = nullptr, *d_uv = nullptr;
    uint32_t n_nodes = 0;
    float root_box[6] = {0, 0, 0, 0, 0, 0};
    std::vector<uint32_t> order;      // slot -> uploaded primitive index
    int levels = 0;
    uint64_t h2d_bytes = 0;
};
// Builds every pending segment with the host builder (in parallel) and stitches the tree together.
void finish_sah_bvh(TopPhase&& top, unsigned n_threads, BvhBuildResult* out);

}  // namespace cpbrt
