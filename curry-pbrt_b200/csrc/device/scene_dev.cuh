// Device-resident scene: how the flattened cpbrt_scene_desc is laid out in HBM.
//
//   prims  48 B per BVH slot, three 16-byte vectors (one LDG.128 each), in BVH order so that a
//          leaf's primitives are contiguous:
//            q0 = p0.x p0.y p0.z p1.x      (sphere: radius, -, -, -)
//            q1 = p1.y p1.z p2.x p2.y
//            q2 = p2.z  upload_index  meta  source
//          meta   bit 0 = shape kind (0 triangle, 1 sphere), bit 1 = has explicit uv triple,
//                 bits 2.. = innermost transform index + 1 (0 = none)
//          source material index, or 0x80000000 | light index for emissive primitives
//   nodes  64 B per Parent node = both children's boxes + links, four 16-byte vectors:
//            q0 = L.min.x L.min.y L.min.z L.max.x
//            q1 = L.max.y L.max.z R.min.x R.min.y
//            q2 = R.min.z R.max.x R.max.y R.max.z
//            q3 = left_ref  right_ref  left_cnt | axis << 30  right_cnt
//          child with cnt > 0 is a Leaf: ref = first BVH slot, cnt = number of primitives; the
//          reference tests no box for leaves (bvh.rs:158-165), so its predicate ignores this
//          box -- it holds the union of the leaf's primitive bounds and is used only by the
//          two-phase traversal to skip primitives that cannot be hit (bvh_trace.cuh);
//          child with cnt == 0 is a Parent: ref = node index, box = that Parent's own `bound`.
//          The reference stores each Parent's bound in the Parent itself (bvh.rs:13-21) and
//          tests it on entry; keeping both children's bounds in the parent lets one 64-byte
//          fetch decide both children, with identical arithmetic.
#pragma once
#include "dmath.cuh"

namespace cpbrt {

struct DevTransform { float m[16]; float m_inv[16]; int32_t outer; int32_t pad[3]; };   // 144 B
struct DevTexture { uint32_t kind, channels; float value[3]; uint32_t width, height; uint32_t pad; uint64_t offset; };
struct DevMaterial { uint32_t type; int32_t tex[9]; int32_t sub[2]; };
struct DevLight {
    uint32_t type;
    uint32_t shape_kind, shape_data; int32_t shape_xform;   // AREA: its own shape
    float v[3]; float L[3];
    int32_t tex;
    float tri[9];        // AREA + triangle: object-space vertices
    float tri_n[3];      // ... and Triangle::normal of them (triangle.rs:78-82), computed once on the host with the same f32 operations
    float radius;        // AREA + sphere
    float area;          // Shape::area of the light's shape
    uint32_t slot;       // AREA: BVH slot of the one primitive that carries this light, 0xFFFFFFFF if none or several
    // INFINITE: Distribution2D over the environment map (infinite_area.rs:11-26), four f32 tables at
    // DevScene::env + env_off: row pdf [h], row cdf [h], per-row pdf [h * w], per-row cdf [h * w]
    uint32_t env_w, env_h;
    uint64_t env_off;
};

static constexpr uint32_t kMetaSphere = 1u, kMetaHasUv = 2u;
static constexpr uint32_t kSourceLight = 0x80000000u;

struct DevCamera {
    float raster_to_camera[16];
    float camera_to_world[16];
    int32_t has_camera_to_world, has_lens;
    float lens_radius, focal_distance;
};

// one 48-byte record per Halton dimension (three 16-byte loads): everything scrambled_radical_inverse needs.
// Digits are consumed in GROUPS of `group_digits` (as many as keep base^digits <= kHaltonGroupMax): one division by
// `group` = base^group_digits and one look-up in a table of already reversed and permuted digit strings replace
// group_digits rounds of divide / look up / multiply-add.  Integer arithmetic, so the accumulator is the same.
struct DevHaltonDim {
    uint64_t magic;        // floor(2^64 / base) + 1: x / base == umul64hi(x, magic) exactly for x < 2^32
    uint32_t base;         // PRIMS[d]
    uint32_t primsum;      // offset of the dimension's permutation block (single digits)
    float tail;            // inv_base * perm(0) / (1 - inv_base), evaluated on the host in f32 (halton.rs:61-63)
    uint32_t group;        // base^group_digits
    uint64_t group_magic;  // floor(2^64 / group) + 1
    uint32_t group_off;    // tables of 1, 2, ... group_digits digits one after the other (sizes base, base^2, ...), in DevSampler::permutation
    uint32_t full_off;     // where the last of them (group_digits digits: the full-group table) starts
    uint32_t group_digits;
    uint32_t pad;
};
static_assert(sizeof(DevHaltonDim) == 48, "three 16-byte loads");
static constexpr int kHaltonMaxDigits = 32;   // base >= 5 from dimension 2 on: at most 28 digits in 64 bits

struct DevSampler {
    uint32_t spp;
    uint32_t exp[2];
    uint64_t scale[2];
    uint64_t mult_inverse[2];
    uint64_t stride;                 // scale[0] * scale[1]
    const uint16_t* permutation;     // digit permutations, block of prime k at primsum[k]; behind them the digit-group tables (DevHaltonDim)
    const uint32_t* primes;          // first 1000 primes
    const uint32_t* primsum;
    const uint64_t* prime_magic;     // floor(2^64 / p) + 1: x / p == umul64hi(x, magic) exactly for x < 2^32
    const DevHaltonDim* dims;        // the same per-dimension data gathered in one record
    const float* inv_pow;            // [dim][digit count] = 1 / (base^count as f32)
    uint32_t table_len;              // entries behind `permutation` (permutation blocks + digit-group tables)
};

struct DevScene {
    // geometry
    const float4* prims;        // 3 per BVH slot
    const float* prim_uv;       // 6 per BVH slot, only if any triangle has explicit uvs (else null)
    // 3 per BVH slot (triangles): {n.xyz, snx.x}{snx.yz, sny.xy}{sny.z, n_local.xyz} -- Triangle::normal
    // (triangle.rs:78-82) and, for an untransformed triangle, the shading frame BSDF::new builds from it
    // (bxdf/mod.rs:83-97, coordinate_system math/mod.rs:67-74), evaluated once on the host with the same
    // f32 operations instead of at every hit
    const float4* prim_frame;
    const float4* nodes;        // 4 per Parent node
    const DevTransform* transforms;
    uint32_t n_prims, n_nodes;
    // root: either a Parent (root_cnt == 0, own bound in root_box) or a single Leaf
    uint32_t root_ref, root_cnt;
    float root_box[6];
    // tiny scenes (<= kFlatMaxPrims primitives): the flat search tables of flat_trace.cuh, null otherwise
    const float4* flat;
    uint32_t flat_n, flat_stride;          // primitives; float4s between the three rotated copies of the per-slot triangles
    uint32_t flat_items, flat_pstride;     // phase-1 items (triangle pairs); float4s between their three rotated copies
    uint64_t flat_item_mask, flat_special_mask;   // valid candidate bits (2 per item); slots that are not plain triangles
    // shading
    const DevMaterial* materials;
    const DevTexture* textures;
    const float* texels;
    const DevLight* lights;
    const float* env;           // distribution tables of the infinite lights
    uint32_t n_lights, n_materials, n_textures, n_transforms;
    uint32_t integrator, max_depth;
    uint32_t width, height;
    DevCamera camera;
    DevSampler sampler;
};

}  // namespace cpbrt
