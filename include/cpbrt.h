/*
 * cpbrt.h -- C-ABI of the B200-native rendering hot path for curry-pbrt.
 *
 * The reference (CurryPseudo/curry-pbrt, Rust) has no FFI; its seam for this path is the
 * whole-frame function
 *     render(scene, sampler, integrator, film, camera) -> Film        (src/render.rs:7-50)
 * whose only caller is render_from_file (src/render.rs:63-82).  Everything that function
 * reaches -- Halton sampling, camera rays, SAH-BVH build and traversal, ray/triangle and
 * ray/sphere intersection, BSDF + texture evaluation, light sampling, shadow rays, the path
 * and direct-lighting integrators and film accumulation -- lives behind this header in
 * hand-written sm_100a CUDA.  The host (parser, Scene/Primitive vocabulary, PNG I/O) stays on
 * the CPU and hands the device a flattened, plain-old-data scene description.
 *
 * Conventions
 *   - every function returns 0 on success or a negative cpbrt_status; nothing unwinds or
 *     aborts across the ABI.  cpbrt_last_error() gives a human-readable message.
 *   - all pointers inside cpbrt_scene_desc are borrowed only for the duration of the call
 *     that receives them.
 *   - matrices are 4x4 f32, row-major (m[r*4+c]), as the reference's nalgebra Matrix4::new
 *     argument order (src/geometry/transform.rs:43-60).
 *   - a cpbrt_ctx may be used from one host thread at a time (Send, not Sync).
 *   - there is no CPU fallback: without a CUDA device cpbrt_create fails with
 *     CPBRT_ERR_NO_DEVICE and every compute entry point fails with CPBRT_ERR_BAD_CTX.
 */
#ifndef CPBRT_H
#define CPBRT_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CPBRT_ABI_VERSION 2

typedef enum cpbrt_status {
    CPBRT_OK = 0,
    CPBRT_ERR_NO_DEVICE = -1,   /* no CUDA device / driver */
    CPBRT_ERR_CUDA = -2,        /* a CUDA runtime call failed (see cpbrt_last_error) */
    CPBRT_ERR_BAD_ARG = -3,     /* null pointer, out-of-range index in the scene description ... */
    CPBRT_ERR_BAD_CTX = -4,     /* ctx is null or in the wrong state (no scene, no BVH ...) */
    CPBRT_ERR_PARSE = -5,       /* scene file could not be read / parsed (reference: panic!) */
    CPBRT_ERR_IO = -6,          /* image file could not be read / written */
    CPBRT_ERR_UNSUPPORTED = -7, /* feature of the reference that this build does not cover */
    CPBRT_ERR_NCCL = -8         /* NCCL call failed */
} cpbrt_status;

/* ------------------------------------------------------------------------------------------
 * Flattened scene description = the reference's Scene{lights, materials, aggregate}
 * (src/scene.rs:10-14) + camera + sampler + integrator + film parameters
 * (src/render.rs:63-79), with every Arc<dyn Trait> replaced by an index.
 * ---------------------------------------------------------------------------------------- */

/* Shape kinds (src/geometry/shape/{triangle,sphere}.rs). */
enum { CPBRT_SHAPE_TRIANGLE = 0, CPBRT_SHAPE_SPHERE = 1 };

/* TransformShape (src/geometry/shape/transform.rs:11-15).  `outer` chains a wrapper applied
 * on top of this one (ObjectInstance under a CTM wraps an already wrapped shape,
 * src/primitive.rs:115-130); -1 = none.  object -> m -> (outer.m) -> world. */
typedef struct cpbrt_transform {
    float m[16];
    float m_inv[16];
    int32_t outer;
    int32_t _pad[3];
} cpbrt_transform;

/* Triangle in object space: the three mesh vertices it indexes (triangle.rs:61-66) and its
 * uv triple (fixed (0,0),(1,0),(1,1) when the mesh has no uvs, triangle.rs:67-77). */
typedef struct cpbrt_triangle {
    float p[3][3];
    float uv[3][2];
} cpbrt_triangle;

/* A shape instance: what Arc<dyn Shape> points at. */
typedef struct cpbrt_shape {
    uint32_t kind;     /* CPBRT_SHAPE_* */
    uint32_t data;     /* triangle: index into triangles[]; sphere: index into sphere_radius[] */
    int32_t  xform;    /* innermost cpbrt_transform, -1 = bare shape */
    int32_t  _pad;
} cpbrt_shape;

/* Primitive{shape, source} (src/primitive.rs:6-16) in pre-BVH order, i.e. the order in which
 * SceneParseStack::parse pushed them (src/scene.rs:84-140), after the camera clipper. */
typedef struct cpbrt_primitive {
    cpbrt_shape shape;
    int32_t material;  /* PrimitiveSource::Material -> index into materials[], else -1 */
    int32_t light;     /* PrimitiveSource::Light    -> index into lights[],    else -1 */
} cpbrt_primitive;

/* Texture<T>::evaluate (src/texture/{constant,image}.rs). */
enum { CPBRT_TEX_CONSTANT = 0, CPBRT_TEX_IMAGE = 1 };
typedef struct cpbrt_texture {
    uint32_t kind;      /* CPBRT_TEX_* */
    uint32_t channels;  /* 1 = Texture<Float>, 3 = Texture<Spectrum> */
    float    value[3];  /* constant value (value[0] for Float) */
    uint32_t width, height;
    uint64_t offset;    /* image: first float in texels[]; row-major, `channels` floats per texel */
} cpbrt_texture;

/* Materials (src/material/{matte,mirror,glass,...}.rs).  tex[] holds texture indices, -1 = slot unused. */
enum {
    CPBRT_MAT_MATTE = 0,       /* tex: 0 Kd(S) 1 sigma(F)                                  matte.rs:17-28 */
    CPBRT_MAT_MIRROR = 1,      /* tex: 0 Kr(S)                                             mirror.rs:15-23 */
    CPBRT_MAT_GLASS = 2,       /* tex: 0 Kr(S) 1 Kt(S) 2 index(F)                          glass.rs:17-29 */
    CPBRT_MAT_PLASTIC = 3,     /* tex: 0 Kd(S) 1 Ks(S) 2 roughness(F)                      plastic.rs:10-31 */
    CPBRT_MAT_UBER = 4,        /* tex: 0 Kd 1 Ks 2 Kr 3 Kt 4 roughness 5 urough 6 vrough 7 eta 8 opacity  uber.rs:16-62 */
    CPBRT_MAT_TRANSLUCENT = 5, /* tex: 0 Kd 1 Ks 2 roughness 3 reflect 4 transmit          translucent.rs:13-57 */
    CPBRT_MAT_MIX = 6          /* tex: 0 amount(S); sub[0], sub[1] = material indices       mix.rs:11-16 */
};
#define CPBRT_MAT_MAX_TEX 9
typedef struct cpbrt_material {
    uint32_t type;
    int32_t  tex[CPBRT_MAT_MAX_TEX];
    int32_t  sub[2];
} cpbrt_material;

/* Lights in scene.lights order (src/scene.rs:91-94,141-147). */
enum {
    CPBRT_LIGHT_AREA = 0,      /* AreaLight{shape, le}            src/light/area.rs:5-15 */
    CPBRT_LIGHT_POINT = 1,     /* PointLight{point, i}            src/light/point.rs:4-26 */
    CPBRT_LIGHT_DISTANT = 2,   /* DistantLight{w, i}              src/light/distant.rs:4-26 */
    CPBRT_LIGHT_INFINITE = 3   /* InfiniteAreaLight{map}          src/light/infinite_area.rs:3-26 */
};
typedef struct cpbrt_light {
    uint32_t type;
    cpbrt_shape shape;   /* AREA: the light's own shape (the Arc it was created from) */
    float    v[3];       /* POINT: world position; DISTANT: w (already transformed) */
    float    L[3];       /* AREA: le; POINT/DISTANT: i */
    int32_t  tex;        /* INFINITE: Spectrum texture index */
} cpbrt_light;

/* Camera (src/camera/{perspective,mod,lens}.rs). */
typedef struct cpbrt_camera {
    float raster_to_camera[16];    /* PerspectiveCamera.raster_to_camera.m   perspective.rs:5-8 */
    float camera_to_raster[16];    /* its inverse as stored (m_inv); used by the host clipper */
    float camera_to_world[16];     /* TransformCamera.transform.m            camera/mod.rs:49-55 */
    int32_t has_camera_to_world;   /* 0 = bare camera (no top-level transform directive) */
    int32_t has_lens;              /* LensCamera wrapper present              lens.rs:2-6 */
    float lens_radius, focal_distance;
} cpbrt_camera;

/* HaltonSampler (src/sampler/halton.rs:71-102) plus the process-global tables it reads.
 * The reference shuffles PERMUTATION from thread_rng() at start-up (halton.rs:216-231), so the
 * table is an INPUT here: the same table goes to the oracle and to the GPU. */
#define CPBRT_N_PRIMES 1000
#define CPBRT_PERM_TABLE_LEN 3682913u   /* sum of the first 1000 primes */
typedef struct cpbrt_sampler {
    uint32_t spp;                  /* sample_per_pixel */
    uint64_t scale[2];             /* smallest 2^a >= xres, 3^b >= yres */
    uint32_t exp[2];               /* a, b */
    uint64_t mult_inverse[2];      /* halton.rs:91-93 */
    const uint16_t* permutation;   /* CPBRT_PERM_TABLE_LEN digits, prime p's block at PRIMSUM[p] */
} cpbrt_sampler;

enum { CPBRT_INTEGRATOR_PATH = 0, CPBRT_INTEGRATOR_DIRECT = 1 };

typedef struct cpbrt_scene_desc {
    uint32_t abi_version;          /* CPBRT_ABI_VERSION */

    uint64_t n_primitives;  const cpbrt_primitive* primitives;
    uint64_t n_triangles;   const cpbrt_triangle*  triangles;
    uint64_t n_spheres;     const float*           sphere_radius;
    uint32_t n_transforms;  const cpbrt_transform* transforms;
    uint32_t n_materials;   const cpbrt_material*  materials;
    uint32_t n_textures;    const cpbrt_texture*   textures;
    uint64_t n_texels;      const float*           texels;
    uint32_t n_lights;      const cpbrt_light*     lights;

    cpbrt_camera  camera;
    cpbrt_sampler sampler;
    uint32_t integrator;           /* CPBRT_INTEGRATOR_* (src/integrator/mod.rs:98-114) */
    uint32_t max_depth;
    uint32_t film_width, film_height;   /* Film::new(resolution) (src/film.rs:27-34) */
} cpbrt_scene_desc;

/* ------------------------------------------------------------------------------------------
 * Batch records for the parity entry points.
 * ---------------------------------------------------------------------------------------- */
typedef struct cpbrt_ray {          /* Ray{o, d, t_max} (src/geometry/ray.rs:7-11), 32 B */
    float o[3]; float t_max;
    float d[3]; float _pad;
} cpbrt_ray;

#define CPBRT_NO_HIT 0xFFFFFFFFu
typedef struct cpbrt_hit {          /* 16 B */
    float    t;
    uint32_t prim;      /* index into the uploaded (pre-BVH) primitive list; CPBRT_NO_HIT = miss */
    float    u, v;      /* ShapePoint.uv of the hit (triangle.rs:83-86, sphere.rs:14-18) */
} cpbrt_hit;

/* BVH dump (BVHAggregate{primitives, nodes, root}, src/aggregate/bvh.rs:6-21) in the
 * reference's own node order (post-order, root last).  Arrays are caller-allocated with room
 * for 2*n_primitives nodes; n_nodes receives the count. */
typedef struct cpbrt_bvh_dump {
    uint64_t  capacity;     /* in: node slots available in the arrays below */
    uint64_t  n_nodes;      /* out */
    uint64_t  root;         /* out */
    uint8_t*  is_leaf;      /* out, per node */
    uint8_t*  axis;         /* out, Parent.axis (0 for leaves) */
    uint32_t* a;            /* out, Leaf: range.start   Parent: left  */
    uint32_t* b;            /* out, Leaf: range.end     Parent: right */
    float*    bound;        /* out, 6 floats per node (min xyz, max xyz); zeros for leaves */
} cpbrt_bvh_dump;

typedef struct cpbrt_render_params {
    uint32_t sample_begin, sample_end;  /* this call renders samples [begin,end) of every pixel;
                                           0,0 = all of sampler.spp (or this rank's share when a
                                           communicator is attached) */
    uint32_t first_hit_only;            /* 1 = stop every path after bounce 0 (first-hit shading parity) */
    uint32_t resolve;                   /* 1 = divide by the valid-sample count (film.rs:16); 0 = leave sums */
    uint32_t samples_per_pass;          /* 0 = auto */
    uint32_t film_on_device;            /* 1 = film_rgb_out is a device pointer */
    uint32_t profile_stages;            /* 1 = bracket every stage with CUDA events and fill ms_trace / ms_shade /
                                           ms_raygen_film and the *_launches counts (a little slower) */
    uint32_t output_srgb8;              /* 1 = film_rgb_out receives W*H*3 BYTES: gamma_correct (src/math/mod.rs:51-57) and the
                                           `clamp(f * 255 + 0.5, 0, 255) as u8` of ImageTexture::into_file
                                           (src/texture/image.rs:124) are applied on the device, fused with the
                                           per-pixel mean; needs resolve = 1 */
} cpbrt_render_params;

typedef struct cpbrt_stats {
    uint64_t camera_samples;        /* W*H*(sample_end-sample_begin) */
    uint64_t rays_closest;          /* Scene::intersect calls of the reference (camera + continuation + MIS rays),
                                       whether answered by a traversal or elided (see rays_closest_traced) */
    uint64_t rays_any;              /* Scene::intersect_predicate calls (shadow rays) */
    uint64_t nan_dropped;           /* samples dropped by li.has_nan() (render.rs:34-37) */
    uint64_t kernel_launches;       /* CUDA kernels launched by this call */
    double   ms_total;              /* device time of the call (CUDA events) */
    double   ms_trace;              /* device time inside the traversal kernels (when profiled) */
    double   ms_shade;
    double   ms_raygen_film;
    double   ms_reduce;             /* NCCL film reduce + resolve */
    uint64_t trace_parent_visits;   /* only filled by cpbrt_trace_* with counting on */
    uint64_t trace_prim_tests;
    uint64_t trace_launches;        /* profile_stages: launches of the closest-hit traversal kernel */
    double   ms_trace_closest;      /* profile_stages: device time inside those launches */
    double   ms_trace_shadow;
    uint64_t rays_closest_traced;   /* closest-hit queries that needed a BVH traversal: MIS rays that miss their
                                       light's primitive and last-depth non-specular rays are answered without one */
    uint64_t shade_items;           /* hits handed to the shading stage, all bounces and material kinds (terminal hits included) */
    uint64_t trace_slow_paths;      /* rays the fast search could not settle by itself and handed to the reference-order traversal
                                       (a hit whose reachability could not be proven, zero / non-finite direction components) */
} cpbrt_stats;

/* ------------------------------------------------------------------------------------------
 * Device entry points.
 * ---------------------------------------------------------------------------------------- */
typedef struct cpbrt_ctx cpbrt_ctx;

/* One context per GPU (one process per GPU in the multi-GPU layout). device_id = CUDA ordinal. */
int  cpbrt_create(cpbrt_ctx** out, int device_id);
void cpbrt_destroy(cpbrt_ctx* ctx);
const char* cpbrt_last_error(const cpbrt_ctx* ctx);   /* ctx may be NULL: last create error */

/* Copies the description to the device.  Replaces `scene` argument of render(). */
int cpbrt_upload_scene(cpbrt_ctx* ctx, const cpbrt_scene_desc* desc);

/* Aggregate::build (src/aggregate/bvh.rs:194-207): exact full-sweep SAH.  prim_order_out[i]
 * (nullable, n_primitives entries) = uploaded index of the primitive in BVH slot i. */
int cpbrt_build_bvh(cpbrt_ctx* ctx, uint32_t* prim_order_out, cpbrt_bvh_dump* nodes_out);

/* Scene::intersect (src/scene.rs:20-22) / Scene::intersect_predicate (scene.rs:17-19) over a
 * ray batch.  flags: CPBRT_TRACE_ON_DEVICE = rays / results are device pointers;
 * CPBRT_TRACE_COUNT = also count Parent visits and primitive tests (slower kernel variant).
 * stats_out (nullable) receives the kernel's device time and, with COUNT, the counters. */
#define CPBRT_TRACE_ON_DEVICE 1
#define CPBRT_TRACE_COUNT 2
/* CPBRT_TRACE_COUNT_OWN = run the SHIPPED traversal (the schedule cpbrt_render uses for this scene) and count its own
 * node steps (one 64-byte Parent record each) and primitive tests (one 48-byte record each) into
 * trace_parent_visits / trace_prim_tests: the work the kernel really does, for its roofline.  Tree walks only
 * (scenes of more than 64 primitives); wins over CPBRT_TRACE_COUNT when both are set. */
#define CPBRT_TRACE_COUNT_OWN 4
int cpbrt_trace_closest(cpbrt_ctx* ctx, const cpbrt_ray* rays, uint64_t n, cpbrt_hit* hits_out,
                        int flags, cpbrt_stats* stats_out);
int cpbrt_trace_any(cpbrt_ctx* ctx, const cpbrt_ray* rays, uint64_t n, uint8_t* occluded_out,
                    int flags, cpbrt_stats* stats_out);

/* HaltonSampler::set_pixel + next_sample x sample[i] + get_sample at dimension dim[i]
 * (src/sampler/halton.rs:108-136).  index_out / value_out nullable. */
int cpbrt_halton(cpbrt_ctx* ctx, const uint32_t* px, const uint32_t* py, const uint32_t* sample,
                 const uint32_t* dim, uint64_t n, uint64_t* index_out, float* value_out);

/* Camera rays of render.rs:26-32 for sample `sample` of every pixel (row-major y*W+x), with
 * the running film offset of samples 0..sample replayed. */
int cpbrt_camera_rays(cpbrt_ctx* ctx, uint32_t sample, cpbrt_ray* rays_out, int on_device);

/* render() (src/render.rs:7-50).  film_rgb_out: W*H*3 f32, linear, row-major, y down = Film.pixels.
 * With a communicator attached, every rank renders its sample share and the sum lands on
 * rank 0 (other ranks' film_rgb_out may be NULL). */
int cpbrt_render(cpbrt_ctx* ctx, const cpbrt_render_params* params, float* film_rgb_out,
                 cpbrt_stats* stats_out);

/* How many levels of the last cpbrt_build_bvh ran level-synchronously on the device (0 = the exact SAH build ran on
 * the host only).  Scenes of >= 131072 primitives build the top of the tree (segments of >= 256 primitives) level by
 * level on the GPU and every smaller subtree in one thread block each; the result is identical to the host build
 * (bvh.rs:24-124). */
int cpbrt_bvh_gpu_levels(const cpbrt_ctx* ctx);
/* Subtrees of the last device build that the HOST builder had to finish (segments longer than one block handles, or
 * CPBRT_BVH_SMALL=0): 0 = the whole SAH build ran on the device; -1 = the last build was a pure host build. */
int cpbrt_bvh_host_segments(const cpbrt_ctx* ctx);

/* Tests / debugging: the traversal layout as it sits in device memory after cpbrt_build_bvh, copied to the host.
 *   nodes_out    16 floats per Parent: {L.min.xyz L.max.x}{L.max.yz R.min.xy}{R.min.z R.max.xyz}
 *                {left ref, right ref, left count | axis << 30, right count} (the last four are uint32 bit patterns;
 *                count 0 = the ref is a Parent index, otherwise a Leaf of `count` primitives starting at slot `ref`;
 *                a Leaf's box is the union of its primitives' bounds).  Parent 0 is the root.
 *   records_out  12 floats per primitive in BVH order (triangle vertices or sphere radius, upload index, meta, source)
 *   frames_out   12 floats per primitive (unit normal and the shading frame BSDF::new derives from it)
 *   uv_out       6 floats per primitive, only written when the scene carries non-default uv (return value 1)
 * Any output may be NULL; *n_nodes_out receives the Parent count (call once with NULL arrays to size them),
 * *resident_out 1 when the last build ran on the device from bounds to layout.  Parents are numbered differently by
 * the device and the host builders (same tree): compare by walking from Parent 0.
 * Returns 0 / 1 (uv present) or a negative CPBRT_ERR_*. */
int cpbrt_bvh_device_layout(cpbrt_ctx* ctx, uint64_t node_capacity, uint64_t* n_nodes_out, float* nodes_out,
                            float* records_out, float* frames_out, float* uv_out, int* resident_out);

/* Checked build only (libcpbrt compiled with -DCPBRT_CHECKED: every index into the path state, the queues, the
 * primitive / node / light / material arrays and the sampler tables is compared with its bound on the device):
 * number of violations since the library was loaded and the site code of the first one.  Returns 1 from a checked
 * build, 0 from a normal one (outputs zeroed), negative CPBRT_ERR_* on failure. */
int cpbrt_check_failures(cpbrt_ctx* ctx, uint64_t* n_failures_out, uint32_t* first_site_out);

/* Bytes this context has copied host -> device since cpbrt_create (counted where the copies are issued):
 * scene arrays, sampler tables, BVH layout, ray batches.  Bookkeeping for the end-to-end measurement. */
uint64_t cpbrt_h2d_bytes(const cpbrt_ctx* ctx);

/* Roofline denominators: read bandwidth of a zero-filled device buffer of `bytes` bytes
 * streamed `repeats` times by a read-only kernel (16-byte loads).  64 MB stays resident in the 126 MB
 * L2 and gives the L2 read peak; a multi-GB buffer gives the HBM read peak.  No reference equivalent. */
int cpbrt_measure_read_bandwidth(cpbrt_ctx* ctx, uint64_t bytes, uint32_t repeats, double* gb_per_s);

/* Multi-GPU: one NCCL communicator over the per-GPU contexts.  id_bytes = ncclUniqueId (128 B).
 * cpbrt_render then renders this rank's share of the sample indices and reduces the film to rank 0.
 * If the communicator is attached BEFORE cpbrt_build_bvh, scenes of >= 65536 primitives are built once,
 * on rank 0, and broadcast to the other ranks (every rank must then call cpbrt_build_bvh; the node dump
 * is only available on rank 0). */
#define CPBRT_COMM_ID_BYTES 128
int cpbrt_comm_unique_id(uint8_t id_bytes_out[CPBRT_COMM_ID_BYTES]);
int cpbrt_comm_init(cpbrt_ctx* ctx, int n_ranks, int rank, const uint8_t id_bytes[CPBRT_COMM_ID_BYTES]);

/* ------------------------------------------------------------------------------------------
 * Host side (no GPU needed): the reference's parser / Scene builder semantics
 * (src/scene_file_parser/{lex,mod}.rs, src/scene.rs, src/render.rs:63-78) producing the description.
 * ---------------------------------------------------------------------------------------- */
typedef struct cpbrt_scene cpbrt_scene;

/* read_scene + camera/sampler/film/integrator factories + SceneBuilder::build_with_clipper.
 * perm_seed seeds the Halton digit-permutation table (see cpbrt_make_permutation). */
int cpbrt_scene_load(const char* pbrt_path, uint64_t perm_seed, cpbrt_scene** out);
/* Same, from an in-memory scene text; base_dir resolves Include / texture paths. */
int cpbrt_scene_parse(const char* text, const char* base_dir, uint64_t perm_seed, cpbrt_scene** out);
void cpbrt_scene_free(cpbrt_scene* scene);
const cpbrt_scene_desc* cpbrt_scene_get_desc(const cpbrt_scene* scene);
const char* cpbrt_scene_film_filename(const cpbrt_scene* scene);   /* Film "string filename" */
const char* cpbrt_host_last_error(void);
/* Overrides (the reference takes these from the file only): */
int cpbrt_scene_set_spp(cpbrt_scene* scene, uint32_t spp);
int cpbrt_scene_set_integrator(cpbrt_scene* scene, uint32_t integrator, uint32_t max_depth);

/* Aggregate::build (src/aggregate/bvh.rs:194-207) on the host, no device involved: the same
 * exact-SAH builder cpbrt_build_bvh uses.  Outputs as for cpbrt_build_bvh. */
int cpbrt_host_build_bvh(const cpbrt_scene_desc* desc, uint32_t* prim_order_out, cpbrt_bvh_dump* nodes_out);

/* Fisher-Yates per prime block driven by a 64-bit LCG (x <- x*6364136223846793005 +
 * 1442695040888963407); stands in for the reference's thread_rng shuffle (halton.rs:216-231). */
void cpbrt_make_permutation(uint64_t seed, uint16_t* out /* CPBRT_PERM_TABLE_LEN */);

/* Tests: the digit-GROUP evaluation of scrambled_radical_inverse that the device uses (csrc/device/halton.cuh:
 * several digits per division and table look-up, tables built by the host for `group_max` = the largest base^digits
 * of a group, 2 .. 65536), run on the host for n (index, dimension) pairs: the integer accumulator ("reversed",
 * halton.rs:56-64, wrapping in 64 bits) and the digit count.  Must equal the one-digit loop for every group_max. */
int cpbrt_host_halton_grouped(const uint16_t* permutation, uint32_t group_max, const uint64_t* index, const uint32_t* dim, uint64_t n,
                              uint64_t* acc_out, uint32_t* digits_out);

/* The contiguous range [*begin, *end) of sample indices rank `rank` of `n_ranks` renders when cpbrt_render is
 * called without an explicit range.  The shares tile [0, spp) and shrink slightly with the rank: a rank that
 * starts at sample s first replays the film-point drift of the s samples before it (src/render.rs:26-31). */
void cpbrt_sample_share(uint32_t spp, int rank, int n_ranks, uint32_t* begin, uint32_t* end);

/* Film::write_image (src/film.rs:35-38 -> src/texture/image.rs:108-127): gamma_correct,
 * clamp(v*255+0.5,0,255) as u8, RGB8 PNG. */
int cpbrt_write_png(const char* path, const float* film_rgb, uint32_t width, uint32_t height);

/* render_from_file (src/render.rs:63-82): parse, clip, build, render on the GPU(s), write the
 * Film's PNG, print its file name.  device_id < 0 = device 0. */
int cpbrt_render_from_file(const char* pbrt_path, int device_id);

#ifdef __cplusplus
}
#endif
#endif /* CPBRT_H */
