// Device half of include/cpbrt.h: context, scene upload, BVH conversion, the wavefront render
// loop and the batch parity entry points.  There is deliberately no CPU path here: every
// entry point fails with an error code when no CUDA device is usable.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/cpbrt.h"
#include "bvh_build.h"
#include "host/scene_host.h"      // build_halton_groups
#include "device/wavefront.cuh"

namespace cpbrt {
// bvh_build_gpu.cu: the levels of the SAH build whose segments hold at least `min_len` primitives
bool gpu_sah_top_phase(const std::vector<BuildBounds>& bounds, uint32_t min_len, int max_levels, cudaStream_t stream, TopPhase* out, std::string* err);
bool gpu_sah_build_resident(const ResidentBuildInput& in, uint32_t min_len, int max_levels, cudaStream_t stream, ResidentBuildOutput* out, std::string* err);
}
using namespace cpbrt;

namespace {

thread_local std::string g_create_error;

struct DevBuf {
    void* p = nullptr; size_t bytes = 0;
    bool view = false;      // a piece of another allocation (the resident BVH build's block): nothing to free
    void release() { if (p && !view) cudaFree(p); p = nullptr; bytes = 0; view = false; }
    void set_view(void* q, size_t b) { release(); p = q; bytes = b; view = q != nullptr; }
};

// ---- NCCL through dlopen: the library is only needed when a communicator is attached -------
typedef struct ncclComm* ncclComm_t;
struct NcclUniqueId { char internal[128]; };
struct NcclApi {
    void* handle = nullptr;
    int (*GetUniqueId)(NcclUniqueId*) = nullptr;
    int (*CommInitRank)(ncclComm_t*, int, NcclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*Reduce)(const void*, void*, size_t, int, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Broadcast)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    bool load(std::string* err) {
        if (handle) return true;
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) { *err = std::string("cannot load libnccl: ") + dlerror(); return false; }
        GetUniqueId = (int (*)(NcclUniqueId*))dlsym(h, "ncclGetUniqueId");
        CommInitRank = (int (*)(ncclComm_t*, int, NcclUniqueId, int))dlsym(h, "ncclCommInitRank");
        CommDestroy = (int (*)(ncclComm_t))dlsym(h, "ncclCommDestroy");
        Reduce = (int (*)(const void*, void*, size_t, int, int, int, ncclComm_t, cudaStream_t))dlsym(h, "ncclReduce");
        Broadcast = (int (*)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t))dlsym(h, "ncclBroadcast");
        AllReduce = (int (*)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t))dlsym(h, "ncclAllReduce");
        GetErrorString = (const char* (*)(int))dlsym(h, "ncclGetErrorString");
        if (!GetUniqueId || !CommInitRank || !CommDestroy || !Reduce || !Broadcast || !AllReduce) {
            // the handle is only published once every entry point resolved: a later call fails the same way
            *err = "libnccl lacks a required symbol"; dlclose(h); return false;
        }
        handle = h;
        return true;
    }
};
NcclApi g_nccl;
const int kNcclFloat32 = 7, kNcclUint8 = 1, kNcclInt32 = 2, kNcclSum = 0, kNcclMax = 2;

}  // namespace

struct cpbrt_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;      // control stream: uploads, builds, batch entry points, start / end of a render
    std::string error;
    uint64_t h2d_bytes = 0;      // bytes copied host -> device by this context so far (counted where the copies are issued)

    // host copy of what the BVH build needs
    bool has_scene = false, has_bvh = false, has_general_materials = false, recursive_direct = false;
    std::vector<cpbrt_primitive> primitives;
    std::vector<cpbrt_triangle> triangles;
    std::vector<float> sphere_radius;
    std::vector<cpbrt_transform> transforms;
    BvhBuildResult bvh;
    int bvh_gpu_levels = 0;   // levels of the last BVH build that ran on the device
    bool bvh_resident = false; // the last build never left the device (no host arena: ctx->bvh holds the order only)
    std::vector<DevLight> host_lights;

    DevScene scene;   // by-value kernel argument
    DevBuf d_bvh_block;       // the resident build's single output allocation; d_prims / d_prim_uv / d_prim_frame / d_nodes are views into it then
    DevBuf d_prims, d_prim_uv, d_prim_frame, d_nodes, d_flat, d_env, d_transforms, d_materials, d_textures, d_texels, d_lights, d_perm, d_primes, d_primsum, d_magic, d_halton_dims, d_halton_inv_pow;

    // wavefront state: up to kMaxLanes independent pipelines ("lanes"), each with its own path state, queues,
    // streams and events, so that consecutive passes of a render overlap on the GPU (cpbrt_render)
    struct WaveLane {
        uint32_t capacity = 0;
        DevBuf d_path, d_kind, d_queues, d_counts;
        PathArrays P; Queues Q;
        cudaStream_t s[3] = {nullptr, nullptr, nullptr};        // main + two auxiliary streams of a bounce
        cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
        cudaEvent_t raygen_done = nullptr, film_done = nullptr; // order of the per-pixel state across lanes
    };
    static constexpr int kMaxLanes = 3;
    static constexpr int kDefaultLanes = 2;
    WaveLane lanes[kMaxLanes];
    DevBuf d_pixels;
    PixelArrays PX;
    Totals* d_totals = nullptr;
    DevBuf d_totals_buf, d_film_out;

    // multi-GPU
    ncclComm_t comm = nullptr; int n_ranks = 1, rank = 0;

    int fail(int code, const char* fmt, ...) {
        char buf[512];
        va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
        error = buf;
        return code;
    }
};

#define CU(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e__ = (call);                                                                        \
        if (e__ != cudaSuccess) return ctx->fail(CPBRT_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e__)); \
    } while (0)

namespace {

int alloc_upload(cpbrt_ctx* ctx, DevBuf& b, const void* src, size_t bytes) {
    const size_t src_bytes = bytes;      // an empty array still gets a 16-byte placeholder, but nothing is read from `src`
    if (bytes == 0) bytes = 16;
    if (b.bytes != bytes || b.view) {   // re-uploading a scene of the same shape reuses its buffers
        b.release();
        CU(cudaMalloc(&b.p, bytes));
        b.bytes = bytes;
    }
    if (src && src_bytes) { CU(cudaMemcpyAsync(b.p, src, src_bytes, cudaMemcpyHostToDevice, ctx->stream)); ctx->h2d_bytes += src_bytes; }
    else CU(cudaMemsetAsync(b.p, 0, bytes, ctx->stream));
    return 0;
}

float host_tri_area(const float p[9]) {   // triangle.rs:110-113
    float a[3] = {p[3] - p[0], p[4] - p[1], p[5] - p[2]}, b[3] = {p[6] - p[0], p[7] - p[1], p[8] - p[2]};
    float c[3] = {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
    return 0.5f * sqrtf((c[0] * c[0] + c[1] * c[1]) + c[2] * c[2]);
}

// Triangle::normal (triangle.rs:78-82): normalize(cross(p0 - p2, p1 - p2)), normalized once more by Normal3f::from;
// nalgebra's dot is (x*x + y*y) + z*z and normalize divides each component by the norm
void host_tri_normal(const float p[9], float out[3]) {
    const float a[3] = {p[0] - p[6], p[1] - p[7], p[2] - p[8]}, b[3] = {p[3] - p[6], p[4] - p[7], p[5] - p[8]};
    float c[3] = {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
    for (int pass = 0; pass < 2; ++pass) {
        const float len = sqrtf((c[0] * c[0] + c[1] * c[1]) + c[2] * c[2]);
        c[0] = c[0] / len; c[1] = c[1] / len; c[2] = c[2] / len;
    }
    out[0] = c[0]; out[1] = c[1]; out[2] = c[2];
}

// Triangle::normal plus the frame BSDF::new(n, n) derives from it (bxdf/mod.rs:83-97): snx, sny from
// coordinate_system (math/mod.rs:67-74), n in that frame (normalised).  Layout: DevScene::prim_frame.
void host_tri_frame(const float p[9], float4 out[3]) {
    float n[3];
    host_tri_normal(p, n);
    float x[3], y[3];
    if (fabsf(n[0]) > fabsf(n[1])) { const float l = sqrtf(n[0] * n[0] + n[2] * n[2]); x[0] = -n[2] / l; x[1] = 0.f / l; x[2] = n[0] / l; }
    else { const float l = sqrtf(n[1] * n[1] + n[2] * n[2]); x[0] = 0.f / l; x[1] = n[2] / l; x[2] = -n[1] / l; }
    y[0] = n[1] * x[2] - n[2] * x[1]; y[1] = n[2] * x[0] - n[0] * x[2]; y[2] = n[0] * x[1] - n[1] * x[0];
    float nl[3] = {(n[0] * x[0] + n[1] * x[1]) + n[2] * x[2], (n[0] * y[0] + n[1] * y[1]) + n[2] * y[2], (n[0] * n[0] + n[1] * n[1]) + n[2] * n[2]};
    const float len = sqrtf((nl[0] * nl[0] + nl[1] * nl[1]) + nl[2] * nl[2]);
    nl[0] = nl[0] / len; nl[1] = nl[1] / len; nl[2] = nl[2] / len;
    out[0] = make_float4(n[0], n[1], n[2], x[0]);
    out[1] = make_float4(x[1], x[2], y[0], y[1]);
    out[2] = make_float4(y[2], nl[0], nl[1], nl[2]);
}

float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

int grid_for(const cpbrt_ctx* ctx, int blocks_per_sm) { return ctx->sm_count * blocks_per_sm; }

// Warp schedule of the traversal kernels (bvh_trace.cuh): vote-scheduled node / primitive steps pay
// off once rays spend many node steps between leaves (1M triangles, incoherent rays: 1.5x), the
// descend-then-test loop is ahead on small trees (33-primitive Cornell box: 1.2x).
// CPBRT_SCHEDULE=vote|loop overrides (used to measure the crossover).
static constexpr uint32_t kVoteMinPrims = 256;    // measured: equal at 264 primitives, vote +12% at 4k, +23% at 40k, +32% at 1M (tools/schedule_crossover.py)
// Scenes of at most kFlatMaxPrims primitives skip the tree walk altogether (flat_trace.cuh).
// CPBRT_SCHEDULE=flat|loop|vote overrides ("flat" only where the tables exist).
// Tiny scenes: SCHED_FUSED additionally traces shadow and continuation rays inside the shading kernels (cpbrt_render
// only; the batch entry points treat it as SCHED_FLAT).  Measured and NOT the default: same film bit for bit, but the
// Cornell box drops from 829 to 685-735 Msamples/s (profiles/r02_fused_sweep.txt) -- the searches run at the warp's
// full instruction count whenever any lane has a ray, while the queues hand the separate trace kernels full warps.
// CPBRT_SCHEDULE=flat|fused|loop|vote overrides.
#ifndef CPBRT_DEFAULT_TINY_SCHEDULE
#define CPBRT_DEFAULT_TINY_SCHEDULE SCHED_FLAT
#endif
int trace_schedule(const cpbrt_ctx* ctx) {
    const char* e = getenv("CPBRT_SCHEDULE");
    const bool can_flat = ctx->scene.flat != nullptr;
    if (e && strcmp(e, "vote") == 0) return SCHED_VOTE;
    if (e && strcmp(e, "loop") == 0) return SCHED_LOOP;
    if (e && strcmp(e, "flat") == 0 && can_flat) return SCHED_FLAT;
    if (e && strcmp(e, "fused") == 0 && can_flat) return SCHED_FUSED;
    if (can_flat) return CPBRT_DEFAULT_TINY_SCHEDULE;
    return ctx->scene.n_prims >= kVoteMinPrims ? SCHED_VOTE : SCHED_LOOP;
}
size_t flat_bytes(const cpbrt_ctx* ctx) { return flat_smem_bytes(ctx->scene.flat_n, ctx->scene.flat_stride, ctx->scene.flat_pstride); }

}  // namespace

extern "C" {

const char* cpbrt_last_error(const cpbrt_ctx* ctx) { return ctx ? ctx->error.c_str() : g_create_error.c_str(); }

int cpbrt_create(cpbrt_ctx** out, int device_id) {
    if (!out) { g_create_error = "null out pointer"; return CPBRT_ERR_BAD_ARG; }
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        g_create_error = std::string("no usable CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
        return CPBRT_ERR_NO_DEVICE;
    }
    if (device_id < 0) device_id = 0;
    if (device_id >= n) { g_create_error = "device_id out of range"; return CPBRT_ERR_BAD_ARG; }
    e = cudaSetDevice(device_id);
    if (e != cudaSuccess) { g_create_error = std::string("cudaSetDevice: ") + cudaGetErrorString(e); return CPBRT_ERR_CUDA; }
    cpbrt_ctx* ctx = new cpbrt_ctx();
    ctx->device = device_id;
    cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device_id);
    e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { g_create_error = std::string("cudaStreamCreate: ") + cudaGetErrorString(e); delete ctx; return CPBRT_ERR_CUDA; }
    for (auto& lane : ctx->lanes) {
        for (int i = 0; i < 3 && e == cudaSuccess; ++i) e = cudaStreamCreateWithFlags(&lane.s[i], cudaStreamNonBlocking);
        for (int i = 0; i < 5 && e == cudaSuccess; ++i) e = cudaEventCreateWithFlags(&lane.ev[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&lane.raygen_done, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&lane.film_done, cudaEventDisableTiming);
    }
    if (e != cudaSuccess) { g_create_error = std::string("cudaStreamCreate / cudaEventCreate: ") + cudaGetErrorString(e); cpbrt_destroy(ctx); return CPBRT_ERR_CUDA; }
    memset(&ctx->scene, 0, sizeof ctx->scene);
    *out = ctx;
    return CPBRT_OK;
}

void cpbrt_destroy(cpbrt_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(ctx->comm);
    DevBuf* bufs[] = {&ctx->d_bvh_block, &ctx->d_prims, &ctx->d_prim_uv, &ctx->d_prim_frame, &ctx->d_nodes, &ctx->d_flat, &ctx->d_env, &ctx->d_transforms, &ctx->d_materials, &ctx->d_textures, &ctx->d_texels,
                      &ctx->d_lights, &ctx->d_perm, &ctx->d_primes, &ctx->d_primsum, &ctx->d_magic, &ctx->d_halton_dims, &ctx->d_halton_inv_pow,
                      &ctx->d_pixels, &ctx->d_totals_buf, &ctx->d_film_out};
    for (DevBuf* b : bufs) b->release();
    for (auto& lane : ctx->lanes) {
        lane.d_path.release(); lane.d_kind.release(); lane.d_queues.release(); lane.d_counts.release();
        for (cudaEvent_t ev : lane.ev) if (ev) cudaEventDestroy(ev);
        if (lane.raygen_done) cudaEventDestroy(lane.raygen_done);
        if (lane.film_done) cudaEventDestroy(lane.film_done);
        for (cudaStream_t st : lane.s) if (st) cudaStreamDestroy(st);
    }
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int cpbrt_upload_scene(cpbrt_ctx* ctx, const cpbrt_scene_desc* d) {
    if (!ctx) return CPBRT_ERR_BAD_CTX;
    if (!d || d->abi_version != CPBRT_ABI_VERSION) return ctx->fail(CPBRT_ERR_BAD_ARG, "null or wrong-version scene description");
    CU(cudaSetDevice(ctx->device));
    if (d->film_width == 0 || d->film_height == 0) return ctx->fail(CPBRT_ERR_BAD_ARG, "empty film");
    if (d->n_primitives >= 0x40000000ull) return ctx->fail(CPBRT_ERR_UNSUPPORTED, "more than 2^30 primitives");
    if (!d->sampler.permutation) return ctx->fail(CPBRT_ERR_BAD_ARG, "sampler.permutation is null");
    // validation: every index must be in range; unsupported features fail loudly
    // (big scenes: all host threads look for the first bad primitive, its message is then produced by the serial checks)
    uint64_t first_check = 0, last_check = d->n_primitives;
    if (d->n_primitives >= (1u << 18)) {
        std::atomic<uint64_t> first_bad(d->n_primitives);
        parallel_for(d->n_primitives, 1 << 16, [&](uint64_t lo, uint64_t hi) {
            for (uint64_t i = lo; i < hi && i < first_bad.load(std::memory_order_relaxed); ++i) {
                const cpbrt_primitive& p = d->primitives[i];
                const bool bad = p.shape.kind > CPBRT_SHAPE_SPHERE ||
                                 (p.shape.kind == CPBRT_SHAPE_TRIANGLE ? p.shape.data >= d->n_triangles : p.shape.data >= d->n_spheres) ||
                                 p.shape.xform >= (int32_t)d->n_transforms || (p.material < 0) == (p.light < 0) ||
                                 p.material >= (int32_t)d->n_materials || p.light >= (int32_t)d->n_lights;
                if (bad) { uint64_t cur = first_bad.load(); while (i < cur && !first_bad.compare_exchange_weak(cur, i)) {} break; }
            }
        });
        first_check = first_bad.load();
        last_check = std::min<uint64_t>(d->n_primitives, first_check + 1);      // nothing to re-check when all passed
    }
    for (uint64_t i = first_check; i < last_check; ++i) {
        const cpbrt_primitive& p = d->primitives[i];
        if (p.shape.kind > CPBRT_SHAPE_SPHERE) return ctx->fail(CPBRT_ERR_BAD_ARG, "primitive %llu: bad shape kind", (unsigned long long)i);
        if (p.shape.kind == CPBRT_SHAPE_TRIANGLE ? p.shape.data >= d->n_triangles : p.shape.data >= d->n_spheres)
            return ctx->fail(CPBRT_ERR_BAD_ARG, "primitive %llu: shape data index out of range", (unsigned long long)i);
        if (p.shape.xform >= (int32_t)d->n_transforms) return ctx->fail(CPBRT_ERR_BAD_ARG, "primitive %llu: transform out of range", (unsigned long long)i);
        if ((p.material < 0) == (p.light < 0)) return ctx->fail(CPBRT_ERR_BAD_ARG, "primitive %llu: needs exactly one of material / light", (unsigned long long)i);
        if (p.material >= (int32_t)d->n_materials || p.light >= (int32_t)d->n_lights) return ctx->fail(CPBRT_ERR_BAD_ARG, "primitive %llu: source out of range", (unsigned long long)i);
    }
    for (uint32_t i = 0; i < d->n_transforms; ++i) {
        int hops = 0;
        for (int32_t x = (int32_t)i; x >= 0; x = d->transforms[x].outer) {
            if (x >= (int32_t)d->n_transforms || ++hops > 8) return ctx->fail(CPBRT_ERR_UNSUPPORTED, "transform chain %u: longer than 8 links or out of range", i);
        }
    }
    bool has_general = false;
    for (uint32_t i = 0; i < d->n_materials; ++i) {
        const cpbrt_material& m = d->materials[i];
        if (m.type > CPBRT_MAT_MIX) return ctx->fail(CPBRT_ERR_BAD_ARG, "material %u: unknown type %u", i, m.type);
        // texture slots every material type reads (include/cpbrt.h); uber's urough / vrough (5, 6) are optional
        static const int need_n[7] = {2, 1, 3, 3, 9, 5, 1};
        for (int k = 0; k < need_n[m.type]; ++k) {
            const bool optional = m.type == CPBRT_MAT_UBER && (k == 5 || k == 6);
            if (m.tex[k] >= (int32_t)d->n_textures || (m.tex[k] < 0 && !optional)) return ctx->fail(CPBRT_ERR_BAD_ARG, "material %u: texture slot %d out of range", i, k);
        }
        if (m.type == CPBRT_MAT_MIX) {
            for (int k = 0; k < 2; ++k) if (m.sub[k] < 0 || m.sub[k] >= (int32_t)d->n_materials) return ctx->fail(CPBRT_ERR_BAD_ARG, "material %u: mix operand %d out of range", i, k);
        }
        if (m.type > CPBRT_MAT_GLASS) has_general = true;
    }
    // mix trees: no cycles, nesting no deeper than the scale stack of a lobe (bsdf_general.cuh)
    {
        std::vector<int> depth(d->n_materials, -1);   // -1 unknown, -2 on the stack
        std::vector<std::pair<uint32_t, int>> stack;
        for (uint32_t i = 0; i < d->n_materials; ++i) {
            if (depth[i] >= 0) continue;
            stack.assign(1, std::make_pair(i, 0));
            while (!stack.empty()) {
                const uint32_t mi = stack.back().first; const int phase = stack.back().second;
                const cpbrt_material& m = d->materials[mi];
                if (m.type != CPBRT_MAT_MIX) { depth[mi] = 0; stack.pop_back(); continue; }
                if (phase < 2) {
                    depth[mi] = -2;
                    stack.back().second = phase + 1;
                    const uint32_t c = (uint32_t)m.sub[phase];
                    if (depth[c] == -2) return ctx->fail(CPBRT_ERR_BAD_ARG, "material %u: mix materials form a cycle", mi);
                    if (depth[c] < 0) stack.push_back(std::make_pair(c, 0));
                } else {
                    depth[mi] = 1 + std::max(depth[m.sub[0]], depth[m.sub[1]]);
                    if (depth[mi] > kMaxScales) return ctx->fail(CPBRT_ERR_UNSUPPORTED, "material %u: mix nesting deeper than %d", mi, kMaxScales);
                    stack.pop_back();
                }
            }
        }
    }
    for (uint32_t i = 0; i < d->n_textures; ++i) {
        const cpbrt_texture& t = d->textures[i];
        if (t.kind == CPBRT_TEX_IMAGE && (t.width == 0 || t.height == 0 || t.offset + (uint64_t)t.width * t.height * t.channels > d->n_texels))
            return ctx->fail(CPBRT_ERR_BAD_ARG, "texture %u: texel range out of bounds", i);
    }
    for (uint32_t i = 0; i < d->n_lights; ++i) {
        const cpbrt_light& l = d->lights[i];
        if (l.type > CPBRT_LIGHT_INFINITE) return ctx->fail(CPBRT_ERR_BAD_ARG, "light %u: unknown type %u", i, l.type);
        if (l.type == CPBRT_LIGHT_INFINITE && (l.tex < 0 || l.tex >= (int32_t)d->n_textures)) return ctx->fail(CPBRT_ERR_BAD_ARG, "light %u: environment texture out of range", i);
        if (l.type == CPBRT_LIGHT_AREA) {
            if (l.shape.kind > CPBRT_SHAPE_SPHERE || (l.shape.kind == CPBRT_SHAPE_TRIANGLE ? l.shape.data >= d->n_triangles : l.shape.data >= d->n_spheres) ||
                l.shape.xform >= (int32_t)d->n_transforms)
                return ctx->fail(CPBRT_ERR_BAD_ARG, "light %u: bad shape", i);
        }
    }
    // directlighting recursing over delta lobes (direct_light.rs:26-31) is a ray TREE per sample: rendered
    // by k_direct_recursive, whose explicit stack bounds the depth
    bool recursive_direct = false;
    if (d->integrator == CPBRT_INTEGRATOR_DIRECT && d->max_depth > 1) {
        for (uint32_t i = 0; i < d->n_materials; ++i) {
            const uint32_t t = d->materials[i].type;
            if (t == CPBRT_MAT_MIRROR || t == CPBRT_MAT_GLASS || t == CPBRT_MAT_UBER) recursive_direct = true;
        }
        if (recursive_direct && d->max_depth > (uint32_t)kMaxDirectDepth)
            return ctx->fail(CPBRT_ERR_UNSUPPORTED, "directlighting over specular materials with maxdepth > %d", kMaxDirectDepth);
    }
    if (d->max_depth > 200) return ctx->fail(CPBRT_ERR_UNSUPPORTED, "max_depth > 200");

    ctx->has_scene = false; ctx->has_bvh = false; ctx->has_general_materials = has_general; ctx->recursive_direct = recursive_direct;
    ctx->primitives.assign(d->primitives, d->primitives + d->n_primitives);
    ctx->triangles.assign(d->triangles, d->triangles + d->n_triangles);
    ctx->sphere_radius.assign(d->sphere_radius, d->sphere_radius + d->n_spheres);
    ctx->transforms.assign(d->transforms, d->transforms + d->n_transforms);

    DevScene& S = ctx->scene;
    memset(&S, 0, sizeof S);
    S.n_prims = (uint32_t)d->n_primitives;
    S.n_lights = d->n_lights;
    S.integrator = d->integrator; S.max_depth = d->max_depth;
    S.width = d->film_width; S.height = d->film_height;
    memcpy(S.camera.raster_to_camera, d->camera.raster_to_camera, 64);
    memcpy(S.camera.camera_to_world, d->camera.camera_to_world, 64);
    S.camera.has_camera_to_world = d->camera.has_camera_to_world; S.camera.has_lens = d->camera.has_lens;
    S.camera.lens_radius = d->camera.lens_radius; S.camera.focal_distance = d->camera.focal_distance;

    // transforms / materials / textures / lights
    {
        std::vector<DevTransform> t(d->n_transforms);
        for (uint32_t i = 0; i < d->n_transforms; ++i) { memcpy(t[i].m, d->transforms[i].m, 64); memcpy(t[i].m_inv, d->transforms[i].m_inv, 64); t[i].outer = d->transforms[i].outer; t[i].pad[0] = t[i].pad[1] = t[i].pad[2] = 0; }
        if (int rc = alloc_upload(ctx, ctx->d_transforms, t.data(), t.size() * sizeof(DevTransform))) return rc;
        std::vector<DevMaterial> m(d->n_materials);
        for (uint32_t i = 0; i < d->n_materials; ++i) { m[i].type = d->materials[i].type; memcpy(m[i].tex, d->materials[i].tex, sizeof m[i].tex); m[i].sub[0] = d->materials[i].sub[0]; m[i].sub[1] = d->materials[i].sub[1]; }
        if (int rc = alloc_upload(ctx, ctx->d_materials, m.data(), m.size() * sizeof(DevMaterial))) return rc;
        std::vector<DevTexture> tx(d->n_textures);
        for (uint32_t i = 0; i < d->n_textures; ++i) {
            const cpbrt_texture& s = d->textures[i];
            tx[i].kind = s.kind; tx[i].channels = s.channels; memcpy(tx[i].value, s.value, 12); tx[i].width = s.width; tx[i].height = s.height; tx[i].pad = 0; tx[i].offset = s.offset;
        }
        if (int rc = alloc_upload(ctx, ctx->d_textures, tx.data(), tx.size() * sizeof(DevTexture))) return rc;
        if (int rc = alloc_upload(ctx, ctx->d_texels, d->texels, d->n_texels * sizeof(float))) return rc;
        std::vector<DevLight> L(d->n_lights);
        for (uint32_t i = 0; i < d->n_lights; ++i) {
            const cpbrt_light& s = d->lights[i];
            DevLight& o = L[i]; memset(&o, 0, sizeof o);
            o.type = s.type; o.shape_kind = s.shape.kind; o.shape_data = s.shape.data; o.shape_xform = s.type == CPBRT_LIGHT_AREA ? s.shape.xform : -1;
            memcpy(o.v, s.v, 12); memcpy(o.L, s.L, 12); o.tex = s.tex;
            if (s.type == CPBRT_LIGHT_AREA) {
                if (s.shape.kind == CPBRT_SHAPE_SPHERE) { o.radius = d->sphere_radius[s.shape.data]; o.area = o.radius * o.radius * 3.14159265358979323846f * 4.f; }   // sphere.rs:106-108
                else { memcpy(o.tri, d->triangles[s.shape.data].p, 36); o.area = host_tri_area(o.tri); host_tri_normal(o.tri, o.tri_n); }
            }
        }
        for (auto& l : L) l.slot = 0xFFFFFFFFu;
        // InfiniteAreaLight::new (infinite_area.rs:10-26): Distribution2D over luminance * sin(theta),
        // with theta = (row + 0.5) / rows exactly as written there (no factor pi)
        std::vector<float> env;
        for (uint32_t i = 0; i < d->n_lights; ++i) {
            if (d->lights[i].type != CPBRT_LIGHT_INFINITE) continue;
            const cpbrt_texture& t = d->textures[d->lights[i].tex];
            const uint32_t w = t.kind == CPBRT_TEX_IMAGE ? t.width : 1, h = t.kind == CPBRT_TEX_IMAGE ? t.height : 1;
            if (t.kind == CPBRT_TEX_IMAGE && t.channels != 3) return ctx->fail(CPBRT_ERR_BAD_ARG, "light %u: environment map must be a spectrum texture", i);
            L[i].env_w = w; L[i].env_h = h; L[i].env_off = env.size();
            env.resize(env.size() + 2ull * h + 2ull * h * w);
            float* row_pdf = env.data() + L[i].env_off; float* row_cdf = row_pdf + h; float* col_pdf = row_cdf + h; float* col_cdf = col_pdf + (size_t)h * w;
            std::vector<float> f(w), fsum(h);
            auto distribution = [](const float* fv, size_t len, float* pdf, float* cdf) {   // Distribution1D::from (distribution.rs:16-33)
                const float n = (float)len; float last = 0.f;
                for (size_t k = 0; k < len; ++k) { cdf[k] = last + fv[k] / n; last = cdf[k]; }
                const float f_sum = cdf[len - 1];
                for (size_t k = 0; k < len; ++k) { cdf[k] /= f_sum; pdf[k] = fv[k] / (n * f_sum); }
                return f_sum;
            };
            for (uint32_t y = 0; y < h; ++y) {
                const float theta = ((float)y + 0.5f) / (float)h;
                for (uint32_t x = 0; x < w; ++x) {
                    const float* px = t.kind == CPBRT_TEX_IMAGE ? d->texels + t.offset + ((size_t)x + (size_t)y * w) * 3 : t.value;
                    const float lum = (0.212671f * px[0] + 0.715160f * px[1]) + 0.072169f * px[2];   // RGBSpectrum::y
                    f[x] = lum * sinf(theta);
                }
                fsum[y] = distribution(f.data(), w, col_pdf + (size_t)y * w, col_cdf + (size_t)y * w);
            }
            distribution(fsum.data(), h, row_pdf, row_cdf);
        }
        if (int rc = alloc_upload(ctx, ctx->d_env, env.empty() ? nullptr : env.data(), env.size() * sizeof(float))) return rc;
        if (int rc = alloc_upload(ctx, ctx->d_lights, L.data(), L.size() * sizeof(DevLight))) return rc;
        ctx->host_lights = L;
    }
    // sampler tables: PRIMS / PRIMSUM (halton.rs:141-214) and the caller's permutation
    {
        std::vector<uint32_t> primes, sums(CPBRT_N_PRIMES, 0);
        for (uint32_t c = 2; primes.size() < CPBRT_N_PRIMES; ++c) {
            bool is_p = true;
            for (uint32_t q = 2; q * q <= c; ++q) if (c % q == 0) { is_p = false; break; }
            if (is_p) primes.push_back(c);
        }
        for (int i = 1; i < CPBRT_N_PRIMES; ++i) sums[i] = sums[i - 1] + primes[i - 1];
        if (int rc = alloc_upload(ctx, ctx->d_primes, primes.data(), primes.size() * 4)) return rc;
        if (int rc = alloc_upload(ctx, ctx->d_primsum, sums.data(), sums.size() * 4)) return rc;
        std::vector<uint64_t> magic(CPBRT_N_PRIMES);
        for (int i = 0; i < CPBRT_N_PRIMES; ++i) magic[i] = 0xFFFFFFFFFFFFFFFFull / primes[i] + 1;   // floor(2^64/p)+1 (p never divides 2^64)
        if (int rc = alloc_upload(ctx, ctx->d_magic, magic.data(), magic.size() * 8)) return rc;
        // Per-dimension constants of scrambled_radical_inverse (halton.rs:56-64), evaluated once here with the
        // reference's f32 operations instead of per sample on the device:
        //   tail[d]       = inv_base * perm_d(0) / (1 - inv_base)        (the infinite tail of permuted zero digits)
        //   inv_pow[d][c] = 1 / (base^c as f32)                           (usize pow, wrapping like release Rust)
        // and the digit-group tables (scene_dev.cuh, DevHaltonDim): behind the caller's permutation table, for every
        // dimension whose base^2 still fits kHaltonGroupMax, the permuted-and-reversed strings of 1, 2, ... g digits.
        {
            uint32_t group_max = 1024;      // measured: tools/variant_sweep.py / DESIGN.md 4.2
            if (const char* e = getenv("CPBRT_HALTON_GROUP")) group_max = (uint32_t)std::max(2, std::min(65536, atoi(e)));
            std::vector<uint16_t> table;        // what follows the caller's CPBRT_PERM_TABLE_LEN entries on the device
            const uint32_t table0 = CPBRT_PERM_TABLE_LEN;
            std::vector<cpbrt_host::HaltonGroupDim> groups;
            cpbrt_host::build_halton_groups(d->sampler.permutation, group_max, &groups, &table);      // host/halton_tables.cpp
            std::vector<DevHaltonDim> dims(CPBRT_N_PRIMES);
            std::vector<float> inv_pow((size_t)CPBRT_N_PRIMES * kHaltonMaxDigits);
            for (int i = 0; i < CPBRT_N_PRIMES; ++i) {
                const uint32_t b = primes[i];
                const float inv_base = 1.f / (float)b;
                memset(&dims[i], 0, sizeof dims[i]);
                dims[i].magic = magic[i]; dims[i].base = b; dims[i].primsum = sums[i];
                dims[i].tail = inv_base * (float)d->sampler.permutation[sums[i]] / (1.f - inv_base);
                dims[i].group = groups[i].group; dims[i].group_digits = groups[i].group_digits; dims[i].group_magic = groups[i].group_magic;
                dims[i].group_off = groups[i].group_off; dims[i].full_off = groups[i].full_off;
                uint64_t pw = 1;
                for (int c = 0; c < kHaltonMaxDigits; ++c) { inv_pow[(size_t)i * kHaltonMaxDigits + c] = 1.f / (float)pw; pw *= primes[i]; }
            }
            (void)table0;
            // the caller's table goes up as it is (7 MB: no host copy), the group tables behind it
            const size_t perm_bytes = (size_t)table0 * 2, total_bytes = perm_bytes + table.size() * 2;
            if (ctx->d_perm.bytes != total_bytes) { ctx->d_perm.release(); CU(cudaMalloc(&ctx->d_perm.p, total_bytes)); ctx->d_perm.bytes = total_bytes; }
            CU(cudaMemcpyAsync(ctx->d_perm.p, d->sampler.permutation, perm_bytes, cudaMemcpyHostToDevice, ctx->stream));
            if (!table.empty()) CU(cudaMemcpyAsync((char*)ctx->d_perm.p + perm_bytes, table.data(), table.size() * 2, cudaMemcpyHostToDevice, ctx->stream));
            ctx->h2d_bytes += total_bytes;
            S.sampler.table_len = (uint32_t)(total_bytes / 2);
            if (int rc = alloc_upload(ctx, ctx->d_halton_dims, dims.data(), dims.size() * sizeof(DevHaltonDim))) return rc;
            if (int rc = alloc_upload(ctx, ctx->d_halton_inv_pow, inv_pow.data(), inv_pow.size() * 4)) return rc;
        }
        DevSampler& sm = S.sampler;
        sm.spp = d->sampler.spp;
        for (int i = 0; i < 2; ++i) { sm.exp[i] = d->sampler.exp[i]; sm.scale[i] = d->sampler.scale[i]; sm.mult_inverse[i] = d->sampler.mult_inverse[i]; }
        sm.stride = sm.scale[0] * sm.scale[1];
        if (sm.scale[0] != (1ull << sm.exp[0])) return ctx->fail(CPBRT_ERR_BAD_ARG, "sampler.scale[0] must be 2^exp[0]");
        sm.permutation = (const uint16_t*)ctx->d_perm.p; sm.primes = (const uint32_t*)ctx->d_primes.p; sm.primsum = (const uint32_t*)ctx->d_primsum.p; sm.prime_magic = (const uint64_t*)ctx->d_magic.p;
        sm.dims = (const DevHaltonDim*)ctx->d_halton_dims.p; sm.inv_pow = (const float*)ctx->d_halton_inv_pow.p;
    }
    S.n_materials = d->n_materials; S.n_textures = d->n_textures; S.n_transforms = d->n_transforms;
    S.transforms = (const DevTransform*)ctx->d_transforms.p;
    S.materials = (const DevMaterial*)ctx->d_materials.p;
    S.textures = (const DevTexture*)ctx->d_textures.p;
    S.texels = (const float*)ctx->d_texels.p;
    S.lights = (const DevLight*)ctx->d_lights.p;
    S.env = (const float*)ctx->d_env.p;
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->has_scene = true;
    return CPBRT_OK;
}

// ---- one build for all ranks ---------------------------------------------------------------------
// With a communicator attached BEFORE the build, rank 0 builds the BVH and its device layout and the
// other ranks receive the finished arrays over NVLink (ncclBroadcast) instead of repeating the same
// build on the same host cores: eight ranks share one CPU, and the host part of a 10-million-primitive
// build would otherwise run eight times at an eighth of the threads.
struct BvhShareHeader { uint32_t n_prims, n_nodes, root_ref, root_cnt, any_uv, status, pad[2]; float root_box[6]; float pad2[2]; };
static bool share_bvh(const cpbrt_ctx* ctx, uint32_t n) {
    if (!ctx->comm || ctx->n_ranks <= 1 || !g_nccl.Broadcast) return false;
    const char* e = getenv("CPBRT_BVH_SHARE");
    return e ? atoi(e) != 0 : n >= (1u << 16);
}
static int bcast(cpbrt_ctx* ctx, void* buf, size_t bytes) {
    if (bytes == 0) return 0;
    const int rc = g_nccl.Broadcast(buf, buf, bytes, kNcclUint8, 0, ctx->comm, ctx->stream);
    if (rc != 0) return ctx->fail(CPBRT_ERR_NCCL, "ncclBroadcast: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "error");
    return 0;
}
// Both sides: header first (sizes + rank 0's status), then every rank's own status is agreed on with one
// all-reduce, and only if ALL ranks are fine do the array broadcasts follow.  A rank-local failure (rank 0's
// build, a mismatching scene, an allocation, a bad argument) therefore ends the call with an error on every
// rank after the same two collectives -- nobody is left waiting inside a broadcast.  `local_rc` is what this
// rank already knows to be wrong (its error text is in ctx->error).  Collectives that fail themselves
// (NCCL errors) leave the communicator unusable in any case.
static int exchange_bvh(cpbrt_ctx* ctx, uint32_t n, int local_rc, bool any_uv, const uint32_t* order_host, uint32_t* prim_order_out) {
    DevScene& S = ctx->scene;
    const bool root = ctx->rank == 0;
    DevBuf d_hdr, d_order;
    BvhShareHeader h; memset(&h, 0, sizeof h);
    if (root) { h.n_prims = n; h.n_nodes = S.n_nodes; h.root_ref = S.root_ref; h.root_cnt = S.root_cnt; h.any_uv = any_uv ? 1u : 0u; h.status = local_rc ? 1u : 0u; memcpy(h.root_box, S.root_box, sizeof h.root_box); }
    const std::string first_error = local_rc ? ctx->error : std::string();
    auto cleanup = [&]() { d_hdr.release(); d_order.release(); };
    // the header buffer carries one extra int32 pair for the status all-reduce
    struct { BvhShareHeader h; int32_t flag[4]; } up; up.h = h; up.flag[0] = up.flag[1] = up.flag[2] = up.flag[3] = 0;
    if (int rc = alloc_upload(ctx, d_hdr, &up, sizeof up)) { cleanup(); return rc; }      // cannot even talk: peers see the NCCL time-out / abort
    if (int rc = bcast(ctx, d_hdr.p, sizeof h)) { cleanup(); return rc; }
    int status = local_rc;
    if (!root) {
        cudaError_t e = cudaMemcpyAsync(&h, d_hdr.p, sizeof h, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) status = status ? status : ctx->fail(CPBRT_ERR_CUDA, "bvh header: %s", cudaGetErrorString(e));
        else if (h.status != 0) status = status ? status : ctx->fail(CPBRT_ERR_BAD_ARG, "rank 0 could not build the shared BVH");
        else if (h.n_prims != n) status = status ? status : ctx->fail(CPBRT_ERR_BAD_ARG, "rank 0 built a BVH over %u primitives, this rank uploaded %u", h.n_prims, n);
        if (!status) {
            S.n_nodes = h.n_nodes; S.root_ref = h.root_ref; S.root_cnt = h.root_cnt; memcpy(S.root_box, h.root_box, sizeof h.root_box);
            any_uv = h.any_uv != 0;
            status = alloc_upload(ctx, ctx->d_prims, nullptr, 3ull * n * sizeof(float4));
            if (!status) status = alloc_upload(ctx, ctx->d_prim_uv, nullptr, any_uv ? 6ull * n * 4 : 0);
            if (!status) status = alloc_upload(ctx, ctx->d_prim_frame, nullptr, 3ull * n * sizeof(float4));
            if (!status) status = alloc_upload(ctx, ctx->d_nodes, nullptr, 4ull * h.n_nodes * sizeof(float4));
            if (!status) status = alloc_upload(ctx, ctx->d_lights, ctx->host_lights.data(), ctx->host_lights.size() * sizeof(DevLight));
        }
    }
    if (!status) status = alloc_upload(ctx, d_order, root ? order_host : nullptr, (size_t)n * 4);
    // agreement: max over ranks of "something is wrong here"
    {
        int32_t* flag = reinterpret_cast<int32_t*>((char*)d_hdr.p + sizeof(BvhShareHeader));
        const int32_t mine = status ? 1 : 0;
        int32_t all = 1;
        cudaError_t e = cudaMemcpyAsync(flag, &mine, 4, cudaMemcpyHostToDevice, ctx->stream);
        int nrc = e == cudaSuccess ? g_nccl.AllReduce(flag, flag + 1, 1, kNcclInt32, kNcclMax, ctx->comm, ctx->stream) : 0;
        if (e == cudaSuccess && nrc == 0) e = cudaMemcpyAsync(&all, flag + 1, 4, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess && nrc == 0) e = cudaStreamSynchronize(ctx->stream);
        if (nrc != 0) { cleanup(); return ctx->fail(CPBRT_ERR_NCCL, "ncclAllReduce: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(nrc) : "error"); }
        if (e != cudaSuccess) { cleanup(); return ctx->fail(CPBRT_ERR_CUDA, "bvh exchange: %s", cudaGetErrorString(e)); }
        if (all != 0) {
            cleanup();
            if (status) { if (!first_error.empty()) ctx->error = first_error; return status; }
            return ctx->fail(CPBRT_ERR_BAD_ARG, "the shared BVH build failed on another rank");
        }
    }
    int rc = bcast(ctx, ctx->d_prims.p, 3ull * n * sizeof(float4));
    if (!rc && any_uv) rc = bcast(ctx, ctx->d_prim_uv.p, 6ull * n * 4);
    if (!rc) rc = bcast(ctx, ctx->d_prim_frame.p, 3ull * n * sizeof(float4));
    if (!rc) rc = bcast(ctx, ctx->d_nodes.p, 4ull * S.n_nodes * sizeof(float4));
    if (!rc) rc = bcast(ctx, ctx->d_lights.p, ctx->host_lights.size() * sizeof(DevLight));
    if (!rc) rc = bcast(ctx, d_order.p, (size_t)n * 4);
    if (!rc) {
        cudaError_t e = cudaSuccess;
        std::vector<uint32_t> order_rx;
        if (!root && n) { order_rx.resize(n); e = cudaMemcpyAsync(order_rx.data(), d_order.p, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream); }
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) rc = ctx->fail(CPBRT_ERR_CUDA, "bvh exchange: %s", cudaGetErrorString(e));
        if (!rc && !root && n) {
            if (prim_order_out) memcpy(prim_order_out, order_rx.data(), (size_t)n * 4);
        }
    }
    cleanup();
    if (!rc && !root) {
        S.prims = (const float4*)ctx->d_prims.p;
        S.prim_uv = any_uv ? (const float*)ctx->d_prim_uv.p : nullptr;
        S.prim_frame = n ? (const float4*)ctx->d_prim_frame.p : nullptr;
        S.nodes = (const float4*)ctx->d_nodes.p;
        S.lights = (const DevLight*)ctx->d_lights.p;
    }
    return rc;
}

// BVH slot of each area light's primitive (when exactly one primitive carries the light)
static int assign_light_slots(cpbrt_ctx* ctx, const std::vector<uint32_t>& order) {
    const uint32_t n = (uint32_t)order.size();
    const size_t nl = ctx->host_lights.size();
    std::vector<std::atomic<uint32_t>> count(nl), slot_of(nl);
    for (size_t i = 0; i < nl; ++i) { count[i].store(0); slot_of[i].store(0xFFFFFFFFu); }
    parallel_for(n, 1 << 16, [&](uint64_t lo, uint64_t hi) {
        for (uint32_t slot = (uint32_t)lo; slot < (uint32_t)hi; ++slot) {
            const cpbrt_primitive& p = ctx->primitives[order[slot]];
            if (p.light >= 0) { slot_of[p.light].store(slot); count[p.light].fetch_add(1); }
        }
    });
    for (size_t i = 0; i < nl; ++i) ctx->host_lights[i].slot = count[i].load() == 1 ? slot_of[i].load() : 0xFFFFFFFFu;
    if (int rc = alloc_upload(ctx, ctx->d_lights, ctx->host_lights.data(), nl * sizeof(DevLight))) return rc;
    ctx->scene.lights = (const DevLight*)ctx->d_lights.p;
    return 0;
}

// The build proper: SAH tree, device layout, uploads.  Three ways to the same tree (bvh.rs:24-124):
//   resident  big scenes, no node dump asked for: bounds, SAH levels, small subtrees, traversal layout and primitive
//             records all on the device (bvh_build_gpu.cu); the host only gets the primitive order back
//   hybrid    the same device phases, pieces downloaded and stitched into the host arena (node dumps, tests;
//             CPBRT_BVH_SMALL=0 additionally leaves the small subtrees to the host builder)
//   host      small scenes, NaN bounds, CPBRT_BVH_BUILD=host
static int build_bvh_local(cpbrt_ctx* ctx, uint32_t n, bool* any_uv_out, bool want_dump) {
    const bool timing = getenv("CPBRT_TIMING") != nullptr;
    auto now = []() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    double t_mark = now();
    auto lap = [&](const char* what) { if (timing) { const double t = now(); fprintf(stderr, "[cpbrt_build_bvh] %-28s %.3f s\n", what, t - t_mark); t_mark = t; } };
    DevScene& S = ctx->scene;
    // CPBRT_BVH_BUILD=host|gpu and CPBRT_BVH_GPU_MINLEN override the defaults (tests use both).
    const char* how = getenv("CPBRT_BVH_BUILD");
    const char* ml = getenv("CPBRT_BVH_GPU_MINLEN");
    const uint32_t min_len = ml ? (uint32_t)std::max(2, atoi(ml)) : 256u;   // measured flat between 16 and 1024 (tools/bvh_build_bench.py)
    bool on_gpu = (how ? strcmp(how, "gpu") == 0 : n >= (1u << 17)) && n >= min_len && n >= 2;
    ctx->bvh_gpu_levels = 0; ctx->bvh_resident = false;
    bool any_uv = false;
    {
        std::atomic<bool> found(false);
        parallel_for(n, 1 << 16, [&](uint64_t lo, uint64_t hi) {
            for (uint64_t i = lo; i < hi && !found.load(std::memory_order_relaxed); ++i) {
                const cpbrt_primitive& p = ctx->primitives[i];
                if (p.shape.kind != CPBRT_SHAPE_TRIANGLE) continue;
                const cpbrt_triangle& t = ctx->triangles[p.shape.data];
                if (!(t.uv[0][0] == 0.f && t.uv[0][1] == 0.f && t.uv[1][0] == 1.f && t.uv[1][1] == 0.f && t.uv[2][0] == 1.f && t.uv[2][1] == 1.f)) found.store(true);
            }
        });
        any_uv = found.load();
    }
    {
        const char* e = getenv("CPBRT_BVH_RESIDENT");
        if (on_gpu && !want_dump && (!e || atoi(e) != 0) && n > (uint32_t)kFlatMaxPrims) {
            ResidentBuildInput in;
            in.prims = ctx->primitives.data(); in.n_prims = n;
            in.tris = ctx->triangles.data(); in.n_tris = (uint32_t)ctx->triangles.size();
            in.radius = ctx->sphere_radius.data(); in.n_spheres = (uint32_t)ctx->sphere_radius.size();
            in.d_transforms = ctx->d_transforms.p; in.any_uv = any_uv;
            ResidentBuildOutput out; std::string why;
            if (gpu_sah_build_resident(in, min_len, 48, ctx->stream, &out, &why)) {
                lap("device build");
                ctx->h2d_bytes += out.h2d_bytes;
                ctx->d_prims.release(); ctx->d_prim_frame.release(); ctx->d_prim_uv.release(); ctx->d_nodes.release();
                ctx->d_bvh_block.release(); ctx->d_bvh_block.p = out.d_base; ctx->d_bvh_block.bytes = out.bytes;
                ctx->d_prims.set_view(out.d_records, 48ull * n);
                ctx->d_prim_frame.set_view(out.d_frames, 48ull * n);
                ctx->d_prim_uv.set_view(out.d_uv, any_uv ? 24ull * n : 0);
                ctx->d_nodes.set_view(out.d_nodes, 64ull * out.n_nodes);
                ctx->bvh = BvhBuildResult(); ctx->bvh.order.swap(out.order);
                ctx->bvh_gpu_levels = out.levels; ctx->bvh_resident = true;
                S.n_nodes = out.n_nodes; S.root_ref = 0; S.root_cnt = 0; memcpy(S.root_box, out.root_box, sizeof S.root_box);
                S.flat = nullptr; S.flat_n = 0; S.flat_stride = 0; S.flat_items = 0; S.flat_pstride = 0; S.flat_item_mask = 0; S.flat_special_mask = 0;
                if (int rc = assign_light_slots(ctx, ctx->bvh.order)) return rc;
                S.prims = (const float4*)ctx->d_prims.p;
                S.prim_uv = any_uv ? (const float*)ctx->d_prim_uv.p : nullptr;
                S.prim_frame = (const float4*)ctx->d_prim_frame.p;
                S.nodes = (const float4*)ctx->d_nodes.p;
                CU(cudaStreamSynchronize(ctx->stream));
                lap("light slots");
                *any_uv_out = any_uv;
                return CPBRT_OK;
            }
            if (timing) fprintf(stderr, "[cpbrt_build_bvh] resident build declined: %s\n", why.c_str());
            cudaGetLastError();
        }
    }
    // what a previous resident build left behind goes first: the four arrays are about to be uploaded afresh
    if (ctx->d_bvh_block.p) { ctx->d_prims.release(); ctx->d_prim_frame.release(); ctx->d_prim_uv.release(); ctx->d_nodes.release(); ctx->d_bvh_block.release(); }
    std::vector<BuildBounds> bounds;
    primitive_bounds(ctx->primitives.data(), n, ctx->triangles.data(), ctx->sphere_radius.data(), ctx->transforms.data(), &bounds);
    {
        if (on_gpu) {
            TopPhase top; std::string why;
            lap("primitive bounds");
            if (gpu_sah_top_phase(bounds, min_len, 48, ctx->stream, &top, &why)) {
                lap("device phases");
                ctx->bvh_gpu_levels = top.levels;
                finish_sah_bvh(std::move(top), 0, &ctx->bvh);
                lap("host stitching");
            } else {
                cudaGetLastError();
                on_gpu = false;
            }
        }
        if (!on_gpu) { build_sah_bvh(bounds, 0, &ctx->bvh); lap("host build"); }
    }
    const BvhBuildResult& R = ctx->bvh;

    // primitive records in BVH order
    std::vector<float4> prims(3ull * n);
    std::vector<float> uvs(any_uv ? 6ull * n : 0);
    std::vector<float4> frames(3ull * n);
    parallel_for(n, 1 << 16, [&](uint64_t slot_lo, uint64_t slot_hi) {
    for (uint32_t slot = (uint32_t)slot_lo; slot < (uint32_t)slot_hi; ++slot) {
        const uint32_t up = R.order[slot];
        const cpbrt_primitive& p = ctx->primitives[up];
        uint32_t meta = ((uint32_t)(p.shape.xform + 1)) << 2;
        float4 q0 = make_float4(0, 0, 0, 0), q1 = q0, q2 = q0;
        if (p.shape.kind == CPBRT_SHAPE_SPHERE) { meta |= kMetaSphere; q0.x = ctx->sphere_radius[p.shape.data]; }
        else {
            const cpbrt_triangle& t = ctx->triangles[p.shape.data];
            q0 = make_float4(t.p[0][0], t.p[0][1], t.p[0][2], t.p[1][0]);
            q1 = make_float4(t.p[1][1], t.p[1][2], t.p[2][0], t.p[2][1]);
            q2.x = t.p[2][2];
            if (any_uv) { meta |= kMetaHasUv; memcpy(&uvs[6ull * slot], t.uv, 24); }
            host_tri_frame(&t.p[0][0], &frames[3ull * slot]);
        }
        q2.y = u2f(up); q2.z = u2f(meta);
        q2.w = u2f(p.material >= 0 ? (uint32_t)p.material : (kSourceLight | (uint32_t)p.light));
        prims[3ull * slot] = q0; prims[3ull * slot + 1] = q1; prims[3ull * slot + 2] = q2;
    }
    });
    lap("primitive records");
    if (int rc = alloc_upload(ctx, ctx->d_prims, prims.data(), prims.size() * sizeof(float4))) return rc;
    if (int rc = alloc_upload(ctx, ctx->d_prim_uv, any_uv ? uvs.data() : nullptr, uvs.size() * 4)) return rc;
    if (int rc = alloc_upload(ctx, ctx->d_prim_frame, frames.data(), frames.size() * sizeof(float4))) return rc;

    // Parent nodes in depth-first order, far... the device layout (see scene_dev.cuh)
    S.n_nodes = 0; S.root_ref = 0; S.root_cnt = 0; memset(S.root_box, 0, sizeof S.root_box);
    std::vector<float4> nodes;
    if (n > 0) {
        const BuildNode& root = R.arena[R.root];
        if (root.leaf) { S.root_ref = root.a; S.root_cnt = root.b - root.a; }
        else {
            std::vector<uint32_t> dev_index(R.arena.size(), 0xFFFFFFFFu), dfs;
            std::vector<uint32_t> stack(1, R.root);
            while (!stack.empty()) {
                uint32_t a = stack.back(); stack.pop_back();
                const BuildNode& nd = R.arena[a];
                if (nd.leaf) continue;
                dev_index[a] = (uint32_t)dfs.size(); dfs.push_back(a);
                stack.push_back(nd.b); stack.push_back(nd.a);
            }
            nodes.resize(4ull * dfs.size());
            auto leaf_bound = [&](const BuildNode& leaf) {   // union of the leaf's primitive bounds
                BuildNode out = leaf;
                for (uint32_t sl = leaf.a; sl < leaf.b; ++sl) {
                    const BuildBounds& pb = bounds[R.order[sl]];
                    for (int c = 0; c < 3; ++c) {
                        out.bound.mn[c] = sl == leaf.a ? pb.mn[c] : std::min(out.bound.mn[c], pb.mn[c]);
                        out.bound.mx[c] = sl == leaf.a ? pb.mx[c] : std::max(out.bound.mx[c], pb.mx[c]);
                    }
                }
                return out;
            };
            parallel_for(dfs.size(), 1 << 15, [&](uint64_t k_lo, uint64_t k_hi) {
            for (uint32_t k = (uint32_t)k_lo; k < (uint32_t)k_hi; ++k) {
                const BuildNode& nd = R.arena[dfs[k]];
                const BuildNode l = R.arena[nd.a].leaf ? leaf_bound(R.arena[nd.a]) : R.arena[nd.a];
                const BuildNode r = R.arena[nd.b].leaf ? leaf_bound(R.arena[nd.b]) : R.arena[nd.b];
                uint32_t lref = l.leaf ? l.a : dev_index[nd.a], lcnt = l.leaf ? l.b - l.a : 0;
                uint32_t rref = r.leaf ? r.a : dev_index[nd.b], rcnt = r.leaf ? r.b - r.a : 0;
                nodes[4ull * k] = make_float4(l.bound.mn[0], l.bound.mn[1], l.bound.mn[2], l.bound.mx[0]);
                nodes[4ull * k + 1] = make_float4(l.bound.mx[1], l.bound.mx[2], r.bound.mn[0], r.bound.mn[1]);
                nodes[4ull * k + 2] = make_float4(r.bound.mn[2], r.bound.mx[0], r.bound.mx[1], r.bound.mx[2]);
                nodes[4ull * k + 3] = make_float4(u2f(lref), u2f(rref), u2f(lcnt | ((uint32_t)nd.axis << 30)), u2f(rcnt));
            }
            });
            S.n_nodes = (uint32_t)dfs.size();
            S.root_ref = 0; S.root_cnt = 0;
            for (int k = 0; k < 3; ++k) { S.root_box[k] = root.bound.mn[k]; S.root_box[3 + k] = root.bound.mx[k]; }
        }
    }
    lap("node layout");
    if (int rc = alloc_upload(ctx, ctx->d_nodes, nodes.empty() ? nullptr : nodes.data(), nodes.size() * sizeof(float4))) return rc;
    // tiny scenes: tables of the flat search (flat_trace.cuh): each untransformed triangle in the three
    // component rotations of the watertight test, and every primitive's world bounds
    S.flat = nullptr; S.flat_n = 0; S.flat_stride = 0; S.flat_items = 0; S.flat_pstride = 0; S.flat_item_mask = 0; S.flat_special_mask = 0;
    if (n > 0 && n <= (uint32_t)kFlatMaxPrims) {
        // phase-1 items: an untransformed triangle is paired with the first later one that shares two of its
        // vertices (bitwise); what is left stays single
        struct Item { uint32_t a, b; int sa0, sa1, ua, ub; };      // slots; vertex numbers of S0, S1, Ua in a, of Ub in b
        std::vector<Item> items;
        std::vector<char> used(n, 0);
        uint64_t special_mask = 0;
        auto tri_of = [&](uint32_t slot) -> const cpbrt_triangle* {
            const cpbrt_primitive& p = ctx->primitives[R.order[slot]];
            return (p.shape.kind == CPBRT_SHAPE_TRIANGLE && p.shape.xform < 0) ? &ctx->triangles[p.shape.data] : nullptr;
        };
        for (uint32_t a = 0; a < n; ++a) {
            const cpbrt_triangle* ta = tri_of(a);
            if (!ta) { special_mask |= 1ull << a; continue; }
            if (used[a]) continue;
            used[a] = 1;
            Item it; it.a = a; it.b = 0xFFFFFFFFu; it.sa0 = 0; it.sa1 = 1; it.ua = 2; it.ub = 2;
            for (uint32_t b = a + 1; b < n && it.b == 0xFFFFFFFFu; ++b) {
                const cpbrt_triangle* tb = tri_of(b);
                if (!tb || used[b]) continue;
                int match_a[3] = {-1, -1, -1}, n_match = 0;      // vertex of b equal to vertex i of a
                for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j)
                    if (match_a[i] < 0 && memcmp(ta->p[i], tb->p[j], 12) == 0) { match_a[i] = j; ++n_match; break; }
                if (n_match != 2) continue;
                int ua = 0; while (match_a[ua] >= 0) ++ua;
                const int s0 = (ua + 1) % 3, s1 = (ua + 2) % 3;
                const int ub = 3 - match_a[s0] - match_a[s1];
                if (match_a[s0] == match_a[s1]) continue;        // degenerate b
                it.b = b; it.sa0 = s0; it.sa1 = s1; it.ua = ua; it.ub = ub; used[b] = 1;
            }
            items.push_back(it);
        }
        if (items.size() <= 32) {
        const uint32_t n_items = (uint32_t)items.size();
        uint32_t stride = 3 * n, pstride = 3 * std::max(1u, n_items);
        while ((stride * 16u) % 128u != 48u) ++stride;          // the three copies start in different shared-memory banks
        while ((pstride * 16u) % 128u != 48u) ++pstride;
        const FlatLayout L = flat_layout(n, stride, pstride);
        std::vector<float4> flat(L.total, make_float4(0, 0, 0, 0));
        uint8_t* tab = reinterpret_cast<uint8_t*>(&flat[L.tab_off]);
        uint64_t item_mask = 0;
        for (uint32_t k = 0; k < n_items; ++k) {
            const Item& it = items[k];
            const cpbrt_triangle& ta = *tri_of(it.a);
            const float* S0 = ta.p[it.sa0]; const float* S1 = ta.p[it.sa1]; const float* Ua = ta.p[it.ua];
            const float* Ub = it.b != 0xFFFFFFFFu ? tri_of(it.b)->p[it.ub] : Ua;
            item_mask |= 1ull << (2 * k); tab[2 * k] = (uint8_t)it.a;
            if (it.b != 0xFFFFFFFFu) { item_mask |= 1ull << (2 * k + 1); tab[2 * k + 1] = (uint8_t)it.b; }
            for (int c = 0; c < 3; ++c) {      // permute(): (v[kz + 1], v[kz + 2], v[kz]) with kz = c
                const int kx = (c + 1) % 3, ky = (c + 2) % 3;
                float4* q = &flat[(size_t)c * pstride + 3ull * k];
                q[0] = make_float4(S0[kx], S0[ky], S0[c], S1[kx]);
                q[1] = make_float4(S1[ky], S1[c], Ua[kx], Ua[ky]);
                q[2] = make_float4(Ua[c], Ub[kx], Ub[ky], Ub[c]);
            }
        }
        for (uint32_t slot = 0; slot < n; ++slot) {
            if (const cpbrt_triangle* t = tri_of(slot)) {
                for (int c = 0; c < 3; ++c) {
                    const int kx = (c + 1) % 3, ky = (c + 2) % 3;
                    float4* q = &flat[L.tri_off + (size_t)c * stride + 3ull * slot];
                    q[0] = make_float4(t->p[0][kx], t->p[0][ky], t->p[0][c], t->p[1][kx]);
                    q[1] = make_float4(t->p[1][ky], t->p[1][c], t->p[2][kx], t->p[2][ky]);
                    q[2] = make_float4(t->p[2][c], 0.f, 0.f, 0.f);
                }
            }
            const BuildBounds& pb = bounds[R.order[slot]];
            flat[L.box_off + 2ull * slot] = make_float4(pb.mn[0], pb.mn[1], pb.mn[2], pb.mx[0]);
            flat[L.box_off + 2ull * slot + 1] = make_float4(pb.mx[1], pb.mx[2], 0.f, 0.f);
        }
        if (int rc = alloc_upload(ctx, ctx->d_flat, flat.data(), flat.size() * sizeof(float4))) return rc;
        S.flat = (const float4*)ctx->d_flat.p; S.flat_n = n; S.flat_stride = stride; S.flat_items = n_items; S.flat_pstride = pstride;
        S.flat_item_mask = item_mask; S.flat_special_mask = special_mask;
        }
        // the margins of the flat search scale with the scene bounds: also needed when the root is a single Leaf
        const BuildNode& root = R.arena[R.root];
        if (root.leaf) {
            for (uint32_t slot = 0; slot < n; ++slot) {
                const BuildBounds& pb = bounds[R.order[slot]];
                for (int k = 0; k < 3; ++k) {
                    S.root_box[k] = slot == 0 ? pb.mn[k] : std::min(S.root_box[k], pb.mn[k]);
                    S.root_box[3 + k] = slot == 0 ? pb.mx[k] : std::max(S.root_box[3 + k], pb.mx[k]);
                }
            }
        }
    }
    if (int rc = assign_light_slots(ctx, R.order)) return rc;
    S.prims = (const float4*)ctx->d_prims.p;
    S.prim_uv = any_uv ? (const float*)ctx->d_prim_uv.p : nullptr;
    S.prim_frame = n ? (const float4*)ctx->d_prim_frame.p : nullptr;
    S.nodes = (const float4*)ctx->d_nodes.p;
    CU(cudaStreamSynchronize(ctx->stream));
    lap("uploads");
    *any_uv_out = any_uv;
    return CPBRT_OK;
}

int cpbrt_build_bvh(cpbrt_ctx* ctx, uint32_t* prim_order_out, cpbrt_bvh_dump* dump) {
    if (!ctx) return CPBRT_ERR_BAD_CTX;
    if (!ctx->has_scene) return ctx->fail(CPBRT_ERR_BAD_CTX, "cpbrt_build_bvh before cpbrt_upload_scene");
    CU(cudaSetDevice(ctx->device));
    ctx->has_bvh = false;
    const uint32_t n = (uint32_t)ctx->primitives.size();
    const bool shared = share_bvh(ctx, n);
    const bool builds = !shared || ctx->rank == 0;
    // every argument / local check happens BEFORE the first collective, and whatever goes wrong on this rank is
    // carried into the exchange as a status instead of an early return (the peers would wait forever)
    int local_rc = 0; bool any_uv = false;
    if (builds) local_rc = build_bvh_local(ctx, n, &any_uv, dump != nullptr);
    else {
        ctx->bvh = BvhBuildResult(); ctx->bvh_gpu_levels = 0;
        if (dump) local_rc = ctx->fail(CPBRT_ERR_BAD_ARG, "the node dump of a shared build is only available on rank 0");
    }
    if (shared) {
        if (int rc = exchange_bvh(ctx, n, local_rc, any_uv, builds && !local_rc ? ctx->bvh.order.data() : nullptr, prim_order_out)) return rc;
    } else if (local_rc) return local_rc;
    ctx->has_bvh = true;
    if (builds && !export_bvh(ctx->bvh, prim_order_out, dump))
        return ctx->fail(CPBRT_ERR_BAD_ARG, "bvh dump capacity %llu < %llu nodes", (unsigned long long)dump->capacity, (unsigned long long)dump->n_nodes);
    return CPBRT_OK;
}

static int ensure_totals(cpbrt_ctx* ctx) {
    if (!ctx->d_totals) {
        if (int rc = alloc_upload(ctx, ctx->d_totals_buf, nullptr, sizeof(Totals) + 4 * sizeof(unsigned long long))) return rc;
        ctx->d_totals = (Totals*)ctx->d_totals_buf.p;
    }
    return 0;
}

static cudaError_t slow_paths_reset(cudaStream_t st) { const unsigned long long z = 0; return cudaMemcpyToSymbolAsync(g_trace_slow_paths, &z, sizeof z, 0, cudaMemcpyHostToDevice, st); }
static cudaError_t slow_paths_read(unsigned long long* out, cudaStream_t st) { return cudaMemcpyFromSymbolAsync(out, g_trace_slow_paths, sizeof *out, 0, cudaMemcpyDeviceToHost, st); }

static int check_overflow(cpbrt_ctx* ctx, const Totals& t) {
    if (t.overflow) return ctx->fail(CPBRT_ERR_UNSUPPORTED, "BVH deeper than the %d-entry traversal stack", kStackSize);
    if (t.bsdf_overflow) return ctx->fail(CPBRT_ERR_UNSUPPORTED, "a material expands to more than %d lobes of one kind", kMaxLobes);
    if (t.replay_mismatch) return ctx->fail(CPBRT_ERR_CUDA, "film-point replay: %llu table-driven radical inverses differ from the digit loop", t.replay_mismatch);
    return 0;
}

int cpbrt_trace_closest(cpbrt_ctx* ctx, const cpbrt_ray* rays, uint64_t n, cpbrt_hit* hits_out, int flags, cpbrt_stats* stats) {
    const int on_device = flags & CPBRT_TRACE_ON_DEVICE; const bool count = (flags & CPBRT_TRACE_COUNT) != 0, count_own = (flags & CPBRT_TRACE_COUNT_OWN) != 0;
    if (!ctx) return CPBRT_ERR_BAD_CTX;
    if (!ctx->has_bvh) return ctx->fail(CPBRT_ERR_BAD_CTX, "cpbrt_trace_closest before cpbrt_build_bvh");
    if ((!rays || !hits_out) && n) return ctx->fail(CPBRT_ERR_BAD_ARG, "null ray / hit buffer");
    if (n >= 0xFFFFFFFFull) return ctx->fail(CPBRT_ERR_UNSUPPORTED, "more than 2^32-1 rays in one batch");
    if (n == 0) { if (stats) memset(stats, 0, sizeof *stats); return CPBRT_OK; }
    CU(cudaSetDevice(ctx->device));
    if (int rc = ensure_totals(ctx)) return rc;
    CU(cudaMemsetAsync(ctx->d_totals, 0, ctx->d_totals_buf.bytes, ctx->stream));
    unsigned long long* counters = (unsigned long long*)(ctx->d_totals + 1);
    const float4* d_rays = (const float4*)rays; float4* d_hits = (float4*)hits_out;
    DevBuf br, bh;
    if (!on_device) {
        if (int rc = alloc_upload(ctx, br, rays, n * sizeof(cpbrt_ray))) return rc;
        if (int rc = alloc_upload(ctx, bh, nullptr, n * sizeof(cpbrt_hit))) { br.release(); return rc; }
        d_rays = (const float4*)br.p; d_hits = (float4*)bh.p;
    }
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    slow_paths_reset(ctx->stream);
    cudaEventRecord(e0, ctx->stream);
    if (n) {
        const int sched = trace_schedule(ctx) == SCHED_FUSED ? SCHED_FLAT : trace_schedule(ctx);
        if (count_own && sched == SCHED_VOTE) k_batch_closest<4><<<grid_for(ctx, 8), 128, 0, ctx->stream>>>(ctx->scene, d_rays, n, d_hits, counters, ctx->d_totals);
        else if (count_own && sched == SCHED_LOOP) k_batch_closest<5><<<grid_for(ctx, 8), 128, 0, ctx->stream>>>(ctx->scene, d_rays, n, d_hits, counters, ctx->d_totals);
        else if (count) k_batch_closest<1><<<grid_for(ctx, 8), 128, 0, ctx->stream>>>(ctx->scene, d_rays, n, d_hits, counters, ctx->d_totals);
        else if (sched == SCHED_VOTE) k_batch_closest<2><<<grid_for(ctx, 8), 128, 0, ctx->stream>>>(ctx->scene, d_rays, n, d_hits, counters, ctx->d_totals);
        else if (sched == SCHED_FLAT) k_batch_closest<3><<<grid_for(ctx, CPBRT_BLOCKS_FLAT), 128, flat_bytes(ctx), ctx->stream>>>(ctx->scene, d_rays, n, d_hits, counters, ctx->d_totals);
        else k_batch_closest<0><<<grid_for(ctx, 8), 128, 0, ctx->stream>>>(ctx->scene, d_rays, n, d_hits, counters, ctx->d_totals);
    }
    cudaEventRecord(e1, ctx->stream);
    cudaError_t err = cudaGetLastError();
    if (err == cudaSuccess && !on_device && n) err = cudaMemcpyAsync(hits_out, d_hits, n * sizeof(cpbrt_hit), cudaMemcpyDeviceToHost, ctx->stream);
    Totals t; unsigned long long c[2] = {0, 0}, slow = 0;
    if (err == cudaSuccess) err = slow_paths_read(&slow, ctx->stream);
    if (err == cudaSuccess) err = cudaMemcpyAsync(&t, ctx->d_totals, sizeof t, cudaMemcpyDeviceToHost, ctx->stream);
    if (err == cudaSuccess) err = cudaMemcpyAsync(c, counters, sizeof c, cudaMemcpyDeviceToHost, ctx->stream);
    if (err == cudaSuccess) err = cudaStreamSynchronize(ctx->stream);
    float ms = 0.f; if (err == cudaSuccess) cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    br.release(); bh.release();
    if (err != cudaSuccess) return ctx->fail(CPBRT_ERR_CUDA, "trace_closest: %s", cudaGetErrorString(err));
    if (stats) { memset(stats, 0, sizeof *stats); stats->rays_closest = n; stats->kernel_launches = n ? 1 : 0; stats->ms_total = ms; stats->ms_trace = ms; stats->trace_parent_visits = c[0]; stats->trace_prim_tests = c[1]; stats->trace_slow_paths = slow; }
    return check_overflow(ctx, t);
}

int cpbrt_trace_any(cpbrt_ctx* ctx, const cpbrt_ray* rays, uint64_t n, uint8_t* out, int flags, cpbrt_stats* stats) {
    const int on_device = flags & CPBRT_TRACE_ON_DEVICE; const bool count = (flags & CPBRT_TRACE_COUNT) != 0, count_own = (flags & CPBRT_TRACE_COUNT_OWN) != 0;
    if (!ctx) return CPBRT_ERR_BAD_CTX;
    if (!ctx->has_bvh) return ctx->fail(CPBRT_ERR_BAD_CTX, "cpbrt_trace_any before cpbrt_build_bvh");
    if ((!rays || !out) && n) return ctx->fail(CPBRT_ERR_BAD_ARG, "null ray / result buffer");
    if (n >= 0xFFFFFFFFull) return ctx->fail(CPBRT_ERR_UNSUPPORTED, "more than 2^32-1 rays in one batch");
    if (n == 0) { if (stats) memset(stats, 0, sizeof *stats); return CPBRT_OK; }
    CU(cudaSetDevice(ctx->device));
    if (int rc = ensure_totals(ctx)) return rc;
    CU(cudaMemsetAsync(ctx->d_totals, 0, ctx->d_totals_buf.bytes, ctx->stream));
    unsigned long long* counters = (unsigned long long*)(ctx->d_totals + 1);
    const float4* d_rays = (const float4*)rays; uint8_t* d_out = out;
    DevBuf br, bo;
    if (!on_device) {
        if (int rc = alloc_upload(ctx, br, rays, n * sizeof(cpbrt_ray))) return rc;
        if (int rc = alloc_upload(ctx, bo, nullptr, n)) { br.release(); return rc; }
        d_rays = (const float4*)br.p; d_out = (uint8_t*)bo.p;
    }
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    slow_paths_reset(ctx->stream);
    cudaEventRecord(e0, ctx->stream);
    if (n) {
        const int sched = trace_schedule(ctx) == SCHED_FUSED ? SCHED_FLAT : trace_schedule(ctx);
        if (count_own && sched == SCHED_VOTE) k_batch_any<4><<<grid_for(ctx, 8), 128, 0, ctx->stream>>>(ctx->scene, d_rays, n, d_out, counters, ctx->d_totals);
        else if (count_own && sched == SCHED_LOOP) k_batch_any<5><<<grid_for(ctx, 8), 128, 0, ctx->stream>>>(ctx->scene, d_rays, n, d_out, counters, ctx->d_totals);
        else if (count) k_batch_any<1><<<grid_for(ctx, 8), 128, 0, ctx->stream>>>(ctx->scene, d_rays, n, d_out, counters, ctx->d_totals);
        else if (sched == SCHED_VOTE) k_batch_any<2><<<grid_for(ctx, 8), 128, 0, ctx->stream>>>(ctx->scene, d_rays, n, d_out, counters, ctx->d_totals);
        else if (sched == SCHED_FLAT) k_batch_any<3><<<grid_for(ctx, CPBRT_BLOCKS_FLAT), 128, flat_bytes(ctx), ctx->stream>>>(ctx->scene, d_rays, n, d_out, counters, ctx->d_totals);
        else k_batch_any<0><<<grid_for(ctx, 8), 128, 0, ctx->stream>>>(ctx->scene, d_rays, n, d_out, counters, ctx->d_totals);
    }
    cudaEventRecord(e1, ctx->stream);
    cudaError_t err = cudaGetLastError();
    if (err == cudaSuccess && !on_device && n) err = cudaMemcpyAsync(out, d_out, n, cudaMemcpyDeviceToHost, ctx->stream);
    Totals t; unsigned long long c[2] = {0, 0}, slow = 0;
    if (err == cudaSuccess) err = slow_paths_read(&slow, ctx->stream);
    if (err == cudaSuccess) err = cudaMemcpyAsync(&t, ctx->d_totals, sizeof t, cudaMemcpyDeviceToHost, ctx->stream);
    if (err == cudaSuccess) err = cudaMemcpyAsync(c, counters, sizeof c, cudaMemcpyDeviceToHost, ctx->stream);
    if (err == cudaSuccess) err = cudaStreamSynchronize(ctx->stream);
    float ms = 0.f; if (err == cudaSuccess) cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    br.release(); bo.release();
    if (err != cudaSuccess) return ctx->fail(CPBRT_ERR_CUDA, "trace_any: %s", cudaGetErrorString(err));
    if (stats) { memset(stats, 0, sizeof *stats); stats->rays_any = n; stats->kernel_launches = n ? 1 : 0; stats->ms_total = ms; stats->ms_trace = ms; stats->trace_parent_visits = c[0]; stats->trace_prim_tests = c[1]; stats->trace_slow_paths = slow; }
    return check_overflow(ctx, t);
}

int cpbrt_halton(cpbrt_ctx* ctx, const uint32_t* px, const uint32_t* py, const uint32_t* sample, const uint32_t* dim, uint64_t n,
                 uint64_t* index_out, float* value_out) {
    if (!ctx) return CPBRT_ERR_BAD_CTX;
    if (!ctx->has_scene) return ctx->fail(CPBRT_ERR_BAD_CTX, "cpbrt_halton before cpbrt_upload_scene");
    if (n == 0) return CPBRT_OK;
    if (!px || !py || !sample || !dim) return ctx->fail(CPBRT_ERR_BAD_ARG, "null input array");
    CU(cudaSetDevice(ctx->device));
    DevBuf b[4], bi, bv;
    const uint32_t* src[4] = {px, py, sample, dim};
    int rc = 0;
    for (int i = 0; i < 4 && !rc; ++i) rc = alloc_upload(ctx, b[i], src[i], n * 4);
    if (!rc && index_out) rc = alloc_upload(ctx, bi, nullptr, n * 8);
    if (!rc && value_out) rc = alloc_upload(ctx, bv, nullptr, n * 4);
    cudaError_t err = cudaSuccess;
    unsigned long long mismatches = 0;
    if (!rc) rc = ensure_totals(ctx);
    if (!rc) {
        unsigned long long* d_mismatch = (unsigned long long*)(ctx->d_totals + 1);
        err = cudaMemsetAsync(d_mismatch, 0, sizeof(unsigned long long), ctx->stream);
        k_batch_halton<<<grid_for(ctx, 4), 256, 0, ctx->stream>>>(ctx->scene, (const uint32_t*)b[0].p, (const uint32_t*)b[1].p, (const uint32_t*)b[2].p,
                                                                  (const uint32_t*)b[3].p, n, (unsigned long long*)bi.p, (float*)bv.p, d_mismatch);
        if (err == cudaSuccess) err = cudaGetLastError();
        if (err == cudaSuccess) err = cudaMemcpyAsync(&mismatches, d_mismatch, sizeof mismatches, cudaMemcpyDeviceToHost, ctx->stream);
        if (err == cudaSuccess && index_out) err = cudaMemcpyAsync(index_out, bi.p, n * 8, cudaMemcpyDeviceToHost, ctx->stream);
        if (err == cudaSuccess && value_out) err = cudaMemcpyAsync(value_out, bv.p, n * 4, cudaMemcpyDeviceToHost, ctx->stream);
        if (err == cudaSuccess) err = cudaStreamSynchronize(ctx->stream);
    }
    for (int i = 0; i < 4; ++i) b[i].release();
    bi.release(); bv.release();
    if (rc) return rc;
    if (err != cudaSuccess) return ctx->fail(CPBRT_ERR_CUDA, "halton: %s", cudaGetErrorString(err));
    if (mismatches) return ctx->fail(CPBRT_ERR_CUDA, "halton: %llu values differ between the interleaved and the one-at-a-time evaluation", mismatches);
    return CPBRT_OK;
}

int cpbrt_camera_rays(cpbrt_ctx* ctx, uint32_t sample, cpbrt_ray* rays_out, int on_device) {
    if (!ctx) return CPBRT_ERR_BAD_CTX;
    if (!ctx->has_scene) return ctx->fail(CPBRT_ERR_BAD_CTX, "cpbrt_camera_rays before cpbrt_upload_scene");
    if (!rays_out) return ctx->fail(CPBRT_ERR_BAD_ARG, "null output");
    CU(cudaSetDevice(ctx->device));
    const size_t npix = (size_t)ctx->scene.width * ctx->scene.height;
    DevBuf b; float4* d = (float4*)rays_out;
    if (!on_device) { if (int rc = alloc_upload(ctx, b, nullptr, npix * sizeof(cpbrt_ray))) return rc; d = (float4*)b.p; }
    k_batch_camera_rays<<<grid_for(ctx, 4), 256, 0, ctx->stream>>>(ctx->scene, sample, d);
    cudaError_t err = cudaGetLastError();
    if (err == cudaSuccess && !on_device) err = cudaMemcpyAsync(rays_out, d, npix * sizeof(cpbrt_ray), cudaMemcpyDeviceToHost, ctx->stream);
    if (err == cudaSuccess) err = cudaStreamSynchronize(ctx->stream);
    b.release();
    if (err != cudaSuccess) return ctx->fail(CPBRT_ERR_CUDA, "camera_rays: %s", cudaGetErrorString(err));
    return CPBRT_OK;
}

// ---- wavefront state ---------------------------------------------------------------------------
static int ensure_wavefront(cpbrt_ctx* ctx, uint32_t n_paths, int n_lanes) {
    const uint32_t npix = ctx->scene.width * ctx->scene.height;
    if (ctx->d_pixels.bytes < (size_t)npix * 32) {
        if (int rc = alloc_upload(ctx, ctx->d_pixels, nullptr, (size_t)npix * 32)) return rc;
    }
    char* pp = (char*)ctx->d_pixels.p;
    ctx->PX.base_index = (uint2*)pp; ctx->PX.film_point = (float2*)(pp + (size_t)npix * 8); ctx->PX.sum = (float4*)(pp + (size_t)npix * 16);
    for (int li = 0; li < n_lanes; ++li) {
        cpbrt_ctx::WaveLane& L = ctx->lanes[li];
        if (L.capacity < n_paths) {
            L.capacity = 0;      // stays 0 if any allocation below fails: the next call starts over instead of trusting freed buffers
            if (int rc = alloc_upload(ctx, L.d_path, nullptr, kPathBytes * n_paths)) return rc;
            if (int rc = alloc_upload(ctx, L.d_kind, nullptr, (size_t)n_paths + 4)) return rc;   // k_classify reads whole 32-bit words
            CU(cudaMemsetAsync(L.d_kind.p, 0xFF, (size_t)n_paths + 4, ctx->stream));
            L.P.shade_kind = (uint8_t*)L.d_kind.p;
            // a path can have a MIS ray and a continuation ray queued at once: 2 entries per path
            if (int rc = alloc_upload(ctx, L.d_queues, nullptr, (size_t)n_paths * 4 * (5 + SHADE_KINDS))) return rc;
            L.capacity = n_paths;
            char* p = (char*)L.d_path.p;
            const size_t rec = (size_t)n_paths * 32;
            PathArrays& P = L.P;
            P.ray_o.base = (float4*)p; P.ray_d.base = (float4*)p + 1; P.hit_t.base = p; p += rec;
            P.beta.base = (float4*)p; P.L.base = (float4*)p + 1; p += rec;
            P.sh_o.base = (float4*)p; P.sh_d.base = (float4*)p + 1; p += rec;
            P.mis_o.base = (float4*)p; P.mis_d.base = (float4*)p + 1; p += rec;
            P.mis_A.base = (float4*)p; P.mis_p.base = (float4*)p + 1; p += rec;
            P.nee_light.base = (float4*)p; P.nee_bsdf.base = (float4*)p + 1; p += rec;
            P.beta_prev.base = (float4*)p; P.state2.base = (float4*)p + 1; P.smp_index.base = p; P.smp_dim.base = p; P.hit_slot.base = p;
#ifdef CPBRT_CHECKED
            {
                const uint32_t cap = n_paths;
                P.ray_o.cap = P.ray_d.cap = P.hit_t.cap = P.beta.cap = P.L.cap = P.sh_o.cap = P.sh_d.cap = P.mis_o.cap = P.mis_d.cap = P.mis_A.cap = P.mis_p.cap =
                    P.nee_light.cap = P.nee_bsdf.cap = P.beta_prev.cap = P.state2.cap = P.smp_index.cap = P.smp_dim.cap = P.hit_slot.cap = cap;
                L.Q.n_paths = n_paths;
            }
#endif
            uint32_t* q = (uint32_t*)L.d_queues.p;
            L.Q.closest[0] = q; q += 2 * (size_t)n_paths; L.Q.closest[1] = q; q += 2 * (size_t)n_paths; L.Q.shadow = q; q += n_paths;
            for (int k = 0; k < SHADE_KINDS; ++k) { L.Q.shade[k] = q; q += n_paths; }
        }
        if (!L.d_counts.p) { if (int rc = alloc_upload(ctx, L.d_counts, nullptr, 64 * 4)) return rc; }
        L.Q.counts = (uint32_t*)L.d_counts.p;
    }
    return ensure_totals(ctx);
}

int cpbrt_render(cpbrt_ctx* ctx, const cpbrt_render_params* params, float* film_out, cpbrt_stats* stats) {
    if (!ctx) return CPBRT_ERR_BAD_CTX;
    if (!ctx->has_bvh) return ctx->fail(CPBRT_ERR_BAD_CTX, "cpbrt_render before cpbrt_build_bvh");
    CU(cudaSetDevice(ctx->device));
    cpbrt_render_params rp; memset(&rp, 0, sizeof rp); rp.resolve = 1;
    if (params) rp = *params;
    DevScene S = ctx->scene;
    if (rp.first_hit_only) { S.integrator = CPBRT_INTEGRATOR_DIRECT; S.max_depth = 0; }
    const uint32_t spp = S.sampler.spp;
    uint32_t s_begin = rp.sample_begin, s_end = rp.sample_end;
    if (s_begin == 0 && s_end == 0) {
        // whole job, or this rank's contiguous share of the sample indices
        cpbrt_sample_share(spp, ctx->rank, ctx->n_ranks, &s_begin, &s_end);
    }
    if (s_end > spp || s_begin > s_end) return ctx->fail(CPBRT_ERR_BAD_ARG, "sample range [%u,%u) outside [0,%u)", s_begin, s_end, spp);
    const bool root = ctx->rank == 0;
    if (root && !film_out) return ctx->fail(CPBRT_ERR_BAD_ARG, "film_rgb_out is null on rank 0");
    if (ctx->comm && !rp.resolve) return ctx->fail(CPBRT_ERR_BAD_ARG, "resolve = 0 is not available with a communicator attached");
    if (rp.output_srgb8 && !rp.resolve) return ctx->fail(CPBRT_ERR_BAD_ARG, "output_srgb8 needs resolve = 1");

    const uint32_t npix = S.width * S.height;
    uint32_t spass = rp.samples_per_pass;
    // up to 32 spp and ~32 M paths per pass (7 GB of path state + 1.3 GB of queues, of 180 GB): long
    // stages amortise the launch gaps and the thinning later bounces (Cornell 900x900: 32 spp per pass
    // measured 1 % ahead of 20, 4 % ahead of 8)
    if (spass == 0) {
        spass = (32u << 20) / npix; if (spass < 1) spass = 1; if (spass > 32) spass = 32;
        // equal passes instead of a ragged last one: a rank's share of a partitioned render is rarely a multiple of 32
        // (131 samples = 4 x 32 + 3 would pay a whole pass of launches for 3 samples; it becomes 33 + 33 + 33 + 32)
        const uint32_t n_samples = s_end - s_begin;
        if (n_samples > 0) {
            const uint32_t n_eq = std::max(1u, (n_samples + spass / 2) / spass);
            const uint32_t even = (n_samples + n_eq - 1) / n_eq;
            if (even <= spass + spass / 4 && (uint64_t)even * npix < 0x7FFFFFFFull) spass = even;
        }
    }
    if ((uint64_t)spass * npix >= 0x7FFFFFFFull) return ctx->fail(CPBRT_ERR_UNSUPPORTED, "film too large for one pass");
    // Two lanes: consecutive passes run as independent pipelines on their own streams and overlap on the GPU.
    // The stages of a pass stress different things -- the traversal kernels saturate the math pipes, the shading
    // kernels wait on memory at 20 resident warps per SM -- so one pass's shading fills the issue slots the other
    // pass's traversal leaves, and the tails / launch gaps of one hide behind the other.  What couples two passes
    // is the per-pixel state only: the film-point drift (k_raygen) and the sample-ordered sums (k_film) are chained
    // across lanes with events, everything else is private to a lane.  profile_stages, the recursive direct-lighting
    // kernel and single-pass renders use one lane on the control stream.  CPBRT_LANES=1|2 overrides.
    const uint32_t n_passes = (s_end - s_begin + spass - 1) / spass;
    int n_lanes = (rp.profile_stages || (ctx->recursive_direct && !rp.first_hit_only) || n_passes < 2) ? 1 : cpbrt_ctx::kDefaultLanes;
    if (const char* e = getenv("CPBRT_LANES")) { if (n_lanes > 1) n_lanes = std::max(1, std::min((int)cpbrt_ctx::kMaxLanes, atoi(e))); }
    n_lanes = (int)std::min<uint32_t>((uint32_t)n_lanes, n_passes);
    if (int rc = ensure_wavefront(ctx, spass * npix, n_lanes)) return rc;
    const PixelArrays PX = ctx->PX;
    Totals* T = ctx->d_totals;
    cudaStream_t st = ctx->stream;
    CU(cudaMemsetAsync(T, 0, sizeof(Totals), st));
    CU(slow_paths_reset(st));
    for (int li = 0; li < n_lanes; ++li) CU(cudaMemsetAsync(ctx->lanes[li].Q.counts, 0, 64 * 4, st));

    // every event of this call lives in one bag that destroys them on ANY way out (error returns included)
    struct EventBag {
        std::vector<cudaEvent_t> all;
        cudaEvent_t make() { cudaEvent_t e = nullptr; cudaEventCreate(&e); all.push_back(e); return e; }
        ~EventBag() { for (cudaEvent_t e : all) if (e) cudaEventDestroy(e); }
    } events;
    cudaEvent_t e0 = events.make(), e1 = events.make(), e2 = events.make();
    // optional per-stage timing: event pairs around each stage, read back after the final sync
    struct Span { cudaEvent_t a, b; int stage; };   // 0 raygen/film, 1 trace closest, 2 trace shadow, 3 shade
    std::vector<Span> spans;
    const bool prof = rp.profile_stages != 0;
    auto span_begin = [&](int stage) { if (prof) { Span s; s.a = events.make(); s.b = events.make(); s.stage = stage; cudaEventRecord(s.a, st); spans.push_back(s); } };      // prof => one lane on `st`
    auto span_end = [&]() { if (prof) cudaEventRecord(spans.back().b, st); };
    cudaEventRecord(e0, st);
    uint64_t launches = 0, trace_launches = 0;
    const int sched = trace_schedule(ctx);
    const bool fused = sched == SCHED_FUSED && !(ctx->recursive_direct && !rp.first_hit_only);
    const bool vote = sched == SCHED_VOTE, flat = sched == SCHED_FLAT || sched == SCHED_FUSED;
    const size_t flat_sm = flat ? flat_bytes(ctx) : 0;
    // blocks per SM of the long kernels: what their launch bounds allow by default; CPBRT_GRID_TRACE / CPBRT_GRID_MATTE
    // shrink the grids (tuning: how two lanes share an SM, tools/variant_sweep.py)
    auto env_blocks = [](const char* name, int dflt) { const char* e = getenv(name); const int v = e ? atoi(e) : 0; return v > 0 ? std::min(v, dflt) : dflt; };
    const int b_closest = env_blocks("CPBRT_GRID_TRACE", flat ? CPBRT_BLOCKS_FLAT_CLOSEST : (vote ? CPBRT_BLOCKS_CLOSEST_VOTE : CPBRT_BLOCKS_CLOSEST_LOOP));
    const int b_shadow = env_blocks("CPBRT_GRID_TRACE", flat ? CPBRT_BLOCKS_FLAT_SHADOW : (vote ? CPBRT_BLOCKS_SHADOW_VOTE : CPBRT_BLOCKS_SHADOW_LOOP));
    const int b_matte = env_blocks("CPBRT_GRID_MATTE", ShadeBlocks<SHADE_MATTE>::value);
    const int g_closest = grid_for(ctx, b_closest), g_shadow = grid_for(ctx, b_shadow), g_shade = grid_for(ctx, 8), g_pix = grid_for(ctx, 4), g_matte = grid_for(ctx, b_matte);
    // one thread per pixel where possible: each walks a dependent chain of radical inverses (all of a
    // rank's skipped samples on the first pass), which only many resident warps can hide
    const int g_raygen = (int)std::min<uint64_t>(((uint64_t)npix + 127) / 128, (uint64_t)ctx->sm_count * 16);
    k_pixel_setup<<<g_pix, 256, 0, st>>>(S, PX); ++launches;
    const uint32_t n_iters = S.integrator == CPBRT_INTEGRATOR_PATH ? S.max_depth + 1 : 2;
    const bool recursive = ctx->recursive_direct && !rp.first_hit_only;
    const bool multi = n_lanes > 1;
    cudaEvent_t e_start = events.make();
    if (multi) cudaEventRecord(e_start, st);       // pixel set-up done: the lanes may start
    uint32_t pass_index = 0;
    for (uint32_t k0 = s_begin; k0 < s_end; k0 += spass, ++pass_index) {
        const uint32_t Sn = std::min(spass, s_end - k0);
        cpbrt_ctx::WaveLane& lane = ctx->lanes[pass_index % n_lanes];
        const cpbrt_ctx::WaveLane* prev_lane = (multi && pass_index > 0) ? &ctx->lanes[(pass_index - 1) % n_lanes] : nullptr;
        const PathArrays P = lane.P; const Queues Q = lane.Q;
        // the pass's main stream; with one lane everything stays on the control stream
        cudaStream_t sm = multi ? lane.s[0] : st;
        if (multi && pass_index < (uint32_t)n_lanes) cudaStreamWaitEvent(sm, e_start, 0);
        // a rank that starts in the middle of the sample range first replays the film-point drift
        const uint32_t skip = pass_index == 0 ? k0 : 0;
        span_begin(0);
        if (prev_lane) cudaStreamWaitEvent(sm, prev_lane->raygen_done, 0);      // film_point: sample order
        if (skip) { k_film_replay<<<g_pix, 256, 0, sm>>>(S, PX, 0, skip, T); ++launches; }
        k_raygen<<<g_raygen, 128, 0, sm>>>(S, PX, P, Q, k0, Sn, 0); ++launches;
        if (multi) cudaEventRecord(lane.raygen_done, sm);
        span_end();
        int cur = 0;
        if (recursive) {
            span_begin(3);
            k_direct_recursive<<<grid_for(ctx, 8), 64, 0, sm>>>(S, P, Sn * npix, T); ++launches;
            span_end();
        }
        // One bounce = closest-hit stage (K1), shadow stage (K2), queue bookkeeping (K3), classify (K4),
        // the per-kind shading kernels (K5..) and their bookkeeping (K9).  K1 and K2 are independent, K4
        // only needs K1, and the shading kernels of different kinds touch different paths: they are put
        // on three streams so that the short, poorly filled kernels (terminal / mirror / glass / classify,
        // ~8 % of a bounce when run back to back) and the tails of the trace kernels overlap
        // with the long ones.  `sm` is the join point.  With profile_stages the bounce runs back to back
        // on one stream so that the per-stage times mean something.
        cudaStream_t sa = prof ? sm : lane.s[1], sb = prof ? sm : lane.s[2];
        cudaEvent_t eK1 = lane.ev[0], eK3 = lane.ev[1], eK4 = lane.ev[2], eS = lane.ev[3], ePrev = lane.ev[4];
        if (!prof && !recursive) cudaEventRecord(ePrev, sm);          // after raygen
        for (uint32_t it = 0; it < n_iters && !recursive; ++it) {
            const int next = cur ^ 1;
            span_begin(1);
            if (flat) k_trace_closest<SCHED_FLAT><<<g_closest, 128, flat_sm, sm>>>(S, P, Q, cur, T);
            else if (vote) k_trace_closest<SCHED_VOTE><<<g_closest, 128, 0, sm>>>(S, P, Q, cur, T);
            else k_trace_closest<SCHED_LOOP><<<g_closest, 128, 0, sm>>>(S, P, Q, cur, T);
            ++trace_launches; ++launches;
            span_end();
            if (!prof) { cudaEventRecord(eK1, sm); cudaStreamWaitEvent(sa, ePrev, 0); }
            if (!fused) {      // fused: shadow rays are traced inside the shading kernels
                span_begin(2);
                if (flat) k_trace_shadow<SCHED_FLAT><<<g_shadow, 128, flat_sm, sa>>>(S, P, Q, T);
                else if (vote) k_trace_shadow<SCHED_VOTE><<<g_shadow, 128, 0, sa>>>(S, P, Q, T);
                else k_trace_shadow<SCHED_LOOP><<<g_shadow, 128, 0, sa>>>(S, P, Q, T);
                ++launches;
                span_end();
            }
            if (!prof) cudaStreamWaitEvent(sa, eK1, 0);
            k_queue_reset<<<1, 32, 0, sa>>>(Q, cur, 0, T); ++launches;
            if (!prof) cudaEventRecord(eK3, sa);
            span_begin(3);
            k_classify<<<g_pix, 256, 0, sm>>>(P, Q, Sn * npix); ++launches;
            if (!prof) { cudaEventRecord(eK4, sm); cudaStreamWaitEvent(sb, eK4, 0); cudaStreamWaitEvent(sb, eK3, 0); }
            k_shade_terminal<<<g_shade, 128, 0, sb>>>(S, P, Q); ++launches;
            if (fused) {
                k_shade<SHADE_MIRROR, true><<<grid_for(ctx, ShadeBlocksFused<SHADE_MIRROR>::value), 128, flat_sm, sb>>>(S, P, Q, next, T); ++launches;
                k_shade<SHADE_GLASS, true><<<grid_for(ctx, ShadeBlocksFused<SHADE_GLASS>::value), 128, flat_sm, sb>>>(S, P, Q, next, T); ++launches;
                if (ctx->has_general_materials) { k_shade<SHADE_GENERAL, true><<<grid_for(ctx, ShadeBlocksFused<SHADE_GENERAL>::value), 128, flat_sm, sb>>>(S, P, Q, next, T); ++launches; }
            } else {
                k_shade<SHADE_MIRROR, false><<<g_shade, 128, 0, sb>>>(S, P, Q, next, T); ++launches;
                k_shade<SHADE_GLASS, false><<<g_shade, 128, 0, sb>>>(S, P, Q, next, T); ++launches;
                if (ctx->has_general_materials) { k_shade<SHADE_GENERAL, false><<<grid_for(ctx, ShadeBlocks<SHADE_GENERAL>::value), 128, 0, sb>>>(S, P, Q, next, T); ++launches; }
            }
            if (!prof) { cudaEventRecord(eS, sb); cudaStreamWaitEvent(sm, eK3, 0); }
            if (fused) { k_shade<SHADE_MATTE, true><<<grid_for(ctx, env_blocks("CPBRT_GRID_MATTE", ShadeBlocksFused<SHADE_MATTE>::value)), 128, flat_sm, sm>>>(S, P, Q, next, T); ++launches; }
            else { k_shade<SHADE_MATTE, false><<<g_matte, 128, 0, sm>>>(S, P, Q, next, T); ++launches; }
            span_end();
            if (!prof) cudaStreamWaitEvent(sm, eS, 0);
            k_queue_reset<<<1, 32, 0, sm>>>(Q, cur, 1, T); ++launches;
            if (!prof) cudaEventRecord(ePrev, sm);
            cur = next;
        }
        span_begin(0);
        if (prev_lane) cudaStreamWaitEvent(sm, prev_lane->film_done, 0);        // per-pixel sums: sample order
        k_film<<<g_pix, 128, 0, sm>>>(S, PX, P, Sn, T); ++launches;
        if (multi) cudaEventRecord(lane.film_done, sm);
        span_end();
    }
    // join: the last pass's k_film waited for the one before it, and so on down the chain
    if (multi && pass_index > 0) cudaStreamWaitEvent(st, ctx->lanes[(pass_index - 1) % n_lanes].film_done, 0);
    cudaEventRecord(e1, st);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return ctx->fail(CPBRT_ERR_CUDA, "render launch: %s", cudaGetErrorString(err));

    // film: single GPU -> resolve straight into the output.  With a communicator every GPU
    // writes its raw {sum, count} accumulator, one ncclReduce sums them on rank 0, and the
    // resolve kernel there consumes the reduced buffer directly.
    const size_t out_bytes = rp.output_srgb8 ? (size_t)npix * 3 : (size_t)npix * (rp.resolve ? 3 : 4) * 4;
    float* d_out = film_out;
    if (!rp.film_on_device) {
        if (ctx->d_film_out.bytes < out_bytes) { if (int rc = alloc_upload(ctx, ctx->d_film_out, nullptr, out_bytes)) return rc; }
        d_out = (float*)ctx->d_film_out.p;
    }
    if (ctx->comm) {
        // the path-state buffer is idle now: first npix*16 bytes = contribution, next = reduce target
        float* contrib = (float*)ctx->lanes[0].d_path.p;
        float* red = contrib + (size_t)npix * 4;
        k_film_resolve<<<g_pix, 256, 0, st>>>(S, PX.sum, contrib, 0); ++launches;
        int nrc = g_nccl.Reduce(contrib, red, (size_t)npix * 4, kNcclFloat32, kNcclSum, 0, ctx->comm, st);
        if (nrc != 0) return ctx->fail(CPBRT_ERR_NCCL, "ncclReduce: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(nrc) : "error");
        if (root) {
            if (rp.output_srgb8) k_film_resolve_srgb8<<<g_pix, 256, 0, st>>>(S, (const float4*)red, (uint8_t*)d_out, 1);
            else k_film_resolve<<<g_pix, 256, 0, st>>>(S, (const float4*)red, d_out, 2);
            ++launches;
        }
    } else {
        if (rp.output_srgb8) k_film_resolve_srgb8<<<g_pix, 256, 0, st>>>(S, PX.sum, (uint8_t*)d_out, 0);
        else k_film_resolve<<<g_pix, 256, 0, st>>>(S, PX.sum, d_out, rp.resolve ? 1 : 0);
        ++launches;
    }
    cudaEventRecord(e2, st);
    Totals t; memset(&t, 0, sizeof t);
    unsigned long long slow = 0;
    err = cudaGetLastError();
    if (err == cudaSuccess) err = slow_paths_read(&slow, st);
    if (err == cudaSuccess) err = cudaMemcpyAsync(&t, T, sizeof t, cudaMemcpyDeviceToHost, st);
    if (err == cudaSuccess && root && !rp.film_on_device) err = cudaMemcpyAsync(film_out, d_out, out_bytes, cudaMemcpyDeviceToHost, st);
    if (err == cudaSuccess) err = cudaStreamSynchronize(st);
    float ms_render = 0.f, ms_film = 0.f;
    if (err == cudaSuccess) { cudaEventElapsedTime(&ms_render, e0, e1); cudaEventElapsedTime(&ms_film, e1, e2); }
    double ms_stage[4] = {0, 0, 0, 0};
    for (auto& sp : spans) {
        float ms = 0.f;
        if (err == cudaSuccess) cudaEventElapsedTime(&ms, sp.a, sp.b);
        ms_stage[sp.stage] += ms;
    }
    if (err != cudaSuccess) return ctx->fail(CPBRT_ERR_CUDA, "render: %s", cudaGetErrorString(err));
    if (stats) {
        memset(stats, 0, sizeof *stats);
        stats->camera_samples = (uint64_t)npix * (s_end - s_begin);
        stats->rays_closest = t.rays_closest + t.elided_closest; stats->rays_closest_traced = t.rays_closest; stats->rays_any = t.rays_any; stats->nan_dropped = t.nan_dropped;
        stats->kernel_launches = launches; stats->ms_total = ms_render + ms_film; stats->shade_items = t.shade_items; stats->trace_slow_paths = slow;
        stats->ms_reduce = ctx->comm ? ms_film : 0.0; stats->ms_raygen_film = ms_film;
        if (prof) {
            stats->ms_raygen_film = ms_stage[0] + ms_film; stats->ms_trace_closest = ms_stage[1]; stats->ms_trace_shadow = ms_stage[2];
            stats->ms_trace = ms_stage[1] + ms_stage[2]; stats->ms_shade = ms_stage[3]; stats->trace_launches = trace_launches;
        }
    }
    return check_overflow(ctx, t);
}

int cpbrt_bvh_gpu_levels(const cpbrt_ctx* ctx) { return ctx ? ctx->bvh_gpu_levels : 0; }

#ifdef CPBRT_CHECKED
// CPBRT_CHECK_SELFTEST=1: one deliberate out-of-range access through the checked accessor, to show that a violation
// is counted and reported (site 101) and does not touch memory it should not
__global__ void k_check_selftest(F4Field f) { if (threadIdx.x == 0) { float4 v = f[f.cap + 7u]; if (v.x == 123456.f) f[0] = v; } }
#endif
int cpbrt_check_failures(cpbrt_ctx* ctx, uint64_t* n_failures_out, uint32_t* first_site_out) {
    if (!ctx) return CPBRT_ERR_BAD_CTX;
#ifdef CPBRT_CHECKED
    CU(cudaSetDevice(ctx->device));
    if (getenv("CPBRT_CHECK_SELFTEST") && ctx->lanes[0].capacity) { k_check_selftest<<<1, 32, 0, ctx->stream>>>(ctx->lanes[0].P.ray_o); }
    unsigned long long n = 0; unsigned int site = 0;
    CU(cudaDeviceSynchronize());
    CU(cudaMemcpyFromSymbol(&n, g_check_failures, sizeof n));
    CU(cudaMemcpyFromSymbol(&site, g_check_first_site, sizeof site));
    if (n_failures_out) *n_failures_out = n;
    if (first_site_out) *first_site_out = site;
    return 1;
#else
    if (n_failures_out) *n_failures_out = 0;
    if (first_site_out) *first_site_out = 0;
    return 0;
#endif
}

int cpbrt_bvh_device_layout(cpbrt_ctx* ctx, uint64_t node_capacity, uint64_t* n_nodes_out, float* nodes_out, float* records_out, float* frames_out, float* uv_out, int* resident_out) {
    if (!ctx) return CPBRT_ERR_BAD_CTX;
    if (!ctx->has_bvh) return ctx->fail(CPBRT_ERR_BAD_CTX, "cpbrt_bvh_device_layout before cpbrt_build_bvh");
    CU(cudaSetDevice(ctx->device));
    const DevScene& S = ctx->scene;
    const uint64_t n = ctx->primitives.size();
    if (n_nodes_out) *n_nodes_out = S.n_nodes;
    if (resident_out) *resident_out = ctx->bvh_resident ? 1 : 0;
    if (nodes_out && node_capacity < S.n_nodes) return ctx->fail(CPBRT_ERR_BAD_ARG, "node capacity %llu < %u Parents", (unsigned long long)node_capacity, S.n_nodes);
    if (nodes_out && S.n_nodes) CU(cudaMemcpyAsync(nodes_out, S.nodes, 64ull * S.n_nodes, cudaMemcpyDeviceToHost, ctx->stream));
    if (records_out && n) CU(cudaMemcpyAsync(records_out, S.prims, 48ull * n, cudaMemcpyDeviceToHost, ctx->stream));
    if (frames_out && n) CU(cudaMemcpyAsync(frames_out, S.prim_frame, 48ull * n, cudaMemcpyDeviceToHost, ctx->stream));
    if (uv_out && n && S.prim_uv) CU(cudaMemcpyAsync(uv_out, S.prim_uv, 24ull * n, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return S.prim_uv ? 1 : 0;
}
int cpbrt_bvh_host_segments(const cpbrt_ctx* ctx) { return (ctx && ctx->bvh_gpu_levels > 0) ? (int)ctx->bvh.host_built_segments : -1; }

uint64_t cpbrt_h2d_bytes(const cpbrt_ctx* ctx) { return ctx ? ctx->h2d_bytes : 0; }

int cpbrt_measure_read_bandwidth(cpbrt_ctx* ctx, uint64_t bytes, uint32_t repeats, double* gb_per_s) {
    if (!ctx) return CPBRT_ERR_BAD_CTX;
    if (!gb_per_s || bytes < 4096 || repeats == 0) return ctx->fail(CPBRT_ERR_BAD_ARG, "bad bandwidth probe arguments");
    CU(cudaSetDevice(ctx->device));
    DevBuf b, sink;
    if (int rc = alloc_upload(ctx, b, nullptr, bytes)) return rc;
    if (int rc = alloc_upload(ctx, sink, nullptr, 16)) { b.release(); return rc; }
    const uint64_t n = bytes / 16;
    const int grid = grid_for(ctx, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int k = 0; k < 3; ++k) k_read_bandwidth<<<grid, 256, 0, ctx->stream>>>((const float4*)b.p, n, (float*)sink.p);   // warm (and fill L2 when it fits)
    cudaEventRecord(e0, ctx->stream);
    for (uint32_t k = 0; k < repeats; ++k) k_read_bandwidth<<<grid, 256, 0, ctx->stream>>>((const float4*)b.p, n, (float*)sink.p);
    cudaEventRecord(e1, ctx->stream);
    cudaError_t err = cudaStreamSynchronize(ctx->stream);
    float ms = 0.f; if (err == cudaSuccess) cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    b.release(); sink.release();
    if (err != cudaSuccess) return ctx->fail(CPBRT_ERR_CUDA, "bandwidth probe: %s", cudaGetErrorString(err));
    *gb_per_s = (double)n * 16.0 * repeats / (ms * 1e-3) / 1e9;
    return CPBRT_OK;
}

int cpbrt_comm_unique_id(uint8_t id_out[CPBRT_COMM_ID_BYTES]) {
    std::string err;
    if (!g_nccl.load(&err)) { g_create_error = err; return CPBRT_ERR_NCCL; }
    NcclUniqueId id;
    if (g_nccl.GetUniqueId(&id) != 0) { g_create_error = "ncclGetUniqueId failed"; return CPBRT_ERR_NCCL; }
    memcpy(id_out, &id, CPBRT_COMM_ID_BYTES);
    return CPBRT_OK;
}

int cpbrt_comm_init(cpbrt_ctx* ctx, int n_ranks, int rank, const uint8_t id_bytes[CPBRT_COMM_ID_BYTES]) {
    if (!ctx) return CPBRT_ERR_BAD_CTX;
    if (n_ranks < 1 || rank < 0 || rank >= n_ranks || !id_bytes) return ctx->fail(CPBRT_ERR_BAD_ARG, "bad communicator arguments");
    CU(cudaSetDevice(ctx->device));
    std::string err;
    if (!g_nccl.load(&err)) return ctx->fail(CPBRT_ERR_NCCL, "%s", err.c_str());
    NcclUniqueId id; memcpy(&id, id_bytes, CPBRT_COMM_ID_BYTES);
    if (ctx->comm) { g_nccl.CommDestroy(ctx->comm); ctx->comm = nullptr; ctx->n_ranks = 1; ctx->rank = 0; }   // re-attaching replaces the old communicator
    int rc = g_nccl.CommInitRank(&ctx->comm, n_ranks, id, rank);
    if (rc != 0) return ctx->fail(CPBRT_ERR_NCCL, "ncclCommInitRank: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "error");
    ctx->n_ranks = n_ranks; ctx->rank = rank;
    return CPBRT_OK;
}

}  // extern "C"
