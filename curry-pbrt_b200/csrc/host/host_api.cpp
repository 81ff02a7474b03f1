// extern "C" surface of the host side (see include/cpbrt.h, "Host side").
#include <cstdio>
#include <cstring>
#include <string>

#include "scene_host.h"
#include "../bvh_build.h"

using namespace cpbrt_host;

static thread_local std::string g_host_error;

extern "C" {

const char* cpbrt_host_last_error(void) { return g_host_error.c_str(); }

static int load_impl(const std::vector<Segment>& segs, uint64_t perm_seed, cpbrt_scene** out) {
    cpbrt_scene* s = new cpbrt_scene();
    try {
        build_scene_from_segments(segs, perm_seed, &s->data);
    } catch (...) {
        delete s;
        throw;
    }
    *out = s;
    return CPBRT_OK;
}

int cpbrt_scene_load(const char* pbrt_path, uint64_t perm_seed, cpbrt_scene** out) {
    if (!pbrt_path || !out) { g_host_error = "null argument"; return CPBRT_ERR_BAD_ARG; }
    try {
        std::vector<Segment> segs = segments_from_tokens(lex_file(pbrt_path));
        return load_impl(segs, perm_seed, out);
    } catch (const std::exception& e) {
        g_host_error = e.what();
        return CPBRT_ERR_PARSE;
    }
}

int cpbrt_scene_parse(const char* text, const char* base_dir, uint64_t perm_seed, cpbrt_scene** out) {
    if (!text || !out) { g_host_error = "null argument"; return CPBRT_ERR_BAD_ARG; }
    try {
        std::string label = join_path(base_dir ? base_dir : "", "<memory>.pbrt");
        std::vector<Segment> segs = segments_from_tokens(lex_text(text, label));
        return load_impl(segs, perm_seed, out);
    } catch (const std::exception& e) {
        g_host_error = e.what();
        return CPBRT_ERR_PARSE;
    }
}

void cpbrt_scene_free(cpbrt_scene* scene) { delete scene; }

const cpbrt_scene_desc* cpbrt_scene_get_desc(const cpbrt_scene* scene) { return scene ? &scene->data.desc : nullptr; }

const char* cpbrt_scene_film_filename(const cpbrt_scene* scene) { return scene ? scene->data.film_filename.c_str() : nullptr; }

int cpbrt_scene_set_spp(cpbrt_scene* scene, uint32_t spp) {
    if (!scene || spp == 0) return CPBRT_ERR_BAD_ARG;
    scene->data.desc.sampler.spp = spp;
    return CPBRT_OK;
}

int cpbrt_scene_set_integrator(cpbrt_scene* scene, uint32_t integrator, uint32_t max_depth) {
    if (!scene || integrator > CPBRT_INTEGRATOR_DIRECT) return CPBRT_ERR_BAD_ARG;
    scene->data.desc.integrator = integrator;
    scene->data.desc.max_depth = max_depth;
    return CPBRT_OK;
}

void cpbrt_make_permutation(uint64_t seed, uint16_t* out) { make_permutation(seed, out); }

int cpbrt_host_halton_grouped(const uint16_t* permutation, uint32_t group_max, const uint64_t* index, const uint32_t* dim, uint64_t n,
                              uint64_t* acc_out, uint32_t* digits_out) {
    if (!permutation || !index || !dim || !acc_out || !digits_out || group_max < 2 || group_max > 65536) { g_host_error = "cpbrt_host_halton_grouped: bad argument"; return CPBRT_ERR_BAD_ARG; }
    std::vector<HaltonGroupDim> dims; std::vector<uint16_t> extra;
    build_halton_groups(permutation, group_max, &dims, &extra);
    const uint32_t* sums = prime_sums_1000();
    for (uint64_t i = 0; i < n; ++i) {
        if (dim[i] >= (uint32_t)CPBRT_N_PRIMES) { g_host_error = "cpbrt_host_halton_grouped: dimension out of range"; return CPBRT_ERR_BAD_ARG; }
        halton_grouped_reference(permutation, dims[dim[i]], extra, sums[dim[i]], index[i], &acc_out[i], &digits_out[i]);
    }
    return CPBRT_OK;
}

// Sample-index share of a rank (see curry-pbrt_b200/partition.py for the reasoning): boundaries
// b_k = round(spp * (1 - q^k) / (1 - q^n)), q = 1 - 1/420, in IEEE double with this exact operation order.
static uint32_t share_boundary(uint32_t spp, int k, int n_ranks) {
    if (k <= 0) return 0;
    if (k >= n_ranks) return spp;
    const double q = 1.0 - 1.0 / 420.0;
    double qk = 1.0, qn = 1.0;
    for (int i = 0; i < n_ranks; ++i) { qn *= q; if (i < k) qk *= q; }
    const double b = (double)spp * ((1.0 - qk) / (1.0 - qn)) + 0.5;
    const long long v = (long long)b;
    return (uint32_t)(v < 0 ? 0 : (v > (long long)spp ? (long long)spp : v));
}
void cpbrt_sample_share(uint32_t spp, int rank, int n_ranks, uint32_t* begin, uint32_t* end) {
    if (n_ranks < 1) n_ranks = 1;
    if (begin) *begin = share_boundary(spp, rank, n_ranks);
    if (end) *end = share_boundary(spp, rank + 1, n_ranks);
}

int cpbrt_write_png(const char* path, const float* film_rgb, uint32_t width, uint32_t height) {
    if (!path || !film_rgb) { g_host_error = "null argument"; return CPBRT_ERR_BAD_ARG; }
    std::vector<uint8_t> rgb(size_t(width) * height * 3);
    film_to_rgb8(film_rgb, rgb.size(), rgb.data());
    std::string err;
    if (!write_png_rgb8(path, rgb.data(), width, height, &err)) { g_host_error = err; return CPBRT_ERR_IO; }
    return CPBRT_OK;
}

// Aggregate::build without a device: the same exact-SAH builder cpbrt_build_bvh runs, so the
// primitive order / node list can be checked (and reused by a host) on a machine with no GPU.
int cpbrt_host_build_bvh(const cpbrt_scene_desc* d, uint32_t* prim_order_out, cpbrt_bvh_dump* nodes_out) {
    if (!d) { g_host_error = "null scene description"; return CPBRT_ERR_BAD_ARG; }
    try {
        std::vector<cpbrt::BuildBounds> bounds;
        cpbrt::primitive_bounds(d->primitives, d->n_primitives, d->triangles, d->sphere_radius, d->transforms, &bounds);
        cpbrt::BvhBuildResult r;
        cpbrt::build_sah_bvh(bounds, 0, &r);
        if (!cpbrt::export_bvh(r, prim_order_out, nodes_out)) { g_host_error = "bvh dump capacity too small"; return CPBRT_ERR_BAD_ARG; }
    } catch (const std::exception& e) {
        g_host_error = e.what();
        return CPBRT_ERR_BAD_ARG;
    }
    return CPBRT_OK;
}

}  // extern "C"

// render_from_file (src/render.rs:63-82): parse, clip, build the aggregate, render on the GPU,
// write the Film's PNG and print its file name.
extern "C" int cpbrt_render_from_file(const char* pbrt_path, int device_id) {
    cpbrt_scene* scene = nullptr;
    int rc = cpbrt_scene_load(pbrt_path, 1, &scene);
    if (rc != CPBRT_OK) return rc;
    cpbrt_ctx* ctx = nullptr;
    rc = cpbrt_create(&ctx, device_id);
    if (rc != CPBRT_OK) { g_host_error = cpbrt_last_error(nullptr); cpbrt_scene_free(scene); return rc; }
    const cpbrt_scene_desc* d = cpbrt_scene_get_desc(scene);
    // the film comes back as the 8-bit sRGB image: gamma_correct and the quantisation run on the device
    std::vector<uint8_t> rgb((size_t)d->film_width * d->film_height * 3);
    cpbrt_render_params rp; memset(&rp, 0, sizeof rp); rp.resolve = 1; rp.output_srgb8 = 1;
    rc = cpbrt_upload_scene(ctx, d);
    if (rc == CPBRT_OK) rc = cpbrt_build_bvh(ctx, nullptr, nullptr);
    if (rc == CPBRT_OK) rc = cpbrt_render(ctx, &rp, (float*)rgb.data(), nullptr);
    if (rc != CPBRT_OK) g_host_error = cpbrt_last_error(ctx);
    if (rc == CPBRT_OK) {
        std::string err;
        if (!write_png_rgb8(scene->data.film_filename, rgb.data(), d->film_width, d->film_height, &err)) { g_host_error = err; rc = CPBRT_ERR_IO; }
        if (rc == CPBRT_OK) printf("%s\n", scene->data.film_filename.c_str());
    }
    cpbrt_destroy(ctx);
    cpbrt_scene_free(scene);
    return rc;
}
