/* A C consumer of include/cpbrt.h (compiled -std=c99 -Wall -Wextra -pedantic -Werror by tests/test_abi.py).
 *
 *   abi_smoke --layout     prints "struct field offset size" for every field of every struct of the header
 *                          (no GPU needed); the test compares it with the ctypes mirror in _abi.py
 *   abi_smoke              builds a scene BY HAND -- a matte square of two triangles at z = 3 facing a 90-degree
 *                          camera at the origin, one point light at the eye -- and walks the whole boundary:
 *                          create -> upload -> build -> trace_closest / trace_any -> halton -> render -> destroy.
 *                          The radiance of the centre pixel has the closed form Kd / pi * I / D^2.
 * Exit code 0 = everything agreed.  This is what a host written in C (or any FFI that speaks C) does;
 * nothing here goes through the Python binding or the C++ host parser. */
#include <math.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/cpbrt.h"

#define FIELD(S, F) printf("%s %s %lu %lu\n", #S, #F, (unsigned long)offsetof(S, F), (unsigned long)sizeof(((S*)0)->F))
#define SIZE(S) printf("%s . 0 %lu\n", #S, (unsigned long)sizeof(S))

static void layout(void) {
    SIZE(cpbrt_transform); FIELD(cpbrt_transform, m); FIELD(cpbrt_transform, m_inv); FIELD(cpbrt_transform, outer); FIELD(cpbrt_transform, _pad);
    SIZE(cpbrt_triangle); FIELD(cpbrt_triangle, p); FIELD(cpbrt_triangle, uv);
    SIZE(cpbrt_shape); FIELD(cpbrt_shape, kind); FIELD(cpbrt_shape, data); FIELD(cpbrt_shape, xform); FIELD(cpbrt_shape, _pad);
    SIZE(cpbrt_primitive); FIELD(cpbrt_primitive, shape); FIELD(cpbrt_primitive, material); FIELD(cpbrt_primitive, light);
    SIZE(cpbrt_texture); FIELD(cpbrt_texture, kind); FIELD(cpbrt_texture, channels); FIELD(cpbrt_texture, value); FIELD(cpbrt_texture, width);
    FIELD(cpbrt_texture, height); FIELD(cpbrt_texture, offset);
    SIZE(cpbrt_material); FIELD(cpbrt_material, type); FIELD(cpbrt_material, tex); FIELD(cpbrt_material, sub);
    SIZE(cpbrt_light); FIELD(cpbrt_light, type); FIELD(cpbrt_light, shape); FIELD(cpbrt_light, v); FIELD(cpbrt_light, L); FIELD(cpbrt_light, tex);
    SIZE(cpbrt_camera); FIELD(cpbrt_camera, raster_to_camera); FIELD(cpbrt_camera, camera_to_raster); FIELD(cpbrt_camera, camera_to_world);
    FIELD(cpbrt_camera, has_camera_to_world); FIELD(cpbrt_camera, has_lens); FIELD(cpbrt_camera, lens_radius); FIELD(cpbrt_camera, focal_distance);
    SIZE(cpbrt_sampler); FIELD(cpbrt_sampler, spp); FIELD(cpbrt_sampler, scale); FIELD(cpbrt_sampler, exp); FIELD(cpbrt_sampler, mult_inverse);
    FIELD(cpbrt_sampler, permutation);
    SIZE(cpbrt_scene_desc); FIELD(cpbrt_scene_desc, abi_version);
    FIELD(cpbrt_scene_desc, n_primitives); FIELD(cpbrt_scene_desc, primitives); FIELD(cpbrt_scene_desc, n_triangles); FIELD(cpbrt_scene_desc, triangles);
    FIELD(cpbrt_scene_desc, n_spheres); FIELD(cpbrt_scene_desc, sphere_radius); FIELD(cpbrt_scene_desc, n_transforms); FIELD(cpbrt_scene_desc, transforms);
    FIELD(cpbrt_scene_desc, n_materials); FIELD(cpbrt_scene_desc, materials); FIELD(cpbrt_scene_desc, n_textures); FIELD(cpbrt_scene_desc, textures);
    FIELD(cpbrt_scene_desc, n_texels); FIELD(cpbrt_scene_desc, texels); FIELD(cpbrt_scene_desc, n_lights); FIELD(cpbrt_scene_desc, lights);
    FIELD(cpbrt_scene_desc, camera); FIELD(cpbrt_scene_desc, sampler); FIELD(cpbrt_scene_desc, integrator); FIELD(cpbrt_scene_desc, max_depth);
    FIELD(cpbrt_scene_desc, film_width); FIELD(cpbrt_scene_desc, film_height);
    SIZE(cpbrt_ray); FIELD(cpbrt_ray, o); FIELD(cpbrt_ray, t_max); FIELD(cpbrt_ray, d); FIELD(cpbrt_ray, _pad);
    SIZE(cpbrt_hit); FIELD(cpbrt_hit, t); FIELD(cpbrt_hit, prim); FIELD(cpbrt_hit, u); FIELD(cpbrt_hit, v);
    SIZE(cpbrt_bvh_dump); FIELD(cpbrt_bvh_dump, capacity); FIELD(cpbrt_bvh_dump, n_nodes); FIELD(cpbrt_bvh_dump, root); FIELD(cpbrt_bvh_dump, is_leaf);
    FIELD(cpbrt_bvh_dump, axis); FIELD(cpbrt_bvh_dump, a); FIELD(cpbrt_bvh_dump, b); FIELD(cpbrt_bvh_dump, bound);
    SIZE(cpbrt_render_params); FIELD(cpbrt_render_params, sample_begin); FIELD(cpbrt_render_params, sample_end); FIELD(cpbrt_render_params, first_hit_only);
    FIELD(cpbrt_render_params, resolve); FIELD(cpbrt_render_params, samples_per_pass); FIELD(cpbrt_render_params, film_on_device);
    FIELD(cpbrt_render_params, profile_stages); FIELD(cpbrt_render_params, output_srgb8);
    SIZE(cpbrt_stats); FIELD(cpbrt_stats, camera_samples); FIELD(cpbrt_stats, rays_closest); FIELD(cpbrt_stats, rays_any); FIELD(cpbrt_stats, nan_dropped);
    FIELD(cpbrt_stats, kernel_launches); FIELD(cpbrt_stats, ms_total); FIELD(cpbrt_stats, ms_trace); FIELD(cpbrt_stats, ms_shade);
    FIELD(cpbrt_stats, ms_raygen_film); FIELD(cpbrt_stats, ms_reduce); FIELD(cpbrt_stats, trace_parent_visits); FIELD(cpbrt_stats, trace_prim_tests);
    FIELD(cpbrt_stats, trace_launches); FIELD(cpbrt_stats, ms_trace_closest); FIELD(cpbrt_stats, ms_trace_shadow); FIELD(cpbrt_stats, rays_closest_traced); FIELD(cpbrt_stats, shade_items); FIELD(cpbrt_stats, trace_slow_paths);
}

#define CHECK(call)                                                                                      \
    do {                                                                                                 \
        int rc__ = (call);                                                                               \
        if (rc__ != CPBRT_OK) { fprintf(stderr, "%s -> %d: %s\n", #call, rc__, cpbrt_last_error(ctx)); return 1; } \
    } while (0)
#define EXPECT(cond)                                                                                     \
    do { if (!(cond)) { fprintf(stderr, "abi_smoke: expectation failed at line %d: %s\n", __LINE__, #cond); return 1; } } while (0)

int main(int argc, char** argv) {
    if (argc > 1 && strcmp(argv[1], "--layout") == 0) { layout(); return 0; }

    enum { W = 33, H = 33, SPP = 16 };
    cpbrt_ctx* ctx = NULL;
    uint16_t* perm = (uint16_t*)malloc((size_t)CPBRT_PERM_TABLE_LEN * sizeof(uint16_t));
    float* film = (float*)calloc((size_t)W * H * 3, sizeof(float));
    if (!perm || !film) return 1;
    cpbrt_make_permutation(1, perm);

    /* the square: vertices 0..3, triangles (0 1 2) and (2 3 0) like the Cornell quads */
    static const float V[4][3] = {{-1.f, -1.f, 3.f}, {1.f, -1.f, 3.f}, {1.f, 1.f, 3.f}, {-1.f, 1.f, 3.f}};
    static const int I[2][3] = {{0, 1, 2}, {2, 3, 0}};
    cpbrt_triangle tris[2];
    cpbrt_primitive prims[2];
    for (int t = 0; t < 2; ++t) {
        for (int k = 0; k < 3; ++k) memcpy(tris[t].p[k], V[I[t][k]], sizeof tris[t].p[k]);
        tris[t].uv[0][0] = 0.f; tris[t].uv[0][1] = 0.f; tris[t].uv[1][0] = 1.f; tris[t].uv[1][1] = 0.f; tris[t].uv[2][0] = 1.f; tris[t].uv[2][1] = 1.f;
        prims[t].shape.kind = CPBRT_SHAPE_TRIANGLE; prims[t].shape.data = (uint32_t)t; prims[t].shape.xform = -1; prims[t].shape._pad = 0;
        prims[t].material = 0; prims[t].light = -1;
    }
    cpbrt_texture tex[2];
    memset(tex, 0, sizeof tex);
    tex[0].kind = CPBRT_TEX_CONSTANT; tex[0].channels = 3; tex[0].value[0] = tex[0].value[1] = tex[0].value[2] = 0.5f;   /* Kd */
    tex[1].kind = CPBRT_TEX_CONSTANT; tex[1].channels = 1; tex[1].value[0] = 0.f;                                       /* sigma */
    cpbrt_material mat;
    mat.type = CPBRT_MAT_MATTE;
    for (int k = 0; k < CPBRT_MAT_MAX_TEX; ++k) mat.tex[k] = -1;
    mat.tex[0] = 0; mat.tex[1] = 1; mat.sub[0] = mat.sub[1] = -1;
    cpbrt_light light;
    memset(&light, 0, sizeof light);
    light.type = CPBRT_LIGHT_POINT; light.shape.xform = -1; light.tex = -1;
    light.L[0] = 10.f; light.L[1] = 5.f; light.L[2] = 2.5f;      /* I; position v = origin */

    cpbrt_scene_desc d;
    memset(&d, 0, sizeof d);
    d.abi_version = CPBRT_ABI_VERSION;
    d.n_primitives = 2; d.primitives = prims;
    d.n_triangles = 2; d.triangles = tris;
    d.n_materials = 1; d.materials = &mat;
    d.n_textures = 2; d.textures = tex;
    d.n_lights = 1; d.lights = &light;
    /* raster (x, y, 0, 1) -> camera (2x/W - 1, 1 - 2y/H, 1): a 90-degree pinhole at the origin looking down +z */
    d.camera.raster_to_camera[0] = 2.f / W; d.camera.raster_to_camera[3] = -1.f;
    d.camera.raster_to_camera[5] = -2.f / H; d.camera.raster_to_camera[7] = 1.f;
    d.camera.raster_to_camera[11] = 1.f; d.camera.raster_to_camera[15] = 1.f;
    d.sampler.spp = SPP;
    d.sampler.scale[0] = 64; d.sampler.exp[0] = 6;       /* smallest 2^a >= 33 */
    d.sampler.scale[1] = 81; d.sampler.exp[1] = 4;       /* smallest 3^b >= 33 */
    /* multiplicative_inverse(scale[1 - i], scale[i]) (halton.rs:5-8, 91-93): 81^-1 mod 64 = 49, 64^-1 mod 81 = 19 */
    d.sampler.mult_inverse[0] = 49; d.sampler.mult_inverse[1] = 19;
    d.sampler.permutation = perm;
    d.integrator = CPBRT_INTEGRATOR_PATH; d.max_depth = 1;
    d.film_width = W; d.film_height = H;
    EXPECT((81u * 49u) % 64u == 1u && (64u * 19u) % 81u == 1u);

    CHECK(cpbrt_create(&ctx, 0));
    CHECK(cpbrt_upload_scene(ctx, &d));
    uint32_t order[2] = {9, 9};
    CHECK(cpbrt_build_bvh(ctx, order, NULL));
    EXPECT((order[0] == 0 && order[1] == 1) || (order[0] == 1 && order[1] == 0));

    cpbrt_ray rays[3];
    memset(rays, 0, sizeof rays);
    rays[0].d[0] = 0.3f; rays[0].d[1] = -0.2f; rays[0].d[2] = 1.f; rays[0].t_max = 1e30f;      /* hits triangle 0 at t = 3 */
    rays[1].d[0] = -0.3f; rays[1].d[1] = 0.2f; rays[1].d[2] = 1.f; rays[1].t_max = 1e30f;      /* hits triangle 1 at t = 3 */
    rays[2].d[0] = 2.f; rays[2].d[2] = 1.f; rays[2].t_max = 1e30f;                             /* misses */
    cpbrt_hit hits[3];
    cpbrt_stats st;
    CHECK(cpbrt_trace_closest(ctx, rays, 3, hits, 0, &st));
    EXPECT(hits[0].prim == 0 && hits[0].t == 3.f);
    EXPECT(hits[1].prim == 1 && hits[1].t == 3.f);
    EXPECT(hits[2].prim == CPBRT_NO_HIT);
    EXPECT(st.rays_closest == 3);
    uint8_t occ[3] = {7, 7, 7};
    CHECK(cpbrt_trace_any(ctx, rays, 3, occ, 0, NULL));
    EXPECT(occ[0] == 1 && occ[1] == 1 && occ[2] == 0);

    {   /* the traversal layout as it sits in device memory: two primitives -> one Parent with two Leaf children, or none */
        uint64_t n_nodes = 99; int resident = 7;
        float nodes[16], records[24], frames[24];
        EXPECT(cpbrt_bvh_device_layout(ctx, 0, &n_nodes, NULL, NULL, NULL, NULL, &resident) >= 0);
        EXPECT(n_nodes <= 1 && resident == 0);               /* a two-primitive scene is built on the host */
        EXPECT(cpbrt_bvh_device_layout(ctx, 1, NULL, nodes, records, frames, NULL, NULL) >= 0);
        EXPECT(records[0] == d.triangles[order[0]].p[0][0] && records[12] == d.triangles[order[1]].p[0][0]);
        EXPECT(cpbrt_bvh_gpu_levels(ctx) == 0 && cpbrt_bvh_host_segments(ctx) == -1);
        uint64_t bad = 5; uint32_t site = 5;
        EXPECT(cpbrt_check_failures(ctx, &bad, &site) == 0 && bad == 0 && site == 0);      /* the normal library: no checks compiled in */
    }

    uint32_t px = 5, py = 7, sample = 3, dim = 2;
    uint64_t index = 0; float value = -1.f;
    CHECK(cpbrt_halton(ctx, &px, &py, &sample, &dim, 1, &index, &value));
    EXPECT(value >= 0.f && value < 1.f && index >= 3ull * 64 * 81);

    CHECK(cpbrt_render(ctx, NULL, film, &st));
    EXPECT(st.camera_samples == (uint64_t)W * H * SPP && st.kernel_launches > 0 && st.nan_dropped == 0);
    {   /* centre pixel: Kd / pi * I / D^2 with D = 3 (the jittered film point stays within the pixel: cos^3 ~ 1) */
        const float* c = film + ((size_t)(H / 2) * W + W / 2) * 3;
        const double want = 0.5 / 3.14159265358979 * 10.0 / 9.0;
        printf("centre pixel %.6f %.6f %.6f, closed form %.6f\n", c[0], c[1], c[2], want);
        EXPECT(fabs(c[0] - want) < 0.01 * want && fabs(c[1] - want / 2) < 0.01 * want && fabs(c[2] - want / 4) < 0.01 * want);
    }
    /* error behaviour: codes, never a crash */
    EXPECT(cpbrt_render(NULL, NULL, film, NULL) == CPBRT_ERR_BAD_CTX);
    d.abi_version = 999;
    EXPECT(cpbrt_upload_scene(ctx, &d) == CPBRT_ERR_BAD_ARG && strlen(cpbrt_last_error(ctx)) > 0);
    printf("h2d bytes %llu\n", (unsigned long long)cpbrt_h2d_bytes(ctx));
    cpbrt_destroy(ctx);
    free(perm); free(film);
    printf("abi_smoke ok\n");
    return 0;
}
