// Host-side scene object: owns the arrays a cpbrt_scene_desc points into.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../../include/cpbrt.h"
#include "hmath.h"
#include "parser.h"

namespace cpbrt_host {

bool read_png_rgb8(const std::string& path, uint32_t* w, uint32_t* h, std::vector<uint8_t>* rgb, std::string* err);
// OpenEXR scanline files (NONE / RLE / ZIPS / ZIP), channels R, G, B as f16 / f32 / u32 -> packed RGB f32 (texture/image/exr.rs:7-61)
bool read_exr_rgb(const std::string& path, uint32_t* w, uint32_t* h, std::vector<float>* rgb, std::string* err);
bool write_png_rgb8(const std::string& path, const uint8_t* rgb, uint32_t w, uint32_t h, std::string* err);
void film_to_rgb8(const float* film, size_t n_floats, uint8_t* out);

// RGBSpectrum::from_sampled (src/spectrum/rgb_spectrum.rs:85-96)
void spectrum_from_sampled(std::vector<std::pair<float, float>> sampled, float rgb[3]);

// Halton tables
const uint32_t* primes_1000();        // PRIMS   (halton.rs:141-203: the first 1000 primes)
const uint32_t* prime_sums_1000();    // PRIMSUM (halton.rs:205-214)
void make_permutation(uint64_t seed, uint16_t* out);
void halton_sampler_params(uint32_t spp, uint32_t xres, uint32_t yres, cpbrt_sampler* out);   // HaltonSampler::new
// digit-group tables of the device's scrambled_radical_inverse (halton_tables.cpp)
struct HaltonGroupDim { uint32_t base = 0, group = 0, group_digits = 0, group_off = 0, full_off = 0; uint64_t group_magic = 0; };
void build_halton_groups(const uint16_t* permutation, uint32_t group_max, std::vector<HaltonGroupDim>* dims, std::vector<uint16_t>* extra);
void halton_grouped_reference(const uint16_t* permutation, const HaltonGroupDim& d, const std::vector<uint16_t>& extra, uint32_t primsum,
                              uint64_t index, uint64_t* acc_out, uint32_t* digits_out);

struct SceneData {
    std::vector<cpbrt_primitive> primitives;
    std::vector<cpbrt_triangle> triangles;
    std::vector<float> sphere_radius;
    std::vector<cpbrt_transform> transforms;
    std::vector<cpbrt_material> materials;
    std::vector<cpbrt_texture> textures;
    std::vector<float> texels;
    std::vector<cpbrt_light> lights;
    std::vector<uint16_t> permutation;
    cpbrt_scene_desc desc;
    std::string film_filename;
    uint64_t n_clipped = 0;   // primitives dropped by the camera clipper (scene.rs:107-115)
    void refresh_desc();      // re-point desc at the vectors
};

// read_scene + the factories of render_from_file (src/render.rs:63-78)
void build_scene_from_segments(const std::vector<Segment>& segments, uint64_t perm_seed, SceneData* out);

}  // namespace cpbrt_host

struct cpbrt_scene {
    cpbrt_host::SceneData data;
};
