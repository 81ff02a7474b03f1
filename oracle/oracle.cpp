// ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle_math.h for the parity statement).
//
// CPU restatement of the reference's hot path, one section per reference file:
//   sampler        src/sampler/mod.rs, src/sampler/halton.rs
//   camera         src/camera/{perspective,mod,lens}.rs
//   shapes         src/geometry/shape/{mod,triangle,sphere,transform}.rs
//   bounds / BVH   src/geometry/bounds.rs, src/aggregate/{bvh,plain}.rs
//   textures       src/texture/{constant,image}.rs
//   BSDF           src/material/*.rs, src/material/bxdf/*.rs, src/math/distribution.rs
//   lights         src/light/{mod,area,point,distant,infinite_area}.rs
//   integrators    src/integrator/{mod,path,direct_light}.rs
//   render / film  src/render.rs, src/film.rs
// It consumes the same flattened cpbrt_scene_desc the GPU library is given.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#include "../include/cpbrt.h"
#include "oracle_math.h"

namespace orc {

// =============================================================================================
// Halton sampler (src/sampler/halton.rs, src/sampler/mod.rs)
// =============================================================================================
struct HaltonTables {
    uint32_t primes[CPBRT_N_PRIMES];
    uint32_t primsum[CPBRT_N_PRIMES];
    HaltonTables() {
        // PRIMS (halton.rs:141-203) = the first 1000 primes; PRIMSUM (halton.rs:205-214)
        int n = 0;
        for (uint32_t c = 2; n < CPBRT_N_PRIMES; ++c) {
            bool is_p = true;
            for (uint32_t d = 2; d * d <= c; ++d) if (c % d == 0) { is_p = false; break; }
            if (is_p) primes[n++] = c;
        }
        primsum[0] = 0;
        for (int i = 1; i < CPBRT_N_PRIMES; ++i) primsum[i] = primsum[i - 1] + primes[i - 1];
    }
};
static const HaltonTables& tables() { static HaltonTables t; return t; }

static uint64_t upow(uint64_t b, uint32_t e) { uint64_t r = 1; while (e--) r *= b; return r; }

// inverse (halton.rs:36-47)
static void digit_inverse(uint64_t x, uint32_t base_index, uint64_t* acc, uint32_t* count) {
    uint64_t base = tables().primes[base_index], a = 0; uint32_t c = 0;
    while (x != 0) { uint64_t digit = x % base; x /= base; a = a * base + digit; ++c; }
    *acc = a; *count = c;
}
// inverse_exp (halton.rs:48-55)
static uint64_t inverse_exp(uint64_t x, uint32_t base_index, uint32_t exp) {
    uint64_t r; uint32_t c;
    digit_inverse(x, base_index, &r, &c);
    if (c < exp) return r * upow(tables().primes[base_index], exp - c);
    return r;
}
// radical_inverse (halton.rs:65-70)
static Float radical_inverse(uint64_t x, uint32_t base_index) {
    uint64_t base = tables().primes[base_index], acc; uint32_t c;
    digit_inverse(x, base_index, &acc, &c);
    Float inv_base_n = 1.f / (Float)upow(base, c);
    return inv_base_n * (Float)acc;
}
// scrambled_radical_inverse (halton.rs:56-64) with scrambled_inverse (halton.rs:24-35)
static Float scrambled_radical_inverse(uint64_t x, uint32_t base_index, const uint16_t* perm) {
    const uint64_t base = tables().primes[base_index];
    const uint16_t* pm = perm + tables().primsum[base_index];
    uint64_t acc = 0; uint32_t c = 0;
    while (x != 0) { uint64_t digit = x % base; x /= base; acc = acc * base + pm[digit]; ++c; }
    Float inv_base_n = 1.f / (Float)upow(base, c);
    Float inv_base = 1.f / (Float)base;
    return inv_base_n * ((Float)acc + inv_base * (Float)pm[0] / (1.f - inv_base));
}

struct Sampler {   // HaltonSampler (halton.rs:71-79) behind the Sampler trait (sampler/mod.rs:5-44)
    const cpbrt_sampler* cfg;
    uint64_t index; uint32_t dim;
    void set_pixel(uint64_t px, uint64_t py) {   // halton.rs:108-119
        uint64_t pix[2] = {px, py}, inv[2];
        for (int i = 0; i < 2; ++i) inv[i] = inverse_exp(pix[i], i, cfg->exp[i]);
        index = 0;
        for (int i = 0; i < 2; ++i) index += inv[i] * cfg->scale[1 - i] * cfg->mult_inverse[i];
        index %= cfg->scale[0] * cfg->scale[1];
        dim = 0;
    }
    void next_sample() { index += cfg->scale[0] * cfg->scale[1]; dim = 0; }   // halton.rs:120-123
    Float get_sample() {   // halton.rs:124-136
        Float r;
        if (dim < 2) r = radical_inverse(index / cfg->scale[dim], dim);
        else if (dim < 1000) r = scrambled_radical_inverse(index, dim, cfg->permutation);
        else {
            // reference: rand::random() (halton.rs:131), not reproducible by design; a counter-based draw keyed on
            // (index, dimension), the same function as the device path (halton.cuh, halton_overflow_sample)
            uint64_t z = index * 0x9E3779B97F4A7C15ull + (uint64_t)dim * 0xD1B54A32D192ED03ull + 0x632BE59BD9B4E019ull;
            z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; z ^= z >> 31;
            r = (Float)(uint32_t)(z >> 40) * (1.f / 16777216.f);
        }
        ++dim;
        return r;
    }
    Float get_1d() { Float r = get_sample(); return r >= 1.f ? (Float)0.9999999999999 : r; }   // mod.rs:10-17 (the constant is 1.0 in f32)
    P2 get_2d() { P2 p; p.x = get_1d(); p.y = get_1d(); return p; }                            // mod.rs:34-36
    uint64_t get_usize(uint64_t max) {   // mod.rs:24-33 (uses the unclamped sample)
        Float f = get_sample() * (Float)max;
        uint64_t r = f32_to_usize(std::trunc(f));
        return r == max ? max - 1 : r;
    }
};

// =============================================================================================
// Rays, shape points (src/geometry/ray.rs, src/geometry/shape/mod.rs:102-196)
// =============================================================================================
struct Ray { V3 o, d; Float t_max; };

struct ShapePoint {
    V3 p, n; P2 uv; V3 p_error;
    V3 point_offset_by_error(V3 w) const {   // shape/mod.rs:119-127
        Float d = dot(vabs(n), p_error);
        V3 offset = n * d;
        if (dot(w, n) < 0.f) offset = -offset;
        return p + offset;
    }
};
struct ShapeIntersect { Float t; ShapePoint sp; };

static Ray ray_from_shape_point(const ShapePoint& sp, V3 d) {   // Ray::new_shape_point_d (ray.rs:27-29)
    Ray r = {sp.point_offset_by_error(d), d, FLOAT_MAX};
    return r;
}
static Ray ray_from_to(V3 from, V3 to) {   // Ray::from_to (ray.rs:30-36)
    Float a_bit = 0.00001f;
    Ray r = {from, to - from, 1.f - a_bit};
    return r;
}
static Ray ray_apply(const Ray& r, const Float* m) {   // Transformable for Ray (ray.rs:45-51)
    Ray o = {apply_point(m, r.o), apply_vector(m, r.d), r.t_max};
    return o;
}
// Transformable for ShapePoint (shape/mod.rs:135-160)
static ShapePoint shape_point_apply(const ShapePoint& s, const Float* m, const Float* m_inv) {
    ShapePoint o;
    V3 p = apply_point(m, s.p);
    o.p = p;
    o.n = apply_normal(m_inv, s.n);
    Float xs = std::fabs(m[0] * p.x) + std::fabs(m[1] * p.y) + std::fabs(m[2] * p.z) + m[3];
    Float ys = std::fabs(m[4] * p.x) + std::fabs(m[5] * p.y) + std::fabs(m[6] * p.z) + m[7];
    Float zs = std::fabs(m[8] * p.x) + std::fabs(m[9] * p.y) + std::fabs(m[10] * p.z) + m[11];
    o.p_error = gamma_n(3) * v3(xs, ys, zs);
    o.uv = s.uv;
    return o;
}

// =============================================================================================
// Scene view over the flattened description
// =============================================================================================
struct Scene;
static bool shape_intersect(const Scene& sc, const cpbrt_shape& sh, const Ray& ray, ShapeIntersect* out);

// ---- Triangle (src/geometry/shape/triangle.rs) ------------------------------------------------
struct Tri { V3 p0, p1, p2; P2 uv0, uv1, uv2; };
static Tri load_tri(const cpbrt_triangle& t) {
    Tri r;
    r.p0 = v3(t.p[0][0], t.p[0][1], t.p[0][2]); r.p1 = v3(t.p[1][0], t.p[1][1], t.p[1][2]); r.p2 = v3(t.p[2][0], t.p[2][1], t.p[2][2]);
    r.uv0.x = t.uv[0][0]; r.uv0.y = t.uv[0][1]; r.uv1.x = t.uv[1][0]; r.uv1.y = t.uv[1][1]; r.uv2.x = t.uv[2][0]; r.uv2.y = t.uv[2][1];
    return r;
}
static V3 tri_normal(const Tri& t) {   // triangle.rs:78-82: normalize, then Normal3f::from normalizes again
    return normalize(normalize(cross(t.p0 - t.p2, t.p1 - t.p2)));
}
static V3 tri_point(const Tri& t, Float b0, Float b1, Float b2) { return b0 * t.p0 + b1 * t.p1 + b2 * t.p2; }   // :87-90
static P2 tri_uv(const Tri& t, Float b0, Float b1, Float b2) {   // :83-86
    P2 r = {(b0 * t.uv0.x + b1 * t.uv1.x) + b2 * t.uv2.x, (b0 * t.uv0.y + b1 * t.uv1.y) + b2 * t.uv2.y};
    return r;
}
static V3 tri_abs_sum(const Tri& t, Float b0, Float b1, Float b2) { return vabs(b0 * t.p0) + vabs(b1 * t.p1) + vabs(b2 * t.p2); }   // :103-106
static Float tri_area(const Tri& t) { return 0.5f * length(cross(t.p1 - t.p0, t.p2 - t.p0)); }   // :110-113

static V3 permute(V3 v, int x, int y, int z) { return v3(v[x], v[y], v[z]); }

// Triangle::intersect (triangle.rs:194-262).  Triangle::intersect_predicate (triangle.rs:127-192)
// evaluates the same conditions on the same values and only skips the output, so the
// predicate is `tri_intersect(..., nullptr)`.
static bool tri_intersect(const Tri& tr, const Ray& ray, ShapeIntersect* out) {
    V3 p0t = tr.p0 - ray.o, p1t = tr.p1 - ray.o, p2t = tr.p2 - ray.o;
    int kz = imax(vabs(ray.d));
    int kx = kz + 1 == 3 ? 0 : kz + 1;
    int ky = kx + 1 == 3 ? 0 : kx + 1;
    V3 d = permute(ray.d, kx, ky, kz);
    p0t = permute(p0t, kx, ky, kz); p1t = permute(p1t, kx, ky, kz); p2t = permute(p2t, kx, ky, kz);
    Float sx = -d.x / d.z, sy = -d.y / d.z, sz = 1.f / d.z;
    p0t.x += sx * p0t.z; p0t.y += sy * p0t.z;
    p1t.x += sx * p1t.z; p1t.y += sy * p1t.z;
    p2t.x += sx * p2t.z; p2t.y += sy * p2t.z;
    Float e0 = p1t.x * p2t.y - p1t.y * p2t.x;
    Float e1 = p2t.x * p0t.y - p2t.y * p0t.x;
    Float e2 = p0t.x * p1t.y - p0t.y * p1t.x;
    if ((e0 < 0.f || e1 < 0.f || e2 < 0.f) && (e0 > 0.f || e1 > 0.f || e2 > 0.f)) return false;
    Float det = e0 + e1 + e2;
    if (det == 0.f) return false;
    p0t.z *= sz; p1t.z *= sz; p2t.z *= sz;
    Float t_scaled = e0 * p0t.z + e1 * p1t.z + e2 * p2t.z;
    if ((det < 0.f && (t_scaled >= 0.f || t_scaled < ray.t_max * det)) ||
        (det > 0.f && (t_scaled <= 0.f || t_scaled > ray.t_max * det))) return false;
    Float inv_det = 1.f / det;
    Float b0 = e0 * inv_det, b1 = e1 * inv_det, b2 = e2 * inv_det;
    Float t = t_scaled * inv_det;
    Float max_zt = vmax(vabs(v3(p0t.z, p1t.z, p2t.z)));
    Float delta_z = gamma_n(3) * max_zt;
    Float max_xt = vmax(vabs(v3(p0t.x, p1t.x, p2t.x)));
    Float max_yt = vmax(vabs(v3(p0t.y, p1t.y, p2t.y)));
    Float delta_x = gamma_n(5) * (max_xt + max_zt);
    Float delta_y = gamma_n(5) * (max_yt + max_zt);
    Float delta_e = 2.f * (gamma_n(2) * max_xt * max_yt + delta_y * max_xt + delta_x * max_yt);
    Float max_e = vmax(vabs(v3(e0, e1, e2)));
    Float delta_t = 3.f * (gamma_n(3) * max_e * max_zt + delta_e * max_zt + delta_z * max_e) * std::fabs(inv_det);
    if (t <= delta_t) return false;
    if (out) {
        out->t = t;
        out->sp.p_error = gamma_n(7) * tri_abs_sum(tr, b0, b1, b2);
        out->sp.p = tri_point(tr, b0, b1, b2);
        out->sp.n = tri_normal(tr);
        out->sp.uv = tri_uv(tr, b0, b1, b2);
    }
    return true;
}

// ---- Sphere (src/geometry/shape/sphere.rs) ----------------------------------------------------
static P2 sphere_uv(Float radius, V3 p) {   // sphere.rs:14-18
    P2 r;
    r.x = (std::atan2(p.y, p.x) + PI) * 0.5f * INV_PI;
    r.y = std::acos(rclamp(p.z / radius, -1.f, 1.f)) * INV_PI;
    return r;
}
static bool solve_quadratic(Float af, Float bf, Float cf, Float* t0o, Float* t1o) {   // sphere.rs:111-132
    double a = af, b = bf, c = cf;
    double discrim = b * b - 4. * a * c;
    if (discrim < 0.) return false;
    double root = std::sqrt(discrim);
    double q = b < 0. ? -0.5 * (b - root) : -0.5 * (b + root);
    double t0 = q / a, t1 = c / q;
    if (t0 > t1) std::swap(t0, t1);
    *t0o = (Float)t0; *t1o = (Float)t1;
    return true;
}
static bool sphere_intersect(Float radius, const Ray& ray, ShapeIntersect* out) {   // sphere.rs:29-53
    Float a = length_sq(ray.d);
    Float b = 2.f * dot(ray.d, ray.o);
    Float c = length_sq(ray.o) - radius * radius;
    Float t0, t1;
    if (!solve_quadratic(a, b, c, &t0, &t1)) return false;
    if (t0 > ray.t_max || t1 < 0.f) return false;
    Float t = t0 < 0.f ? t1 : t0;
    if (t > ray.t_max) return false;
    if (out) {
        V3 p = ray.o + ray.d * t;
        p = p * (radius / length(p));
        out->t = t;
        out->sp.p = p;
        out->sp.n = normalize(normalize(p));
        out->sp.p_error = gamma_n(5) * vabs(p);
        out->sp.uv = sphere_uv(radius, p);
    }
    return true;
}

// =============================================================================================
// Bounds (src/geometry/bounds.rs)
// =============================================================================================
struct Bounds3 {
    V3 mn, mx;
    V3 corner(int i) const { return v3((i & 1) ? mx.x : mn.x, (i & 2) ? mx.y : mn.y, (i & 4) ? mx.z : mn.z); }   // :58-66
    V3 center() const { return mn + (mx - mn) / 2.f; }                                                             // :70-72
    Float surface_area() const { V3 d = mx - mn; return 2.f * (d.x * d.y + d.x * d.z + d.y * d.z); }              // :255-258
};
static Bounds3 bounds_new(V3 a, V3 b) {   // :45-57
    Bounds3 r;
    r.mn = v3(rmin(a.x, b.x), rmin(a.y, b.y), rmin(a.z, b.z));
    r.mx = v3(rmax(a.x, b.x), rmax(a.y, b.y), rmax(a.z, b.z));
    return r;
}
static Bounds3 bounds_or(Bounds3 r, V3 p) {   // :172-185
    r.mn = v3(rmin(r.mn.x, p.x), rmin(r.mn.y, p.y), rmin(r.mn.z, p.z));
    r.mx = v3(rmax(r.mx.x, p.x), rmax(r.mx.y, p.y), rmax(r.mx.z, p.z));
    return r;
}
static Bounds3 bounds_union(Bounds3 r, const Bounds3& o) {   // :186-223
    r.mn = v3(rmin(r.mn.x, o.mn.x), rmin(r.mn.y, o.mn.y), rmin(r.mn.z, o.mn.z));
    r.mx = v3(rmax(r.mx.x, o.mx.x), rmax(r.mx.y, o.mx.y), rmax(r.mx.z, o.mx.z));
    return r;
}
static Bounds3 bounds_apply(const Bounds3& b, const Float* m) {   // :292-300
    V3 p0 = apply_point(m, b.mn);
    Bounds3 r = {p0, p0};
    for (int i = 1; i < 8; ++i) r = bounds_or(r, apply_point(m, b.corner(i)));
    return r;
}

// RayIntersectCache (bounds.rs:326-368)
struct RayCache {
    Ray ray; V3 inv_d; int pos[3], neg[3];
    explicit RayCache(const Ray& r) : ray(r) {
        inv_d = v3(1.f / r.d.x, 1.f / r.d.y, 1.f / r.d.z);
        for (int i = 0; i < 3; ++i) { pos[i] = r.d[i] >= 0.f ? 1 : 0; neg[i] = 1 - pos[i]; }
    }
};
// Bounds3f::intersect_predicate_cached (bounds.rs:306-323)
static bool bounds_hit(const Bounds3& b, const RayCache& rc) {
    const V3* mm[2] = {&b.mn, &b.mx};
    Float t_min = 0.f, t_max = 0.f;
    for (int c = 0; c < 3; ++c) {
        Float lo = ((*mm[rc.neg[c]])[c] - rc.ray.o[c]) * rc.inv_d[c];
        Float hi = ((*mm[rc.pos[c]])[c] - rc.ray.o[c]) * rc.inv_d[c] * (1.f + 2.f * gamma_n(3));
        if (c == 0) { t_min = lo; t_max = hi; continue; }
        if (t_min > hi || lo > t_max) return false;
        t_min = rmin(t_min, lo);
        t_max = rmax(t_max, hi);
    }
    return t_min < rc.ray.t_max && t_max > 0.f;
}

// =============================================================================================
// Scene + BVH (src/aggregate/bvh.rs)
// =============================================================================================
struct BvhNode {   // BVHNode (bvh.rs:13-21)
    bool leaf; uint8_t axis; uint32_t a, b;   // leaf: range [a,b) ; parent: left = a, right = b
    Bounds3 bound;
};
struct TraceCounters { uint64_t parent_visits = 0, prim_tests = 0; };

struct Scene {
    const cpbrt_scene_desc* d;
    std::vector<uint32_t> order;      // BVH slot -> uploaded primitive index (= self.primitives after build)
    std::vector<BvhNode> nodes;
    uint32_t root = 0;

    const cpbrt_primitive& prim_at(uint32_t slot) const { return d->primitives[order[slot]]; }
};

static Bounds3 shape_bound(const Scene& sc, const cpbrt_shape& sh) {
    Bounds3 b;
    if (sh.kind == CPBRT_SHAPE_SPHERE) {   // sphere.rs:22-27
        Float r = sc.d->sphere_radius[sh.data];
        b = bounds_new(v3(-r, -r, -r), v3(r, r, r));
    } else {                                 // triangle.rs:114-119
        Tri t = load_tri(sc.d->triangles[sh.data]);
        b = bounds_or(bounds_new(t.p0, t.p1), t.p2);
    }
    for (int32_t x = sh.xform; x >= 0; x = sc.d->transforms[x].outer) b = bounds_apply(b, sc.d->transforms[x].m);   // transform.rs:52-54
    return b;
}

// Shape::intersect through any TransformShape wrappers (transform.rs:55-59), outermost first
static bool shape_intersect_chain(const Scene& sc, const cpbrt_shape& sh, int32_t xf, const Ray& ray, ShapeIntersect* out) {
    if (xf < 0) {
        if (sh.kind == CPBRT_SHAPE_SPHERE) return sphere_intersect(sc.d->sphere_radius[sh.data], ray, out);
        return tri_intersect(load_tri(sc.d->triangles[sh.data]), ray, out);
    }
    return false;   // not used: see shape_intersect
}
static bool shape_intersect(const Scene& sc, const cpbrt_shape& sh, const Ray& ray, ShapeIntersect* out) {
    // collect the chain innermost -> outermost; the ray goes through the inverses outermost first
    int32_t chain[16]; int n = 0;
    for (int32_t x = sh.xform; x >= 0 && n < 16; x = sc.d->transforms[x].outer) chain[n++] = x;
    Ray r = ray;
    for (int i = n - 1; i >= 0; --i) r = ray_apply(r, sc.d->transforms[chain[i]].m_inv);
    if (!shape_intersect_chain(sc, sh, -1, r, out)) return false;
    if (out) for (int i = 0; i < n; ++i) out->sp = shape_point_apply(out->sp, sc.d->transforms[chain[i]].m, sc.d->transforms[chain[i]].m_inv);
    return true;
}

// OrderedFloat total order used by the SAH sort key (bvh.rs:51-54): NaN is greatest, -0 == +0
static bool ordered_less(Float a, Float b) {
    bool an = a != a, bn = b != b;
    if (an || bn) return !an && bn;
    return a < b;
}

struct BvhBuilder {
    Scene& sc;
    std::vector<Bounds3> pb;   // bound of the primitive currently in each slot
    explicit BvhBuilder(Scene& s) : sc(s) {}

    // create_sah_node (bvh.rs:24-124)
    uint32_t build(uint32_t start, uint32_t end) {
        uint32_t len = end - start;
        if (len <= 1) {
            BvhNode n; n.leaf = true; n.axis = 0; n.a = start; n.b = end; n.bound = Bounds3();
            sc.nodes.push_back(n);
            return (uint32_t)sc.nodes.size() - 1;
        }
        Bounds3 full = pb[start];
        for (uint32_t i = start + 1; i < end; ++i) full = bounds_union(full, pb[i]);
        bool have = false; Float min_cost = 0.f; uint32_t min_axis = 0, min_i = 0;
        std::vector<uint32_t> best, idx(len);
        std::vector<Bounds3> fwd(len - 1), bwd(len - 1);
        std::vector<Float> key(len);
        for (uint32_t axis = 0; axis < 3; ++axis) {
            for (uint32_t i = 0; i < len; ++i) { idx[i] = i; key[i] = pb[start + i].center()[axis]; }
            std::stable_sort(idx.begin(), idx.end(), [&](uint32_t a, uint32_t b) { return ordered_less(key[a], key[b]); });
            for (uint32_t i = 0; i + 1 < len; ++i) {
                const Bounds3& f = pb[start + idx[i]];
                const Bounds3& b = pb[start + idx[len - 1 - i]];
                fwd[i] = i ? bounds_union(fwd[i - 1], f) : f;
                bwd[i] = i ? bounds_union(bwd[i - 1], b) : b;
            }
            std::reverse(bwd.begin(), bwd.end());
            for (uint32_t i = 0; i + 1 < len; ++i) {
                Float cost = 0.125f + ((Float)(i + 1) * fwd[i].surface_area() + (Float)(len - 1 - i) * bwd[i].surface_area()) / full.surface_area();
                if (have) {
                    if (cost < min_cost) {
                        min_cost = cost;
                        if (axis != min_axis) { min_axis = axis; best = idx; }
                        min_i = i;
                    }
                } else {
                    have = true; min_cost = cost; best = idx; min_axis = axis; min_i = i;
                }
            }
        }
        if (min_cost > (Float)len) {
            BvhNode n; n.leaf = true; n.axis = 0; n.a = start; n.b = end; n.bound = Bounds3();
            sc.nodes.push_back(n);
            return (uint32_t)sc.nodes.size() - 1;
        }
        // the cycle-following swaps of bvh.rs:100-112 realise new[j] = old[best[j]]
        {
            std::vector<uint32_t> o(len); std::vector<Bounds3> b(len);
            for (uint32_t j = 0; j < len; ++j) { o[j] = sc.order[start + best[j]]; b[j] = pb[start + best[j]]; }
            for (uint32_t j = 0; j < len; ++j) { sc.order[start + j] = o[j]; pb[start + j] = b[j]; }
        }
        best.clear(); best.shrink_to_fit(); idx.clear(); idx.shrink_to_fit();
        fwd.clear(); fwd.shrink_to_fit(); bwd.clear(); bwd.shrink_to_fit(); key.clear(); key.shrink_to_fit();
        uint32_t left = build(start, start + min_i + 1);
        uint32_t right = build(start + min_i + 1, end);
        BvhNode n; n.leaf = false; n.axis = (uint8_t)min_axis; n.a = left; n.b = right; n.bound = full;
        sc.nodes.push_back(n);
        return (uint32_t)sc.nodes.size() - 1;
    }
};

static void scene_build(Scene& sc) {   // Aggregate::build (bvh.rs:194-207)
    uint64_t n = sc.d->n_primitives;
    sc.order.resize(n);
    sc.nodes.clear();
    if (n == 0) return;
    BvhBuilder b(sc);
    b.pb.resize(n);
    for (uint64_t i = 0; i < n; ++i) { sc.order[i] = (uint32_t)i; b.pb[i] = shape_bound(sc, sc.d->primitives[i].shape); }
    sc.root = b.build(0, (uint32_t)n);
}

struct Hit { bool hit; uint32_t slot; ShapeIntersect si; };

// intersect_through_bound (bvh.rs:151-191)
static void bvh_closest(const Scene& sc, uint32_t node, RayCache& rc, Hit* result, TraceCounters* cnt) {
    const BvhNode& n = sc.nodes[node];
    if (n.leaf) {
        for (uint32_t s = n.a; s < n.b; ++s) {
            if (cnt) ++cnt->prim_tests;
            ShapeIntersect si;
            if (shape_intersect(sc, sc.prim_at(s).shape, rc.ray, &si)) {
                rc.ray.t_max = si.t;
                result->hit = true; result->slot = s; result->si = si;
            }
        }
        return;
    }
    if (cnt) ++cnt->parent_visits;
    if (bounds_hit(n.bound, rc)) {
        uint32_t left = n.a, right = n.b;
        if (rc.pos[n.axis] != 1) std::swap(left, right);   // bvh.rs:173-178
        bvh_closest(sc, right, rc, result, cnt);
        bvh_closest(sc, left, rc, result, cnt);
    }
}
// intersect_predicate_through_bound (bvh.rs:125-150)
static bool bvh_any(const Scene& sc, uint32_t node, const RayCache& rc, TraceCounters* cnt) {
    const BvhNode& n = sc.nodes[node];
    if (n.leaf) {
        for (uint32_t s = n.a; s < n.b; ++s) {
            if (cnt) ++cnt->prim_tests;
            if (shape_intersect(sc, sc.prim_at(s).shape, rc.ray, nullptr)) return true;
        }
        return false;
    }
    if (cnt) ++cnt->parent_visits;
    return bounds_hit(n.bound, rc) && (bvh_any(sc, n.a, rc, cnt) || bvh_any(sc, n.b, rc, cnt));
}
static Hit scene_intersect(const Scene& sc, const Ray& ray, TraceCounters* cnt = nullptr) {   // bvh.rs:215-221
    Hit h; h.hit = false; h.slot = 0;
    if (sc.nodes.empty()) return h;
    RayCache rc(ray);
    bvh_closest(sc, sc.root, rc, &h, cnt);
    return h;
}
static bool scene_intersect_predicate(const Scene& sc, const Ray& ray, TraceCounters* cnt = nullptr) {   // bvh.rs:208-214
    if (sc.nodes.empty()) return false;
    RayCache rc(ray);
    return bvh_any(sc, sc.root, rc, cnt);
}
// PlainAggregate::intersect (plain.rs:17-36): linear scan in upload order, each primitive
// behind its own bound test; used as an independent cross-check of the BVH.
static Hit plain_intersect(const Scene& sc, const Ray& ray) {
    Hit best; best.hit = false; best.slot = 0;
    RayCache rc(ray);
    for (uint64_t i = 0; i < sc.d->n_primitives; ++i) {
        const cpbrt_shape& sh = sc.d->primitives[i].shape;
        ShapeIntersect si;
        bool h = bounds_hit(shape_bound(sc, sh), rc) && shape_intersect(sc, sh, rc.ray, &si);
        if (best.hit) {
            if (h) { rc.ray.t_max = si.t; if (si.t < best.si.t) { best.slot = (uint32_t)i; best.si = si; } }
        } else if (h) { best.hit = true; best.slot = (uint32_t)i; best.si = si; }
    }
    return best;
}

// =============================================================================================
// Textures (src/texture/constant.rs, src/texture/image.rs:92-101)
// =============================================================================================
static void tex_eval(const Scene& sc, int32_t id, P2 uv, Float out[3]) {
    const cpbrt_texture& t = sc.d->textures[id];
    if (t.kind == CPBRT_TEX_CONSTANT) { out[0] = t.value[0]; out[1] = t.value[1]; out[2] = t.value[2]; return; }
    Float u = uv.x, v = 1.f - uv.y;
    uint64_t ix = rmin<uint64_t>(f32_to_usize(u * (Float)t.width), t.width - 1);
    uint64_t iy = rmin<uint64_t>(f32_to_usize(v * (Float)t.height), t.height - 1);
    const float* px = sc.d->texels + t.offset + (ix + iy * t.width) * t.channels;
    out[0] = px[0];
    if (t.channels == 3) { out[1] = px[1]; out[2] = px[2]; } else { out[1] = out[2] = 0.f; }
}
static Spectrum tex_spectrum(const Scene& sc, int32_t id, P2 uv) { Float v[3]; tex_eval(sc, id, uv, v); return spec3(v); }
static Float tex_float(const Scene& sc, int32_t id, P2 uv) { Float v[3]; tex_eval(sc, id, uv, v); return v[0]; }

// =============================================================================================
// BxDFs and BSDF (src/material/bxdf/*.rs)
// =============================================================================================
// Distribution1D::from + sample_remap (src/math/distribution.rs:16-53) for tiny tables
static void distribution_sample_remap(const Float* f, int len, Float u, int* idx, Float* pdf_out, Float* remap) {
    Float cdf[16], pdf[16];
    Float n = (Float)len, last = 0.f;
    for (int i = 0; i < len; ++i) { cdf[i] = last + f[i] / n; last = cdf[i]; }
    Float f_sum = cdf[len - 1];
    for (int i = 0; i < len; ++i) { cdf[i] /= f_sum; pdf[i] = f[i] / (n * f_sum); }
    for (int i = 0; i < len; ++i)
        if (u <= cdf[i]) { *idx = i; *pdf_out = pdf[i]; *remap = (cdf[i] - u) / pdf[i]; return; }
    *idx = len - 1; *pdf_out = pdf[len - 1]; *remap = 1.f;
}

struct FresnelDielectric {   // specular.rs:5-44
    Float eta_i, eta_t;
    Float evaluate(Float cos_i) const {
        Float ei = cos_i > 0.f ? eta_i : eta_t, et = cos_i > 0.f ? eta_t : eta_i;
        cos_i = std::fabs(cos_i);
        Float sin_i = std::sqrt(1.f - cos_i * cos_i);
        Float sin_t = sin_i * ei / et;
        if (sin_t >= 1.f) return 1.f;
        Float cos_t = std::sqrt(1.f - sin_t * sin_t);
        Float t0 = et * cos_i, t1 = ei * cos_t;
        Float r_perl = (t0 - t1) / (t0 + t1);
        t0 = ei * cos_i; t1 = et * cos_t;
        Float r_perp = (t0 - t1) / (t0 + t1);
        return (r_perl * r_perl + r_perp * r_perp) / 2.f;
    }
};

struct TrowbridgeReitz {   // microfacet.rs:16-121
    Float ax, ay;
    static Float roughness_to_alpha(Float rough) {
        rough = rmax(rough, 1e-3f);
        Float x = std::log(rough), x2 = x * x;
        return 1.62142f + 0.819955f * x + 0.1734f * x2 + 0.0171201f * x * x2 + 0.000640711f * x2 * x2;
    }
    Float d(V3 wh) const {
        Float t2 = tan_2_theta(wh);
        if (t2 != t2 || std::isinf(t2)) return 0.f;
        Float c2 = cos_2_theta(wh), c4 = c2 * c2;
        Float e = (cos_2_phi(wh) / (ax * ax) + sin_2_phi(wh) / (ay * ay)) * t2;
        return 1.f / (PI * ax * ay * c4 * (1.f + e) * (1.f + e));
    }
    Float lambda(V3 w) const {
        Float at = std::fabs(tan_theta(w));
        if (at != at || std::isinf(at)) return 0.f;
        Float alpha = std::sqrt(cos_2_phi(w) * ax * ax + sin_2_phi(w) * ay * ay);
        Float a = alpha * at;
        return (-1.f + std::sqrt(1.f + a * a)) / 2.f;
    }
    Float g(V3 wo, V3 wi) const { return 1.f / (1.f + lambda(wo) + lambda(wi)); }
    Float g1(V3 w) const { return 1.f / (1.f + lambda(w)); }
    V3 sample_wh(V3 wo, P2 u, Float* pdf) const {
        bool flip = wo.z < 0.f;
        V3 wi = flip ? -wo : wo;
        V3 ws = v3(ax * wi.x, ay * wi.y, wi.z);
        Float cti = cos_theta(ws);
        Float slope_x, slope_y;
        if (cti > 0.9999f) {
            Float r = std::sqrt(u.x / (1.f - u.x));
            Float phi = 6.28318530718f * u.y;
            slope_x = r * std::cos(phi); slope_y = r * std::sin(phi);
        } else {
            Float st = std::sqrt(rmax(0.f, 1.f - cti * cti));
            Float tt = st / cti;
            Float a = 1.f / tt;
            Float G1 = 2.f / (1.f + std::sqrt(1.f + 1.f / (a * a)));
            Float A = 2.f * u.x / G1 - 1.f;
            Float tmp = 1.f / (A * A - 1.f);
            if (tmp > 1e10f) tmp = 1e10f;
            Float B = tt;
            Float D = std::sqrt(rmax(B * B * tmp * tmp - (A * A - B * B) * tmp, 0.f));
            Float sx1 = B * tmp - D, sx2 = B * tmp + D;
            slope_x = (A < 0.f || sx2 > 1.f / tt) ? sx1 : sx2;
            Float S, u2;
            if (u.y > 0.5f) { S = 1.f; u2 = 2.f * (u.y - 0.5f); } else { S = -1.f; u2 = 2.f * (0.5f - u.y); }
            Float z = (u2 * (u2 * (u2 * 0.27385f - 0.73369f) + 0.46341f)) / (u2 * (u2 * (u2 * 0.093073f + 0.309420f) - 1.f) + 0.597999f);
            slope_y = S * z * std::sqrt(1.f + slope_x * slope_x);
        }
        Float tmp = cos_phi(ws) * slope_x - sin_phi(ws) * slope_y;
        slope_y = sin_phi(ws) * slope_x + cos_phi(ws) * slope_y;
        slope_x = tmp;
        slope_x *= ax; slope_y *= ay;
        V3 wh = normalize(v3(-slope_x, -slope_y, 1.f));
        if (flip) wh = -wh;
        *pdf = d(wh) * g1(wo) * std::fabs(dot(wo, wh)) / std::fabs(cos_theta(wo));
        return wh;
    }
};

enum LobeKind { LAMBERT_R, LAMBERT_T, OREN_NAYAR, MICRO_R, MICRO_T };
struct Lobe {   // a non-delta BxDF, optionally wrapped in ScaledBxDF layers (bxdf/mod.rs:49-69)
    LobeKind kind;
    Spectrum r;
    Float a, b;                    // Oren-Nayar
    TrowbridgeReitz dist;          // microfacet
    FresnelDielectric fresnel;     // microfacet
    Float eta_a, eta_b;            // microfacet transmission
    Spectrum scales[4]; int n_scales;
    bool transmit() const { return kind == LAMBERT_T || kind == MICRO_T; }

    bool f_inner(V3 wo, V3 wi, Spectrum* out) const {
        switch (kind) {
            case LAMBERT_R: case LAMBERT_T: *out = r * INV_PI; return true;   // lambertian.rs:14-36
            case OREN_NAYAR: {   // oren_nayar.rs:18-30
                Float sin_alpha, tan_beta;
                if (cos_theta(wi) < cos_theta(wo)) { sin_alpha = sin_theta(wi); tan_beta = tan_theta(wo); }
                else { sin_alpha = sin_theta(wo); tan_beta = tan_theta(wi); }
                *out = r * ((a + b * rmax(0.f, cos_delta_phi(wi, wo)) * sin_alpha * tan_beta) / PI);
                return true;
            }
            case MICRO_R: {   // microfacet.rs:143-165
                Float co = std::fabs(cos_theta(wo)), ci = std::fabs(cos_theta(wi));
                V3 wi_n = normalize(wi), wo_n = normalize(wo);
                V3 wh = wi_n + wo_n;
                if ((wh.x == 0.f && wh.y == 0.f && wh.z == 0.f) || co == 0.f || ci == 0.f) return false;
                wh = normalize(wh);
                Float fr = fresnel.evaluate(dot(wi_n, wh.z < 0.f ? -wh : wh));
                *out = r * dist.d(wh) * dist.g(wo_n, wi_n) * fr / (4.f * co * ci);
                return true;
            }
            case MICRO_T: {   // microfacet.rs:206-242
                if (wo.z * wi.z > 0.f) return false;
                Float co = cos_theta(wo), ci = cos_theta(wi);
                if (ci == 0.f || co == 0.f) return false;
                Float eta = cos_theta(wo) > 0.f ? eta_b / eta_a : eta_a / eta_b;
                V3 wh = normalize(wo + wi * eta);
                if (wh.z < 0.f) wh = -wh;
                Float sqrt_denom = dot(wo, wh) + eta * dot(wi, wh);
                Float fr = fresnel.evaluate(dot(wo, wh));
                Float factor = 1.f / eta;
                Float v = std::fabs(dist.d(wh) * dist.g(wo, wi) * eta * eta * std::fabs(dot(wi, wh)) * std::fabs(dot(wo, wh)) * factor * factor /
                                    (ci * co * sqrt_denom * sqrt_denom));
                *out = (spec(1.f) - spec(fr)) * r * v;
                return true;
            }
        }
        return false;
    }
    bool f(V3 wo, V3 wi, Spectrum* out) const {
        if (!f_inner(wo, wi, out)) return false;
        for (int i = 0; i < n_scales; ++i) *out = *out * scales[i];
        return true;
    }
    Float pdf(V3, V3 wi) const { return wi.z * INV_PI; }   // BxDF::pdf default (bxdf/mod.rs:38-40); no lobe overrides it
    // sample_f: default (bxdf/mod.rs:20-37) or the microfacet overrides
    bool sample_f(V3 wo, P2 u, V3* wi, Spectrum* fo, Float* pdf_out) const {
        bool has_f;
        if (kind == MICRO_R) {   // microfacet.rs:166-179
            if (wo.z == 0.f) { *wi = v3(0, 0, 0); *pdf_out = 0.f; return false; }
            Float pdf; V3 wh = dist.sample_wh(wo, u, &pdf);
            if (dot(wo, wh) < 0.f) { *wi = v3(0, 0, 0); *pdf_out = 0.f; return false; }
            *wi = -wo + 2.f * dot(wo, wh) * wh;
            if (wi->z * wo.z < 0.f) { *wi = v3(0, 0, 0); *pdf_out = 0.f; return false; }
            has_f = f_inner(wo, *wi, fo);
            *pdf_out = pdf / (4.f * dot(wo, wh));
        } else if (kind == MICRO_T) {   // microfacet.rs:246-266
            if (wo.z == 0.f) { *wi = v3(0, 0, 0); *pdf_out = 0.f; return false; }
            Float pdf; V3 wh = dist.sample_wh(wo, u, &pdf);
            if (dot(wo, wh) < 0.f) { *wi = v3(0, 0, 0); *pdf_out = 0.f; return false; }
            Float eta_i, eta_o;
            if (cos_theta(wo) > 0.f) { eta_i = eta_a / eta_b; eta_o = eta_b / eta_a; } else { eta_i = eta_b / eta_a; eta_o = eta_a / eta_b; }
            if (!refract(wo, normalize(wh), eta_i, wi)) { *wi = v3(0, 0, 0); *pdf_out = 0.f; return false; }
            Float sqrt_denom = dot(wo, wh) + eta_o * dot(*wi, wh);
            Float dwh_dwi = std::fabs(eta_o * eta_o * dot(*wi, wh)) / (sqrt_denom * sqrt_denom);
            has_f = f_inner(wo, *wi, fo);
            *pdf_out = pdf * dwh_dwi;
        } else {
            Float pdf; V3 w = cosine_sample_hemisphere(u, &pdf);
            if (!transmit()) { if (wo.z < 0.f) w.z *= -1.f; } else { if (wo.z > 0.f) w.z *= -1.f; }
            *wi = w;
            has_f = f_inner(wo, w, fo);
            *pdf_out = pdf;
        }
        if (has_f) for (int i = 0; i < n_scales; ++i) *fo = *fo * scales[i];
        return has_f;
    }
};

enum DeltaKind { SPEC_R_NOOP, SPEC_R_DIELECTRIC, SPEC_T };
struct DeltaLobe {   // specular.rs:45-102 (+ ScaledDeltaBxDF, bxdf/mod.rs:294-300)
    DeltaKind kind; Spectrum r; Float eta_a, eta_b;
    Spectrum scales[4]; int n_scales;
    bool sample_f(V3 wo, V3* wi, Spectrum* s) const {
        if (kind == SPEC_T) {
            Float ei = cos_theta(wo) > 0.f ? eta_a : eta_b, et = cos_theta(wo) > 0.f ? eta_b : eta_a;
            Float eta = ei / et;
            Float sg = std::copysign(1.f, wo.z); if (wo.z != wo.z) sg = wo.z;   // f32::signum
            if (!refract(wo, normalize(v3(0.f, 0.f, sg)), eta, wi)) return false;
            FresnelDielectric fd = {eta_a, eta_b};
            Float ft = 1.f - fd.evaluate(cos_theta(*wi));
            if (ft == 0.f) return false;
            *s = r * ft / std::fabs(cos_theta(*wi));
        } else {
            *wi = v3(-wo.x, -wo.y, wo.z);
            Float c = cos_theta(*wi);
            Float fr = 1.f;
            if (kind == SPEC_R_DIELECTRIC) { FresnelDielectric fd = {eta_a, eta_b}; fr = fd.evaluate(c); }
            if (fr == 0.f) return false;
            *s = r * fr / std::fabs(c);
        }
        for (int i = 0; i < n_scales; ++i) *s = *s * scales[i];
        return true;
    }
};

struct BSDF {   // bxdf/mod.rs:71-270
    V3 n, sn, snx, sny;
    Lobe lobes[8]; int n_lobes;          // self.bxdfs (reflect_bxdfs / transmit_bxdfs are its two subsets, in order)
    DeltaLobe deltas[8]; int n_deltas;   // self.delta_bxdfs

    void init(V3 normal, V3 shading) {   // BSDF::new (:83-97)
        sn = shading;
        coordinate_system(sn, &snx, &sny);
        n = normalize(v3(dot(normal, snx), dot(normal, sny), dot(normal, sn)));
        n_lobes = 0; n_deltas = 0;
    }
    V3 local_to_world(V3 w) const {   // :98-108
        return normalize(v3((snx.x * w.x + sny.x * w.y) + sn.x * w.z, (snx.y * w.x + sny.y * w.y) + sn.y * w.z, (snx.z * w.x + sny.z * w.y) + sn.z * w.z));
    }
    V3 world_to_local(V3 w) const { return normalize(v3(dot(w, snx), dot(w, sny), dot(w, sn))); }   // :109-111
    Lobe& add_lobe(LobeKind k, Spectrum r) { Lobe& l = lobes[n_lobes++]; l.kind = k; l.r = r; l.n_scales = 0; return l; }
    DeltaLobe& add_delta(DeltaKind k, Spectrum r, Float ea, Float eb) { DeltaLobe& l = deltas[n_deltas++]; l.kind = k; l.r = r; l.eta_a = ea; l.eta_b = eb; l.n_scales = 0; return l; }
    bool is_all_delta() const { return n_lobes == 0; }   // :215-217

    // sample_no_delta_f + choose_no_delta_f (:136-159)
    bool sample_no_delta_f(V3 wo, P2 u, V3* wi, Spectrum* f, Float* pdf) const {
        if (n_lobes == 0) { *wi = v3(0, 0, 0); *pdf = 0.f; return false; }
        Float fl = u.x * (Float)n_lobes;                     // sample_usize_remap (math/mod.rs:81-86)
        Float tr = std::trunc(fl), remap = fl - tr;
        uint64_t index = rmin<uint64_t>(f32_to_usize(tr), (uint64_t)n_lobes - 1);
        P2 u2 = {remap, u.y};
        V3 wo_l = world_to_local(wo), wi_l; Float p;
        bool has = lobes[index].sample_f(wo_l, u2, &wi_l, f, &p);
        *wi = local_to_world(wi_l);
        *pdf = p / (Float)n_lobes;
        return has;
    }
    // sample_delta_f (:160-175).  `ok` false = the reference would panic on an empty table.
    bool sample_delta_f(V3 wo, Float u, V3* wi, Spectrum* f, Float* pdf) const {
        V3 wo_l = world_to_local(wo);
        if (n_deltas == 0) { *wi = v3(0, 0, 0); *pdf = 0.f; return false; }
        V3 w[8]; Spectrum s[8]; Float lum[8]; int n = 0;
        for (int i = 0; i < n_deltas; ++i) if (deltas[i].sample_f(wo_l, &w[n], &s[n])) { lum[n] = s[n].y(); ++n; }
        if (n == 0) { *wi = v3(0, 0, 0); *pdf = 0.f; return false; }   // reference: assert_ne!(len, 0) panics
        int i; Float p, remap;
        distribution_sample_remap(lum, n, u, &i, &p, &remap);
        *wi = local_to_world(w[i]); *f = s[i]; *pdf = p;
        return true;
    }
    // no_delta_f_pdf (:176-198)
    bool no_delta_f_pdf(V3 wo_w, V3 wi_w, Spectrum* f, Float* pdf) const {
        if (n_lobes == 0) { *pdf = 0.f; return false; }
        V3 wo = world_to_local(wo_w), wi = world_to_local(wi_w);
        bool reflect = dot(wo, n) * dot(wi, n) > 0.f;
        bool has = false; Spectrum acc = spec(0.f); Float p = 0.f; int len = 0;
        for (int i = 0; i < n_lobes; ++i) {
            if (lobes[i].transmit() == reflect) continue;
            ++len;
            Spectrum tf;
            if (lobes[i].f(wo, wi, &tf)) { acc = (has ? acc : spec(0.f)) + tf; has = true; p += lobes[i].pdf(wo, wi); }
        }
        *f = acc; *pdf = p / (Float)len;
        return has;
    }
    // sample_f (:199-214)
    bool sample_f(V3 wo, Sampler& smp, V3* wi, Spectrum* f, Float* pdf, bool* is_delta) const {
        Float w[2] = {(Float)n_lobes, (Float)n_deltas};
        int i; Float p, remap;
        distribution_sample_remap(w, 2, smp.get_1d(), &i, &p, &remap);
        bool has; Float ip;
        if (i == 0) { P2 u = {remap, 0.f}; u.y = smp.get_1d(); has = sample_no_delta_f(wo, u, wi, f, &ip); *is_delta = false; }
        else { has = sample_delta_f(wo, remap, wi, f, &ip); *is_delta = true; }
        *pdf = p * ip;
        return has;
    }
    // BSDF::mix (:218-269)
    void mix(const BSDF& rhs, Spectrum scale) {
        Spectrum s1 = scale, s2 = spec(1.f) - s1;
        if (s1.is_black()) { *this = rhs; return; }
        if (s2.is_black()) return;
        for (int i = 0; i < n_lobes; ++i) lobes[i].scales[lobes[i].n_scales++] = s1;
        for (int i = 0; i < rhs.n_lobes && n_lobes < 8; ++i) { lobes[n_lobes] = rhs.lobes[i]; lobes[n_lobes].scales[lobes[n_lobes].n_scales++] = s2; ++n_lobes; }
        for (int i = 0; i < n_deltas; ++i) deltas[i].scales[deltas[i].n_scales++] = s1;
        for (int i = 0; i < rhs.n_deltas && n_deltas < 8; ++i) { deltas[n_deltas] = rhs.deltas[i]; deltas[n_deltas].scales[deltas[n_deltas].n_scales++] = s2; ++n_deltas; }
    }
};

static void add_microfacet_r(BSDF& b, Spectrum ks, Float ru, Float rv, Float eta) {
    Lobe& l = b.add_lobe(MICRO_R, ks);
    l.dist.ax = ru; l.dist.ay = rv; l.fresnel.eta_i = 1.f; l.fresnel.eta_t = eta;
}

// Material::compute_scattering_functions (src/material/{matte,mirror,glass,plastic,uber,translucent,mix}.rs)
static void compute_scattering(const Scene& sc, int32_t mat, const ShapePoint& sp, BSDF* b) {
    const cpbrt_material& m = sc.d->materials[mat];
    P2 uv = sp.uv;
    b->init(sp.n, sp.n);
    switch (m.type) {
        case CPBRT_MAT_MATTE: {   // matte.rs:17-28
            Spectrum kd = tex_spectrum(sc, m.tex[0], uv);
            Float sigma = rclamp(tex_float(sc, m.tex[1], uv), 0.f, 90.f);
            if (sigma == 0.f) b->add_lobe(LAMBERT_R, kd);
            else {   // OrenNayar::new (oren_nayar.rs:8-16)
                Lobe& l = b->add_lobe(OREN_NAYAR, kd);
                Float s = sigma * (PI / 180.f), s2 = s * s;
                l.a = 1.f - (s2 / (2.f * (s + 0.33f)));
                l.b = 0.45f * s2 / (s2 + 0.09f);
            }
            break;
        }
        case CPBRT_MAT_MIRROR:   // mirror.rs:15-23
            b->add_delta(SPEC_R_NOOP, tex_spectrum(sc, m.tex[0], uv), 1.f, 1.f);
            break;
        case CPBRT_MAT_GLASS: {  // glass.rs:17-29
            Spectrum r = tex_spectrum(sc, m.tex[0], uv), t = tex_spectrum(sc, m.tex[1], uv);
            Float eta = tex_float(sc, m.tex[2], uv);
            b->add_delta(SPEC_R_DIELECTRIC, r, 1.f, eta);
            b->add_delta(SPEC_T, t, 1.f, eta);
            break;
        }
        case CPBRT_MAT_PLASTIC: {   // plastic.rs:10-31
            Spectrum kd = tex_spectrum(sc, m.tex[0], uv);
            if (!kd.is_black()) b->add_lobe(LAMBERT_R, kd);
            Spectrum ks = tex_spectrum(sc, m.tex[1], uv);
            if (!ks.is_black()) {
                Float rough = TrowbridgeReitz::roughness_to_alpha(tex_float(sc, m.tex[2], uv));
                add_microfacet_r(*b, ks, rough, rough, 1.5f);
            }
            break;
        }
        case CPBRT_MAT_UBER: {   // uber.rs:16-62
            Float eta = tex_float(sc, m.tex[7], uv);
            Spectrum opacity = tex_spectrum(sc, m.tex[8], uv);
            Spectrum t = spec(1.f) - opacity;
            if (!t.is_black()) b->add_delta(SPEC_T, t, 1.f, 1.f);
            Spectrum kd = opacity * tex_spectrum(sc, m.tex[0], uv);
            if (!kd.is_black()) b->add_lobe(LAMBERT_R, kd);
            Spectrum ks = opacity * tex_spectrum(sc, m.tex[1], uv);
            if (!ks.is_black()) {
                Float ru = m.tex[5] >= 0 ? tex_float(sc, m.tex[5], uv) : tex_float(sc, m.tex[4], uv);
                Float rv = m.tex[6] >= 0 ? tex_float(sc, m.tex[6], uv) : ru;
                ru = TrowbridgeReitz::roughness_to_alpha(ru); rv = TrowbridgeReitz::roughness_to_alpha(rv);
                add_microfacet_r(*b, ks, ru, rv, eta);
            }
            Spectrum kr = opacity * tex_spectrum(sc, m.tex[2], uv);
            if (!kr.is_black()) b->add_delta(SPEC_R_DIELECTRIC, kr, 1.f, eta);
            Spectrum kt = opacity * tex_spectrum(sc, m.tex[3], uv);
            if (!kt.is_black()) b->add_delta(SPEC_T, kt, 1.f, eta);
            break;
        }
        case CPBRT_MAT_TRANSLUCENT: {   // translucent.rs:13-57
            const Float eta = 1.5f;
            Spectrum r = tex_spectrum(sc, m.tex[3], uv), t = tex_spectrum(sc, m.tex[4], uv);
            if (r.is_black() && t.is_black()) break;
            Spectrum kd = tex_spectrum(sc, m.tex[0], uv);
            if (!kd.is_black()) {
                if (!r.is_black()) b->add_lobe(LAMBERT_R, r * kd);
                if (!t.is_black()) b->add_lobe(LAMBERT_T, t * kd);
            }
            Spectrum ks = tex_spectrum(sc, m.tex[1], uv);
            if (!ks.is_black()) {
                Float rough = TrowbridgeReitz::roughness_to_alpha(tex_float(sc, m.tex[2], uv));
                if (!r.is_black()) add_microfacet_r(*b, r * ks, rough, rough, eta);
                if (!t.is_black()) {
                    Lobe& l = b->add_lobe(MICRO_T, t * ks);
                    l.dist.ax = rough; l.dist.ay = rough; l.eta_a = 1.f; l.eta_b = eta; l.fresnel.eta_i = 1.f; l.fresnel.eta_t = eta;
                }
            }
            break;
        }
        case CPBRT_MAT_MIX: {   // mix.rs:11-16
            Spectrum scale = tex_spectrum(sc, m.tex[0], uv);
            compute_scattering(sc, m.sub[0], sp, b);
            BSDF other;
            compute_scattering(sc, m.sub[1], sp, &other);
            b->mix(other, scale);
            break;
        }
    }
}

// =============================================================================================
// Shape sampling for area lights (shape/mod.rs:17-63, triangle.rs:120-126, sphere.rs:54-105)
// =============================================================================================
static Float shape_area(const Scene& sc, const cpbrt_shape& sh) {
    if (sh.kind == CPBRT_SHAPE_SPHERE) { Float r = sc.d->sphere_radius[sh.data]; return r * r * PI * 4.f; }   // sphere.rs:106-108
    return tri_area(load_tri(sc.d->triangles[sh.data]));
}
static void bare_sample(const Scene& sc, const cpbrt_shape& sh, Sampler& smp, ShapePoint* sp, Float* pdf) {
    if (sh.kind == CPBRT_SHAPE_SPHERE) {   // sphere.rs:54-65
        Float radius = sc.d->sphere_radius[sh.data];
        V3 d = uniform_sample_hemisphere(smp.get_2d());
        V3 p = d * radius;
        sp->n = normalize(d);
        p = p * (radius / length(p));
        sp->p = p; sp->p_error = gamma_n(5) * vabs(p); sp->uv = sphere_uv(radius, p);
    } else {   // triangle.rs:120-126
        Tri t = load_tri(sc.d->triangles[sh.data]);
        P2 b = uniform_sample_triangle(smp.get_2d());
        Float b0 = b.x, b1 = b.y, b2 = 1.f - b.x - b.y;
        sp->p = tri_point(t, b0, b1, b2); sp->n = tri_normal(t); sp->uv = tri_uv(t, b0, b1, b2);
        sp->p_error = gamma_n(6) * tri_abs_sum(t, b0, b1, b2);
    }
    *pdf = 1.f / shape_area(sc, sh);
}
// default_sample_by_point (shape/mod.rs:24-41), in the bare shape's space
static void bare_default_sample_by_point(const Scene& sc, const cpbrt_shape& sh, V3 point, Sampler& smp, ShapePoint* sp, Float* pdf) {
    Float p0;
    bare_sample(sc, sh, smp, sp, &p0);
    V3 wi = sp->p - point;
    if (length_sq(wi) == 0.f) { *pdf = 0.f; return; }
    Float p = p0 * length_sq(wi) / -dot(normalize(wi), sp->n);
    *pdf = (p != p || std::isinf(p)) ? 0.f : p;
}
static void bare_sample_by_point(const Scene& sc, const cpbrt_shape& sh, V3 point, Sampler& smp, ShapePoint* sp, Float* pdf) {
    if (sh.kind != CPBRT_SHAPE_SPHERE) { bare_default_sample_by_point(sc, sh, point, smp, sp, pdf); return; }
    // Sphere::sample_by_point (sphere.rs:66-95)
    Float radius = sc.d->sphere_radius[sh.data];
    Float distance_2 = length_sq(point), radius_2 = radius * radius;
    if (distance_2 <= radius_2) { bare_default_sample_by_point(sc, sh, point, smp, sp, pdf); return; }
    Float distance = std::sqrt(distance_2);
    V3 z = point / distance, x, y;
    coordinate_system(z, &x, &y);
    P2 u = smp.get_2d();
    Float sin_theta_max_2 = radius_2 / distance_2;
    Float cos_theta_max = std::sqrt(rmax(1.f - sin_theta_max_2, 0.f));
    Float cos_t = (1.f - u.x) + u.x * cos_theta_max;
    Float sin_t = std::sqrt(rmax(1.f - cos_t * cos_t, 0.f));
    Float phi = u.y * 2.f * PI;
    Float ds = distance * cos_t - std::sqrt(rmax(radius_2 - distance_2 * sin_t * sin_t, 0.f));
    Float cos_alpha = (distance_2 + radius_2 - ds * ds) / (2.f * distance * radius);
    Float sin_alpha = std::sqrt(rmax(1.f - cos_alpha * cos_alpha, 0.f));
    V3 d = cos_alpha * z + sin_alpha * std::cos(phi) * x + sin_alpha * std::sin(phi) * y;
    sp->n = normalize(d);
    V3 p = d * radius;
    sp->p = p; sp->p_error = gamma_n(5) * vabs(p); sp->uv = sphere_uv(radius, p);
    *pdf = 1.f / (2.f * PI * (1.f - cos_theta_max));
}
static Float bare_default_by_point_pdf(const Scene& sc, const cpbrt_shape& sh, V3 point, const ShapePoint& sp) {   // shape/mod.rs:42-52
    V3 d = point - sp.p;
    Float d2 = length_sq(d), dist = std::sqrt(d2);
    Float pdf = d2 / (std::fabs(dot(d / dist, sp.n)) * shape_area(sc, sh));
    return (pdf != pdf || std::isinf(pdf)) ? 0.f : pdf;
}
static Float bare_by_point_pdf(const Scene& sc, const cpbrt_shape& sh, V3 point, const ShapePoint& sp) {
    if (sh.kind != CPBRT_SHAPE_SPHERE) return bare_default_by_point_pdf(sc, sh, point, sp);
    Float radius = sc.d->sphere_radius[sh.data];   // sphere.rs:96-105
    Float distance_2 = length_sq(point);
    if (distance_2 < radius * radius) return bare_default_by_point_pdf(sc, sh, point, sp);
    Float sin_theta_max_2 = radius * radius / distance_2;
    Float cos_theta_max = std::sqrt(1.f - sin_theta_max_2);
    return 1.f / (2.f * PI * (1.f - cos_theta_max));
}
// TransformShape::sample_by_point / by_point_pdf (transform.rs:60-70), through the chain
static void shape_sample_by_point(const Scene& sc, const cpbrt_shape& sh, V3 point, Sampler& smp, ShapePoint* sp, Float* pdf) {
    int32_t chain[16]; int n = 0;
    for (int32_t x = sh.xform; x >= 0 && n < 16; x = sc.d->transforms[x].outer) chain[n++] = x;
    for (int i = n - 1; i >= 0; --i) point = apply_point(sc.d->transforms[chain[i]].m_inv, point);
    bare_sample_by_point(sc, sh, point, smp, sp, pdf);
    for (int i = 0; i < n; ++i) *sp = shape_point_apply(*sp, sc.d->transforms[chain[i]].m, sc.d->transforms[chain[i]].m_inv);
}
static Float shape_by_point_pdf(const Scene& sc, const cpbrt_shape& sh, V3 point, ShapePoint sp) {
    int32_t chain[16]; int n = 0;
    for (int32_t x = sh.xform; x >= 0 && n < 16; x = sc.d->transforms[x].outer) chain[n++] = x;
    for (int i = n - 1; i >= 0; --i) {
        const cpbrt_transform& t = sc.d->transforms[chain[i]];
        point = apply_point(t.m_inv, point);
        sp = shape_point_apply(sp, t.m_inv, t.m);   // ShapePoint::apply(&inverse_transform)
    }
    return bare_by_point_pdf(sc, sh, point, sp);
}

// =============================================================================================
// Lights (src/light/*.rs)
// =============================================================================================
struct LightSample { V3 wi; bool has_li; Spectrum li; Float pdf; Ray vis; };

static Ray visibility_new(const ShapePoint& from, const ShapePoint& to) {   // light/mod.rs:103-107
    V3 f = from.point_offset_by_error(to.p - from.p);
    V3 t = to.point_offset_by_error(f - to.p);
    return ray_from_to(f, t);
}

// Distribution2D over the environment map (infinite_area.rs:11-26, distribution.rs:74-130)
struct EnvDistribution {
    uint32_t w = 0, h = 0;
    std::vector<Float> row_pdf, row_cdf, col_pdf, col_cdf;   // row distribution over h rows; per-row over w
    Float row_fsum = 0.f;
};
struct SceneLights { std::vector<EnvDistribution> env; };   // index = light id (empty for non-infinite)

static void build_dist(const Float* f, size_t len, Float* pdf, Float* cdf, Float* f_sum) {   // Distribution1D::from
    Float n = (Float)len, last = 0.f;
    for (size_t i = 0; i < len; ++i) { cdf[i] = last + f[i] / n; last = cdf[i]; }
    *f_sum = cdf[len - 1];
    for (size_t i = 0; i < len; ++i) { cdf[i] /= *f_sum; pdf[i] = f[i] / (n * *f_sum); }
}
static void dist_sample_remap(const Float* pdf, const Float* cdf, size_t len, Float u, size_t* i, Float* p, Float* remap) {
    for (size_t k = 0; k < len; ++k) if (u <= cdf[k]) { *i = k; *p = pdf[k]; *remap = (cdf[k] - u) / pdf[k]; return; }
    *i = len - 1; *p = pdf[len - 1]; *remap = 1.f;
}

struct World {
    Scene scene;
    SceneLights lights;
};

static void build_env(const Scene& sc, World& w) {
    w.lights.env.resize(sc.d->n_lights);
    for (uint32_t li = 0; li < sc.d->n_lights; ++li) {
        const cpbrt_light& l = sc.d->lights[li];
        if (l.type != CPBRT_LIGHT_INFINITE) continue;
        const cpbrt_texture& t = sc.d->textures[l.tex];
        EnvDistribution& e = w.lights.env[li];
        e.w = t.kind == CPBRT_TEX_IMAGE ? t.width : 1; e.h = t.kind == CPBRT_TEX_IMAGE ? t.height : 1;
        std::vector<Float> f(size_t(e.w) * e.h), fsum(e.h);
        for (uint32_t y = 0; y < e.h; ++y)
            for (uint32_t x = 0; x < e.w; ++x) {
                Spectrum s;
                if (t.kind == CPBRT_TEX_IMAGE) s = spec3(sc.d->texels + t.offset + (x + size_t(y) * e.w) * 3); else s = spec3(t.value);
                Float theta = ((Float)y + 0.5f) / (Float)e.h;
                f[x + size_t(y) * e.w] = s.y() * std::sin(theta);
            }
        e.col_pdf.resize(f.size()); e.col_cdf.resize(f.size()); e.row_pdf.resize(e.h); e.row_cdf.resize(e.h);
        for (uint32_t y = 0; y < e.h; ++y) build_dist(&f[size_t(y) * e.w], e.w, &e.col_pdf[size_t(y) * e.w], &e.col_cdf[size_t(y) * e.w], &fsum[y]);
        build_dist(fsum.data(), e.h, e.row_pdf.data(), e.row_cdf.data(), &e.row_fsum);
    }
}

static P2 spherical_to_normalize_phi_theta(V3 w) {   // math/mod.rs:135-141
    Float p = std::atan2(w.y, w.x);
    P2 r;
    r.x = (p < 0.f ? p + 2.f * PI : p) / 2.f * INV_PI;
    r.y = std::acos(rclamp(w.z, -1.f, 1.f)) * INV_PI;
    return r;
}
static V3 normalize_phi_theta_to_spherical(P2 pt) {   // math/mod.rs:143-151
    Float theta = pt.y * PI, phi = pt.x * PI * 2.f;
    Float ct = std::cos(theta), st = std::sin(theta), cp = std::cos(phi), sp = std::sin(phi);
    return v3(st * cp, st * sp, ct);
}
static bool light_le_out_scene(const World& w, uint32_t li, const Ray& ray, Spectrum* le) {   // infinite_area.rs:35-39
    const cpbrt_light& l = w.scene.d->lights[li];
    if (l.type != CPBRT_LIGHT_INFINITE) return false;
    P2 uv = spherical_to_normalize_phi_theta(normalize(ray.d));
    uv.y = 1.f - uv.y;
    *le = tex_spectrum(w.scene, l.tex, uv);
    return true;
}
static Float env_continuous_pdf(const EnvDistribution& e, P2 u) {   // Distribution2D::continuous_pdf (distribution.rs:119-126)
    Float nx = (Float)e.h, ny = (Float)e.w;   // n = (row_distribution.n, rows[0].n)
    uint64_t r = f32_to_usize(rmin(u.x * nx, nx - 1.f)), c = f32_to_usize(rmin(u.y * ny, ny - 1.f));
    return e.row_pdf[r] * e.col_pdf[r * e.w + c] * nx * ny;
}
static Float light_out_scene_pdf(const World& w, uint32_t li, const Ray& ray) {   // infinite_area.rs:40-51
    const cpbrt_light& l = w.scene.d->lights[li];
    if (l.type != CPBRT_LIGHT_INFINITE) return 0.f;
    P2 uv = spherical_to_normalize_phi_theta(normalize(ray.d));
    uv.y = 1.f - uv.y;
    Float pdf = env_continuous_pdf(w.lights.env[li], uv);
    Float st = std::sin(uv.y * PI);
    return st != 0.f ? pdf / (2.f * PI * PI * st) : 0.f;
}

static bool light_is_delta(const cpbrt_light& l) { return l.type == CPBRT_LIGHT_POINT || l.type == CPBRT_LIGHT_DISTANT; }

static LightSample light_sample_li(const World& w, uint32_t li, const ShapePoint& point, Sampler& smp) {
    const Scene& sc = w.scene;
    const cpbrt_light& l = sc.d->lights[li];
    LightSample r; r.has_li = false; r.pdf = 0.f; r.wi = v3(0, 0, 0); r.li = spec(0.f);
    switch (l.type) {
        case CPBRT_LIGHT_AREA: {   // area.rs:30-51
            ShapePoint lp; Float pdf;
            shape_sample_by_point(sc, l.shape, point.p, smp, &lp, &pdf);
            r.vis = visibility_new(point, lp);
            if (length_sq(lp.p - point.p) == 0.f || pdf == 0.f) return r;
            r.wi = normalize(lp.p - point.p); r.has_li = true; r.li = spec3(l.L); r.pdf = pdf;
            return r;
        }
        case CPBRT_LIGHT_POINT: {   // light/mod.rs:90-98 + point.rs:33-38
            smp.get_2d();
            V3 lpnt = v3(l.v[0], l.v[1], l.v[2]);
            V3 wi = lpnt - point.p;
            r.li = spec3(l.L) / length_sq(wi);
            ShapePoint to; to.p = lpnt; to.n = normalize(-wi); to.uv.x = to.uv.y = 0.f; to.p_error = v3(0, 0, 0);
            r.wi = normalize(wi); r.has_li = true; r.pdf = 1.f; r.vis = visibility_new(point, to);
            return r;
        }
        case CPBRT_LIGHT_DISTANT: {   // distant.rs:28-34
            smp.get_2d();
            V3 wv = v3(l.v[0], l.v[1], l.v[2]);
            r.wi = -wv; r.has_li = true; r.li = spec3(l.L); r.pdf = 1.f;
            Ray vr = {point.point_offset_by_error(-wv), -wv, FLOAT_MAX};   // VisibilityTester::new_od
            r.vis = vr;
            return r;
        }
        case CPBRT_LIGHT_INFINITE: {   // infinite_area.rs:52-72
            const EnvDistribution& e = w.lights.env[li];
            P2 u = smp.get_2d();
            size_t row, col; Float rp, cp, xr, yr;
            dist_sample_remap(e.row_pdf.data(), e.row_cdf.data(), e.h, u.x, &row, &rp, &xr);   // sample_continuous
            Float row_pdf = rp * (Float)e.h, x = ((Float)row + xr) / (Float)e.h;
            dist_sample_remap(&e.col_pdf[row * e.w], &e.col_cdf[row * e.w], e.w, u.y, &col, &cp, &yr);
            Float col_pdf = cp * (Float)e.w, y = ((Float)col + yr) / (Float)e.w;
            Float pdf = row_pdf * col_pdf;
            P2 uv = {x, 1.f - y};
            V3 wi = normalize_phi_theta_to_spherical(uv);
            Float st = std::sin(uv.y * PI);
            pdf = st != 0.f ? pdf / (2.f * PI * PI * st) : 0.f;
            r.wi = wi; r.has_li = true; r.li = tex_spectrum(sc, l.tex, uv); r.pdf = pdf;
            Ray vr = {point.point_offset_by_error(wi), wi, FLOAT_MAX};
            r.vis = vr;
            return r;
        }
    }
    return r;
}

// =============================================================================================
// Integrators (src/integrator/*.rs)
// =============================================================================================
struct RenderCounters { uint64_t rays_closest = 0, rays_any = 0, nan_dropped = 0, rays_unused_last = 0; TraceCounters closest, any; bool count_visits = false; };

static Hit trace(const World& w, const Ray& r, RenderCounters& c) { ++c.rays_closest; return scene_intersect(w.scene, r, c.count_visits ? &c.closest : nullptr); }

// uniform_sample_one_light (integrator/mod.rs:13-97)
static Spectrum uniform_sample_one_light(const World& w, const ShapePoint& sp, const BSDF& bsdf, V3 wo, Sampler& smp, RenderCounters& c) {
    const Scene& sc = w.scene;
    Spectrum l = spec(0.f);
    uint32_t n_lights = sc.d->n_lights;
    V3 n = sp.n;
    if (n_lights != 0) {
        uint32_t li = (uint32_t)smp.get_usize(n_lights);
        const cpbrt_light& light = sc.d->lights[li];
        LightSample ls = light_sample_li(w, li, sp, smp);
        if (ls.has_li && ls.pdf != 0.f) {
            Spectrum f; Float f_pdf;
            if (bsdf.no_delta_f_pdf(wo, ls.wi, &f, &f_pdf) && f_pdf != 0.f) {
                ++c.rays_any;
                if (!scene_intersect_predicate(sc, ls.vis, c.count_visits ? &c.any : nullptr)) {
                    Spectrum ld;
                    if (light_is_delta(light)) ld = ls.li * f * std::fabs(dot(n, ls.wi)) / ls.pdf;
                    else ld = ls.li * f * std::fabs(dot(n, ls.wi)) * power_heuristic(ls.pdf, f_pdf) / ls.pdf;
                    l = l + ld;
                }
            }
        }
        if (!light_is_delta(light)) {
            P2 u = smp.get_2d();
            V3 wi; Spectrum f; Float f_pdf;
            if (bsdf.sample_no_delta_f(wo, u, &wi, &f, &f_pdf) && f_pdf != 0.f) {
                Ray ray = ray_from_shape_point(sp, wi);
                Spectrum ld = spec(0.f);
                Hit h = trace(w, ray, c);
                if (h.hit) {
                    const cpbrt_primitive& hp = sc.prim_at(h.slot);
                    if (hp.light >= 0 && (uint32_t)hp.light == li) {   // Arc::ptr_eq (:63)
                        // light.le_pdf (light/mod.rs:38-40): only AreaLight can be hit
                        Float li_pdf = shape_by_point_pdf(sc, light.shape, sp.p, h.si.sp);
                        if (li_pdf != 0.f) ld = spec3(light.L) * f * std::fabs(dot(n, wi)) * power_heuristic(f_pdf, li_pdf) / f_pdf;
                    }
                } else {
                    Spectrum le;
                    if (light_le_out_scene(w, li, ray, &le)) {
                        Float le_pdf = light_out_scene_pdf(w, li, ray);
                        if (le_pdf != 0.f) ld = le * f * std::fabs(dot(n, wi)) * power_heuristic(f_pdf, le_pdf) / f_pdf;
                    }
                }
                l = l + ld;
            }
        }
    }
    l = l * (Float)n_lights;
    return l;
}

static Spectrum add_out_scene(const World& w, const Ray& ray, Spectrum l, Spectrum beta) {
    for (uint32_t i = 0; i < w.scene.d->n_lights; ++i) {
        Spectrum le;
        l = l + beta * (light_le_out_scene(w, i, ray, &le) ? le : spec(0.f));
    }
    return l;
}

// PathIntegrator::li (path.rs:13-66)
static Spectrum path_li(const World& w, Ray ray, Sampler& smp, uint32_t max_depth, RenderCounters& c) {
    const Scene& sc = w.scene;
    Spectrum l = spec(0.f), beta = spec(1.f);
    uint32_t bounce = 0; bool specular_bounce = false;
    for (;;) {
        Hit h = trace(w, ray, c);
        // bookkeeping only: this ray's result cannot influence the estimate (no emission added, loop ends below)
        if (bounce >= max_depth && !(bounce == 0 || specular_bounce)) ++c.rays_unused_last;
        if (bounce == 0 || specular_bounce) {
            if (h.hit) {
                const cpbrt_primitive& hp = sc.prim_at(h.slot);
                l = l + beta * (hp.light >= 0 ? spec3(sc.d->lights[hp.light].L) : spec(0.f));   // intersect.le() (primitive.rs:81-86)
            } else {
                l = add_out_scene(w, ray, l, beta);
            }
        }
        if (!h.hit) break;
        if (bounce >= max_depth) break;
        const cpbrt_primitive& hp = sc.prim_at(h.slot);
        if (hp.material < 0) break;   // compute_scattering_functions -> None for emissive primitives
        const ShapePoint& sp = h.si.sp;
        BSDF bsdf;
        compute_scattering(sc, hp.material, sp, &bsdf);
        V3 wo = -ray.d;
        if (!bsdf.is_all_delta()) l = l + beta * uniform_sample_one_light(w, sp, bsdf, wo, smp, c);
        V3 wi; Spectrum f; Float f_pdf; bool is_delta;
        if (!bsdf.sample_f(wo, smp, &wi, &f, &f_pdf, &is_delta)) break;
        if (f_pdf == 0.f) break;
        beta = beta * (f * std::fabs(dot(wi, sp.n)) / f_pdf);
        ray = ray_from_shape_point(sp, wi);
        specular_bounce = is_delta;
        if (bounce > 3) {
            Float q = rmax(0.05f, 1.f - beta.y());
            if (smp.get_1d() < q) break;
            beta = beta / (1.f - q);
        }
        ++bounce;
    }
    return l;
}

// DirectLightIntegrator::li_depth (direct_light.rs:12-41)
static Spectrum direct_li(const World& w, const Ray& ray, Sampler& smp, uint32_t max_depth, uint32_t depth, RenderCounters& c) {
    const Scene& sc = w.scene;
    Spectrum l = spec(0.f);
    Hit h = trace(w, ray, c);
    if (h.hit) {
        const cpbrt_primitive& hp = sc.prim_at(h.slot);
        if (hp.material >= 0) {
            BSDF bsdf;
            compute_scattering(sc, hp.material, h.si.sp, &bsdf);
            V3 wo = -normalize(ray.d);
            const ShapePoint& sp = h.si.sp;
            l = l + uniform_sample_one_light(w, sp, bsdf, wo, smp, c);
            if (depth + 1 < max_depth) {
                V3 wo_l = bsdf.world_to_local(wo);
                for (int i = 0; i < bsdf.n_deltas; ++i) {   // sample_all_delta_f (bxdf/mod.rs:127-135)
                    V3 wi_l; Spectrum s;
                    if (!bsdf.deltas[i].sample_f(wo_l, &wi_l, &s)) continue;
                    V3 wi = bsdf.local_to_world(wi_l);
                    Ray r2 = ray_from_shape_point(sp, wi);
                    l = l + direct_li(w, r2, smp, max_depth, depth + 1, c) * s * std::fabs(dot(sp.n, wi));
                }
            }
        } else {
            l = l + (hp.light >= 0 ? spec3(sc.d->lights[hp.light].L) : spec(0.f));
        }
    } else {
        l = add_out_scene(w, ray, l, spec(1.f));
    }
    return l;
}

// =============================================================================================
// Camera (src/camera/*.rs)
// =============================================================================================
static Ray camera_generate_ray(const cpbrt_camera& cam, P2 film, Sampler& smp) {
    V3 cp = apply_point(cam.raster_to_camera, v3(film.x, film.y, 0.f));   // perspective.rs:47-51
    Ray ray = {v3(0, 0, 0), normalize(cp), FLOAT_MAX};
    if (cam.has_lens) {   // lens.rs:24-33
        P2 disk = concentric_sample_disk(smp.get_2d());
        P2 lens = {cam.lens_radius * disk.x, cam.lens_radius * disk.y};
        Float ft = cam.focal_distance / ray.d.z;
        V3 focus = ray.o + ray.d * ft;
        V3 o = v3(lens.x, lens.y, 0.f);
        Ray lr = {o, normalize(focus - o), FLOAT_MAX};
        ray = lr;
    }
    if (cam.has_camera_to_world) ray = ray_apply(ray, cam.camera_to_world);   // camera/mod.rs:66-68
    return ray;
}

// =============================================================================================
// render (src/render.rs:7-50) + film (src/film.rs)
// =============================================================================================
struct RenderOptions {
    uint32_t integrator, max_depth;
    uint32_t x0, y0, x1, y1;   // pixel rectangle (the reference always renders the whole film)
    uint32_t threads;
};

static void render_pixel(const World& w, const RenderOptions& opt, uint32_t px, uint32_t py, Sampler& smp, Float out[3], RenderCounters& c) {
    const cpbrt_scene_desc* d = w.scene.d;
    smp.set_pixel(px, py);
    P2 film = {(Float)px, (Float)py};                       // render.rs:26
    Spectrum sum = spec(0.f); uint32_t kept = 0;
    for (uint32_t s = 0; s < d->sampler.spp; ++s) {
        P2 u = smp.get_2d();
        film.x += u.x - 0.5f; film.y += u.y - 0.5f;         // render.rs:30-31 (running offset, Q1)
        Ray ray = camera_generate_ray(d->camera, film, smp);
        Spectrum li = opt.integrator == CPBRT_INTEGRATOR_PATH ? path_li(w, ray, smp, opt.max_depth, c)
                                                              : direct_li(w, ray, smp, opt.max_depth, 0, c);
        if (li.has_nan()) ++c.nan_dropped; else { sum = sum + li; ++kept; }   // render.rs:34-40, film.rs:11-15
        smp.next_sample();
    }
    sum = sum / (Float)kept;                                // film.rs:16
    out[0] = 0.f + sum.c[0]; out[1] = 0.f + sum.c[1]; out[2] = 0.f + sum.c[2];   // add_sample into a zeroed tile/film
}

static void render(const World& w, const RenderOptions& opt, Float* film, RenderCounters* total) {
    const cpbrt_scene_desc* d = w.scene.d;
    const uint32_t W = d->film_width, tile = 16;   // film.rs:40
    struct TileJob { uint32_t x0, y0, x1, y1; };
    std::vector<TileJob> jobs;
    for (uint32_t tx = opt.x0 / tile; tx * tile < opt.x1; ++tx)
        for (uint32_t ty = opt.y0 / tile; ty * tile < opt.y1; ++ty) {
            TileJob j = {std::max(tx * tile, opt.x0), std::max(ty * tile, opt.y0), std::min((tx + 1) * tile, opt.x1), std::min((ty + 1) * tile, opt.y1)};
            if (j.x0 < j.x1 && j.y0 < j.y1) jobs.push_back(j);
        }
    std::atomic<size_t> next(0);
    std::mutex mu;
    auto worker = [&]() {
        RenderCounters c; c.count_visits = total->count_visits;
        Sampler smp; smp.cfg = &d->sampler; smp.index = 0; smp.dim = 0;   // sampler.box_clone() (render.rs:22)
        for (;;) {
            size_t j = next.fetch_add(1);
            if (j >= jobs.size()) break;
            const TileJob& t = jobs[j];
            for (uint32_t x = t.x0; x < t.x1; ++x)          // index_inside: x outer, y inner (bounds.rs:266-274)
                for (uint32_t y = t.y0; y < t.y1; ++y)
                    render_pixel(w, opt, x, y, smp, film + (size_t(y) * W + x) * 3, c);
        }
        std::lock_guard<std::mutex> g(mu);
        total->rays_closest += c.rays_closest; total->rays_any += c.rays_any; total->nan_dropped += c.nan_dropped; total->rays_unused_last += c.rays_unused_last;
        total->closest.parent_visits += c.closest.parent_visits; total->closest.prim_tests += c.closest.prim_tests;
        total->any.parent_visits += c.any.parent_visits; total->any.prim_tests += c.any.prim_tests;
    };
    uint32_t nt = opt.threads ? opt.threads : std::max(1u, std::thread::hardware_concurrency());
    std::vector<std::thread> pool;
    for (uint32_t i = 1; i < nt; ++i) pool.emplace_back(worker);
    worker();
    for (auto& t : pool) t.join();
}

}  // namespace orc

// =============================================================================================
// C exports for the test harness (ctypes)
// =============================================================================================
using namespace orc;

extern "C" {

struct orc_world { World w; };

orc_world* orc_create(const cpbrt_scene_desc* desc) {
    orc_world* o = new orc_world();
    o->w.scene.d = desc;
    scene_build(o->w.scene);
    build_env(o->w.scene, o->w);
    return o;
}
void orc_destroy(orc_world* o) { delete o; }

uint64_t orc_bvh_node_count(const orc_world* o) { return o->w.scene.nodes.size(); }
uint64_t orc_bvh_root(const orc_world* o) { return o->w.scene.root; }
void orc_bvh_order(const orc_world* o, uint32_t* out) { memcpy(out, o->w.scene.order.data(), o->w.scene.order.size() * sizeof(uint32_t)); }
void orc_bvh_nodes(const orc_world* o, uint8_t* is_leaf, uint8_t* axis, uint32_t* a, uint32_t* b, float* bound) {
    const std::vector<BvhNode>& n = o->w.scene.nodes;
    for (size_t i = 0; i < n.size(); ++i) {
        is_leaf[i] = n[i].leaf; axis[i] = n[i].axis; a[i] = n[i].a; b[i] = n[i].b;
        float* bb = bound + i * 6;
        if (n[i].leaf) { for (int k = 0; k < 6; ++k) bb[k] = 0.f; }
        else { bb[0] = n[i].bound.mn.x; bb[1] = n[i].bound.mn.y; bb[2] = n[i].bound.mn.z; bb[3] = n[i].bound.mx.x; bb[4] = n[i].bound.mx.y; bb[5] = n[i].bound.mx.z; }
    }
}

void orc_halton(const orc_world* o, const uint32_t* px, const uint32_t* py, const uint32_t* sample, const uint32_t* dim,
                uint64_t n, uint64_t* index_out, float* value_out) {
    Sampler s; s.cfg = &o->w.scene.d->sampler;
    for (uint64_t i = 0; i < n; ++i) {
        s.set_pixel(px[i], py[i]);
        s.index += (uint64_t)sample[i] * (s.cfg->scale[0] * s.cfg->scale[1]);
        s.dim = dim[i];
        if (index_out) index_out[i] = s.index;
        if (value_out) value_out[i] = s.get_sample();
    }
}
// standalone helpers for the reference's own unit tests
uint64_t orc_halton_pixel_index(const cpbrt_sampler* cfg, uint32_t px, uint32_t py) { Sampler s; s.cfg = cfg; s.set_pixel(px, py); return s.index; }
float orc_radical_inverse(uint64_t x, uint32_t base_index) { return radical_inverse(x, base_index); }

void orc_camera_rays(const orc_world* o, uint32_t sample, cpbrt_ray* out) {
    const cpbrt_scene_desc* d = o->w.scene.d;
    Sampler s; s.cfg = &d->sampler;
    for (uint32_t y = 0; y < d->film_height; ++y)
        for (uint32_t x = 0; x < d->film_width; ++x) {
            s.set_pixel(x, y);
            P2 film = {(Float)x, (Float)y};
            Ray r = {};
            for (uint32_t k = 0; k <= sample; ++k) {
                P2 u = s.get_2d();
                film.x += u.x - 0.5f; film.y += u.y - 0.5f;
                if (k == sample) r = camera_generate_ray(d->camera, film, s);
                s.next_sample();
            }
            cpbrt_ray& q = out[size_t(y) * d->film_width + x];
            q.o[0] = r.o.x; q.o[1] = r.o.y; q.o[2] = r.o.z; q.t_max = r.t_max;
            q.d[0] = r.d.x; q.d[1] = r.d.y; q.d[2] = r.d.z; q._pad = 0.f;
        }
}
void orc_camera_ray_at(const orc_world* o, float fx, float fy, cpbrt_ray* out) {   // generate_ray_without_sampler
    cpbrt_camera cam = o->w.scene.d->camera; cam.has_lens = 0;
    Sampler s; s.cfg = &o->w.scene.d->sampler; s.index = 0; s.dim = 0;
    P2 film = {fx, fy};
    Ray r = camera_generate_ray(cam, film, s);
    out->o[0] = r.o.x; out->o[1] = r.o.y; out->o[2] = r.o.z; out->t_max = r.t_max;
    out->d[0] = r.d.x; out->d[1] = r.d.y; out->d[2] = r.d.z; out->_pad = 0.f;
}

static Ray to_ray(const cpbrt_ray& q) { Ray r = {v3(q.o[0], q.o[1], q.o[2]), v3(q.d[0], q.d[1], q.d[2]), q.t_max}; return r; }
static void fill_hit(const Scene& sc, const Hit& h, bool slot_is_upload_index, cpbrt_hit* out) {
    if (!h.hit) { out->t = 0.f; out->prim = CPBRT_NO_HIT; out->u = out->v = 0.f; return; }
    out->t = h.si.t;
    out->prim = slot_is_upload_index ? h.slot : sc.order[h.slot];
    out->u = h.si.sp.uv.x; out->v = h.si.sp.uv.y;
}
// counters: [0] parent visits, [1] primitive tests (accumulated over the batch); nullable
void orc_trace_closest(const orc_world* o, const cpbrt_ray* rays, uint64_t n, cpbrt_hit* out, uint64_t* counters, uint32_t threads) {
    const Scene& sc = o->w.scene;
    uint32_t nt = threads ? threads : 1;
    std::vector<TraceCounters> cs(nt);
    auto work = [&](uint32_t tid) {
        for (uint64_t i = tid; i < n; i += nt) { Hit h = scene_intersect(sc, to_ray(rays[i]), counters ? &cs[tid] : nullptr); fill_hit(sc, h, false, &out[i]); }
    };
    std::vector<std::thread> pool;
    for (uint32_t t = 1; t < nt; ++t) pool.emplace_back(work, t);
    work(0);
    for (auto& t : pool) t.join();
    if (counters) for (auto& c : cs) { counters[0] += c.parent_visits; counters[1] += c.prim_tests; }
}
void orc_trace_any(const orc_world* o, const cpbrt_ray* rays, uint64_t n, uint8_t* out, uint64_t* counters, uint32_t threads) {
    const Scene& sc = o->w.scene;
    uint32_t nt = threads ? threads : 1;
    std::vector<TraceCounters> cs(nt);
    auto work = [&](uint32_t tid) {
        for (uint64_t i = tid; i < n; i += nt) out[i] = scene_intersect_predicate(sc, to_ray(rays[i]), counters ? &cs[tid] : nullptr) ? 1 : 0;
    };
    std::vector<std::thread> pool;
    for (uint32_t t = 1; t < nt; ++t) pool.emplace_back(work, t);
    work(0);
    for (auto& t : pool) t.join();
    if (counters) for (auto& c : cs) { counters[0] += c.parent_visits; counters[1] += c.prim_tests; }
}
void orc_trace_closest_plain(const orc_world* o, const cpbrt_ray* rays, uint64_t n, cpbrt_hit* out) {
    for (uint64_t i = 0; i < n; ++i) { Hit h = plain_intersect(o->w.scene, to_ray(rays[i])); fill_hit(o->w.scene, h, true, &out[i]); }
}

// Renders pixels [x0,x1) x [y0,y1) (0,0,0,0 = whole film) into film (W*H*3, untouched outside
// the rectangle).  integrator / max_depth < 0 = take them from the description.
// stats_out: [0] rays_closest [1] rays_any [2] nan_dropped, and when count_visits != 0 also
// [3] closest Parent visits [4] closest primitive tests [5] any-hit Parent visits [6] any-hit
// primitive tests (the algorithmic traversal work of SURVEY 8d); [7] closest-hit rays whose result
// cannot affect the estimate (traced at bounce == max_depth after a non-specular bounce);
// returns render seconds.
double orc_render(const orc_world* o, int integrator, int max_depth, uint32_t x0, uint32_t y0, uint32_t x1, uint32_t y1,
                  uint32_t threads, float* film, uint64_t* stats_out, int count_visits) {
    const cpbrt_scene_desc* d = o->w.scene.d;
    RenderOptions opt;
    opt.integrator = integrator < 0 ? d->integrator : (uint32_t)integrator;
    opt.max_depth = max_depth < 0 ? d->max_depth : (uint32_t)max_depth;
    if (x1 == 0 && y1 == 0) { x1 = d->film_width; y1 = d->film_height; }
    opt.x0 = x0; opt.y0 = y0; opt.x1 = std::min(x1, d->film_width); opt.y1 = std::min(y1, d->film_height);
    opt.threads = threads;
    RenderCounters c; c.count_visits = count_visits != 0;
    auto t0 = std::chrono::steady_clock::now();
    render(o->w, opt, film, &c);
    double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (stats_out) {
        stats_out[0] = c.rays_closest; stats_out[1] = c.rays_any; stats_out[2] = c.nan_dropped; stats_out[7] = c.rays_unused_last;
        if (count_visits) { stats_out[3] = c.closest.parent_visits; stats_out[4] = c.closest.prim_tests; stats_out[5] = c.any.parent_visits; stats_out[6] = c.any.prim_tests; }
    }
    return sec;
}

}  // extern "C"
