// Ray / shape intersection on the device: Triangle::intersect (+ intersect_predicate),
// Sphere::intersect and the TransformShape wrapper
// (src/geometry/shape/triangle.rs:127-262, sphere.rs:29-53,111-132, transform.rs:47-59,
//  src/geometry/shape/mod.rs:119-160, src/geometry/ray.rs:45-51).
#pragma once
#include "scene_dev.cuh"

namespace cpbrt {

struct Ray { V3 o, d; float t_max; };

struct ShapePoint { V3 p, n; P2 uv; V3 p_error; };

// point_offset_by_error (shape/mod.rs:119-127)
CP_DEV V3 point_offset_by_error(const ShapePoint& s, V3 w) {
    float d = dot(vabs(s.n), s.p_error);
    V3 off = s.n * d;
    if (dot(w, s.n) < 0.f) off = -off;
    return s.p + off;
}
// Transformable for ShapePoint (shape/mod.rs:135-160)
CP_DEV ShapePoint shape_point_apply(const ShapePoint& s, const float* __restrict__ m, const float* __restrict__ mi) {
    ShapePoint o;
    V3 p = apply_point(m, s.p);
    o.p = p;
    o.n = apply_normal(mi, s.n);
    float xs = fabsf(m[0] * p.x) + fabsf(m[1] * p.y) + fabsf(m[2] * p.z) + m[3];
    float ys = fabsf(m[4] * p.x) + fabsf(m[5] * p.y) + fabsf(m[6] * p.z) + m[7];
    float zs = fabsf(m[8] * p.x) + fabsf(m[9] * p.y) + fabsf(m[10] * p.z) + m[11];
    o.p_error = gamma_n(3) * v3(xs, ys, zs);
    o.uv = s.uv;
    return o;
}

// Per-ray constants of the watertight test (triangle.rs:202-212): they depend on ray.d only,
// so a ray that meets many untransformed triangles computes them once.
// kx = kz + 1 and ky = kx + 1 (mod 3): the permutation is the rotation that brings the dominant
// axis to z, so the axis index alone selects every component (two loop-invariant predicates
// instead of six index comparisons per primitive test).
struct TriRaySetup { int kz; float sx, sy, sz; };
CP_DEV V3 permute(V3 v, const TriRaySetup& s) {
    const bool z0 = s.kz == 0, z1 = s.kz == 1;
    return v3(z0 ? v.y : (z1 ? v.z : v.x), z0 ? v.z : (z1 ? v.x : v.y), z0 ? v.x : (z1 ? v.y : v.z));
}
CP_DEV TriRaySetup tri_ray_setup(V3 d) {
    TriRaySetup s;
    V3 a = vabs(d);
    int kz = 0; float m = a.x;                 // nalgebra imax: first maximum wins
    if (a.y > m) { m = a.y; kz = 1; }
    if (a.z > m) { kz = 2; }
    s.kz = kz;
    const V3 dp = permute(d, s);
    s.sx = -dp.x / dp.z; s.sy = -dp.y / dp.z; s.sz = 1.f / dp.z;
    return s;
}

struct TriVerts { V3 p0, p1, p2; };

// Triangle::intersect up to the acceptance decision; b[] and *t are the barycentrics and the
// distance of an accepted hit.  intersect_predicate (triangle.rs:127-192) is the same chain of
// conditions on the same values.
// `p0t`, `p1t`, `p2t` = the vertices relative to the ray origin, already permuted (triangle.rs:202-205).
CP_DEV bool tri_intersect_permuted(V3 p0t, V3 p1t, V3 p2t, float t_max, const TriRaySetup& rs, float* t_out, float b[3]) {
    p0t.x += rs.sx * p0t.z; p0t.y += rs.sy * p0t.z;
    p1t.x += rs.sx * p1t.z; p1t.y += rs.sy * p1t.z;
    p2t.x += rs.sx * p2t.z; p2t.y += rs.sy * p2t.z;
    float e0 = p1t.x * p2t.y - p1t.y * p2t.x;
    float e1 = p2t.x * p0t.y - p2t.y * p0t.x;
    float e2 = p0t.x * p1t.y - p0t.y * p1t.x;
    if ((e0 < 0.f || e1 < 0.f || e2 < 0.f) && (e0 > 0.f || e1 > 0.f || e2 > 0.f)) return false;
    float det = e0 + e1 + e2;
    if (det == 0.f) return false;
    p0t.z *= rs.sz; p1t.z *= rs.sz; p2t.z *= rs.sz;
    float t_scaled = e0 * p0t.z + e1 * p1t.z + e2 * p2t.z;
    if ((det < 0.f && (t_scaled >= 0.f || t_scaled < t_max * det)) ||
        (det > 0.f && (t_scaled <= 0.f || t_scaled > t_max * det))) return false;
    float inv_det = 1.f / det;
    float t = t_scaled * inv_det;
    float max_zt = vmax(vabs(v3(p0t.z, p1t.z, p2t.z)));
    float delta_z = gamma_n(3) * max_zt;
    float max_xt = vmax(vabs(v3(p0t.x, p1t.x, p2t.x)));
    float max_yt = vmax(vabs(v3(p0t.y, p1t.y, p2t.y)));
    float delta_x = gamma_n(5) * (max_xt + max_zt);
    float delta_y = gamma_n(5) * (max_yt + max_zt);
    float delta_e = 2.f * (gamma_n(2) * max_xt * max_yt + delta_y * max_xt + delta_x * max_yt);
    float max_e = vmax(vabs(v3(e0, e1, e2)));
    float delta_t = 3.f * (gamma_n(3) * max_e * max_zt + delta_e * max_zt + delta_z * max_e) * fabsf(inv_det);
    if (t <= delta_t) return false;
    *t_out = t;
    if (b) { b[0] = e0 * inv_det; b[1] = e1 * inv_det; b[2] = e2 * inv_det; }
    return true;
}
CP_DEV bool tri_intersect_core(const TriVerts& tv, V3 o, float t_max, const TriRaySetup& rs, float* t_out, float b[3]) {
    return tri_intersect_permuted(permute(tv.p0 - o, rs), permute(tv.p1 - o, rs), permute(tv.p2 - o, rs), t_max, rs, t_out, b);
}

// Barycentrics of a hit that the acceptance test above has already accepted: the same e0, e1, e2 and
// det (same operations, so the same bits) without the distance and error-bound arithmetic that only
// decides acceptance (triangle.rs:213-262 computes b from exactly these values).
CP_DEV void tri_barycentrics(const TriVerts& tv, V3 o, V3 d, float b[3]) {
    const V3 a = vabs(d);
    int kz = 0; float m = a.x;
    if (a.y > m) { m = a.y; kz = 1; }
    if (a.z > m) { kz = 2; }
    TriRaySetup rs; rs.kz = kz;
    const V3 dp = permute(d, rs);
    const float sx = -dp.x / dp.z, sy = -dp.y / dp.z;
    V3 p0t = permute(tv.p0 - o, rs), p1t = permute(tv.p1 - o, rs), p2t = permute(tv.p2 - o, rs);
    p0t.x += sx * p0t.z; p0t.y += sy * p0t.z;
    p1t.x += sx * p1t.z; p1t.y += sy * p1t.z;
    p2t.x += sx * p2t.z; p2t.y += sy * p2t.z;
    const float e0 = p1t.x * p2t.y - p1t.y * p2t.x;
    const float e1 = p2t.x * p0t.y - p2t.y * p0t.x;
    const float e2 = p0t.x * p1t.y - p0t.y * p1t.x;
    const float det = e0 + e1 + e2;
    const float inv_det = 1.f / det;
    b[0] = e0 * inv_det; b[1] = e1 * inv_det; b[2] = e2 * inv_det;
}

CP_DEV V3 tri_normal(const TriVerts& t) { return normalize(normalize(cross(t.p0 - t.p2, t.p1 - t.p2))); }   // triangle.rs:78-82
CP_DEV V3 tri_point(const TriVerts& t, float b0, float b1, float b2) { return b0 * t.p0 + b1 * t.p1 + b2 * t.p2; }
CP_DEV V3 tri_abs_sum(const TriVerts& t, float b0, float b1, float b2) { return vabs(b0 * t.p0) + vabs(b1 * t.p1) + vabs(b2 * t.p2); }
CP_DEV float tri_area(const TriVerts& t) { return 0.5f * length(cross(t.p1 - t.p0, t.p2 - t.p0)); }
CP_DEV P2 tri_uv(const float* __restrict__ uv6, float b0, float b1, float b2) {
    P2 r;
    if (uv6) { r.x = (b0 * uv6[0] + b1 * uv6[2]) + b2 * uv6[4]; r.y = (b0 * uv6[1] + b1 * uv6[3]) + b2 * uv6[5]; }
    else { r.x = (b0 * 0.f + b1 * 1.f) + b2 * 1.f; r.y = (b0 * 0.f + b1 * 0.f) + b2 * 1.f; }   // (0,0),(1,0),(1,1) (triangle.rs:67-77)
    return r;
}

// solve_quadratic (sphere.rs:111-132), f64 as in the reference
CP_DEV bool solve_quadratic(float af, float bf, float cf, float* t0o, float* t1o) {
    double a = af, b = bf, c = cf;
    double discrim = b * b - 4. * a * c;
    if (discrim < 0.) return false;
    double root = sqrt(discrim);
    double q = b < 0. ? -0.5 * (b - root) : -0.5 * (b + root);
    double t0 = q / a, t1 = c / q;
    if (t0 > t1) { double tmp = t0; t0 = t1; t1 = tmp; }
    *t0o = (float)t0; *t1o = (float)t1;
    return true;
}
// Sphere::intersect acceptance (sphere.rs:29-40)
CP_DEV bool sphere_intersect_core(float radius, const Ray& ray, float* t_out) {
    float a = length_sq(ray.d);
    float b = 2.f * dot(ray.d, ray.o);
    float c = length_sq(ray.o) - radius * radius;
    float t0, t1;
    if (!solve_quadratic(a, b, c, &t0, &t1)) return false;
    if (t0 > ray.t_max || t1 < 0.f) return false;
    float t = t0 < 0.f ? t1 : t0;
    if (t > ray.t_max) return false;
    *t_out = t;
    return true;
}
CP_DEV P2 sphere_uv(float radius, V3 p) {   // sphere.rs:14-18
    P2 r;
    r.x = (atan2f(p.y, p.x) + kPi) * 0.5f * kInvPi;
    r.y = acosf(rclamp(p.z / radius, -1.f, 1.f)) * kInvPi;
    return r;
}
CP_DEV ShapePoint sphere_shape_point(float radius, const Ray& ray, float t) {   // sphere.rs:41-51
    ShapePoint s;
    V3 p = ray.o + ray.d * t;
    p = p * (radius / length(p));
    s.p = p;
    s.n = normalize(normalize(p));
    s.p_error = gamma_n(5) * vabs(p);
    s.uv = sphere_uv(radius, p);
    return s;
}

// Ray through the inverse transforms of a TransformShape chain, outermost wrapper first
// (transform.rs:47-59; Ray::apply ray.rs:45-51).  Chains are short (an instanced, transformed
// shape has two links).
CP_DEV Ray ray_to_object(const DevScene& sc, int32_t xform, Ray r) {
    int32_t chain[8]; int n = 0;
    for (int32_t x = xform; x >= 0 && n < 8; x = sc.transforms[x].outer) chain[n++] = x;
    for (int i = n - 1; i >= 0; --i) {
        const float* mi = sc.transforms[chain[i]].m_inv;
        r.o = apply_point(mi, r.o);
        r.d = apply_vector(mi, r.d);
    }
    return r;
}
CP_DEV ShapePoint shape_point_to_world(const DevScene& sc, int32_t xform, ShapePoint s) {
    for (int32_t x = xform; x >= 0; x = sc.transforms[x].outer) { CP_CHECK((uint32_t)x < sc.n_transforms, 202); s = shape_point_apply(s, sc.transforms[x].m, sc.transforms[x].m_inv); }
    return s;
}

struct PrimRecord { float4 q0, q1, q2; };
CP_DEV PrimRecord load_prim(const DevScene& sc, uint32_t slot) {
    PrimRecord r;
    CP_CHECK(slot < sc.n_prims, 201);
    const float4* p = sc.prims + 3ull * slot;
    r.q0 = __ldg(p); r.q1 = __ldg(p + 1); r.q2 = __ldg(p + 2);
    return r;
}
CP_DEV TriVerts prim_tri(const PrimRecord& r) {
    TriVerts t;
    t.p0 = v3(r.q0.x, r.q0.y, r.q0.z); t.p1 = v3(r.q0.w, r.q1.x, r.q1.y); t.p2 = v3(r.q1.z, r.q1.w, r.q2.x);
    return t;
}
CP_DEV uint32_t prim_upload_index(const PrimRecord& r) { return __float_as_uint(r.q2.y); }
CP_DEV uint32_t prim_meta(const PrimRecord& r) { return __float_as_uint(r.q2.z); }
CP_DEV uint32_t prim_source(const PrimRecord& r) { return __float_as_uint(r.q2.w); }

// Primitive::intersect acceptance for one BVH slot: true and *t on a hit with t <= ray.t_max.
// `rs` is the watertight setup of the WORLD ray (valid for untransformed triangles).
CP_DEV bool prim_intersect_t(const DevScene& sc, const PrimRecord& pr, const Ray& ray, const TriRaySetup& rs, float* t) {
    const uint32_t meta = prim_meta(pr);
    if (meta <= kMetaHasUv) {   // untransformed triangle: the common case
        return tri_intersect_core(prim_tri(pr), ray.o, ray.t_max, rs, t, nullptr);
    }
    const int32_t xform = (int32_t)(meta >> 2) - 1;
    Ray r = xform >= 0 ? ray_to_object(sc, xform, ray) : ray;
    if (meta & kMetaSphere) return sphere_intersect_core(pr.q0.x, r, t);
    TriRaySetup rs2 = tri_ray_setup(r.d);
    return tri_intersect_core(prim_tri(pr), r.o, r.t_max, rs2, t, nullptr);
}

// Full ShapeIntersect of a slot the traversal already chose (world-space ShapePoint).
CP_DEV ShapePoint prim_shape_point(const DevScene& sc, uint32_t slot, const PrimRecord& pr, const Ray& ray_in, float t_hit) {
    const uint32_t meta = prim_meta(pr);
    const int32_t xform = (int32_t)(meta >> 2) - 1;
    Ray r = ray_in; r.t_max = kFloatMax;
    if (xform >= 0) r = ray_to_object(sc, xform, r);
    ShapePoint s;
    if (meta & kMetaSphere) {
        s = sphere_shape_point(pr.q0.x, r, t_hit);
    } else {
        TriVerts tv = prim_tri(pr);
        float b[3];
        tri_barycentrics(tv, r.o, r.d, b);   // same arithmetic as the accepted test, minus what only gates acceptance
        s.p_error = gamma_n(7) * tri_abs_sum(tv, b[0], b[1], b[2]);
        s.p = tri_point(tv, b[0], b[1], b[2]);
        if (sc.prim_frame) { const float4 f0 = __ldg(&sc.prim_frame[3ull * slot]); s.n = v3(f0.x, f0.y, f0.z); }
        else s.n = tri_normal(tv);
        s.uv = tri_uv((meta & kMetaHasUv) ? sc.prim_uv + 6ull * slot : nullptr, b[0], b[1], b[2]);
    }
    if (xform >= 0) s = shape_point_to_world(sc, xform, s);
    return s;
}

}  // namespace cpbrt
