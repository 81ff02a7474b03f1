// ORACLE -- TEST INFRASTRUCTURE ONLY.
// CPU restatement of curry-pbrt's rendering hot path, used to check the CUDA path and as the
// CPU baseline in bench.py.  Nothing under curry-pbrt_b200/ may include, link or call this.
//
// Parity status: the Rust reference cannot be built in this image (no cargo/rustc, 18
// un-vendored crates), so this restatement is pinned only by the reference's two unit tests
// (halton_works, perspective_works) and by values derived from its sources; BVH order, hit
// ids, shading and film are "parity unpinned" against a real reference run.
// What the restatement is additionally held to (tests/test_oracle.py), none of it a reference run:
// an independent brute-force aggregate, exact integer / rational arithmetic for the scrambled
// radical inverse, float64 Moeller-Trumbore + analytic-sphere geometry over all primitives, the
// closed form of the camera, and analytic radiometry (white furnace for matte / mirror / glass,
// point, distant, small-area and spherical emitters) -- these pin its arithmetic and physics, not
// the reference's quirks.  The quirks are held by a second, independent reading of the Rust:
// oracle/{front,sah,trace,path}_numpy.py (Python integers + scalar np.float32, nothing shared with
// this file) against which this restatement is bit-equal in every pixel, BVH node, hit, t and
// visit count on the four file-based BASELINE scenes (tests/test_oracle.py,
// *_against_independent_numpy_restatement).  Two readings that agree are still not a reference run.
//
// Value types and small math.  Follows src/math/mod.rs, src/geometry/{point,vector,normal}.rs and
// the nalgebra 0.20 arithmetic they sit on (dot = (x*x'+y*y')+z*z', normalize = v / sqrt(dot),
// Matrix4*Vector4 accumulated left to right, no fused multiply-add anywhere).
// Build with -ffp-contract=off.
#pragma once
#include <cmath>
#include <cstdint>
#include <limits>

namespace orc {

typedef float Float;

static const Float MACHINE_EPSILON = 1.1920929e-07f * 0.5f;   // src/def.rs:4  f32::EPSILON * 0.5
static const Float INV_PI = 0.31830988618379067154f;           // src/math/mod.rs:213
static const Float PI = 3.14159265358979323846f;               // src/math/mod.rs:216
static const Float FLOAT_MAX = std::numeric_limits<float>::max();

// src/math/mod.rs:36-49 -- note which operand wins when the comparison is false (NaN, ties)
template <class T> inline T rmin(T lhs, T rhs) { return rhs < lhs ? rhs : lhs; }
template <class T> inline T rmax(T lhs, T rhs) { return rhs > lhs ? rhs : lhs; }
// src/math/mod.rs:22-30
inline Float rclamp(Float t, Float lo, Float hi) { return t < lo ? lo : (t > hi ? hi : t); }
// src/math/mod.rs:76-79
inline Float gamma_n(int n) { Float e = (Float)n * MACHINE_EPSILON; return e / (1.f - e); }
// src/math/mod.rs:32-34
inline Float power_heuristic(Float f, Float g) { return (f * f) / (f * f + g * g); }

struct V3 {
    Float x, y, z;
    Float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
inline V3 v3(Float x, Float y, Float z) { V3 r = {x, y, z}; return r; }
inline V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline V3 operator-(V3 a) { return v3(-a.x, -a.y, -a.z); }
inline V3 operator*(V3 a, Float s) { return v3(a.x * s, a.y * s, a.z * s); }
inline V3 operator*(Float s, V3 a) { return v3(s * a.x, s * a.y, s * a.z); }
inline V3 operator/(V3 a, Float s) { return v3(a.x / s, a.y / s, a.z / s); }
inline Float dot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline Float length_sq(V3 a) { return dot(a, a); }
inline Float length(V3 a) { return std::sqrt(dot(a, a)); }
inline V3 normalize(V3 a) { return a / length(a); }
inline V3 cross(V3 a, V3 b) { return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
inline V3 vabs(V3 a) { return v3(std::fabs(a.x), std::fabs(a.y), std::fabs(a.z)); }
// nalgebra imax: first maximum wins (strict >)
inline int imax(V3 a) { int i = 0; Float m = a.x; if (a.y > m) { m = a.y; i = 1; } if (a.z > m) { i = 2; } return i; }
inline Float vmax(V3 a) { Float m = a.x; if (a.y > m) m = a.y; if (a.z > m) m = a.z; return m; }

struct P2 { Float x, y; };

struct Spectrum {   // RGBSpectrum (src/spectrum/rgb_spectrum.rs:8-97)
    Float c[3];
    Float y() const { return (0.212671f * c[0] + 0.715160f * c[1]) + 0.072169f * c[2]; }
    bool is_black() const { return c[0] == 0.f && c[1] == 0.f && c[2] == 0.f; }
    bool has_nan() const { return c[0] != c[0] || c[1] != c[1] || c[2] != c[2]; }
};
inline Spectrum spec(Float v) { Spectrum s = {{v, v, v}}; return s; }
inline Spectrum spec3(const Float* v) { Spectrum s = {{v[0], v[1], v[2]}}; return s; }
inline Spectrum operator+(Spectrum a, Spectrum b) { Spectrum r = {{a.c[0] + b.c[0], a.c[1] + b.c[1], a.c[2] + b.c[2]}}; return r; }
inline Spectrum operator-(Spectrum a, Spectrum b) { Spectrum r = {{a.c[0] - b.c[0], a.c[1] - b.c[1], a.c[2] - b.c[2]}}; return r; }
inline Spectrum operator*(Spectrum a, Spectrum b) { Spectrum r = {{a.c[0] * b.c[0], a.c[1] * b.c[1], a.c[2] * b.c[2]}}; return r; }
inline Spectrum operator*(Spectrum a, Float s) { Spectrum r = {{a.c[0] * s, a.c[1] * s, a.c[2] * s}}; return r; }
inline Spectrum operator/(Spectrum a, Float s) { Spectrum r = {{a.c[0] / s, a.c[1] / s, a.c[2] / s}}; return r; }

// 4x4 row-major matrix applied the way nalgebra does it (point.rs:10-14, vector.rs:11-15)
inline V3 apply_point(const Float* m, V3 p) {
    Float r[4];
    for (int i = 0; i < 4; ++i) r[i] = ((m[i * 4] * p.x + m[i * 4 + 1] * p.y) + m[i * 4 + 2] * p.z) + m[i * 4 + 3] * 1.f;
    if (r[3] != 0.f) return v3(r[0] / r[3], r[1] / r[3], r[2] / r[3]);
    return v3(r[0], r[1], r[2]);   // the reference unwraps a None here (panic); never reached with affine matrices
}
inline V3 apply_vector(const Float* m, V3 v) {
    Float r[3];
    for (int i = 0; i < 3; ++i) r[i] = ((m[i * 4] * v.x + m[i * 4 + 1] * v.y) + m[i * 4 + 2] * v.z) + m[i * 4 + 3] * 0.f;
    return v3(r[0], r[1], r[2]);
}
// Normal3f::apply (normal.rs:32-37): (m_inv^T)_3x3 * n, not renormalised
inline V3 apply_normal(const Float* m_inv, V3 n) {
    return v3((m_inv[0] * n.x + m_inv[4] * n.y) + m_inv[8] * n.z,
              (m_inv[1] * n.x + m_inv[5] * n.y) + m_inv[9] * n.z,
              (m_inv[2] * n.x + m_inv[6] * n.y) + m_inv[10] * n.z);
}

// src/math/mod.rs:67-74
inline void coordinate_system(V3 z, V3* x, V3* y) {
    if (std::fabs(z.x) > std::fabs(z.y)) *x = v3(-z.z, 0.f, z.x) / std::sqrt(z.x * z.x + z.z * z.z);
    else *x = v3(0.f, z.z, -z.y) / std::sqrt(z.y * z.y + z.z * z.z);
    *y = cross(z, *x);
}
// src/math/mod.rs:98-110
inline P2 concentric_sample_disk(P2 u) {
    P2 o = {2.f * u.x - 1.f, 2.f * u.y - 1.f};
    if (o.x == 0.f || o.y == 0.f) { P2 z = {0.f, 0.f}; return z; }
    Float r, theta;
    if (std::fabs(o.x) > std::fabs(o.y)) { r = o.x; theta = (PI / 4.f) * (o.y / o.x); }
    else { r = o.y; theta = (PI / 2.f) - (PI / 4.f) * (o.x / o.y); }
    P2 p = {r * std::cos(theta), r * std::sin(theta)};
    return p;
}
// src/math/mod.rs:111-116 (full sphere despite its name, Q21)
inline V3 uniform_sample_hemisphere(P2 u) {
    Float z = 1.f - 2.f * u.x;
    Float r = std::sqrt(rmax(0.f, 1.f - z * z));
    Float phi = 2.f * PI * u.y;
    return v3(r * std::cos(phi), r * std::sin(phi), z);
}
// src/math/mod.rs:117-121
inline V3 cosine_sample_hemisphere(P2 u, Float* pdf) {
    P2 d = concentric_sample_disk(u);
    Float z = std::sqrt(1.f - (d.x * d.x + d.y * d.y));
    *pdf = z * INV_PI;
    return v3(d.x, d.y, z);
}
// src/math/mod.rs:123-126
inline P2 uniform_sample_triangle(P2 u) { Float su0 = std::sqrt(u.x); P2 r = {1.f - su0, u.y * su0}; return r; }

// shading-frame trigonometry, src/math/mod.rs:152-201
inline Float cos_theta(V3 w) { return w.z; }
inline Float cos_2_theta(V3 w) { return w.z * w.z; }
inline Float sin_2_theta(V3 w) { return rmax(1.f - cos_2_theta(w), 0.f); }
inline Float sin_theta(V3 w) { return std::sqrt(sin_2_theta(w)); }
inline Float tan_theta(V3 w) { return sin_theta(w) / cos_theta(w); }
inline Float tan_2_theta(V3 w) { return sin_2_theta(w) / cos_2_theta(w); }
inline Float cos_phi(V3 w) { Float s = sin_theta(w); return s == 0.f ? 1.f : rclamp(w.x / s, -1.f, 1.f); }
inline Float sin_phi(V3 w) { Float s = sin_theta(w); return s == 0.f ? 0.f : rclamp(w.y / s, -1.f, 1.f); }
inline Float cos_2_phi(V3 w) { Float c = cos_phi(w); return c * c; }
inline Float sin_2_phi(V3 w) { Float s = sin_phi(w); return s * s; }
inline Float cos_delta_phi(V3 wa, V3 wb) {
    return rclamp((wa.x * wb.x + wa.y * wb.y) / std::sqrt((wa.x * wa.x + wa.y * wa.y) * (wb.x * wb.x + wb.y * wb.y)), -1.f, 1.f);
}
// src/math/mod.rs:202-212
inline bool refract(V3 wo, V3 n, Float eta, V3* wt) {
    Float cos_o = dot(wo, n);
    Float sin2_o = 1.f - cos_o * cos_o;
    Float sin2_i = sin2_o * eta * eta;
    if (sin2_i > 1.f) return false;
    Float cos_i = std::sqrt(rmax(1.f - sin2_i, 0.f));
    *wt = eta * (-wo) + (eta * cos_o - cos_i) * n;
    return true;
}

// Rust `f as usize`: saturating, NaN -> 0
inline uint64_t f32_to_usize(Float f) {
    if (!(f > 0.f)) return 0;
    if (f >= 18446744073709551616.f) return ~0ull;
    return (uint64_t)f;
}

}  // namespace orc
