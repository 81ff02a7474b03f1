// Device-side value types and small math for the render hot path.
//
// The integer outputs of this path (BVH slot of the closest hit, Halton indices) must be
// bit-identical to what the reference computes with Rust f32 arithmetic, which never fuses a
// multiply with an add.  This translation unit is therefore compiled with -fmad=false
// -prec-div=true -prec-sqrt=true and the expressions below keep the reference's association
// order (nalgebra: dot = (x*x' + y*y') + z*z', normalize = v / sqrt(dot), matrix*vector
// accumulated left to right).  Reference: src/math/mod.rs, src/geometry/{point,vector,normal}.rs.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace cpbrt {

#define CP_DEV __device__ __forceinline__

// Checked build (-DCPBRT_CHECKED, tools/variant_sweep.py "checked"): every index into the path state, the queues,
// the primitive / node / light / material arrays and the sampler tables is compared with its bound; a violation is
// counted (and the first site remembered) instead of trapping, so that a whole workload runs to its end and the
// count is read back with cpbrt_check_failures.  compute-sanitizer is closed on the GPU pool this was developed on;
// this is the bounds-checking half of what memcheck would have given.  Costs nothing in the normal build.
#ifdef CPBRT_CHECKED
__device__ unsigned long long g_check_failures;
__device__ unsigned int g_check_first_site;
#define CP_CHECK(cond, site) do { if (!(cond)) { if (atomicAdd(&g_check_failures, 1ull) == 0ull) g_check_first_site = (unsigned int)(site); } } while (0)
#else
#define CP_CHECK(cond, site) do { } while (0)
#endif

static constexpr float kMachineEpsilon = 1.1920929e-07f * 0.5f;   // src/def.rs:4
static constexpr float kInvPi = 0.31830988618379067154f;          // src/math/mod.rs:213
static constexpr float kPi = 3.14159265358979323846f;             // src/math/mod.rs:216
static constexpr float kFloatMax = 3.402823466e+38f;

// reference min/max (src/math/mod.rs:36-49): the left operand survives a false comparison
CP_DEV float rmin(float l, float r) { return r < l ? r : l; }
CP_DEV float rmax(float l, float r) { return r > l ? r : l; }
CP_DEV float rclamp(float t, float lo, float hi) { return t < lo ? lo : (t > hi ? hi : t); }
// gamma(n) (src/math/mod.rs:76-79), folded at compile time with the same f32 operations
__host__ __device__ constexpr float gamma_n(int n) { return ((float)n * kMachineEpsilon) / (1.f - (float)n * kMachineEpsilon); }
CP_DEV float power_heuristic(float f, float g) { return (f * f) / (f * f + g * g); }

struct V3 { float x, y, z; };
CP_DEV V3 v3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
CP_DEV float comp(V3 v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : v.z); }
CP_DEV V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
CP_DEV V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
CP_DEV V3 operator-(V3 a) { return v3(-a.x, -a.y, -a.z); }
CP_DEV V3 operator*(V3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
CP_DEV V3 operator*(float s, V3 a) { return v3(s * a.x, s * a.y, s * a.z); }
// IEEE a / s, same bits as the plain operator.  nvcc's inline division leaves a ZERO dividend to an
// out-of-line path that the whole warp then walks through (FCHK), and zero components are the
// common case here (axis-aligned normals, black channels): 0 / s is answered as 0 * s (same
// signed zero) whenever s is an ordinary number, everything else takes the normal sequence.
CP_DEV bool div_ordinary(float s) { const float a = fabsf(s); return a >= 1e-30f && a <= 1e30f; }
CP_DEV float div_by(float a, float s, bool ordinary) {
    const bool z = ordinary && a == 0.f;
    float num = z ? 1.f : a;
    asm("" : "+f"(num));   // opaque: keeps the compiler from folding this back into a / s
    const float q = num / s;
    return z ? a * s : q;
}
CP_DEV V3 operator/(V3 a, float s) { const bool o = div_ordinary(s); return v3(div_by(a.x, s, o), div_by(a.y, s, o), div_by(a.z, s, o)); }
CP_DEV float dot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
CP_DEV float length_sq(V3 a) { return dot(a, a); }
CP_DEV float length(V3 a) { return sqrtf(dot(a, a)); }
CP_DEV V3 normalize(V3 a) { return a / length(a); }
CP_DEV V3 cross(V3 a, V3 b) { return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
CP_DEV V3 vabs(V3 a) { return v3(fabsf(a.x), fabsf(a.y), fabsf(a.z)); }
CP_DEV float vmax(V3 a) { float m = a.x; if (a.y > m) m = a.y; if (a.z > m) m = a.z; return m; }

struct P2 { float x, y; };

struct Spec { float r, g, b; };   // RGBSpectrum (src/spectrum/rgb_spectrum.rs:8-97)
CP_DEV Spec spec(float v) { Spec s; s.r = v; s.g = v; s.b = v; return s; }
CP_DEV Spec spec3(float r, float g, float b) { Spec s; s.r = r; s.g = g; s.b = b; return s; }
CP_DEV Spec operator+(Spec a, Spec b) { return spec3(a.r + b.r, a.g + b.g, a.b + b.b); }
CP_DEV Spec operator-(Spec a, Spec b) { return spec3(a.r - b.r, a.g - b.g, a.b - b.b); }
CP_DEV Spec operator*(Spec a, Spec b) { return spec3(a.r * b.r, a.g * b.g, a.b * b.b); }
CP_DEV Spec operator*(Spec a, float s) { return spec3(a.r * s, a.g * s, a.b * s); }
CP_DEV Spec operator/(Spec a, float s) { const bool o = div_ordinary(s); return spec3(div_by(a.r, s, o), div_by(a.g, s, o), div_by(a.b, s, o)); }
CP_DEV float lum(Spec s) { return (0.212671f * s.r + 0.715160f * s.g) + 0.072169f * s.b; }   // y()
CP_DEV bool is_black(Spec s) { return s.r == 0.f && s.g == 0.f && s.b == 0.f; }
CP_DEV bool has_nan(Spec s) { return s.r != s.r || s.g != s.g || s.b != s.b; }

// Point3f::apply / Vector3f::apply / Normal3f::apply on a row-major 4x4 (point.rs:10-14,
// vector.rs:11-15, normal.rs:32-37)
CP_DEV V3 apply_point(const float* __restrict__ m, V3 p) {
    float r0 = ((m[0] * p.x + m[1] * p.y) + m[2] * p.z) + m[3] * 1.f;
    float r1 = ((m[4] * p.x + m[5] * p.y) + m[6] * p.z) + m[7] * 1.f;
    float r2 = ((m[8] * p.x + m[9] * p.y) + m[10] * p.z) + m[11] * 1.f;
    float r3 = ((m[12] * p.x + m[13] * p.y) + m[14] * p.z) + m[15] * 1.f;
    // x / 1 == x bit for bit: affine matrices (every Translate / Scale / Rotate / LookAt) skip the three IEEE divisions
    if (r3 != 0.f && r3 != 1.f) return v3(r0 / r3, r1 / r3, r2 / r3);
    return v3(r0, r1, r2);
}
CP_DEV V3 apply_vector(const float* __restrict__ m, V3 v) {
    return v3(((m[0] * v.x + m[1] * v.y) + m[2] * v.z) + m[3] * 0.f,
              ((m[4] * v.x + m[5] * v.y) + m[6] * v.z) + m[7] * 0.f,
              ((m[8] * v.x + m[9] * v.y) + m[10] * v.z) + m[11] * 0.f);
}
CP_DEV V3 apply_normal(const float* __restrict__ mi, V3 n) {
    return v3((mi[0] * n.x + mi[4] * n.y) + mi[8] * n.z, (mi[1] * n.x + mi[5] * n.y) + mi[9] * n.z, (mi[2] * n.x + mi[6] * n.y) + mi[10] * n.z);
}

CP_DEV void coordinate_system(V3 z, V3* x, V3* y) {   // src/math/mod.rs:67-74
    if (fabsf(z.x) > fabsf(z.y)) *x = v3(-z.z, 0.f, z.x) / sqrtf(z.x * z.x + z.z * z.z);
    else *x = v3(0.f, z.z, -z.y) / sqrtf(z.y * z.y + z.z * z.z);
    *y = cross(z, *x);
}
CP_DEV P2 concentric_sample_disk(P2 u) {   // src/math/mod.rs:98-110
    P2 o; o.x = 2.f * u.x - 1.f; o.y = 2.f * u.y - 1.f;
    P2 r; r.x = 0.f; r.y = 0.f;
    if (o.x == 0.f || o.y == 0.f) return r;
    float rad, theta;
    if (fabsf(o.x) > fabsf(o.y)) { rad = o.x; theta = (kPi / 4.f) * (o.y / o.x); }
    else { rad = o.y; theta = (kPi / 2.f) - (kPi / 4.f) * (o.x / o.y); }
    float sn, cs;
    sincosf(theta, &sn, &cs);          // one range reduction for both (same accuracy as sinf / cosf)
    r.x = rad * cs; r.y = rad * sn;
    return r;
}
CP_DEV V3 uniform_sample_sphere_q21(P2 u) {   // uniform_sample_hemisphere (math/mod.rs:111-116)
    float z = 1.f - 2.f * u.x;
    float r = sqrtf(rmax(0.f, 1.f - z * z));
    float phi = 2.f * kPi * u.y;
    float sn, cs;
    sincosf(phi, &sn, &cs);
    return v3(r * cs, r * sn, z);
}
CP_DEV V3 cosine_sample_hemisphere(P2 u, float* pdf) {   // math/mod.rs:117-121
    P2 d = concentric_sample_disk(u);
    float z = sqrtf(1.f - (d.x * d.x + d.y * d.y));
    *pdf = z * kInvPi;
    return v3(d.x, d.y, z);
}
CP_DEV P2 uniform_sample_triangle(P2 u) { float su0 = sqrtf(u.x); P2 r; r.x = 1.f - su0; r.y = u.y * su0; return r; }   // :123-126

CP_DEV float cos_theta(V3 w) { return w.z; }
CP_DEV float sin_2_theta(V3 w) { return rmax(1.f - w.z * w.z, 0.f); }
CP_DEV float sin_theta(V3 w) { return sqrtf(sin_2_theta(w)); }
CP_DEV float tan_theta(V3 w) { return sin_theta(w) / cos_theta(w); }
CP_DEV float cos_delta_phi(V3 wa, V3 wb) {
    return rclamp((wa.x * wb.x + wa.y * wb.y) / sqrtf((wa.x * wa.x + wa.y * wa.y) * (wb.x * wb.x + wb.y * wb.y)), -1.f, 1.f);
}
CP_DEV bool refract(V3 wo, V3 n, float eta, V3* wt) {   // math/mod.rs:202-212
    float cos_o = dot(wo, n);
    float sin2_o = 1.f - cos_o * cos_o;
    float sin2_i = sin2_o * eta * eta;
    if (sin2_i > 1.f) return false;
    float cos_i = sqrtf(rmax(1.f - sin2_i, 0.f));
    *wt = eta * (-wo) + (eta * cos_o - cos_i) * n;
    return true;
}

// Rust `f32 as usize` (saturating, NaN -> 0), for values that fit 32 bits after the caller's min()
CP_DEV uint32_t f32_to_u32_sat(float f) { return __float2uint_rz(f); }   // cvt.rzi.u32.f32 saturates; NaN -> 0

}  // namespace cpbrt
