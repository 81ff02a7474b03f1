// Host-side linear algebra used while building the scene description (camera matrices, CTM,
// clipper).  Arithmetic order follows what the reference gets from nalgebra 0.20 so that the
// matrices handed to the device are the ones the Rust host would have produced
// (src/geometry/transform.rs, point.rs:10-14, vector.rs:11-15, bounds.rs:292-300).
// Compile with -ffp-contract=off: Rust never contracts a*b+c into an fma.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

namespace cpbrt_host {

struct V3 {
    float x, y, z;
    float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
    float& at(int i) { return i == 0 ? x : (i == 1 ? y : z); }
};
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a) { return {-a.x, -a.y, -a.z}; }
// nalgebra 3-vector dot: (x*x' + y*y') + z*z'
inline float dot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline float norm(V3 a) { return std::sqrt(dot(a, a)); }
inline V3 normalize(V3 a) { float n = norm(a); return {a.x / n, a.y / n, a.z / n}; }
inline V3 cross(V3 a, V3 b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}

// reference min/max (src/math/mod.rs:36-49): NaN-asymmetric on purpose.
inline float rmin(float l, float r) { return r < l ? r : l; }
inline float rmax(float l, float r) { return r > l ? r : l; }

struct M4 {
    float m[4][4];   // row-major
    static M4 identity() {
        M4 r; std::memset(&r, 0, sizeof r);
        for (int i = 0; i < 4; ++i) r.m[i][i] = 1.f;
        return r;
    }
    static M4 rows(float a00, float a01, float a02, float a03, float a10, float a11, float a12, float a13,
                   float a20, float a21, float a22, float a23, float a30, float a31, float a32, float a33) {
        M4 r = {{{a00, a01, a02, a03}, {a10, a11, a12, a13}, {a20, a21, a22, a23}, {a30, a31, a32, a33}}};
        return r;
    }
    M4 transpose() const {
        M4 r;
        for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) r.m[i][j] = m[j][i];
        return r;
    }
};

// Matrix4 * Matrix4, nalgebra static path: per result column, axpy over the columns of the
// left factor: ((a_i0 b_0j + a_i1 b_1j) + a_i2 b_2j) + a_i3 b_3j.
inline M4 mul(const M4& a, const M4& b) {
    M4 r;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j)
            r.m[i][j] = ((a.m[i][0] * b.m[0][j] + a.m[i][1] * b.m[1][j]) + a.m[i][2] * b.m[2][j]) + a.m[i][3] * b.m[3][j];
    return r;
}

// Matrix4::try_inverse for 4x4: nalgebra's closed-form cofactor expansion on the column-major
// element slice (the classic gluInvertMatrix formulation), then scale by 1/det.
inline bool inverse(const M4& a, M4* out) {
    float s[16], inv[16];
    for (int c = 0; c < 4; ++c) for (int r = 0; r < 4; ++r) s[c * 4 + r] = a.m[r][c];
    inv[0] = s[5] * s[10] * s[15] - s[5] * s[11] * s[14] - s[9] * s[6] * s[15] + s[9] * s[7] * s[14] + s[13] * s[6] * s[11] - s[13] * s[7] * s[10];
    inv[4] = -s[4] * s[10] * s[15] + s[4] * s[11] * s[14] + s[8] * s[6] * s[15] - s[8] * s[7] * s[14] - s[12] * s[6] * s[11] + s[12] * s[7] * s[10];
    inv[8] = s[4] * s[9] * s[15] - s[4] * s[11] * s[13] - s[8] * s[5] * s[15] + s[8] * s[7] * s[13] + s[12] * s[5] * s[11] - s[12] * s[7] * s[9];
    inv[12] = -s[4] * s[9] * s[14] + s[4] * s[10] * s[13] + s[8] * s[5] * s[14] - s[8] * s[6] * s[13] - s[12] * s[5] * s[10] + s[12] * s[6] * s[9];
    inv[1] = -s[1] * s[10] * s[15] + s[1] * s[11] * s[14] + s[9] * s[2] * s[15] - s[9] * s[3] * s[14] - s[13] * s[2] * s[11] + s[13] * s[3] * s[10];
    inv[5] = s[0] * s[10] * s[15] - s[0] * s[11] * s[14] - s[8] * s[2] * s[15] + s[8] * s[3] * s[14] + s[12] * s[2] * s[11] - s[12] * s[3] * s[10];
    inv[9] = -s[0] * s[9] * s[15] + s[0] * s[11] * s[13] + s[8] * s[1] * s[15] - s[8] * s[3] * s[13] - s[12] * s[1] * s[11] + s[12] * s[3] * s[9];
    inv[13] = s[0] * s[9] * s[14] - s[0] * s[10] * s[13] - s[8] * s[1] * s[14] + s[8] * s[2] * s[13] + s[12] * s[1] * s[10] - s[12] * s[2] * s[9];
    inv[2] = s[1] * s[6] * s[15] - s[1] * s[7] * s[14] - s[5] * s[2] * s[15] + s[5] * s[3] * s[14] + s[13] * s[2] * s[7] - s[13] * s[3] * s[6];
    inv[6] = -s[0] * s[6] * s[15] + s[0] * s[7] * s[14] + s[4] * s[2] * s[15] - s[4] * s[3] * s[14] - s[12] * s[2] * s[7] + s[12] * s[3] * s[6];
    inv[10] = s[0] * s[5] * s[15] - s[0] * s[7] * s[13] - s[4] * s[1] * s[15] + s[4] * s[3] * s[13] + s[12] * s[1] * s[7] - s[12] * s[3] * s[5];
    inv[14] = -s[0] * s[5] * s[14] + s[0] * s[6] * s[13] + s[4] * s[1] * s[14] - s[4] * s[2] * s[13] - s[12] * s[1] * s[6] + s[12] * s[2] * s[5];
    inv[3] = -s[1] * s[6] * s[11] + s[1] * s[7] * s[10] + s[5] * s[2] * s[11] - s[5] * s[3] * s[10] - s[9] * s[2] * s[7] + s[9] * s[3] * s[6];
    inv[7] = s[0] * s[6] * s[11] - s[0] * s[7] * s[10] - s[4] * s[2] * s[11] + s[4] * s[3] * s[10] + s[8] * s[2] * s[7] - s[8] * s[3] * s[6];
    inv[11] = -s[0] * s[5] * s[11] + s[0] * s[7] * s[9] + s[4] * s[1] * s[11] - s[4] * s[3] * s[9] - s[8] * s[1] * s[7] + s[8] * s[3] * s[5];
    inv[15] = s[0] * s[5] * s[10] - s[0] * s[6] * s[9] - s[4] * s[1] * s[10] + s[4] * s[2] * s[9] + s[8] * s[1] * s[6] - s[8] * s[2] * s[5];
    float det = s[0] * inv[0] + s[1] * inv[4] + s[2] * inv[8] + s[3] * inv[12];
    if (det == 0.f) return false;
    float inv_det = 1.f / det;
    for (int c = 0; c < 4; ++c) for (int r = 0; r < 4; ++r) out->m[r][c] = inv[c * 4 + r] * inv_det;
    return true;
}

// Transform{m, m_inv} (transform.rs:7-10)
struct Xf {
    M4 m, m_inv;
    static Xf identity() { return {M4::identity(), M4::identity()}; }
    Xf inverse() const { return {m_inv, m}; }
    // Transformable for Transform (transform.rs:148-152): new = other * self
    Xf apply(const Xf& t) const { return {mul(t.m, m), mul(m_inv, t.m_inv)}; }
};

// Point3f::apply (point.rs:10-14): homogeneous multiply then divide by w (w == 0 -> panic in
// the reference; reported through *ok).
inline V3 apply_point(const M4& a, V3 p, bool* ok = nullptr) {
    float r[4];
    for (int i = 0; i < 4; ++i)
        r[i] = ((a.m[i][0] * p.x + a.m[i][1] * p.y) + a.m[i][2] * p.z) + a.m[i][3] * 1.f;
    if (r[3] == 0.f) { if (ok) *ok = false; return {r[0], r[1], r[2]}; }
    return {r[0] / r[3], r[1] / r[3], r[2] / r[3]};
}
// Vector3f::apply (vector.rs:11-15)
inline V3 apply_vector(const M4& a, V3 v) {
    float r[3];
    for (int i = 0; i < 3; ++i)
        r[i] = ((a.m[i][0] * v.x + a.m[i][1] * v.y) + a.m[i][2] * v.z) + a.m[i][3] * 0.f;
    return {r[0], r[1], r[2]};
}

struct B3 {   // Bounds3f
    V3 mn, mx;
    V3 corner(int i) const {   // bounds.rs:58-66
        return {(i & 1) ? mx.x : mn.x, (i & 2) ? mx.y : mn.y, (i & 4) ? mx.z : mn.z};
    }
};
inline B3 bounds_of_points(V3 a, V3 b) {   // Bounds::new (bounds.rs:45-57)
    B3 r = {a, a};
    for (int i = 0; i < 3; ++i) { r.mn.at(i) = rmin(r.mn[i], b[i]); r.mx.at(i) = rmax(r.mx[i], b[i]); }
    return r;
}
inline B3 bounds_or_point(B3 r, V3 p) {    // BitOr<&Point> (bounds.rs:172-185)
    for (int i = 0; i < 3; ++i) { r.mn.at(i) = rmin(r.mn[i], p[i]); r.mx.at(i) = rmax(r.mx[i], p[i]); }
    return r;
}
inline B3 bounds_apply(const B3& b, const M4& a) {   // Transformable for Bounds3f (bounds.rs:292-300)
    V3 p0 = apply_point(a, b.mn);
    B3 r = {p0, p0};
    for (int i = 1; i < 8; ++i) r = bounds_or_point(r, apply_point(a, b.corner(i)));
    return r;
}

}  // namespace cpbrt_host
