"""ORACLE -- TEST INFRASTRUCTURE ONLY.

A second, independent restatement of the reference's BVH search and triangle test, scalar
np.float32 arithmetic written directly from the Rust (shares nothing with oracle/oracle.cpp).
Slow by design (a few hundred rays per second): the tests run it over a few thousand rays and hold
the C++ oracle -- and through it the device kernels -- to the same primitive and the same bits of t.

  RayIntersectCache::from          src/geometry/ray.rs (inverse_d, is_positive_d = d >= 0)
  Bounds3f::intersect_component    src/geometry/bounds.rs:306-311   far slab widened by 1 + 2*gamma(3)
  intersect_predicate_cached       src/geometry/bounds.rs:312-323   the interval is *widened* axis by
                                   axis (min of the mins, max of the maxes) -- the reference's loose test
  intersect_through_bound          src/aggregate/bvh.rs:151-191     far child first (by the split axis
                                   and the sign of d), a Leaf updates t_max on every hit, later hits win
  intersect_predicate_through_bound  bvh.rs:125-150                 left then right, first hit ends it
  Triangle::intersect              src/geometry/shape/triangle.rs:194-262 (and :127-192, the predicate)
  gamma(n)                         src/math/mod.rs:76-79, MACHINE_EPSILON = f32::EPSILON / 2 (def.rs:4)
  min / max                        src/math/mod.rs:36-49 (`rhs < lhs ? rhs : lhs`)

  Sphere::intersect, solve_quadratic  src/geometry/shape/sphere.rs:29-53,114-136 (f64 quadratic, p rescaled
                                   onto the sphere, normal normalised twice, p_error = gamma(5)|p|)
  TransformShape::{bound,intersect} src/geometry/shape/transform.rs:52-59: the ray goes through the inverse, the
                                   hit through ShapePoint::apply (shape/mod.rs:135-160: p_error is REPLACED by
                                   gamma(3) * (|m_i0 p'_x| + |m_i1 p'_y| + |m_i2 p'_z| + m_i3) of the transformed p')
  Point3f / Vector3f / Normal3f::apply  src/geometry/{point,vector,normal}.rs; nalgebra's matrix * vector adds the
                                   columns in order: ((m_i0 x + m_i1 y) + m_i2 z) + m_i3 w

The tree comes from oracle/sah_numpy.py (`SahTree`); `tris` are the (n, 3, 3) float32 vertices in
upload order, or a list of `TrianglePrim` / `SpherePrim`."""
import numpy as np

F = np.float32
_EPS = F(np.finfo(np.float32).eps) * F(0.5)


def gamma(n):
    ne = F(n) * _EPS
    return ne / (F(1) - ne)


_G2, _G3, _G5 = gamma(2), gamma(3), gamma(5)
_WIDEN = F(1) + F(2) * _G3


def _min(lhs, rhs):
    return rhs if rhs < lhs else lhs


def _max(lhs, rhs):
    return rhs if rhs > lhs else lhs


class _Ray:
    def __init__(self, o, d, t_max):
        self.o = [F(x) for x in o]
        self.d = [F(x) for x in d]
        self.t_max = F(t_max)
        self.inv = [F(1) / x for x in self.d]
        self.pos = [1 if x >= 0 else 0 for x in self.d]
        self.neg = [1 - p for p in self.pos]


def _slab(bound, ray):
    """bound = (lo[3], hi[3]) as lists of np.float32."""
    def component(c):
        return ((bound[ray.neg[c]][c] - ray.o[c]) * ray.inv[c],
                (bound[ray.pos[c]][c] - ray.o[c]) * ray.inv[c] * _WIDEN)
    t_min, t_max = component(0)
    for c in (1, 2):
        n_min, n_max = component(c)
        if t_min > n_max or n_min > t_max:
            return False
        t_min = _min(t_min, n_min)
        t_max = _max(t_max, n_max)
    return bool(t_min < ray.t_max and t_max > 0)


def _imax_abs(v):
    a = [abs(x) for x in v]
    best, idx = a[0], 0
    for i in (1, 2):
        if a[i] > best:
            best, idx = a[i], i
    return idx


def triangle_intersect(p, ray, barycentrics=None):
    """Returns t (np.float32) or None; `barycentrics` (a list) receives b0, b1, b2 = e * inv_det.
    `Triangle::intersect` and `intersect_predicate` accept and reject exactly the same rays (the
    same sequence up to the barycentrics)."""
    o = ray.o
    pt = [[p[k][c] - o[c] for c in range(3)] for k in range(3)]
    kz = _imax_abs(ray.d)
    kx = 0 if kz + 1 == 3 else kz + 1
    ky = 0 if kx + 1 == 3 else kx + 1
    d = [ray.d[kx], ray.d[ky], ray.d[kz]]
    pt = [[q[kx], q[ky], q[kz]] for q in pt]
    sx = -d[0] / d[2]
    sy = -d[1] / d[2]
    sz = F(1) / d[2]
    for q in pt:
        q[0] = q[0] + sx * q[2]
        q[1] = q[1] + sy * q[2]
    p0, p1, p2 = pt
    e0 = p1[0] * p2[1] - p1[1] * p2[0]
    e1 = p2[0] * p0[1] - p2[1] * p0[0]
    e2 = p0[0] * p1[1] - p0[1] * p1[0]
    if (e0 < 0 or e1 < 0 or e2 < 0) and (e0 > 0 or e1 > 0 or e2 > 0):
        return None
    det = e0 + e1 + e2
    if det == 0:
        return None
    for q in pt:
        q[2] = q[2] * sz
    t_scaled = e0 * p0[2] + e1 * p1[2] + e2 * p2[2]
    if (det < 0 and (t_scaled >= 0 or t_scaled < ray.t_max * det)) or \
       (det > 0 and (t_scaled <= 0 or t_scaled > ray.t_max * det)):
        return None
    inv_det = F(1) / det
    t = t_scaled * inv_det
    max_zt = max(abs(p0[2]), abs(p1[2]), abs(p2[2]))
    delta_z = _G3 * max_zt
    max_xt = max(abs(p0[0]), abs(p1[0]), abs(p2[0]))
    max_yt = max(abs(p0[1]), abs(p1[1]), abs(p2[1]))
    delta_x = _G5 * (max_xt + max_zt)
    delta_y = _G5 * (max_yt + max_zt)
    delta_e = F(2) * (_G2 * max_xt * max_yt + delta_y * max_xt + delta_x * max_yt)
    max_e = max(abs(e0), abs(e1), abs(e2))
    delta_t = F(3) * (_G3 * max_e * max_zt + delta_e * max_zt + delta_z * max_e) * abs(inv_det)
    if t <= delta_t:
        return None
    if barycentrics is not None:
        barycentrics[:] = [e0 * inv_det, e1 * inv_det, e2 * inv_det]
    return t


def mat_rows(m16):
    return [[F(m16[r * 4 + c]) for c in range(4)] for r in range(4)]


def mat_vec(m, v):
    return [((m[i][0] * v[0] + m[i][1] * v[1]) + m[i][2] * v[2]) + m[i][3] * v[3] for i in range(len(m))]


def mat_mul(a, b):
    """nalgebra gemm: each column of the product is a gemv."""
    cols = [[sum_in_order([a[i][k] * b[k][j] for k in range(len(b))]) for i in range(len(a))] for j in range(len(b[0]))]
    return [[cols[j][i] for j in range(len(cols))] for i in range(len(a))]


def sum_in_order(terms):
    acc = terms[0]
    for t in terms[1:]:
        acc = acc + t
    return acc


def apply_point(m, p):
    h = mat_vec(m, [p[0], p[1], p[2], F(1)])
    return [h[0] / h[3], h[1] / h[3], h[2] / h[3]]          # Point3::from_homogeneous


def apply_vector(m, v):
    return mat_vec(m, [v[0], v[1], v[2], F(0)])[:3]


def normal_matrix(m_inv):
    """Matrix3x4::identity() * (m_inv.transpose() * Matrix4x3::identity())  (normal.rs:33-36)."""
    eye34 = [[F(1 if r == c else 0) for c in range(4)] for r in range(3)]
    eye43 = [[F(1 if r == c else 0) for c in range(3)] for r in range(4)]
    transposed = [[m_inv[c][r] for c in range(4)] for r in range(4)]
    return mat_mul(eye34, mat_mul(transposed, eye43))


def transformed_bounds(lo, hi, m):
    """Bounds3f::apply (bounds.rs:292-300): the union of the eight transformed corners."""
    box = (lo, hi)
    corners = [apply_point(m, [box[(i >> c) & 1][c] for c in range(3)]) for i in range(8)]
    out_lo, out_hi = list(corners[0]), list(corners[0])
    for q in corners[1:]:
        for c in range(3):
            out_lo[c] = _min(out_lo[c], q[c])
            out_hi[c] = _max(out_hi[c], q[c])
    return out_lo, out_hi


class TrianglePrim:
    def __init__(self, verts):
        self.p = [[F(x) for x in v] for v in verts]

    def bounds(self):
        return [min(v[c] for v in self.p) for c in range(3)], [max(v[c] for v in self.p) for c in range(3)]

    def intersect(self, ray):
        """(t, barycentrics) or None."""
        b = []
        t = triangle_intersect(self.p, ray, b)
        return None if t is None else (t, b)


def _libm(name, nargs=1):
    import ctypes
    import ctypes.util
    fn = getattr(ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6"), name)
    fn.restype = ctypes.c_float
    fn.argtypes = [ctypes.c_float] * nargs
    return lambda *x: F(fn(*[float(v) for v in x]))


_atan2f, _acosf = _libm("atan2f", 2), _libm("acosf")
_PI, _INV_PI = F(3.14159265358979323846), F(0.31830988618379067154)


class BareSpherePrim:
    """Sphere{radius} in its own space (sphere.rs:14-53); intersect returns (t, (p, n, p_error, uv))."""

    def __init__(self, radius):
        self.radius = F(radius)

    def bounds(self):
        r = self.radius
        lo, hi = [-r, -r, -r], [r, r, r]
        return [_min(lo[c], hi[c]) for c in range(3)], [_max(lo[c], hi[c]) for c in range(3)]

    def intersect(self, ray):
        o, d = ray.o, ray.d
        a = (d[0] * d[0] + d[1] * d[1]) + d[2] * d[2]
        b = F(2) * ((d[0] * o[0] + d[1] * o[1]) + d[2] * o[2])
        c = ((o[0] * o[0] + o[1] * o[1]) + o[2] * o[2]) - self.radius * self.radius
        a, b, c = np.float64(a), np.float64(b), np.float64(c)
        discrim = b * b - np.float64(4) * a * c
        if discrim < 0:
            return None
        root = np.sqrt(discrim)
        q = np.float64(-0.5) * (b - root) if b < 0 else np.float64(-0.5) * (b + root)
        t0, t1 = q / a, c / q
        if t0 > t1:
            t0, t1 = t1, t0
        t0, t1 = F(t0), F(t1)
        if t0 > ray.t_max or t1 < 0:
            return None
        t = t1 if t0 < 0 else t0
        if t > ray.t_max:
            return None
        p = [o[k] + d[k] * t for k in range(3)]
        k = self.radius / np.sqrt((p[0] * p[0] + p[1] * p[1]) + p[2] * p[2])
        p = [x * k for x in p]

        def normalize(v):
            length = np.sqrt((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2])
            return [x / length for x in v]
        n = normalize(normalize(p))
        g5 = gamma(5)
        uv = [(_atan2f(p[1], p[0]) + _PI) * F(0.5) * _INV_PI, _acosf(min(max(p[2] / self.radius, F(-1)), F(1))) * _INV_PI]
        return t, (p, n, [g5 * abs(x) for x in p], uv)


class TransformedPrim:
    """TransformShape{shape, transform} (shape/transform.rs:52-59).  `to_shape_point` turns the inner primitive's hit
    detail into (p, n, p_error, uv) in the inner space (identity for spheres and nested wrappers)."""

    def __init__(self, inner, m16, m_inv16, to_shape_point=None):
        self.inner = inner
        self.m, self.m_inv = mat_rows(m16), mat_rows(m_inv16)
        self.m3 = normal_matrix(self.m_inv)
        self.to_shape_point = to_shape_point or (lambda detail: detail)

    def bounds(self):
        lo, hi = self.inner.bounds()
        return transformed_bounds(lo, hi, self.m)

    def intersect(self, ray):
        """(t, (p, n, p_error, uv)) in the outer space, or None."""
        hit = self.inner.intersect(_Ray(apply_point(self.m_inv, ray.o), apply_vector(self.m_inv, ray.d), ray.t_max))
        if hit is None:
            return None
        p, n, _dropped, uv = self.to_shape_point(hit[1])
        # ShapeIntersect::apply(transform): the inner p_error is dropped
        pw = apply_point(self.m, p)
        nw = [(self.m3[i][0] * n[0] + self.m3[i][1] * n[1]) + self.m3[i][2] * n[2] for i in range(3)]
        m = self.m
        g3 = gamma(3)
        err = [g3 * (((abs(m[i][0] * pw[0]) + abs(m[i][1] * pw[1])) + abs(m[i][2] * pw[2])) + m[i][3]) for i in range(3)]
        return hit[0], (pw, nw, err, uv)


class SpherePrim(TransformedPrim):
    """TransformShape{Sphere{radius}, transform} -- one wrapper, as `Shape "sphere"` under a CTM yields."""

    def __init__(self, radius, m16, m_inv16):
        TransformedPrim.__init__(self, BareSpherePrim(radius), m16, m_inv16)
        self.radius = F(radius)


class Searcher:
    def __init__(self, tree, tris):
        prims = tris if isinstance(tris, list) else [TrianglePrim(v) for v in np.asarray(tris, F)]
        self.tree = tree
        self.order = [int(i) for i in tree.order]
        self.prims = [prims[i] for i in self.order]   # BVH slot -> primitive
        self.bounds = [([F(x) for x in b[:3]], [F(x) for x in b[3:]]) for b in tree.bound]
        self.visits = 0
        self.tests = 0
        self.hit_b = None      # barycentrics (triangle) or (p, n, p_error) (sphere) of the last accepted hit of the current closest() call

    def closest(self, o, d, t_max=np.finfo(np.float32).max):
        """(upload index of the primitive hit, t) or (None, None)."""
        import sys
        sys.setrecursionlimit(max(sys.getrecursionlimit(), 4 * len(self.order) + 1000))
        with np.errstate(all="ignore"):
            ray = _Ray(o, d, t_max)
            slot = self._closest(self.tree.root, ray)
        return (None, None) if slot is None else (self.order[slot], ray.t_max)

    def _closest(self, node, ray):
        t = self.tree
        result = None
        if t.is_leaf[node]:
            for slot in range(t.a[node], t.b[node]):
                self.tests += 1
                hit = self.prims[slot].intersect(ray)
                if hit is not None:
                    ray.t_max = hit[0]
                    result = slot
                    self.hit_b = hit[1]
            return result
        self.visits += 1
        if _slab(self.bounds[node], ray):
            left, right = (t.a[node], t.b[node]) if ray.pos[t.axis[node]] == 1 else (t.b[node], t.a[node])
            r = self._closest(right, ray)
            if r is not None:
                result = r
            r = self._closest(left, ray)
            if r is not None:
                result = r
        return result

    def any_hit(self, o, d, t_max=np.finfo(np.float32).max):
        with np.errstate(all="ignore"):
            return self._any(self.tree.root, _Ray(o, d, t_max))

    def _any(self, node, ray):
        t = self.tree
        if t.is_leaf[node]:
            for slot in range(t.a[node], t.b[node]):
                self.tests += 1
                if self.prims[slot].intersect(ray) is not None:
                    return True
            return False
        self.visits += 1
        return _slab(self.bounds[node], ray) and (self._any(t.a[node], ray) or self._any(t.b[node], ray))
