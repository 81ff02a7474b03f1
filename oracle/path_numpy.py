"""ORACLE -- TEST INFRASTRUCTURE ONLY.

A second, independent restatement of the reference's render loop and path integrator for scenes
of matte triangles under diffuse area lights (the Cornell box without its glass ball and mirror),
in scalar np.float32 written directly from the Rust; it shares nothing with oracle/oracle.cpp.
The tests render a few hundred paths with it and hold the C++ oracle to the same image and the same
ray counts.  What it takes from outside, as inputs:
  * the flattened scene (triangles, Kd, emitted radiance) -- `cpbrt_scene_desc`;
  * the Halton value of (pixel, sample, dimension) and the raster -> camera-ray map, from callables
    (the sampler and the camera are pinned on their own: exact-arithmetic radical inverses, the
    reference's two unit tests, the closed-form camera);
  * the BVH and its search: oracle/sah_numpy.py, oracle/trace_numpy.py.

  render loop, film-point drift, NaN drop, mean   src/render.rs:20-44, src/film.rs:11-18
  Sampler::get_1d / get_2d / get_usize            src/sampler/mod.rs:10-36
  PathIntegrator::li                              src/integrator/path.rs:13-66
  DirectLightIntegrator::li_depth                 src/integrator/direct_light.rs:12-41
  uniform_sample_one_light                        src/integrator/mod.rs:13-97
  AreaLight::sample_li / le / pdf                 src/light/area.rs:17-51
  VisibilityTester, Ray::from_to                  src/light/mod.rs:100-115, src/geometry/ray.rs:30-36
  Shape::default_sample_by_point / by_point_pdf   src/geometry/shape/mod.rs:23-53
  Triangle::sample / area / normal / interpolate  src/geometry/shape/triangle.rs:76-126,255-261
  ShapePoint::point_offset_by_error               src/geometry/shape/mod.rs:117-124
  MatteMaterial, LambertianReflection, BxDF       src/material/matte.rs, bxdf/lambertian.rs, bxdf/mod.rs:17-44
  MirrorMaterial, GlassMaterial, specular lobes   src/material/{mirror,glass}.rs, bxdf/specular.rs, math/mod.rs:202-211 (refract)
  ConstantTexture, ImageTexture::evaluate         src/texture/{constant,image}.rs (nearest texel, v flipped)
  PointLight, DistantLight, DeltaLight adaptor    src/light/{point,distant}.rs, light/mod.rs:76-98
  Sphere::{sample_by_point,by_point_pdf}          src/geometry/shape/sphere.rs:66-109 (cone sampling, outside the sphere)
  BSDF::{new, sample_f, sample_no_delta_f, no_delta_f_pdf, *_to_*}   src/material/bxdf/mod.rs:74-222
  Distribution1D::{from, sample_remap}            src/math/distribution.rs:16-56
  coordinate_system, cosine / concentric / triangle sampling, power_heuristic   src/math/mod.rs
  RGBSpectrum::y                                  src/spectrum/rgb_spectrum.rs:82-84

Vector arithmetic follows nalgebra for 3-vectors: dot = (x*x' + y*y') + z*z', norm = sqrt(dot),
normalize divides each component by the norm."""
import math

import numpy as np

F = np.float32
ZERO, ONE, HALF = F(0), F(1), F(0.5)
PI = F(3.14159265358979323846)          # src/math/mod.rs:216
INV_PI = F(0.31830988618379067154)      # src/math/mod.rs:213
_EPS = F(np.finfo(np.float32).eps) * HALF


def _libm_f32(name):
    """f32::cos / f32::sin compile to calls of the platform's cosf / sinf; numpy's float32 kernels are
    SIMD approximations that differ from them in the last bit now and then."""
    import ctypes
    import ctypes.util
    fn = getattr(ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6"), name)
    fn.restype = ctypes.c_float
    fn.argtypes = [ctypes.c_float]
    return lambda x: F(fn(float(x)))


_cosf, _sinf = _libm_f32("cosf"), _libm_f32("sinf")


def gamma(n):
    ne = F(n) * _EPS
    return ne / (ONE - ne)


# ---- 3-vectors as lists of np.float32 --------------------------------------------------------
def add(a, b): return [a[0] + b[0], a[1] + b[1], a[2] + b[2]]
def sub(a, b): return [a[0] - b[0], a[1] - b[1], a[2] - b[2]]
def mul(a, b): return [a[0] * b[0], a[1] * b[1], a[2] * b[2]]
def scale(a, s): return [a[0] * s, a[1] * s, a[2] * s]
def unscale(a, s): return [a[0] / s, a[1] / s, a[2] / s]
def neg(a): return [-a[0], -a[1], -a[2]]
def vabs(a): return [abs(a[0]), abs(a[1]), abs(a[2])]
def dot(a, b): return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]
def norm2(a): return dot(a, a)
def normalize(a): return unscale(a, np.sqrt(norm2(a)))
def cross(a, b): return [a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]]
def has_nan(a): return bool(np.isnan(a[0]) or np.isnan(a[1]) or np.isnan(a[2]))
def lum(s): return F(0.212671) * s[0] + F(0.715160) * s[1] + F(0.072169) * s[2]


def _min(lhs, rhs): return rhs if rhs < lhs else lhs
def _max(lhs, rhs): return rhs if rhs > lhs else lhs


def power_heuristic(f, g):
    return (f * f) / (f * f + g * g)


def coordinate_system(z):
    if abs(z[0]) > abs(z[1]):
        x = unscale([-z[2], ZERO, z[0]], np.sqrt(z[0] * z[0] + z[2] * z[2]))
    else:
        x = unscale([ZERO, z[2], -z[1]], np.sqrt(z[1] * z[1] + z[2] * z[2]))
    return x, cross(z, x)


def concentric_sample_disk(u):
    ux, uy = F(2) * u[0] - ONE, F(2) * u[1] - ONE
    if ux == 0 or uy == 0:
        return ZERO, ZERO
    if abs(ux) > abs(uy):
        r, theta = ux, (PI / F(4)) * (uy / ux)
    else:
        r, theta = uy, (PI / F(2)) - (PI / F(4)) * (ux / uy)
    return r * _cosf(theta), r * _sinf(theta)


def cosine_sample_hemisphere(u):
    dx, dy = concentric_sample_disk(u)
    z = np.sqrt(ONE - (dx * dx + dy * dy))
    return [dx, dy, z], z * INV_PI


def sample_usize_remap(u, length):
    f = u * F(length)
    tr = np.trunc(f)
    return min(int(tr), length - 1), f - tr


def distribution_sample_remap(f, u):
    """Distribution1D::from(f).sample_remap(u) -> (index, pdf, remap)."""
    n = F(len(f))
    cdf = []
    for x in f:
        cdf.append((cdf[-1] if cdf else ZERO) + F(x) / n)
    f_sum = cdf[-1]
    pdf = []
    for i in range(len(f)):
        cdf[i] = cdf[i] / f_sum
        pdf.append(F(f[i]) / (n * f_sum))
    for i, c in enumerate(cdf):
        if u <= c:
            return i, pdf[i], (c - u) / pdf[i]
    return len(f) - 1, pdf[-1], ONE


# ---- sampler front (src/sampler/mod.rs) ------------------------------------------------------
class PixelSampler:
    """`value(dim)` is HaltonSampler::get_sample for the current (pixel, sample)."""

    def __init__(self, value):
        self.value = value
        self.dim = 0

    def get_sample(self):
        r = self.value(self.dim)
        self.dim += 1
        return r

    def get_1d(self):
        r = self.get_sample()
        return F(0.9999999999999) if r >= 1 else r

    def get_2d(self):
        x = self.get_1d()
        return [x, self.get_1d()]

    def get_usize(self, mx):
        f = self.get_sample() * F(mx)
        r = int(np.trunc(f))
        return mx - 1 if r == mx else r


# ---- geometry --------------------------------------------------------------------------------
class ShapePoint:
    def __init__(self, p, n, p_error, uv=None):
        self.p, self.n, self.p_error, self.uv = p, n, p_error, uv      # uv None: never looked at (constant textures)

    def offset_by_error(self, w):
        d = dot(vabs(self.n), self.p_error)
        offset = scale(self.n, d)
        if dot(w, self.n) < 0:
            offset = neg(offset)
        return add(self.p, offset)


class Triangle:
    def __init__(self, p, uv=None):
        self.p = [[F(x) for x in v] for v in p]
        self.uv = [[F(0), F(0)], [F(1), F(0)], [F(1), F(1)]] if uv is None else [[F(x) for x in v] for v in uv]

    def normal(self):
        p0, p1, p2 = self.p
        return normalize(normalize(cross(sub(p0, p2), sub(p1, p2))))     # .normalize(), then Normal3f::from normalizes again (normal.rs:8-12)

    def area(self):
        p0, p1, p2 = self.p
        return HALF * np.sqrt(norm2(cross(sub(p1, p0), sub(p2, p0))))

    def interpolate(self, b0, b1, b2):
        p0, p1, p2 = self.p
        return add(add(scale(p0, b0), scale(p1, b1)), scale(p2, b2))

    def abs_sum(self, b0, b1, b2):
        p0, p1, p2 = self.p
        return add(add(vabs(scale(p0, b0)), vabs(scale(p1, b1))), vabs(scale(p2, b2)))

    def hit_point(self, b):
        uv0, uv1, uv2 = self.uv
        uv = [(b[0] * uv0[c] + b[1] * uv1[c]) + b[2] * uv2[c] for c in range(2)]
        return ShapePoint(self.interpolate(*b), self.normal(), scale(self.abs_sum(*b), gamma(7)), uv)

    def sample(self, sampler):
        u = sampler.get_2d()
        su0 = np.sqrt(u[0])
        b0, b1 = ONE - su0, u[1] * su0
        b2 = ONE - b0 - b1
        return ShapePoint(self.interpolate(b0, b1, b2), self.normal(), scale(self.abs_sum(b0, b1, b2), gamma(6))), ONE / self.area()

    def sample_by_point(self, point, sampler):
        sp, pdf = self.sample(sampler)
        wi = sub(sp.p, point)
        if norm2(wi) == 0:
            return sp, ZERO
        pdf = pdf * norm2(wi) / -dot(normalize(wi), sp.n)
        if np.isnan(pdf) or np.isinf(pdf):
            return sp, ZERO
        return sp, pdf

    def by_point_pdf(self, point, sp):
        d = sub(point, sp.p)
        d2 = norm2(d)
        dist = np.sqrt(d2)
        pdf = d2 / (abs(dot(unscale(d, dist), sp.n)) * self.area())
        return ZERO if (np.isnan(pdf) or np.isinf(pdf)) else pdf


def fresnel_dielectric(eta_a, eta_b, cos_i):
    """FresnelDielectric{eta_i: eta_a, eta_t: eta_b}.evaluate (specular.rs:25-45)."""
    eta_i, eta_t = (eta_a, eta_b) if cos_i > 0 else (eta_b, eta_a)
    cos_i = abs(cos_i)
    sin_i = np.sqrt(ONE - cos_i * cos_i)
    sin_t = sin_i * eta_i / eta_t
    if sin_t >= 1:
        return ONE
    cos_t = np.sqrt(ONE - sin_t * sin_t)
    t0, t1 = eta_t * cos_i, eta_i * cos_t
    r_perl = (t0 - t1) / (t0 + t1)
    t0, t1 = eta_i * cos_i, eta_t * cos_t
    r_perp = (t0 - t1) / (t0 + t1)
    return (r_perl * r_perl + r_perp * r_perp) / F(2)


def refract(wo, n, eta):
    cos_o = dot(wo, n)
    sin2_o = ONE - cos_o * cos_o
    sin2_i = sin2_o * eta * eta
    if sin2_i > 1:
        return None
    cos_i = np.sqrt(_max(ONE - sin2_i, ZERO))
    return add(scale(neg(wo), eta), scale(n, eta * cos_o - cos_i))


class SpecularReflection:
    def __init__(self, r, eta=None):
        self.r, self.eta = r, eta            # eta None = FresnelNoOp, else FresnelDielectric(1, eta)

    def sample_f(self, wo):
        wi = [-wo[0], -wo[1], wo[2]]
        cos_i = wi[2]
        fr = ONE if self.eta is None else fresnel_dielectric(ONE, self.eta, cos_i)
        if fr == 0:
            return None
        return wi, unscale(scale(self.r, fr), abs(cos_i))


class SpecularTransmission:
    def __init__(self, t, eta_a, eta_b):
        self.t, self.eta_a, self.eta_b = t, eta_a, eta_b

    def sample_f(self, wo):
        eta_i, eta_t = (self.eta_a, self.eta_b) if wo[2] > 0 else (self.eta_b, self.eta_a)
        eta = eta_i / eta_t
        wi = refract(wo, normalize([ZERO, ZERO, np.copysign(ONE, wo[2])]), eta)
        if wi is None:
            return None
        ft = ONE - fresnel_dielectric(self.eta_a, self.eta_b, wi[2])
        if ft == 0:
            return None
        return wi, unscale(scale(self.t, ft), abs(wi[2]))


class BSDF:
    """BSDF::new(n, n) with `kd` (a LambertianReflection, or None) and a list of delta lobes."""

    def __init__(self, n, kd=None, deltas=()):
        self.sn = n
        self.snx, self.sny = coordinate_system(n)
        self.n = normalize([dot(n, self.snx), dot(n, self.sny), dot(n, self.sn)])
        self.kd = kd
        self.n_bxdfs = 0 if kd is None else 1
        self.deltas = list(deltas)

    def is_all_delta(self):
        return self.n_bxdfs == 0

    def local_to_world(self, w):
        x, y, z = self.snx, self.sny, self.sn
        return normalize([x[0] * w[0] + y[0] * w[1] + z[0] * w[2],
                          x[1] * w[0] + y[1] * w[1] + z[1] * w[2],
                          x[2] * w[0] + y[2] * w[1] + z[2] * w[2]])

    def world_to_local(self, w):
        return normalize([dot(w, self.snx), dot(w, self.sny), dot(w, self.sn)])

    def f(self):
        return scale(self.kd, INV_PI)

    def sample_no_delta_f(self, wo, u):
        if self.n_bxdfs == 0:
            return [ZERO, ZERO, ZERO], None, ZERO
        index, remap = sample_usize_remap(u[0], 1)
        wo_local = self.world_to_local(wo)
        wi, pdf = cosine_sample_hemisphere([remap, u[1]])
        if wo_local[2] < 0:
            wi[2] = wi[2] * F(-1)
        return self.local_to_world(wi), self.f(), pdf / F(1)

    def sample_delta_f(self, wo, u):
        wo_local = self.world_to_local(wo)
        if not self.deltas:
            return [ZERO, ZERO, ZERO], None, ZERO
        wi_f = [r for r in (lobe.sample_f(wo_local) for lobe in self.deltas) if r is not None]
        assert wi_f, "the reference panics here (Distribution1D::from of an empty table)"
        i, pdf, _ = distribution_sample_remap([lum(f) for _, f in wi_f], u)
        wi_local, f = wi_f[i]
        return self.local_to_world(wi_local), f, pdf

    def no_delta_f_pdf(self, wo, wi):
        if self.n_bxdfs == 0:
            return None, ZERO
        wo, wi = self.world_to_local(wo), self.world_to_local(wi)
        if not (dot(wo, self.n) * dot(wi, self.n) > 0):
            return None, ZERO / F(0)        # no transmissive lobe: pdf 0 / len 0 = NaN, f None
        return add([ZERO, ZERO, ZERO], self.f()), (ZERO + wi[2] * INV_PI) / F(1)

    def sample_f(self, wo, sampler):
        i, pdf, remap = distribution_sample_remap([self.n_bxdfs, len(self.deltas)], sampler.get_1d())
        if i == 0:
            wi, f, internal = self.sample_no_delta_f(wo, [remap, sampler.get_1d()])
            return wi, f, pdf * internal, False
        wi, f, internal = self.sample_delta_f(wo, remap)
        return wi, f, pdf * internal, True


MAT_MATTE, MAT_MIRROR, MAT_GLASS = 0, 1, 2


def make_bsdf(material, sp):
    kind, a, b, eta = material
    n = sp.n
    if kind == MAT_MATTE:
        assert b(sp.uv)[0] == 0, "sigma = 0 (LambertianReflection) only"
        return BSDF(n, kd=a(sp.uv))
    if kind == MAT_MIRROR:
        return BSDF(n, deltas=[SpecularReflection(a(sp.uv))])
    index = eta(sp.uv)[0]
    return BSDF(n, deltas=[SpecularReflection(a(sp.uv), index), SpecularTransmission(b(sp.uv), ONE, index)])


def f32_to_usize(f):
    """`f as usize`: truncates, saturates, NaN -> 0."""
    if np.isnan(f) or f <= 0:
        return 0
    return min(int(f), (1 << 64) - 1)


def texture(desc, index):
    """Texture::evaluate as a callable of uv: constant (texture/constant.rs) or nearest image texel
    (ImageTexture::evaluate, texture/image.rs:92-101; texels as the host decoded them)."""
    t = desc.textures[index]
    if t.kind == 0:
        value = [F(x) for x in t.value]
        return lambda uv: value
    width, height, channels, offset = int(t.width), int(t.height), int(t.channels), int(t.offset)
    texels = np.ctypeslib.as_array(desc.texels, shape=(int(desc.n_texels),))

    def evaluate(uv):
        u, v = uv[0], ONE - uv[1]
        i = min(f32_to_usize(u * F(width)), width - 1)
        j = min(f32_to_usize(v * F(height)), height - 1)
        at = offset + (j * width + i) * channels
        return [F(x) for x in texels[at:at + channels]]
    return evaluate


# ---- the scene -------------------------------------------------------------------------------
class SphereLightShape:
    """TransformShape{Sphere} as a light's shape: sample_by_point / by_point_pdf go through the inverse transform
    (transform.rs:60-70), Sphere::sample_by_point is the cone sampling of sphere.rs:66-99 (outside the sphere)."""

    def __init__(self, prim, trace):
        self.prim, self.trace = prim, trace

    def _to_world(self, p, n):
        t, q = self.trace, self.prim
        pw = t.apply_point(q.m, p)
        nw = [(q.m3[i][0] * n[0] + q.m3[i][1] * n[1]) + q.m3[i][2] * n[2] for i in range(3)]
        g3 = gamma(3)
        err = [g3 * (((abs(q.m[i][0] * pw[0]) + abs(q.m[i][1] * pw[1])) + abs(q.m[i][2] * pw[2])) + q.m[i][3]) for i in range(3)]
        return ShapePoint(pw, nw, err)

    def sample_by_point(self, point, sampler):
        r = self.prim.radius
        point = self.trace.apply_point(self.prim.m_inv, point)
        d2 = norm2(point)
        r2 = r * r
        assert d2 > r2, "shading point inside the emitter: default_sample_by_point, not restated"
        dist = np.sqrt(d2)
        z = unscale(point, dist)
        x, y = coordinate_system(z)
        u = sampler.get_2d()
        sin_max2 = r2 / d2
        cos_max = np.sqrt(_max(ONE - sin_max2, ZERO))
        cos_t = (ONE - u[0]) + u[0] * cos_max
        sin_t = np.sqrt(_max(ONE - cos_t * cos_t, ZERO))
        phi = u[1] * F(2) * PI
        ds = dist * cos_t - np.sqrt(_max(r2 - d2 * sin_t * sin_t, ZERO))
        cos_a = (d2 + r2 - ds * ds) / (F(2) * dist * r)
        sin_a = np.sqrt(_max(ONE - cos_a * cos_a, ZERO))
        d = add(add(scale(z, cos_a), scale(x, sin_a * _cosf(phi))), scale(y, sin_a * _sinf(phi)))
        n = normalize(d)
        p = scale(d, r)
        return self._to_world(p, n), ONE / (F(2) * PI * (ONE - cos_max))

    def by_point_pdf(self, point, sp):
        r = self.prim.radius
        point = self.trace.apply_point(self.prim.m_inv, point)
        d2 = norm2(point)
        assert not d2 < r * r
        cos_max = np.sqrt(ONE - r * r / d2)
        return ONE / (F(2) * PI * (ONE - cos_max))


class AreaLight:
    is_delta = False

    def __init__(self, shape, le):
        self.shape, self.le = shape, le

    def sample_li(self, scene, sp, sampler):
        """(wi, li, pdf, shadow ray) or None."""
        lp, pdf = self.shape.sample_by_point(sp.p, sampler)
        if norm2(sub(lp.p, sp.p)) == 0 or pdf == 0:
            return None
        frm = sp.offset_by_error(sub(lp.p, sp.p))
        to = lp.offset_by_error(sub(frm, lp.p))
        return normalize(sub(lp.p, sp.p)), self.le, pdf, (frm, sub(to, frm), ONE - F(0.00001))


class PointLight:
    """DeltaLight adaptor (light/mod.rs:80-98: two numbers drawn and dropped, pdf 1) over PointLight (point.rs)."""
    is_delta = True

    def __init__(self, p, i):
        self.p, self.i = p, i

    def sample_li(self, scene, sp, sampler):
        sampler.get_2d()
        wi = sub(self.p, sp.p)
        li = unscale(self.i, norm2(wi))
        to = ShapePoint(self.p, normalize(neg(wi)), [ZERO, ZERO, ZERO])
        frm = sp.offset_by_error(sub(to.p, sp.p))
        end = to.offset_by_error(sub(frm, to.p))
        return normalize(wi), li, ONE, (frm, sub(end, frm), ONE - F(0.00001))


class DistantLight:
    is_delta = True

    def __init__(self, w, i):
        self.w, self.i = w, i

    def sample_li(self, scene, sp, sampler):
        sampler.get_2d()
        d = neg(self.w)
        return d, self.i, ONE, (sp.offset_by_error(d), d, F(np.finfo(np.float32).max))


class Scene:
    def __init__(self, desc, abi, sah, trace):
        n = int(desc.n_primitives)
        self.tris, self.material, self.light_of = [], [], []
        prims = []
        for i in range(n):
            pr = desc.primitives[i]
            if pr.shape.kind == abi.SHAPE_TRIANGLE:
                assert pr.shape.xform < 0
                verts = [[desc.triangles[pr.shape.data].p[k][c] for c in range(3)] for k in range(3)]
                prims.append(trace.TrianglePrim(verts))
                self.tris.append(Triangle(verts, [[desc.triangles[pr.shape.data].uv[k][c] for c in range(2)] for k in range(3)]))
            else:
                assert pr.shape.kind == abi.SHAPE_SPHERE and pr.shape.xform >= 0
                t = desc.transforms[pr.shape.xform]
                assert t.outer < 0
                prims.append(trace.SpherePrim(desc.sphere_radius[pr.shape.data], list(t.m), list(t.m_inv)))
                self.tris.append(None)
            self.light_of.append(int(pr.light))
            self.material.append(self._material(desc, pr.material) if pr.material >= 0 else None)
        self.lights = []
        for k in range(int(desc.n_lights)):
            lt = desc.lights[k]
            if lt.type == 1:
                self.lights.append(PointLight([F(x) for x in lt.v], [F(x) for x in lt.L]))
            elif lt.type == 2:
                self.lights.append(DistantLight([F(x) for x in lt.v], [F(x) for x in lt.L]))
            else:
                assert lt.type == 0, "infinite area lights are not restated"
                prim = [i for i in range(n) if self.light_of[i] == k]
                assert len(prim) == 1
                shape = self.tris[prim[0]] if self.tris[prim[0]] is not None else SphereLightShape(prims[prim[0]], trace)
                self.lights.append(AreaLight(shape, [F(x) for x in lt.L]))
        bounds = [q.bounds() for q in prims]
        lo = np.array([b[0] for b in bounds], np.float32)
        hi = np.array([b[1] for b in bounds], np.float32)
        self.search = trace.Searcher(sah.build(lo, hi), prims)
        self.max_depth = int(desc.max_depth)
        self.rays_closest = 0
        self.rays_any = 0

    @staticmethod
    def _material(desc, index):
        m = desc.materials[index]
        if m.type == MAT_MATTE:
            return MAT_MATTE, texture(desc, m.tex[0]), texture(desc, m.tex[1]), None
        if m.type == MAT_MIRROR:
            return MAT_MIRROR, texture(desc, m.tex[0]), None, None
        assert m.type == MAT_GLASS
        return MAT_GLASS, texture(desc, m.tex[0]), texture(desc, m.tex[1]), texture(desc, m.tex[2])

    def intersect(self, o, d, t_max=np.finfo(np.float32).max):
        self.rays_closest += 1
        prim, _ = self.search.closest(o, d, t_max)
        if prim is None:
            return None, None
        if self.tris[prim] is None:
            return prim, ShapePoint(*self.search.hit_b)
        return prim, self.tris[prim].hit_point(self.search.hit_b)

    def occluded(self, o, d, t_max):
        self.rays_any += 1
        return self.search.any_hit(o, d, t_max)


def uniform_sample_one_light(scene, sp, bsdf, wo, sampler):
    l = [ZERO, ZERO, ZERO]
    n = sp.n
    nl = len(scene.lights)
    if nl:
        chosen = sampler.get_usize(nl)
        light = scene.lights[chosen]
        # sample the light
        sample = light.sample_li(scene, sp, sampler)
        if sample is not None:
            wi, li, li_pdf, shadow = sample
            if li_pdf != 0:
                f, f_pdf = bsdf.no_delta_f_pdf(wo, wi)
                if f is not None and f_pdf != 0:
                    if not scene.occluded(*shadow):
                        ld = scale(mul(li, f), abs(dot(n, wi)))
                        if not light.is_delta:
                            ld = scale(ld, power_heuristic(li_pdf, f_pdf))
                        l = add(l, unscale(ld, li_pdf))
        if not light.is_delta:
            # sample the BSDF
            wi, f, f_pdf = bsdf.sample_no_delta_f(wo, sampler.get_2d())
            if f is not None and f_pdf != 0:
                ld = [ZERO, ZERO, ZERO]
                prim, hit = scene.intersect(sp.offset_by_error(wi), wi)
                if prim is not None and scene.light_of[prim] == chosen:
                    li_pdf = light.shape.by_point_pdf(sp.p, hit)
                    if li_pdf != 0:
                        ld = unscale(scale(scale(mul(light.le, f), abs(dot(n, wi))), power_heuristic(f_pdf, li_pdf)), f_pdf)
                l = add(l, ld)
    return scale(l, F(nl))


def path_li(scene, o, d, sampler):
    l = [ZERO, ZERO, ZERO]
    beta = [ONE, ONE, ONE]
    bounce = 0
    specular_bounce = False
    while True:
        prim, sp = scene.intersect(o, d)
        if bounce == 0 or specular_bounce:
            if prim is not None and scene.light_of[prim] >= 0:
                l = add(l, mul(beta, scene.lights[scene.light_of[prim]].le))
        if prim is None or bounce >= scene.max_depth or scene.material[prim] is None:
            break
        bsdf = make_bsdf(scene.material[prim], sp)
        wo = neg(d)
        if not bsdf.is_all_delta():
            l = add(l, mul(beta, uniform_sample_one_light(scene, sp, bsdf, wo, sampler)))
        wi, f, f_pdf, is_delta = bsdf.sample_f(wo, sampler)
        if f is None or f_pdf == 0:
            break
        beta = mul(beta, unscale(scale(f, abs(dot(wi, sp.n))), f_pdf))
        o, d = sp.offset_by_error(wi), wi
        specular_bounce = is_delta
        if bounce > 3:
            q = _max(F(0.05), ONE - lum(beta))
            if sampler.get_1d() < q:
                break
            beta = unscale(beta, ONE - q)
        bounce += 1
    return l


def direct_li(scene, o, d, sampler, depth=0):
    """DirectLightIntegrator::li_depth (direct_light.rs:12-41): one-light estimate at the hit, then the recursion
    over every delta lobe.  uniform_sample_one_light runs (and draws its five numbers) on all-delta surfaces too."""
    l = [ZERO, ZERO, ZERO]
    prim, sp = scene.intersect(o, d)
    if prim is None:
        return l
    if scene.material[prim] is None:
        return add(l, scene.lights[scene.light_of[prim]].le)
    bsdf = make_bsdf(scene.material[prim], sp)
    wo = neg(normalize(d))
    l = add(l, uniform_sample_one_light(scene, sp, bsdf, wo, sampler))
    if depth + 1 < scene.max_depth:
        wo_local = bsdf.world_to_local(wo)
        for lobe in bsdf.deltas:
            r = lobe.sample_f(wo_local)
            if r is None:
                continue
            wi, s_ = bsdf.local_to_world(r[0]), r[1]
            inner = direct_li(scene, sp.offset_by_error(wi), wi, sampler, depth + 1)
            l = add(l, scale(mul(inner, s_), abs(dot(sp.n, wi))))
    return l


def render(scene, width, height, spp, halton_value, camera_ray_at, rect=None, integrator="path"):
    """film[h, w, 3] over rect = (x0, y0, x1, y1); halton_value(x, y, sample) -> callable dim -> f32;
    camera_ray_at(fx, fy) -> (o[3], d[3])."""
    x0, y0, x1, y1 = rect or (0, 0, width, height)
    film = np.zeros((height, width, 3), np.float32)
    with np.errstate(all="ignore"):
        for y in range(y0, y1):
            for x in range(x0, x1):
                fx, fy = F(x), F(y)
                kept = []
                for s in range(spp):
                    sampler = PixelSampler(halton_value(x, y, s))
                    u = sampler.get_2d()
                    off = [u[0] - HALF, u[1] - HALF]
                    fx, fy = fx + off[0], fy + off[1]
                    o, d = camera_ray_at(fx, fy)
                    li = (path_li if integrator == "path" else direct_li)(scene, [F(v) for v in o], [F(v) for v in d], sampler)
                    if not has_nan(li):
                        kept.append(li)
                total = [ZERO, ZERO, ZERO]
                for li in kept:
                    total = add(total, li)
                film[y, x] = unscale(total, F(len(kept)))
    return film
