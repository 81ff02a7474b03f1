"""ORACLE -- TEST INFRASTRUCTURE ONLY.

A second, independent restatement of the reference's render loop, both integrators, every material,
every light and the shapes' sampling, in scalar np.float32 written directly from the Rust; it shares
nothing with oracle/oracle.cpp.  The tests render a few hundred to a few thousand paths with it and
hold the C++ oracle -- and, on the GPU box, the device -- to the same image and the same ray counts.
Inputs: the flattened scene (`cpbrt_scene_desc`: vertices, matrices, texels as the host decoded them),
the Halton value of (pixel, sample, dimension) and the raster -> camera-ray map as callables
(oracle/front_numpy.py restates both), the BVH and its search (oracle/sah_numpy.py, trace_numpy.py).

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
  OrenNayar, TrowbridgeReitz, Microfacet{Reflection,Transmission}   src/material/bxdf/{oren_nayar,microfacet}.rs
  plastic / uber / translucent / mix, BSDF::mix   src/material/{plastic,uber,translucent,mix}.rs, bxdf/mod.rs:46-66,218-269
  MirrorMaterial, GlassMaterial, specular lobes   src/material/{mirror,glass}.rs, bxdf/specular.rs, math/mod.rs:202-211 (refract)
  ConstantTexture, ImageTexture::evaluate         src/texture/{constant,image}.rs (nearest texel, v flipped)
  PointLight, DistantLight, DeltaLight adaptor    src/light/{point,distant}.rs, light/mod.rs:76-98
  InfiniteAreaLight, Distribution2D               src/light/infinite_area.rs, math/distribution.rs:69-139, math/mod.rs:136-152
  Sphere::{sample_by_point,by_point_pdf}          src/geometry/shape/sphere.rs:66-109 (cone sampling, outside the sphere)
  BSDF::{new, sample_f, sample_no_delta_f, no_delta_f_pdf, *_to_*}   src/material/bxdf/mod.rs:74-222
  Distribution1D::{from, sample_remap}            src/math/distribution.rs:16-56
  coordinate_system, cosine / concentric / triangle sampling, power_heuristic   src/math/mod.rs
  RGBSpectrum::y                                  src/spectrum/rgb_spectrum.rs:82-84

Vector arithmetic follows nalgebra for 3-vectors: dot = (x*x' + y*y') + z*z', norm = sqrt(dot),
normalize divides each component by the norm."""

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


_cosf, _sinf, _logf, _acosf = _libm_f32("cosf"), _libm_f32("sinf"), _libm_f32("logf"), _libm_f32("acosf")


def _atan2f(y, x):
    import ctypes
    import ctypes.util
    fn = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6").atan2f
    fn.restype = ctypes.c_float
    fn.argtypes = [ctypes.c_float, ctypes.c_float]
    return F(fn(float(y), float(x)))


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


# shading-frame trigonometry (math/mod.rs:152-201)
def cos2_theta(w): return w[2] * w[2]
def sin2_theta(w): return _max(ONE - cos2_theta(w), ZERO)
def sin_theta(w): return np.sqrt(sin2_theta(w))
def tan_theta(w): return sin_theta(w) / w[2]
def tan2_theta(w): return sin2_theta(w) / cos2_theta(w)


def clamp(t, lo, hi):
    return lo if t < lo else (hi if t > hi else t)


def cos_phi(w):
    st = sin_theta(w)
    return ONE if st == 0 else clamp(w[0] / st, F(-1), ONE)


def sin_phi(w):
    st = sin_theta(w)
    return ZERO if st == 0 else clamp(w[1] / st, F(-1), ONE)


def cos_delta_phi(wa, wb):
    return clamp((wa[0] * wb[0] + wa[1] * wb[1]) / np.sqrt((wa[0] * wa[0] + wa[1] * wa[1]) * (wb[0] * wb[0] + wb[1] * wb[1])), F(-1), ONE)


def is_black(s):
    return bool(s[0] == 0 and s[1] == 0 and s[2] == 0)


class Lobe:
    """BxDF defaults (bxdf/mod.rs:17-44): cosine sampling flipped by the lobe's side, pdf = wi.z / pi (signed)."""
    transmit = False

    def sample_f(self, wo, u):
        wi, pdf = cosine_sample_hemisphere(u)
        if (not self.transmit and wo[2] < 0) or (self.transmit and wo[2] > 0):
            wi[2] = wi[2] * F(-1)
        return wi, self.f(wo, wi), pdf

    def pdf(self, wo, wi):
        return wi[2] * INV_PI


class Lambertian(Lobe):
    def __init__(self, r, transmit=False):
        self.r, self.transmit = r, transmit

    def f(self, wo, wi):
        return scale(self.r, INV_PI)


class OrenNayar(Lobe):
    def __init__(self, r, sigma):
        sigma = sigma * (PI / F(180))                  # f32::to_radians
        sigma2 = sigma * sigma
        self.r = r
        self.a = ONE - (sigma2 / (F(2) * (sigma + F(0.33))))
        self.b = F(0.45) * sigma2 / (sigma2 + F(0.09))

    def f(self, wo, wi):
        if wi[2] < wo[2]:
            sin_alpha, tan_beta = sin_theta(wi), tan_theta(wo)
        else:
            sin_alpha, tan_beta = sin_theta(wo), tan_theta(wi)
        return scale(self.r, (self.a + self.b * _max(ZERO, cos_delta_phi(wi, wo)) * sin_alpha * tan_beta) / PI)


class TrowbridgeReitz:
    def __init__(self, ax, ay):
        self.ax, self.ay = ax, ay

    @staticmethod
    def roughness_to_alpha(rough):
        rough = _max(rough, F(1e-3))
        x = _logf(rough)
        x2 = x * x
        return F(1.62142) + F(0.819955) * x + F(0.1734) * x2 + F(0.0171201) * x * x2 + F(0.000640711) * x2 * x2

    def d(self, wh):
        t2 = tan2_theta(wh)
        if np.isnan(t2) or np.isinf(t2):
            return ZERO
        c2 = cos2_theta(wh)
        c4 = c2 * c2
        cp, sp_ = cos_phi(wh), sin_phi(wh)
        e = ((cp * cp) / (self.ax * self.ax) + (sp_ * sp_) / (self.ay * self.ay)) * t2
        return ONE / (PI * self.ax * self.ay * c4 * (ONE + e) * (ONE + e))

    def lam(self, w):
        att = abs(tan_theta(w))
        if np.isnan(att) or np.isinf(att):
            return ZERO
        cp, sp_ = cos_phi(w), sin_phi(w)
        alpha = np.sqrt((cp * cp) * self.ax * self.ax + (sp_ * sp_) * self.ay * self.ay)
        at = alpha * att
        return (F(-1) + np.sqrt(ONE + at * at)) / F(2)

    def g(self, wo, wi):
        return ONE / (ONE + self.lam(wo) + self.lam(wi))

    def g1(self, w):
        return ONE / (ONE + self.lam(w))

    def sample_wh(self, wo, u):
        flip = wo[2] < 0
        wi = neg(wo) if flip else list(wo)
        ws = [self.ax * wi[0], self.ay * wi[1], wi[2]]
        cos_i = ws[2]
        if cos_i > F(0.9999):
            r = np.sqrt(u[0] / (ONE - u[0]))
            phi = F(6.28318530718) * u[1]
            sx, sy = r * _cosf(phi), r * _sinf(phi)
        else:
            sin_t = np.sqrt(_max(ZERO, ONE - cos_i * cos_i))
            tan_t = sin_t / cos_i
            a = ONE / tan_t
            g1 = F(2) / (ONE + np.sqrt(ONE + ONE / (a * a)))
            A = F(2) * u[0] / g1 - ONE
            tmp = ONE / (A * A - ONE)
            if tmp > F(1e10):
                tmp = F(1e10)
            B = tan_t
            D = np.sqrt(_max(B * B * tmp * tmp - (A * A - B * B) * tmp, ZERO))
            sx1, sx2 = B * tmp - D, B * tmp + D
            sx = sx1 if (A < 0 or sx2 > ONE / tan_t) else sx2
            if u[1] > HALF:
                S, u2 = ONE, F(2) * (u[1] - HALF)
            else:
                S, u2 = F(-1), F(2) * (HALF - u[1])
            z = (u2 * (u2 * (u2 * F(0.27385) - F(0.73369)) + F(0.46341))) / (u2 * (u2 * (u2 * F(0.093073) + F(0.309420)) - ONE) + F(0.597999))
            sy = S * z * np.sqrt(ONE + sx * sx)
        tmp = cos_phi(ws) * sx - sin_phi(ws) * sy
        sy = sin_phi(ws) * sx + cos_phi(ws) * sy
        sx = tmp
        sx = sx * self.ax
        sy = sy * self.ay
        wh = normalize([-sx, -sy, ONE])
        if flip:
            wh = neg(wh)
        return wh, self.d(wh) * self.g1(wo) * abs(dot(wo, wh)) / abs(wo[2])


_NONE3 = [ZERO, ZERO, ZERO]


class MicrofacetReflection(Lobe):
    """pdf() is the trait default wi.z / pi: the reference does not override it (bxdf/microfacet.rs)."""

    def __init__(self, r, dist, eta):
        self.r, self.dist, self.eta = r, dist, eta            # FresnelDielectric(1, eta)

    def f(self, wo, wi):
        cos_o, cos_i = abs(wo[2]), abs(wi[2])
        wi, wo = normalize(wi), normalize(wo)
        wh = add(wi, wo)
        if (wh[0] == 0 and wh[1] == 0 and wh[2] == 0) or cos_o == 0 or cos_i == 0:
            return None
        wh = normalize(wh)
        fr = fresnel_dielectric(ONE, self.eta, dot(wi, neg(wh) if wh[2] < 0 else wh))
        return unscale(scale(scale(scale(self.r, self.dist.d(wh)), self.dist.g(wo, wi)), fr), F(4) * cos_o * cos_i)

    def sample_f(self, wo, u):
        if wo[2] == 0:
            return list(_NONE3), None, ZERO
        wh, pdf = self.dist.sample_wh(wo, u)
        if dot(wo, wh) < 0:
            return list(_NONE3), None, ZERO
        wi = add(neg(wo), scale(wh, F(2) * dot(wo, wh)))
        if wi[2] * wo[2] < 0:
            return list(_NONE3), None, ZERO
        return wi, self.f(wo, wi), pdf / (F(4) * dot(wo, wh))


class MicrofacetTransmission(Lobe):
    transmit = True

    def __init__(self, t, dist, eta_a, eta_b):
        self.t, self.dist, self.eta_a, self.eta_b = t, dist, eta_a, eta_b

    def f(self, wo, wi):
        if wo[2] * wi[2] > 0:
            return None
        cos_o, cos_i = wo[2], wi[2]
        if cos_i == 0 or cos_o == 0:
            return None
        eta = self.eta_b / self.eta_a if wo[2] > 0 else self.eta_a / self.eta_b
        wh = normalize(add(wo, scale(wi, eta)))
        if wh[2] < 0:
            wh = neg(wh)
        sd = dot(wo, wh) + eta * dot(wi, wh)
        fr = fresnel_dielectric(self.eta_a, self.eta_b, dot(wo, wh))
        factor = ONE / eta
        k = abs(self.dist.d(wh) * self.dist.g(wo, wi) * eta * eta * abs(dot(wi, wh)) * abs(dot(wo, wh)) * factor * factor
                / (cos_i * cos_o * sd * sd))
        return scale(mul([ONE - fr, ONE - fr, ONE - fr], self.t), k)

    def sample_f(self, wo, u):
        if wo[2] == 0:
            return list(_NONE3), None, ZERO
        wh, pdf = self.dist.sample_wh(wo, u)
        if dot(wo, wh) < 0:
            return list(_NONE3), None, ZERO
        if wo[2] > 0:
            eta_i, eta_o = self.eta_a / self.eta_b, self.eta_b / self.eta_a
        else:
            eta_i, eta_o = self.eta_b / self.eta_a, self.eta_a / self.eta_b
        wi = refract(wo, normalize(wh), eta_i)
        if wi is None:
            return list(_NONE3), None, ZERO
        sd = dot(wo, wh) + eta_o * dot(wi, wh)
        dwh_dwi = abs(eta_o * eta_o * dot(wi, wh)) / (sd * sd)
        return wi, self.f(wo, wi), pdf * dwh_dwi


class Scaled(Lobe):
    """ScaledBxDF (bxdf/mod.rs:46-66)."""

    def __init__(self, inner, s):
        self.inner, self.s, self.transmit = inner, s, inner.transmit

    def f(self, wo, wi):
        f = self.inner.f(wo, wi)
        return None if f is None else mul(f, self.s)

    def sample_f(self, wo, u):
        wi, f, pdf = self.inner.sample_f(wo, u)
        return wi, (None if f is None else mul(f, self.s)), pdf

    def pdf(self, wo, wi):
        return self.inner.pdf(wo, wi)


class ScaledDelta:
    def __init__(self, inner, s):
        self.inner, self.s = inner, s

    def sample_f(self, wo):
        r = self.inner.sample_f(wo)
        return None if r is None else (r[0], mul(r[1], self.s))


class BSDF:
    """BSDF::new(n, n) (bxdf/mod.rs:74-96) with its lobes in insertion order."""

    def __init__(self, n, lobes=(), deltas=()):
        self.sn = n
        self.snx, self.sny = coordinate_system(n)
        self.n = normalize([dot(n, self.snx), dot(n, self.sny), dot(n, self.sn)])
        self.lobes = list(lobes)
        self.deltas = list(deltas)

    def is_all_delta(self):
        return not self.lobes

    def local_to_world(self, w):
        x, y, z = self.snx, self.sny, self.sn
        return normalize([x[0] * w[0] + y[0] * w[1] + z[0] * w[2],
                          x[1] * w[0] + y[1] * w[1] + z[1] * w[2],
                          x[2] * w[0] + y[2] * w[1] + z[2] * w[2]])

    def world_to_local(self, w):
        return normalize([dot(w, self.snx), dot(w, self.sny), dot(w, self.sn)])

    def sample_no_delta_f(self, wo, u):
        if not self.lobes:
            return list(_NONE3), None, ZERO
        index, remap = sample_usize_remap(u[0], len(self.lobes))
        wi_local, f, pdf = self.lobes[index].sample_f(self.world_to_local(wo), [remap, u[1]])
        return self.local_to_world(wi_local), f, pdf / F(len(self.lobes))

    def sample_delta_f(self, wo, u):
        wo_local = self.world_to_local(wo)
        if not self.deltas:
            return list(_NONE3), None, ZERO
        wi_f = [r for r in (lobe.sample_f(wo_local) for lobe in self.deltas) if r is not None]
        assert wi_f, "the reference panics here (Distribution1D::from of an empty table)"
        i, pdf, _ = distribution_sample_remap([lum(f) for _, f in wi_f], u)
        wi_local, f = wi_f[i]
        return self.local_to_world(wi_local), f, pdf

    def no_delta_f_pdf(self, wo, wi):
        if not self.lobes:
            return None, ZERO
        wo, wi = self.world_to_local(wo), self.world_to_local(wi)
        reflect = dot(wo, self.n) * dot(wi, self.n) > 0
        side = [lobe for lobe in self.lobes if lobe.transmit != reflect]
        f, pdf = None, ZERO
        for lobe in side:
            this_f = lobe.f(wo, wi)
            if this_f is not None:
                f = add(f if f is not None else list(_NONE3), this_f)
                pdf = pdf + lobe.pdf(wo, wi)
        return f, pdf / F(len(side))

    def sample_f(self, wo, sampler):
        i, pdf, remap = distribution_sample_remap([len(self.lobes), len(self.deltas)], sampler.get_1d())
        if i == 0:
            wi, f, internal = self.sample_no_delta_f(wo, [remap, sampler.get_1d()])
            return wi, f, pdf * internal, False
        wi, f, internal = self.sample_delta_f(wo, remap)
        return wi, f, pdf * internal, True

    def mix(self, rhs, s1):
        s2 = [ONE - s1[0], ONE - s1[1], ONE - s1[2]]
        if is_black(s1):
            return rhs
        if is_black(s2):
            return self
        self.lobes = [Scaled(x, s1) for x in self.lobes] + [Scaled(x, s2) for x in rhs.lobes]
        self.deltas = [ScaledDelta(x, s1) for x in self.deltas] + [ScaledDelta(x, s2) for x in rhs.deltas]
        return self


MAT_MATTE, MAT_MIRROR, MAT_GLASS, MAT_PLASTIC, MAT_UBER, MAT_TRANSLUCENT, MAT_MIX = range(7)


def make_bsdf(material, sp):
    """Material::compute_scattering_functions (material/{matte,mirror,glass,plastic,uber,translucent,mix}.rs)."""
    kind, tex, sub = material
    n, uv = sp.n, sp.uv
    if kind == MAT_MIX:
        return make_bsdf(sub[0], sp).mix(make_bsdf(sub[1], sp), tex[0](uv))
    bsdf = BSDF(n)
    if kind == MAT_MATTE:
        kd, sigma = tex[0](uv), clamp(tex[1](uv)[0], ZERO, F(90))
        bsdf.lobes.append(Lambertian(kd) if sigma == 0 else OrenNayar(kd, sigma))
    elif kind == MAT_MIRROR:
        bsdf.deltas.append(SpecularReflection(tex[0](uv)))
    elif kind == MAT_GLASS:
        index = tex[2](uv)[0]
        bsdf.deltas += [SpecularReflection(tex[0](uv), index), SpecularTransmission(tex[1](uv), ONE, index)]
    elif kind == MAT_PLASTIC:
        kd, ks = tex[0](uv), tex[1](uv)
        if not is_black(kd):
            bsdf.lobes.append(Lambertian(kd))
        if not is_black(ks):
            rough = TrowbridgeReitz.roughness_to_alpha(tex[2](uv)[0])
            bsdf.lobes.append(MicrofacetReflection(ks, TrowbridgeReitz(rough, rough), F(1.5)))
    elif kind == MAT_UBER:
        eta, opacity = tex[7](uv)[0], tex[8](uv)
        t = [ONE - opacity[0], ONE - opacity[1], ONE - opacity[2]]
        if not is_black(t):
            bsdf.deltas.append(SpecularTransmission(t, ONE, ONE))
        kd = mul(opacity, tex[0](uv))
        if not is_black(kd):
            bsdf.lobes.append(Lambertian(kd))
        ks = mul(opacity, tex[1](uv))
        if not is_black(ks):
            ru = tex[5](uv)[0] if tex[5] is not None else tex[4](uv)[0]
            rv = tex[6](uv)[0] if tex[6] is not None else ru
            ru, rv = TrowbridgeReitz.roughness_to_alpha(ru), TrowbridgeReitz.roughness_to_alpha(rv)
            bsdf.lobes.append(MicrofacetReflection(ks, TrowbridgeReitz(ru, rv), eta))
        kr = mul(opacity, tex[2](uv))
        if not is_black(kr):
            bsdf.deltas.append(SpecularReflection(kr, eta))
        kt = mul(opacity, tex[3](uv))
        if not is_black(kt):
            bsdf.deltas.append(SpecularTransmission(kt, ONE, eta))
    else:
        assert kind == MAT_TRANSLUCENT
        eta = F(1.5)
        r, t = tex[3](uv), tex[4](uv)
        if is_black(r) and is_black(t):
            return bsdf
        kd = tex[0](uv)
        if not is_black(kd):
            if not is_black(r):
                bsdf.lobes.append(Lambertian(mul(r, kd)))
            if not is_black(t):
                bsdf.lobes.append(Lambertian(mul(t, kd), transmit=True))
        ks = tex[1](uv)
        if not is_black(ks):
            rough = TrowbridgeReitz.roughness_to_alpha(tex[2](uv)[0])
            dist = TrowbridgeReitz(rough, rough)
            if not is_black(r):
                bsdf.lobes.append(MicrofacetReflection(mul(r, ks), dist, eta))
            if not is_black(t):
                bsdf.lobes.append(MicrofacetTransmission(mul(t, ks), dist, ONE, eta))
    return bsdf


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
class WrappedShape:
    """TransformShape{shape, transform} as a light's shape (transform.rs:60-70): sample_by_point / by_point_pdf hand the
    inner shape the point -- and, for the pdf, the shape point -- carried through the inverse transform; the sampled
    point comes back through ShapePoint::apply."""

    def __init__(self, inner, prim, trace):
        self.inner, self.prim, self.trace = inner, prim, trace
        self.m3_inverse = trace.normal_matrix(prim.m)        # Normal3f::apply with the inverse transform

    def _apply(self, m, m3, sp):
        pw = self.trace.apply_point(m, sp.p)
        nw = [(m3[i][0] * sp.n[0] + m3[i][1] * sp.n[1]) + m3[i][2] * sp.n[2] for i in range(3)]
        g3 = gamma(3)
        err = [g3 * (((abs(m[i][0] * pw[0]) + abs(m[i][1] * pw[1])) + abs(m[i][2] * pw[2])) + m[i][3]) for i in range(3)]
        return ShapePoint(pw, nw, err)

    def sample_by_point(self, point, sampler):
        sp, pdf = self.inner.sample_by_point(self.trace.apply_point(self.prim.m_inv, point), sampler)
        return self._apply(self.prim.m, self.prim.m3, sp), pdf

    def by_point_pdf(self, point, sp):
        return self.inner.by_point_pdf(self.trace.apply_point(self.prim.m_inv, point), self._apply(self.prim.m_inv, self.m3_inverse, sp))


class SphereShape:
    """Sphere{radius} in its own space as a light's shape.  Sphere::sample_by_point (sphere.rs:66-99): cone sampling
    from outside, Shape::default_sample_by_point over Sphere::sample (a uniform direction, sphere.rs:54-65) from inside
    or on the surface; Sphere::by_point_pdf (sphere.rs:100-109)."""

    def __init__(self, radius):
        self.radius = F(radius)

    def area(self):
        return self.radius * self.radius * PI * F(4)

    def _sample(self, sampler):
        r = self.radius
        u = sampler.get_2d()
        z = ONE - F(2) * u[0]
        rr = np.sqrt(_max(ZERO, ONE - z * z))
        phi = F(2) * PI * u[1]
        d = [rr * _cosf(phi), rr * _sinf(phi), z]
        p = scale(d, r)
        n = normalize(d)
        p = scale(p, r / np.sqrt(norm2(p)))
        return ShapePoint(p, n, scale(vabs(p), gamma(5))), ONE / self.area()

    def sample_by_point(self, point, sampler):
        r = self.radius
        d2 = norm2(point)
        r2 = r * r
        if d2 <= r2:
            sp, pdf = self._sample(sampler)
            wi = sub(sp.p, point)
            if norm2(wi) == 0:
                return sp, ZERO
            pdf = pdf * norm2(wi) / -dot(normalize(wi), sp.n)
            return sp, (ZERO if (np.isnan(pdf) or np.isinf(pdf)) else pdf)
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
        p = scale(d, r)
        return ShapePoint(p, normalize(d), scale(vabs(p), gamma(5))), ONE / (F(2) * PI * (ONE - cos_max))

    def by_point_pdf(self, point, sp):
        r = self.radius
        d2 = norm2(point)
        if d2 < r * r:
            d = sub(point, sp.p)
            dd = norm2(d)
            pdf = dd / (abs(dot(unscale(d, np.sqrt(dd)), sp.n)) * self.area())
            return ZERO if (np.isnan(pdf) or np.isinf(pdf)) else pdf
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


class Distribution1D:
    """math/distribution.rs:16-67; the running sum is sequential f32 (np.cumsum accumulates in order)."""

    def __init__(self, f):
        f = np.asarray(f, np.float32)
        self.n = F(len(f))
        cdf = np.cumsum(f / self.n, dtype=np.float32)
        self.f_sum = cdf[-1]
        self.cdf = cdf / self.f_sum
        self.pdf = f / (self.n * self.f_sum)

    def sample_remap(self, u):
        i = int(np.searchsorted(self.cdf, u, side="left"))       # first i with u <= cdf[i]
        if i >= len(self.cdf) or not (u <= self.cdf[i]):
            return len(self.pdf) - 1, self.pdf[-1], ONE
        return i, self.pdf[i], (self.cdf[i] - u) / self.pdf[i]

    def sample_continuous(self, u):
        i, pdf, remap = self.sample_remap(u)
        return i, pdf * self.n, (F(i) + remap) / self.n


def _spherical_to_uv(w):
    """spherical_to_normalize_phi_theta (math/mod.rs:136-142), then the caller's v flip."""
    p = _atan2f(w[1], w[0])
    u = (p + F(2) * PI if p < 0 else p) / F(2) * INV_PI
    v = _acosf(clamp(w[2], F(-1), ONE)) * INV_PI
    return [u, ONE - v]


class InfiniteLight:
    """InfiniteAreaLight (light/infinite_area.rs): the table weights a texel by sin((row + 0.5) / rows) -- no pi --
    and Distribution2D's first coordinate runs over the rows."""
    is_delta = False
    le = None

    def __init__(self, desc, tex_index):
        t = desc.textures[tex_index]
        self.evaluate = texture(desc, tex_index)
        if t.kind == 0:
            width = height = 1
            pixels = np.array([[list(t.value)]], np.float32)
        else:
            width, height = int(t.width), int(t.height)
            assert int(t.channels) == 3
            texels = np.ctypeslib.as_array(desc.texels, shape=(int(desc.n_texels),))
            pixels = texels[int(t.offset):int(t.offset) + width * height * 3].reshape(height, width, 3).astype(np.float32)
        y = (F(0.212671) * pixels[..., 0] + F(0.715160) * pixels[..., 1]).astype(np.float32) + F(0.072169) * pixels[..., 2]
        rows = []
        for j in range(height):
            theta = (F(j) + HALF) / F(height)
            rows.append(Distribution1D((y[j] * _sinf(theta)).astype(np.float32)))
        self.rows = rows
        self.row_dist = Distribution1D([r.f_sum for r in rows])
        self.n = (self.row_dist.n, rows[0].n)

    def _solid_angle_pdf(self, pdf, uv):
        sin_t = _sinf(uv[1] * PI)
        return pdf / (F(2) * PI * PI * sin_t) if sin_t != 0 else ZERO

    def sample_li(self, scene, sp, sampler):
        u = sampler.get_2d()
        row, row_pdf, x = self.row_dist.sample_continuous(u[0])
        _, col_pdf, y = self.rows[row].sample_continuous(u[1])
        pdf = row_pdf * col_pdf
        uv = [x, ONE - y]
        theta, phi = uv[1] * PI, uv[0] * PI * F(2)
        wi = [_sinf(theta) * _cosf(phi), _sinf(theta) * _sinf(phi), _cosf(theta)]
        return wi, self.evaluate(uv), self._solid_angle_pdf(pdf, uv), (sp.offset_by_error(wi), wi, F(np.finfo(np.float32).max))

    def le_out_scene(self, d):
        return self.evaluate(_spherical_to_uv(normalize(d)))

    def out_scene_pdf(self, d):
        uv = _spherical_to_uv(normalize(d))
        n = self.n
        i = f32_to_usize(_min(uv[0] * n[0], n[0] - ONE))
        j = f32_to_usize(_min(uv[1] * n[1], n[1] - ONE))
        pdf = self.row_dist.pdf[i] * self.rows[i].pdf[j] * n[0] * n[1]
        return self._solid_angle_pdf(pdf, uv)


def add_out_scene(scene, l, beta, d):
    """`for light in lights { l += beta * light.le_out_scene(&ray) }`: None counts as black, and is still added."""
    for light in scene.lights:
        le = light.le_out_scene(d) if isinstance(light, InfiniteLight) else [ZERO, ZERO, ZERO]
        l = add(l, mul(beta, le))
    return l


class Scene:
    def __init__(self, desc, abi, sah, trace):
        n = int(desc.n_primitives)
        self.tris, self.material, self.light_of = [], [], []
        prims = []
        for i in range(n):
            pr = desc.primitives[i]
            chain = []                       # TransformShape wrappers, innermost first (cpbrt_transform.outer)
            x = int(pr.shape.xform)
            while x >= 0:
                chain.append(desc.transforms[x])
                x = int(desc.transforms[x].outer)
            if pr.shape.kind == abi.SHAPE_TRIANGLE:
                verts = [[desc.triangles[pr.shape.data].p[k][c] for c in range(3)] for k in range(3)]
                tri = Triangle(verts, [[desc.triangles[pr.shape.data].uv[k][c] for c in range(2)] for k in range(3)])
                obj = trace.TrianglePrim(verts)

                def to_sp(b, tri=tri):
                    sp = tri.hit_point(b)
                    return sp.p, sp.n, sp.p_error, sp.uv
                for k, t in enumerate(chain):
                    obj = trace.TransformedPrim(obj, list(t.m), list(t.m_inv), to_sp if k == 0 else None)
                self.tris.append(tri if not chain else None)
            else:
                assert pr.shape.kind == abi.SHAPE_SPHERE
                if chain:
                    obj = trace.SpherePrim(desc.sphere_radius[pr.shape.data], list(chain[0].m), list(chain[0].m_inv))
                else:
                    obj = trace.BareSpherePrim(desc.sphere_radius[pr.shape.data])
                for t in chain[1:]:
                    obj = trace.TransformedPrim(obj, list(t.m), list(t.m_inv))
                self.tris.append(None)
            prims.append(obj)
            self.light_of.append(int(pr.light))
            self.material.append(self._material(desc, pr.material) if pr.material >= 0 else None)
        self.lights = []
        for k in range(int(desc.n_lights)):
            lt = desc.lights[k]
            if lt.type == 1:
                self.lights.append(PointLight([F(x) for x in lt.v], [F(x) for x in lt.L]))
            elif lt.type == 2:
                self.lights.append(DistantLight([F(x) for x in lt.v], [F(x) for x in lt.L]))
            elif lt.type == 3:
                self.lights.append(InfiniteLight(desc, int(lt.tex)))
            else:
                # AreaLight{shape}: the Arc the light was created from (scene.rs:86-90) -- for a shape declared inside an
                # ObjectBegin block that is the un-instanced shape, whatever transform its instances are seen under
                assert lt.type == 0
                sh = lt.shape
                if sh.kind == abi.SHAPE_TRIANGLE:
                    shape = Triangle([[desc.triangles[sh.data].p[k][c] for c in range(3)] for k in range(3)])
                else:
                    shape = SphereShape(desc.sphere_radius[sh.data])
                x = int(sh.xform)
                while x >= 0:
                    t = desc.transforms[x]
                    shape = WrappedShape(shape, trace.TransformedPrim(None, list(t.m), list(t.m_inv)), trace)
                    x = int(t.outer)
                self.lights.append(AreaLight(shape, [F(x) for x in lt.L]))
        if prims:
            bounds = [q.bounds() for q in prims]
            lo = np.array([b[0] for b in bounds], np.float32)
            hi = np.array([b[1] for b in bounds], np.float32)
            self.search = trace.Searcher(sah.build(lo, hi), prims)
        else:
            self.search = None               # BVHAggregate::build returns before making a node; every query misses (bvh.rs:195-197,209-211)
        self.max_depth = int(desc.max_depth)
        self.rays_closest = 0
        self.rays_any = 0
        self.nan_dropped = 0

    @staticmethod
    def _material(desc, index):
        m = desc.materials[index]
        sub = tuple(Scene._material(desc, m.sub[k]) for k in range(2)) if m.type == MAT_MIX else None
        return int(m.type), [texture(desc, t) if t >= 0 else None for t in m.tex], sub

    def intersect(self, o, d, t_max=np.finfo(np.float32).max):
        self.rays_closest += 1
        if self.search is None:
            return None, None
        prim, _ = self.search.closest(o, d, t_max)
        if prim is None:
            return None, None
        if self.tris[prim] is None:
            return prim, ShapePoint(*self.search.hit_b)
        return prim, self.tris[prim].hit_point(self.search.hit_b)

    def occluded(self, o, d, t_max):
        self.rays_any += 1
        return self.search is not None and self.search.any_hit(o, d, t_max)


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
                if prim is not None:
                    if scene.light_of[prim] == chosen:
                        li_pdf = light.shape.by_point_pdf(sp.p, hit)
                        if li_pdf != 0:
                            ld = unscale(scale(scale(mul(light.le, f), abs(dot(n, wi))), power_heuristic(f_pdf, li_pdf)), f_pdf)
                elif isinstance(light, InfiniteLight):
                    le_pdf = light.out_scene_pdf(wi)
                    if le_pdf != 0:
                        ld = unscale(scale(scale(mul(light.le_out_scene(wi), f), abs(dot(n, wi))), power_heuristic(f_pdf, le_pdf)), f_pdf)
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
            if prim is None:
                l = add_out_scene(scene, l, beta, d)
            else:
                # `l += beta * intersect.le()`: le() is None off the emitters and None multiplies as black -- the term is
                # still added, so a NaN or infinite beta poisons l here even when nothing emits
                le = scene.lights[scene.light_of[prim]].le if scene.light_of[prim] >= 0 else [ZERO, ZERO, ZERO]
                l = add(l, mul(beta, le))
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
        return add_out_scene(scene, l, [ONE, ONE, ONE], d)
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


def render(scene, width, height, spp, halton_value, camera_ray_at, rect=None, integrator="path", camera_draws=False):
    """film[h, w, 3] over rect = (x0, y0, x1, y1); halton_value(x, y, sample) -> callable dim -> f32;
    camera_ray_at(fx, fy) -> (o[3], d[3]); with camera_draws (a thin lens) it is called as (fx, fy, sampler)."""
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
                    o, d = camera_ray_at(fx, fy, sampler) if camera_draws else camera_ray_at(fx, fy)
                    li = (path_li if integrator == "path" else direct_li)(scene, [F(v) for v in o], [F(v) for v in d], sampler)
                    if not has_nan(li):
                        kept.append(li)
                    else:
                        scene.nan_dropped += 1
                total = [ZERO, ZERO, ZERO]
                for li in kept:
                    total = add(total, li)
                film[y, x] = unscale(total, F(len(kept)))
    return film
