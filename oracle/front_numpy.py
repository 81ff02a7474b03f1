"""ORACLE -- TEST INFRASTRUCTURE ONLY.

The front of the render loop, restated a second time from the Rust alone (shares nothing with
oracle/oracle.cpp): the Halton sampler in Python integers + np.float32, and the perspective camera
under its LookAt transform.  With these two, oracle/path_numpy.py needs nothing from the C++
oracle: scene description in, image out.

  HaltonSampler::{new,set_pixel,next_sample,get_sample}   src/sampler/halton.rs:81-136
  multiplicative_inverse / extended_gcd                   halton.rs:5-19 (`x as usize % n`: a negative
                                                          Bezout coefficient wraps modulo 2^64 first)
  inverse / inverse_exp / radical_inverse                 halton.rs:39-70
  scrambled_inverse / scrambled_radical_inverse           halton.rs:25-38,57-65
  PRIMS, PRIMSUM                                          halton.rs:138-212 (the first 1000 primes, their prefix sums)
  PERMUTATION                                             an input: the reference shuffles it from thread_rng()
  PerspectiveCamera::generate_ray_without_sampler         src/camera/perspective.rs:47-51
  TransformCamera::generate_ray, Ray::apply               src/camera/mod.rs:66-68, src/geometry/ray.rs:45-51
  Point3f / Vector3f::apply                               src/geometry/{point,vector}.rs (nalgebra column order)

usize arithmetic is 64-bit and wraps (release build); `usize as Float` rounds to nearest even."""
import numpy as np

F = np.float32
_MASK = (1 << 64) - 1


def first_primes(count):
    sieve = np.ones(8000, bool)
    sieve[:2] = False
    for i in range(2, 90):
        if sieve[i]:
            sieve[i * i::i] = False
    primes = [int(p) for p in np.nonzero(sieve)[0][:count]]
    assert len(primes) == count
    return primes


PRIMS = first_primes(1000)
PRIMSUM = [0] * 1000
for _i in range(1, 1000):
    PRIMSUM[_i] = PRIMSUM[_i - 1] + PRIMS[_i - 1]


def usize_to_f32(n):
    """`n as f32`: one rounding, to nearest even (np.float32(int) would round twice through f64)."""
    if n < (1 << 24):
        return F(n)
    shift = n.bit_length() - 24
    q, r = divmod(n, 1 << shift)
    half = 1 << (shift - 1)
    if r > half or (r == half and (q & 1)):
        q += 1
    return F(q) * F(2.0 ** shift)


def extended_gcd(a, b):
    if b == 0:
        return 1, 0
    d = a // b
    xp, yp = extended_gcd(b, a % b)
    return yp, xp - d * yp


def multiplicative_inverse(a, n):
    x, _ = extended_gcd(a, n)
    return (x & _MASK) % n          # `x as usize % n`


def inverse(x, base_index):
    acc, digits, base = 0, 0, PRIMS[base_index]
    while x != 0:
        acc = (acc * base + x % base) & _MASK
        x //= base
        digits += 1
    return acc, digits


def inverse_exp(x, base_index, exp):
    r, digits = inverse(x, base_index)
    return (r * PRIMS[base_index] ** (exp - digits)) & _MASK if digits < exp else r


def radical_inverse(x, base_index):
    base = PRIMS[base_index]
    acc, digits = inverse(x, base_index)
    inv_base_n = F(1) / usize_to_f32((base ** digits) & _MASK)
    return inv_base_n * usize_to_f32(acc)


class HaltonSampler:
    def __init__(self, sample_per_pixel, resolution, permutation):
        self.spp = sample_per_pixel
        self.permutation = permutation
        self.scale, self.exp = [1, 1], [0, 0]
        for i in range(2):
            while self.scale[i] < resolution[i]:
                self.scale[i] *= PRIMS[i]
                self.exp[i] += 1
        self.mult_inverse = [multiplicative_inverse(self.scale[1], self.scale[0]),
                             multiplicative_inverse(self.scale[0], self.scale[1])]
        self.index = 0
        self.dim = 0

    def set_pixel(self, x, y):
        pixel = (x, y)
        self.index = 0
        for i in range(2):
            self.index = (self.index + inverse_exp(pixel[i], i, self.exp[i]) * self.scale[1 - i] * self.mult_inverse[i]) & _MASK
        self.index %= self.scale[0] * self.scale[1]
        self.dim = 0

    def next_sample(self):
        self.index = (self.index + self.scale[0] * self.scale[1]) & _MASK
        self.dim = 0

    def scrambled_radical_inverse(self, x, base_index):
        base = PRIMS[base_index]
        perm = self.permutation
        off = PRIMSUM[base_index]
        acc, digits = 0, 0
        while x != 0:
            acc = (acc * base + int(perm[off + x % base])) & _MASK
            x //= base
            digits += 1
        inv_base_n = F(1) / usize_to_f32((base ** digits) & _MASK)
        inv_base = F(1) / usize_to_f32(base)
        return inv_base_n * (usize_to_f32(acc) + inv_base * usize_to_f32(int(perm[off])) / (F(1) - inv_base))

    def get_sample(self):
        if self.dim < 2:
            r = radical_inverse(self.index // self.scale[self.dim], self.dim)
        else:
            assert self.dim < 1000, "dimensions >= 1000 are rand::random() in the reference"
            r = self.scrambled_radical_inverse(self.index, self.dim)
        self.dim += 1
        return r


# ---- camera ----------------------------------------------------------------------------------
def _rows(m16):
    return [[F(m16[r * 4 + c]) for c in range(4)] for r in range(4)]


def _mat_vec(m, v):
    return [((m[i][0] * v[0] + m[i][1] * v[1]) + m[i][2] * v[2]) + m[i][3] * v[3] for i in range(4)]


def _apply_point(m, p):
    h = _mat_vec(m, [p[0], p[1], p[2], F(1)])
    return [h[0] / h[3], h[1] / h[3], h[2] / h[3]]


def _libm(name):
    import ctypes
    import ctypes.util
    fn = getattr(ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6"), name)
    fn.restype = ctypes.c_float
    fn.argtypes = [ctypes.c_float]
    return lambda x: F(fn(float(x)))


_cosf, _sinf = _libm("cosf"), _libm("sinf")
_PI = F(3.14159265358979323846)


def _concentric_sample_disk(u):
    """src/math/mod.rs:98-110."""
    ux, uy = F(2) * u[0] - F(1), F(2) * u[1] - F(1)
    if ux == 0 or uy == 0:
        return F(0), F(0)
    if abs(ux) > abs(uy):
        r, theta = ux, (_PI / F(4)) * (uy / ux)
    else:
        r, theta = uy, (_PI / F(2)) - (_PI / F(4)) * (ux / uy)
    return r * _cosf(theta), r * _sinf(theta)


class PerspectiveCamera:
    """TransformCamera{LensCamera{PerspectiveCamera}}: raster_to_camera and camera_to_world as the host built them
    (4x4, row-major); with a lens (camera/lens.rs:24-33) generate_ray draws two numbers from `sampler`."""

    def __init__(self, raster_to_camera16, camera_to_world16=None, lens_radius=None, focal_distance=None):
        self.raster_to_camera = _rows(raster_to_camera16)
        self.camera_to_world = None if camera_to_world16 is None else _rows(camera_to_world16)
        self.lens_radius = None if lens_radius is None else F(lens_radius)
        self.focal_distance = None if focal_distance is None else F(focal_distance)

    def generate_ray(self, fx, fy, sampler=None):
        c = _apply_point(self.raster_to_camera, [F(fx), F(fy), F(0)])
        length = np.sqrt((c[0] * c[0] + c[1] * c[1]) + c[2] * c[2])
        o, d = [F(0), F(0), F(0)], [c[0] / length, c[1] / length, c[2] / length]
        if self.lens_radius is not None:
            lx, ly = _concentric_sample_disk(sampler.get_2d())
            lx, ly = self.lens_radius * lx, self.lens_radius * ly
            ft = self.focal_distance / d[2]
            focus = [o[k] + d[k] * ft for k in range(3)]
            o = [lx, ly, F(0)]
            v = [focus[k] - o[k] for k in range(3)]
            length = np.sqrt((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2])
            d = [v[0] / length, v[1] / length, v[2] / length]
        if self.camera_to_world is not None:
            o = _apply_point(self.camera_to_world, o)
            d = _mat_vec(self.camera_to_world, [d[0], d[1], d[2], F(0)])[:3]
        return o, d
