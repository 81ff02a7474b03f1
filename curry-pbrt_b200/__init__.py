"""curry-pbrt rendering hot path on B200 -- Python binding over the C-ABI in include/cpbrt.h.

The vocabulary follows the reference: ``render_from_file(path)`` (src/render.rs:63-82) is the
entry point; ``Scene`` is what ``SceneBuilder::build_with_clipper`` returns flattened to plain
arrays; ``Renderer.render()`` replaces ``render(scene, sampler, integrator, film, camera)``
(src/render.rs:7-50).

This module computes nothing itself: every call goes into ``libcpbrt.so`` (hand-written CUDA for
sm_100a plus the C++ host parser).  If the library is missing the import raises -- there is no
Python or CPU fallback.
"""
import ctypes as C
import os

import numpy as np

from . import _abi as abi

_HERE = os.path.dirname(os.path.abspath(__file__))
# CPBRT_LIB: load another build of the same library (tools/variant_sweep.py compares tuning variants)
LIB_PATH = os.environ.get("CPBRT_LIB") or os.path.join(_HERE, "libcpbrt.so")


class CpbrtError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("cpbrt error %d: %s" % (code, message))
        self.code = code


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "%s is missing: build it with `python curry-pbrt_b200/build.py` "
            "(there is no fallback implementation)" % LIB_PATH)
    return abi.bind(C.CDLL(LIB_PATH))


lib = _load()


def _host_check(rc):
    if rc != abi.OK:
        raise CpbrtError(rc, lib.cpbrt_host_last_error().decode(errors="replace"))


class Scene:
    """Host-side scene: parser + SceneParseStack + camera clipper (src/scene.rs:41-200)."""

    def __init__(self, handle):
        self._h = handle

    @classmethod
    def load(cls, path, perm_seed=1):
        h = C.c_void_p()
        _host_check(lib.cpbrt_scene_load(os.fsencode(path), perm_seed, C.byref(h)))
        return cls(h)

    @classmethod
    def parse(cls, text, base_dir=".", perm_seed=1):
        h = C.c_void_p()
        _host_check(lib.cpbrt_scene_parse(text.encode(), os.fsencode(base_dir), perm_seed, C.byref(h)))
        return cls(h)

    @property
    def desc(self):
        return lib.cpbrt_scene_get_desc(self._h).contents

    @property
    def desc_ptr(self):
        return lib.cpbrt_scene_get_desc(self._h)

    @property
    def film_filename(self):
        return lib.cpbrt_scene_film_filename(self._h).decode()

    @property
    def resolution(self):
        d = self.desc
        return d.film_width, d.film_height

    def set_spp(self, spp):
        _host_check(lib.cpbrt_scene_set_spp(self._h, spp))

    def set_integrator(self, integrator, max_depth):
        _host_check(lib.cpbrt_scene_set_integrator(self._h, integrator, max_depth))

    def close(self):
        if self._h:
            lib.cpbrt_scene_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _count_flag(count):
    return 4 if count == "own" else (2 if count else 0)


class Renderer:
    """One GPU context: upload, BVH build (Aggregate::build) and the device entry points."""

    def __init__(self, device=0):
        self._ctx = C.c_void_p()
        rc = lib.cpbrt_create(C.byref(self._ctx), device)
        if rc != abi.OK:
            raise CpbrtError(rc, lib.cpbrt_last_error(None).decode(errors="replace"))
        self._scene = None
        self.rank, self.n_ranks = 0, 1
        self.n_primitives = 0
        self.width = self.height = 0

    def _check(self, rc):
        if rc != abi.OK:
            raise CpbrtError(rc, lib.cpbrt_last_error(self._ctx).decode(errors="replace"))

    def upload(self, scene):
        self._check(lib.cpbrt_upload_scene(self._ctx, scene.desc_ptr))
        self._scene = scene
        d = scene.desc
        self.n_primitives = d.n_primitives
        self.width, self.height = d.film_width, d.film_height

    def build_bvh(self, want_nodes=False):
        n = int(self.n_primitives)
        order = np.zeros(n, np.uint32)
        if not want_nodes:
            self._check(lib.cpbrt_build_bvh(self._ctx, order.ctypes.data_as(C.POINTER(C.c_uint32)), None))
            return order
        cap = max(2 * n, 1)
        arrs = dict(is_leaf=np.zeros(cap, np.uint8), axis=np.zeros(cap, np.uint8), a=np.zeros(cap, np.uint32),
                    b=np.zeros(cap, np.uint32), bound=np.zeros((cap, 6), np.float32))
        dump = abi.BvhDump(capacity=cap,
                           is_leaf=arrs["is_leaf"].ctypes.data_as(C.POINTER(C.c_uint8)),
                           axis=arrs["axis"].ctypes.data_as(C.POINTER(C.c_uint8)),
                           a=arrs["a"].ctypes.data_as(C.POINTER(C.c_uint32)),
                           b=arrs["b"].ctypes.data_as(C.POINTER(C.c_uint32)),
                           bound=arrs["bound"].ctypes.data_as(C.POINTER(C.c_float)))
        self._check(lib.cpbrt_build_bvh(self._ctx, order.ctypes.data_as(C.POINTER(C.c_uint32)), C.byref(dump)))
        nn = int(dump.n_nodes)
        nodes = {k: v[:nn] for k, v in arrs.items()}
        nodes["root"] = int(dump.root)
        return order, nodes

    def trace_closest(self, rays, count=False):
        """count=True: reference-order traversal with its visit counters; count="own": the shipped
        traversal counting its own node steps / primitive tests (include/cpbrt.h, CPBRT_TRACE_COUNT*)."""
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 8)
        hits = np.zeros((rays.shape[0], 4), np.float32)
        stats = abi.Stats()
        self._check(lib.cpbrt_trace_closest(self._ctx, rays.ctypes.data, rays.shape[0], hits.ctypes.data,
                                            _count_flag(count), C.byref(stats)))
        out = dict(t=hits[:, 0].copy(), prim=hits[:, 1].view(np.uint32).copy(), u=hits[:, 2].copy(), v=hits[:, 3].copy())
        return (out, stats) if count else out

    def trace_any(self, rays, count=False):
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 8)
        out = np.zeros(rays.shape[0], np.uint8)
        stats = abi.Stats()
        self._check(lib.cpbrt_trace_any(self._ctx, rays.ctypes.data, rays.shape[0], out.ctypes.data,
                                        _count_flag(count), C.byref(stats)))
        return (out, stats) if count else out

    def trace_closest_device(self, rays_ptr, n, hits_ptr, count=False):
        """Batch already resident in HBM (device pointers); returns Stats (ms_trace = kernel time)."""
        stats = abi.Stats()
        self._check(lib.cpbrt_trace_closest(self._ctx, C.c_void_p(rays_ptr), n, C.c_void_p(hits_ptr), 1 | _count_flag(count), C.byref(stats)))
        return stats

    def trace_any_device(self, rays_ptr, n, out_ptr, count=False):
        stats = abi.Stats()
        self._check(lib.cpbrt_trace_any(self._ctx, C.c_void_p(rays_ptr), n, C.c_void_p(out_ptr), 1 | _count_flag(count), C.byref(stats)))
        return stats

    def halton(self, px, py, sample, dim):
        px, py, sample, dim = (np.ascontiguousarray(a, np.uint32) for a in (px, py, sample, dim))
        n = px.shape[0]
        index = np.zeros(n, np.uint64)
        value = np.zeros(n, np.float32)
        self._check(lib.cpbrt_halton(self._ctx, px.ctypes.data, py.ctypes.data, sample.ctypes.data, dim.ctypes.data, n,
                                     index.ctypes.data, value.ctypes.data))
        return index, value

    def camera_rays(self, sample=0):
        rays = np.zeros((self.width * self.height, 8), np.float32)
        self._check(lib.cpbrt_camera_rays(self._ctx, sample, rays.ctypes.data, 0))
        return rays

    def render(self, sample_begin=0, sample_end=0, first_hit_only=False, samples_per_pass=0, film=None, device_film_ptr=None,
               profile_stages=False, srgb8=False):
        """render() of src/render.rs:7-50.  Returns (film[H, W, 3] float32, Stats); with srgb8 the
        film is the uint8 image the reference writes to its PNG (gamma + quantise on the device)."""
        params = abi.RenderParams(sample_begin=sample_begin, sample_end=sample_end, first_hit_only=int(first_hit_only),
                                  resolve=1, samples_per_pass=samples_per_pass, film_on_device=0,
                                  profile_stages=int(profile_stages), output_srgb8=int(srgb8))
        stats = abi.Stats()
        if device_film_ptr is not None:
            params.film_on_device = 1
            self._check(lib.cpbrt_render(self._ctx, C.byref(params), C.c_void_p(device_film_ptr), C.byref(stats)))
            return None, stats
        if self.rank != 0:
            self._check(lib.cpbrt_render(self._ctx, C.byref(params), None, C.byref(stats)))
            return None, stats
        if film is None:
            film = np.zeros((self.height, self.width, 3), np.uint8 if srgb8 else np.float32)
        self._check(lib.cpbrt_render(self._ctx, C.byref(params), film.ctypes.data, C.byref(stats)))
        return film, stats

    def bvh_gpu_levels(self):
        """Levels of the last cpbrt_build_bvh that ran on the device (0: pure host build)."""
        return int(lib.cpbrt_bvh_gpu_levels(self._ctx))

    def bvh_host_segments(self):
        """Subtrees of the last device build finished by the host builder (0: whole build on the device; -1: host build)."""
        return int(lib.cpbrt_bvh_host_segments(self._ctx))

    def bvh_device_layout(self):
        """The traversal layout in device memory after build_bvh (cpbrt_bvh_device_layout): dict with
        nodes[n_nodes, 16] float32, records[n, 12], frames[n, 12], uv[n, 6] or None, resident (bool)."""
        n_nodes, resident = C.c_uint64(), C.c_int()
        self._check(lib.cpbrt_bvh_device_layout(self._ctx, 0, C.byref(n_nodes), None, None, None, None, C.byref(resident)))
        n = int(self.n_primitives)
        nodes = np.zeros((n_nodes.value, 16), np.float32)
        records, frames, uv = np.zeros((n, 12), np.float32), np.zeros((n, 12), np.float32), np.zeros((n, 6), np.float32)
        rc = lib.cpbrt_bvh_device_layout(self._ctx, n_nodes.value, None, nodes.ctypes.data, records.ctypes.data, frames.ctypes.data,
                                         uv.ctypes.data, None)
        self._check(min(rc, 0))
        return dict(nodes=nodes, records=records, frames=frames, uv=uv if rc == 1 else None, resident=bool(resident.value))

    def check_failures(self):
        """(checked_build, violations, first_site): bounds violations counted by a -DCPBRT_CHECKED build of the library."""
        n, site = C.c_uint64(), C.c_uint32()
        rc = lib.cpbrt_check_failures(self._ctx, C.byref(n), C.byref(site))
        self._check(min(rc, 0))
        return bool(rc), int(n.value), int(site.value)

    def h2d_bytes(self):
        """Bytes this context has copied host -> device so far (counted where the copies are issued)."""
        return int(lib.cpbrt_h2d_bytes(self._ctx))

    def read_bandwidth(self, nbytes, repeats=20):
        """GB/s of a read-only streaming kernel over a device buffer (roofline denominators)."""
        v = C.c_double()
        self._check(lib.cpbrt_measure_read_bandwidth(self._ctx, nbytes, repeats, C.byref(v)))
        return v.value

    def comm_init(self, n_ranks, rank, unique_id):
        buf = (C.c_uint8 * abi.COMM_ID_BYTES).from_buffer_copy(bytes(unique_id))
        self._check(lib.cpbrt_comm_init(self._ctx, n_ranks, rank, buf))
        self.rank, self.n_ranks = rank, n_ranks

    def close(self):
        if self._ctx:
            lib.cpbrt_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def comm_unique_id():
    buf = (C.c_uint8 * abi.COMM_ID_BYTES)()
    rc = lib.cpbrt_comm_unique_id(buf)
    if rc != abi.OK:
        raise CpbrtError(rc, lib.cpbrt_last_error(None).decode(errors="replace"))
    return bytes(buf)


def write_png(path, film):
    film = np.ascontiguousarray(film, np.float32)
    h, w = film.shape[:2]
    _host_check(lib.cpbrt_write_png(os.fsencode(path), film.ctypes.data_as(C.POINTER(C.c_float)), w, h))


def render_from_file(path, device=0):
    """render_from_file (src/render.rs:63-82): writes the Film's PNG and prints its name."""
    rc = lib.cpbrt_render_from_file(os.fsencode(path), device)
    if rc != abi.OK:
        raise CpbrtError(rc, lib.cpbrt_host_last_error().decode(errors="replace"))
