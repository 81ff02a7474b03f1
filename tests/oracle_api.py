"""ctypes binding of oracle/liboracle.so -- the CPU restatement used as the checker.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this.
"""
import ctypes as C
import importlib.util
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load_module(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


_build = _load_module("oracle_build", os.path.join(ROOT, "oracle", "build.py"))


def _bind():
    lib = C.CDLL(_build.build())
    vp, u32, u64, i32 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int
    lib.orc_create.restype = vp; lib.orc_create.argtypes = [vp]
    lib.orc_destroy.argtypes = [vp]
    lib.orc_bvh_node_count.restype = u64; lib.orc_bvh_node_count.argtypes = [vp]
    lib.orc_bvh_root.restype = u64; lib.orc_bvh_root.argtypes = [vp]
    lib.orc_bvh_order.argtypes = [vp, vp]
    lib.orc_bvh_nodes.argtypes = [vp, vp, vp, vp, vp, vp]
    lib.orc_halton.argtypes = [vp, vp, vp, vp, vp, u64, vp, vp]
    lib.orc_halton_pixel_index.restype = u64; lib.orc_halton_pixel_index.argtypes = [vp, u32, u32]
    lib.orc_radical_inverse.restype = C.c_float; lib.orc_radical_inverse.argtypes = [u64, u32]
    lib.orc_camera_rays.argtypes = [vp, u32, vp]
    lib.orc_camera_ray_at.argtypes = [vp, C.c_float, C.c_float, vp]
    lib.orc_trace_closest.argtypes = [vp, vp, u64, vp, vp, u32]
    lib.orc_trace_any.argtypes = [vp, vp, u64, vp, vp, u32]
    lib.orc_trace_closest_plain.argtypes = [vp, vp, u64, vp]
    lib.orc_render.restype = C.c_double
    lib.orc_render.argtypes = [vp, i32, i32, u32, u32, u32, u32, u32, vp, vp, i32]
    return lib


lib = _bind()


class OracleWorld:
    """The oracle's view of one scene (BVH built on construction)."""

    def __init__(self, scene):
        self.scene = scene   # keeps the description alive
        self._w = lib.orc_create(C.cast(scene.desc_ptr, C.c_void_p))
        d = scene.desc
        self.n_primitives = int(d.n_primitives)
        self.width, self.height = d.film_width, d.film_height

    def close(self):
        if self._w:
            lib.orc_destroy(self._w)
            self._w = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def bvh_order(self):
        out = np.zeros(self.n_primitives, np.uint32)
        lib.orc_bvh_order(self._w, out.ctypes.data)
        return out

    def bvh_nodes(self):
        n = int(lib.orc_bvh_node_count(self._w))
        r = dict(is_leaf=np.zeros(n, np.uint8), axis=np.zeros(n, np.uint8), a=np.zeros(n, np.uint32),
                 b=np.zeros(n, np.uint32), bound=np.zeros((n, 6), np.float32))
        lib.orc_bvh_nodes(self._w, r["is_leaf"].ctypes.data, r["axis"].ctypes.data, r["a"].ctypes.data,
                          r["b"].ctypes.data, r["bound"].ctypes.data)
        r["root"] = int(lib.orc_bvh_root(self._w))
        return r

    def halton(self, px, py, sample, dim):
        px, py, sample, dim = (np.ascontiguousarray(a, np.uint32) for a in (px, py, sample, dim))
        n = px.shape[0]
        index = np.zeros(n, np.uint64); value = np.zeros(n, np.float32)
        lib.orc_halton(self._w, px.ctypes.data, py.ctypes.data, sample.ctypes.data, dim.ctypes.data, n,
                       index.ctypes.data, value.ctypes.data)
        return index, value

    def camera_rays(self, sample=0):
        rays = np.zeros((self.width * self.height, 8), np.float32)
        lib.orc_camera_rays(self._w, sample, rays.ctypes.data)
        return rays

    def camera_ray_at(self, fx, fy):
        ray = np.zeros(8, np.float32)
        lib.orc_camera_ray_at(self._w, fx, fy, ray.ctypes.data)
        return ray

    def trace_closest(self, rays, count=False, threads=8):
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 8)
        hits = np.zeros((rays.shape[0], 4), np.float32)
        counters = np.zeros(2, np.uint64)
        lib.orc_trace_closest(self._w, rays.ctypes.data, rays.shape[0], hits.ctypes.data,
                              counters.ctypes.data if count else None, threads)
        out = dict(t=hits[:, 0].copy(), prim=hits[:, 1].view(np.uint32).copy(), u=hits[:, 2].copy(), v=hits[:, 3].copy())
        return (out, counters) if count else out

    def trace_any(self, rays, count=False, threads=8):
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 8)
        out = np.zeros(rays.shape[0], np.uint8)
        counters = np.zeros(2, np.uint64)
        lib.orc_trace_any(self._w, rays.ctypes.data, rays.shape[0], out.ctypes.data,
                          counters.ctypes.data if count else None, threads)
        return (out, counters) if count else out

    def trace_closest_plain(self, rays):
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 8)
        hits = np.zeros((rays.shape[0], 4), np.float32)
        lib.orc_trace_closest_plain(self._w, rays.ctypes.data, rays.shape[0], hits.ctypes.data)
        return dict(t=hits[:, 0].copy(), prim=hits[:, 1].view(np.uint32).copy())

    def render(self, integrator=-1, max_depth=-1, rect=(0, 0, 0, 0), threads=0, film=None, count_visits=False):
        """Returns (film[H, W, 3], seconds, dict(rays_closest, rays_any, nan_dropped[, visit counts]))."""
        if film is None:
            film = np.zeros((self.height, self.width, 3), np.float32)
        st = np.zeros(8, np.uint64)
        sec = lib.orc_render(self._w, integrator, max_depth, rect[0], rect[1], rect[2], rect[3], threads,
                             film.ctypes.data, st.ctypes.data, int(count_visits))
        out = dict(rays_closest=int(st[0]), rays_any=int(st[1]), nan_dropped=int(st[2]), rays_unused_last=int(st[7]))
        if count_visits:
            out.update(closest_parent_visits=int(st[3]), closest_prim_tests=int(st[4]),
                       any_parent_visits=int(st[5]), any_prim_tests=int(st[6]))
        return film, sec, out
