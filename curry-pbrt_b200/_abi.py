"""ctypes mirror of include/cpbrt.h (struct layouts and function prototypes).

Only the declarations live here; nothing in this module computes anything.
"""
import ctypes as C

ABI_VERSION = 2
N_PRIMES = 1000
PERM_TABLE_LEN = 3682913
NO_HIT = 0xFFFFFFFF
COMM_ID_BYTES = 128

OK = 0
ERR_NO_DEVICE, ERR_CUDA, ERR_BAD_ARG, ERR_BAD_CTX = -1, -2, -3, -4
ERR_PARSE, ERR_IO, ERR_UNSUPPORTED, ERR_NCCL = -5, -6, -7, -8

SHAPE_TRIANGLE, SHAPE_SPHERE = 0, 1
TEX_CONSTANT, TEX_IMAGE = 0, 1
MAT_MATTE, MAT_MIRROR, MAT_GLASS, MAT_PLASTIC, MAT_UBER, MAT_TRANSLUCENT, MAT_MIX = range(7)
LIGHT_AREA, LIGHT_POINT, LIGHT_DISTANT, LIGHT_INFINITE = range(4)
INTEGRATOR_PATH, INTEGRATOR_DIRECT = 0, 1
MAT_MAX_TEX = 9


class Transform(C.Structure):
    _fields_ = [("m", C.c_float * 16), ("m_inv", C.c_float * 16), ("outer", C.c_int32), ("_pad", C.c_int32 * 3)]


class Triangle(C.Structure):
    _fields_ = [("p", (C.c_float * 3) * 3), ("uv", (C.c_float * 2) * 3)]


class Shape(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("data", C.c_uint32), ("xform", C.c_int32), ("_pad", C.c_int32)]


class Primitive(C.Structure):
    _fields_ = [("shape", Shape), ("material", C.c_int32), ("light", C.c_int32)]


class Texture(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("channels", C.c_uint32), ("value", C.c_float * 3),
                ("width", C.c_uint32), ("height", C.c_uint32), ("offset", C.c_uint64)]


class Material(C.Structure):
    _fields_ = [("type", C.c_uint32), ("tex", C.c_int32 * MAT_MAX_TEX), ("sub", C.c_int32 * 2)]


class Light(C.Structure):
    _fields_ = [("type", C.c_uint32), ("shape", Shape), ("v", C.c_float * 3), ("L", C.c_float * 3), ("tex", C.c_int32)]


class Camera(C.Structure):
    _fields_ = [("raster_to_camera", C.c_float * 16), ("camera_to_raster", C.c_float * 16),
                ("camera_to_world", C.c_float * 16), ("has_camera_to_world", C.c_int32),
                ("has_lens", C.c_int32), ("lens_radius", C.c_float), ("focal_distance", C.c_float)]


class Sampler(C.Structure):
    _fields_ = [("spp", C.c_uint32), ("scale", C.c_uint64 * 2), ("exp", C.c_uint32 * 2),
                ("mult_inverse", C.c_uint64 * 2), ("permutation", C.POINTER(C.c_uint16))]


class SceneDesc(C.Structure):
    _fields_ = [
        ("abi_version", C.c_uint32),
        ("n_primitives", C.c_uint64), ("primitives", C.POINTER(Primitive)),
        ("n_triangles", C.c_uint64), ("triangles", C.POINTER(Triangle)),
        ("n_spheres", C.c_uint64), ("sphere_radius", C.POINTER(C.c_float)),
        ("n_transforms", C.c_uint32), ("transforms", C.POINTER(Transform)),
        ("n_materials", C.c_uint32), ("materials", C.POINTER(Material)),
        ("n_textures", C.c_uint32), ("textures", C.POINTER(Texture)),
        ("n_texels", C.c_uint64), ("texels", C.POINTER(C.c_float)),
        ("n_lights", C.c_uint32), ("lights", C.POINTER(Light)),
        ("camera", Camera),
        ("sampler", Sampler),
        ("integrator", C.c_uint32),
        ("max_depth", C.c_uint32),
        ("film_width", C.c_uint32), ("film_height", C.c_uint32),
    ]


class Ray(C.Structure):
    _fields_ = [("o", C.c_float * 3), ("t_max", C.c_float), ("d", C.c_float * 3), ("_pad", C.c_float)]


class Hit(C.Structure):
    _fields_ = [("t", C.c_float), ("prim", C.c_uint32), ("u", C.c_float), ("v", C.c_float)]


class BvhDump(C.Structure):
    _fields_ = [("capacity", C.c_uint64), ("n_nodes", C.c_uint64), ("root", C.c_uint64),
                ("is_leaf", C.POINTER(C.c_uint8)), ("axis", C.POINTER(C.c_uint8)),
                ("a", C.POINTER(C.c_uint32)), ("b", C.POINTER(C.c_uint32)), ("bound", C.POINTER(C.c_float))]


class RenderParams(C.Structure):
    _fields_ = [("sample_begin", C.c_uint32), ("sample_end", C.c_uint32), ("first_hit_only", C.c_uint32),
                ("resolve", C.c_uint32), ("samples_per_pass", C.c_uint32), ("film_on_device", C.c_uint32),
                ("profile_stages", C.c_uint32), ("output_srgb8", C.c_uint32)]


class Stats(C.Structure):
    _fields_ = [("camera_samples", C.c_uint64), ("rays_closest", C.c_uint64), ("rays_any", C.c_uint64),
                ("nan_dropped", C.c_uint64), ("kernel_launches", C.c_uint64),
                ("ms_total", C.c_double), ("ms_trace", C.c_double), ("ms_shade", C.c_double),
                ("ms_raygen_film", C.c_double), ("ms_reduce", C.c_double),
                ("trace_parent_visits", C.c_uint64), ("trace_prim_tests", C.c_uint64),
                ("trace_launches", C.c_uint64), ("ms_trace_closest", C.c_double), ("ms_trace_shadow", C.c_double),
                ("rays_closest_traced", C.c_uint64), ("shade_items", C.c_uint64), ("trace_slow_paths", C.c_uint64)]


# name -> (restype, argtypes); every symbol include/cpbrt.h declares
HOST_SYMBOLS = {
    "cpbrt_scene_load": (C.c_int, [C.c_char_p, C.c_uint64, C.POINTER(C.c_void_p)]),
    "cpbrt_scene_parse": (C.c_int, [C.c_char_p, C.c_char_p, C.c_uint64, C.POINTER(C.c_void_p)]),
    "cpbrt_scene_free": (None, [C.c_void_p]),
    "cpbrt_scene_get_desc": (C.POINTER(SceneDesc), [C.c_void_p]),
    "cpbrt_scene_film_filename": (C.c_char_p, [C.c_void_p]),
    "cpbrt_host_last_error": (C.c_char_p, []),
    "cpbrt_scene_set_spp": (C.c_int, [C.c_void_p, C.c_uint32]),
    "cpbrt_scene_set_integrator": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32]),
    "cpbrt_make_permutation": (None, [C.c_uint64, C.POINTER(C.c_uint16)]),
    "cpbrt_host_halton_grouped": (C.c_int, [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]),
    "cpbrt_sample_share": (None, [C.c_uint32, C.c_int, C.c_int, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    "cpbrt_write_png": (C.c_int, [C.c_char_p, C.POINTER(C.c_float), C.c_uint32, C.c_uint32]),
    "cpbrt_render_from_file": (C.c_int, [C.c_char_p, C.c_int]),
    "cpbrt_host_build_bvh": (C.c_int, [C.POINTER(SceneDesc), C.POINTER(C.c_uint32), C.POINTER(BvhDump)]),
}
DEVICE_SYMBOLS = {
    "cpbrt_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int]),
    "cpbrt_destroy": (None, [C.c_void_p]),
    "cpbrt_last_error": (C.c_char_p, [C.c_void_p]),
    "cpbrt_upload_scene": (C.c_int, [C.c_void_p, C.POINTER(SceneDesc)]),
    "cpbrt_build_bvh": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(BvhDump)]),
    "cpbrt_trace_closest": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_int, C.POINTER(Stats)]),
    "cpbrt_trace_any": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_int, C.POINTER(Stats)]),
    "cpbrt_halton": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]),
    "cpbrt_camera_rays": (C.c_int, [C.c_void_p, C.c_uint32, C.c_void_p, C.c_int]),
    "cpbrt_render": (C.c_int, [C.c_void_p, C.POINTER(RenderParams), C.c_void_p, C.POINTER(Stats)]),
    "cpbrt_bvh_gpu_levels": (C.c_int, [C.c_void_p]),
    "cpbrt_bvh_host_segments": (C.c_int, [C.c_void_p]),
    "cpbrt_check_failures": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)]),
    "cpbrt_bvh_device_layout": (C.c_int, [C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]),
    "cpbrt_h2d_bytes": (C.c_uint64, [C.c_void_p]),
    "cpbrt_measure_read_bandwidth": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint32, C.POINTER(C.c_double)]),
    "cpbrt_comm_unique_id": (C.c_int, [C.c_void_p]),
    "cpbrt_comm_init": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
}
ALL_SYMBOLS = dict(HOST_SYMBOLS, **DEVICE_SYMBOLS)


def bind(lib):
    """Attach restype/argtypes for every declared symbol; raises AttributeError if one is missing."""
    for name, (res, args) in ALL_SYMBOLS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib
