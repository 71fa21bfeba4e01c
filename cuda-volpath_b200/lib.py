"""ctypes binding of libvolpath_b200.so -- the C ABI declared in include/volpath.h.

There is no CPU path: if the shared library is missing this module raises at load time, and every
compute call fails with the library's own error when no CUDA device is present."""
import ctypes
import os

from .param import Param

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VOLPATH_B200_LIB") or os.path.join(_HERE, "libvolpath_b200.so")  # override: kernel-variant experiments

c_fp = ctypes.POINTER(ctypes.c_float)
c_vp = ctypes.c_void_p
c_int = ctypes.c_int
c_uint = ctypes.c_uint
c_u64p = ctypes.POINTER(ctypes.c_ulonglong)

VP_OK = 0
VOXEL_U8, VOXEL_F16, VOXEL_F32 = 0, 1, 2
MEM_HOST, MEM_DEVICE = 0, 1
BOUNDS_VOXEL, BOUNDS_CELL, BOUNDS_EXACT = 1, 2, 4
MODE_PARITY, MODE_FAST, MODE_WAVE = 0, 1, 2


class Float3(ctypes.Structure):
    _fields_ = [("x", ctypes.c_float), ("y", ctypes.c_float), ("z", ctypes.c_float)]


class SkyState(ctypes.Structure):
    """vp_sky_state (include/volpath.h): the host-side state of the reference's Skydome after prepareForRender()."""
    _fields_ = [("configs", ctypes.c_double * 9 * 11), ("radiances", ctypes.c_double * 11),
                ("emission_correction_factor_sky", ctypes.c_double * 11), ("lambdas", ctypes.c_float * 7),
                ("weights", ctypes.c_float * 3 * 7), ("gamma_scale", ctypes.c_float), ("sun_dir", ctypes.c_float * 3),
                ("ground_rgb", ctypes.c_float * 3), ("sunsky_scale", ctypes.c_float)]


class Dim3(ctypes.Structure):
    _fields_ = [("x", c_uint), ("y", c_uint), ("z", c_uint)]


class Extent(ctypes.Structure):
    _fields_ = [("width", ctypes.c_size_t), ("height", ctypes.c_size_t), ("depth", ctypes.c_size_t)]


# every symbol include/volpath.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "vp_last_error": (ctypes.c_char_p, []),
    "vp_version": (ctypes.c_char_p, []),
    "vp_create": (c_int, [c_int, ctypes.POINTER(c_vp)]),
    "vp_destroy": (c_int, [c_vp]),
    "vp_upload_volume": (c_int, [c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_fp, c_fp, c_int]),
    "vp_generate_cloud": (c_int, [c_vp, c_int, c_int, c_int, c_uint, c_int, c_fp, c_fp, c_int, c_int]),
    "vp_dense_volume": (c_vp, [c_vp]),
    "vp_release_dense": (c_int, [c_vp]),
    "vp_set_julia": (c_int, [c_vp]),
    "vp_set_filter": (c_int, [c_vp, c_int]),
    "vp_set_envmap": (c_int, [c_vp, c_fp, c_int, c_int]),
    "vp_set_sun": (c_int, [c_vp, c_fp, c_fp]),
    "vp_set_env_sampling": (c_int, [c_vp, c_int]),
    "vp_bake_sunsky": (c_int, [c_vp, c_vp, c_int, c_int]),
    "vp_get_envmap": (c_int, [c_vp, c_fp, ctypes.POINTER(c_int)]),
    "vp_set_inv_view": (c_int, [c_vp, c_fp]),
    "vp_precompute_opacity": (c_int, [c_vp, c_fp]),
    "vp_precompute_opacity_sharded": (c_int, [c_vp, c_fp]),
    "vp_free_volume": (c_int, [c_vp]),
    "vp_render": (c_int, [c_vp, c_vp, c_int, c_int, c_int, ctypes.POINTER(Param), c_int, c_vp]),
    "vp_render_to_host": (c_int, [c_vp, c_vp, c_int, c_int, c_int, ctypes.POINTER(Param), c_int]),
    "vp_resolve": (c_int, [c_vp, c_vp, c_vp, c_int, ctypes.c_float, ctypes.c_float, c_vp]),
    "vp_accumulate": (c_int, [c_vp, c_vp, c_vp, c_int, c_vp]),
    "vp_sync": (c_int, [c_vp]),
    "vp_nccl_available": (c_int, []),
    "vp_nccl_unique_id": (c_int, [ctypes.c_char_p]),
    "vp_nccl_init": (c_int, [c_vp, c_int, c_int, ctypes.c_char_p]),
    "vp_reduce_nccl": (c_int, [c_vp, c_vp, c_vp, c_int, c_int, c_vp]),
    "vp_nccl_destroy": (c_int, [c_vp]),
    "vp_reduce": (c_int, [ctypes.POINTER(c_vp), ctypes.POINTER(c_vp), c_int, c_int, c_int]),
    "vp_ipc_export": (c_int, [c_vp, c_vp, ctypes.c_char_p]),
    "vp_reduce_ipc": (c_int, [c_vp, c_vp, ctypes.c_char_p, c_int, c_int, c_vp]),
    "vp_ipc_close": (c_int, [c_vp]),
    "vp_get_bounds_voxel": (c_int, [c_vp, c_fp]),
    "vp_get_bounds_cell": (c_int, [c_vp, c_fp, ctypes.POINTER(c_int)]),
    "vp_get_half_tables": (c_int, [c_vp, c_vp, c_vp, c_fp, ctypes.POINTER(c_int)]),
    "vp_get_opacity": (c_int, [c_vp, c_fp]),
    "vp_get_opacity_fast": (c_int, [c_vp, c_fp]),
    "vp_opacity_build_ms": (c_int, [c_vp, c_fp]),
    "vp_fetch_density": (c_int, [c_vp, c_fp, c_int, c_int, c_fp]),
    "vp_volume_stats": (c_int, [c_vp, c_u64p]),
    "vp_rng_sequence": (c_int, [c_vp, c_uint, c_uint, c_uint, c_int, c_fp, ctypes.POINTER(c_uint)]),
    "vp_philox2x32": (c_int, [c_vp, c_uint, c_uint, c_uint, ctypes.POINTER(c_uint)]),
    "vp_set_stats": (c_int, [c_vp, c_int]),
    "vp_render_counters": (c_int, [c_vp, c_u64p, c_int]),
    "vp_last_kernel_ms": (c_int, [c_vp, c_fp]),
    "vp_launch_count": (c_int, [c_vp, c_u64p]),
    "vp_dev_alloc": (c_vp, [ctypes.c_size_t]),
    "vp_dev_free": (c_int, [c_vp]),
    "vp_dev_zero": (c_int, [c_vp, ctypes.c_size_t]),
    "vp_dev_to_host": (c_int, [c_vp, c_vp, ctypes.c_size_t]),
    "vp_host_to_dev": (c_int, [c_vp, c_vp, ctypes.c_size_t]),
    # reference-named shims (src/volumeRender.cpp:117-128, 347-356)
    "init_cuda": (None, [c_vp, Extent, ctypes.c_bool, ctypes.POINTER(Float3), ctypes.POINTER(Float3)]),
    "set_texture_filter_mode": (None, [ctypes.c_bool]),
    "free_cuda_buffers": (None, []),
    "precompute_opacity": (None, [c_fp]),
    "init_envmap": (None, [c_vp, c_int, c_int]),
    "free_envmap": (None, []),
    "set_sun": (None, [c_fp, c_fp]),
    "copy_inv_view_matrix": (None, [c_fp, ctypes.c_size_t]),
    "copy_inv_model_matrix": (None, [c_fp, ctypes.c_size_t]),
    "init_rng": (None, [Dim3, Dim3, c_int, c_int]),
    "free_rng": (None, []),
    "scale": (None, [c_vp, c_vp, c_int, ctypes.c_float]),
    "gamma_correct": (None, [c_vp, c_vp, c_int, ctypes.c_float, ctypes.c_float]),
    "render_kernel": (None, [Dim3, Dim3, c_vp, c_int, ctypes.POINTER(Param)]),
    "vp_shim_set_mode": (None, [c_int]),
    "vp_shim_sync": (c_int, []),
    "vp_shim_context": (c_vp, []),
}

_lib = None


def bind(lib):
    """Attach restype/argtypes for every declared symbol (raises AttributeError on a missing export)."""
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib


def load():
    """Load libvolpath_b200.so (built by __graft_entry__.build() / csrc/Makefile).  No fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "libvolpath_b200.so is not built (%s); run `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C cuda-volpath_b200/csrc`.  There is no CPU fallback." % LIB_PATH)
        _lib = bind(ctypes.CDLL(LIB_PATH))
    return _lib


class VolpathError(RuntimeError):
    pass


def check(rc):
    if rc != VP_OK:
        raise VolpathError("volpath error %d: %s" % (rc, load().vp_last_error().decode(errors="replace")))
