"""ctypes bindings of the TEST oracles (never imported by the product package):
  Oracle   -- oracle/libvolpath_oracle.so, the CPU restatement
  RefHost  -- oracle/_ref/libvolpath_ref_host*.so, the reference kernel source compiled by g++
  RefCuda  -- oracle/_ref/libvolpath_ref_cuda*.so, the reference kernel rebuilt for sm_100
All three expose the same small surface: set_volume / set_filter / set_envmap / set_sun / set_inv_view /
precompute_opacity / render(first_frame, n_frames, Param) -> float32 [H, W, 4] sum."""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
REF_DIR = os.path.join(ORACLE_DIR, "_ref")

c_fp = ctypes.POINTER(ctypes.c_float)
c_vp = ctypes.c_void_p


def _fp(a):
    return a.ctypes.data_as(c_fp)


def build_oracle():
    so = os.path.join(ORACLE_DIR, "libvolpath_oracle.so")
    src = [os.path.join(ORACLE_DIR, f) for f in ("volpath_oracle.cpp", "tex_emul.h")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "libvolpath_oracle.so"], stdout=subprocess.DEVNULL)
    return so


def have_ref(name):
    return os.path.exists(os.path.join(REF_DIR, name))


class Oracle:
    def __init__(self):
        L = ctypes.CDLL(build_oracle())
        self.L = L
        L.vo_create.restype = c_vp
        L.vo_hash.restype = ctypes.c_uint32
        L.vo_hash.argtypes = [ctypes.c_uint32]
        L.vo_julia_density.restype = ctypes.c_float
        L.vo_julia_density.argtypes = [ctypes.c_float] * 3
        for fn in ("vo_destroy", "vo_set_julia"):
            getattr(L, fn).argtypes = [c_vp]
        L.vo_set_volume.argtypes = [c_vp, c_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_fp, c_fp]
        L.vo_set_filter.argtypes = [c_vp, ctypes.c_int]
        L.vo_set_envmap.argtypes = [c_vp, c_fp, ctypes.c_int, ctypes.c_int]
        L.vo_set_sun.argtypes = [c_vp, c_fp, c_fp]
        L.vo_set_env_sampling.argtypes = [c_vp, ctypes.c_int]
        L.vo_set_inv_view.argtypes = [c_vp, c_fp]
        L.vo_precompute_opacity.argtypes = [c_vp, c_fp]
        L.vo_get_bounds.argtypes = [c_vp, c_vp]
        L.vo_get_opacity.argtypes = [c_vp, c_fp]
        L.vo_render.argtypes = [c_vp, c_fp, ctypes.c_int, ctypes.c_int, c_vp, c_vp]
        L.vo_trace_path.argtypes = [c_vp, ctypes.c_uint, ctypes.c_uint, ctypes.c_int, c_vp, c_fp]
        L.vo_bounds_u8.argtypes = [c_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float, c_vp]
        L.vo_bounds_f32.argtypes = [c_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float, c_vp]
        L.vo_bounds_brute_u8.argtypes = [c_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_vp]
        L.vo_bounds_brute_f32.argtypes = [c_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_vp]
        L.vo_bound_radius_voxels.argtypes = [ctypes.c_int, ctypes.c_float]
        L.vo_fbm_cloud_f32.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_uint32, c_fp]
        L.vo_rng_sequence.argtypes = [ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_int, c_fp, c_vp]
        L.vo_philox4x32_10.argtypes = [c_vp, c_vp, c_vp]
        L.vo_scale.argtypes = [c_fp, c_fp, ctypes.c_int, ctypes.c_float]
        L.vo_gamma_correct.argtypes = [c_fp, c_fp, ctypes.c_int, ctypes.c_float, ctypes.c_float]
        self.h = L.vo_create()
        self.dims = None
        self.quantized = False

    def close(self):
        if self.h:
            self.L.vo_destroy(self.h)
            self.h = None

    # --- scene -----------------------------------------------------------------------------
    def set_volume(self, vol, quantized, box=None, linear=True):
        vol = np.ascontiguousarray(vol)
        nz, ny, nx = vol.shape
        assert vol.dtype == (np.uint8 if quantized else np.float32)
        self.dims, self.quantized = (nx, ny, nz), bool(quantized)
        bmin = bmax = None
        if box is not None:
            lo, hi = np.asarray(box[0], np.float32), np.asarray(box[1], np.float32)
            bmin, bmax = _fp(lo), _fp(hi)
        rc = self.L.vo_set_volume(self.h, vol.ctypes.data, nx, ny, nz, int(quantized), bmin, bmax)
        assert rc == 0
        self.L.vo_set_filter(self.h, int(linear))

    def set_julia(self):
        self.L.vo_set_julia(self.h)

    def set_filter(self, linear):
        self.L.vo_set_filter(self.h, int(linear))

    def set_envmap(self, env):
        env = np.ascontiguousarray(env, np.float32)
        self.L.vo_set_envmap(self.h, _fp(env), env.shape[1], env.shape[0])

    def set_sun(self, sun_dir, sun_power):
        d, p = np.ascontiguousarray(sun_dir, np.float32), np.ascontiguousarray(sun_power, np.float32)
        self.L.vo_set_sun(self.h, _fp(d), _fp(p))

    def set_env_sampling(self, enable):
        self.L.vo_set_env_sampling(self.h, int(enable))

    def set_inv_view(self, m12):
        m = np.ascontiguousarray(m12, np.float32)
        self.L.vo_set_inv_view(self.h, _fp(m))

    def precompute_opacity(self, sun_dir):
        d = np.ascontiguousarray(sun_dir, np.float32)
        self.L.vo_precompute_opacity(self.h, _fp(d))

    def bounds(self):
        nx, ny, nz = self.dims
        out = np.empty((nz, ny, nx, 2), np.uint8 if self.quantized else np.float32)
        self.L.vo_get_bounds(self.h, out.ctypes.data)
        return out

    def opacity(self):
        nx, ny, nz = self.dims
        out = np.empty((nz, ny, nx), np.float32)
        self.L.vo_get_opacity(self.h, _fp(out))
        return out

    def render(self, param, first_frame, n_frames, accum=None, stats=False):
        if accum is None:
            accum = np.zeros((param.height, param.width, 4), np.float32)
        st = np.zeros(8, np.uint64)
        self.L.vo_render(self.h, _fp(accum), first_frame, n_frames, ctypes.addressof(param),
                         st.ctypes.data if stats else None)
        return (accum, st) if stats else accum

    def trace_path(self, param, x, y, frame):
        out = np.zeros(4, np.float32)
        self.L.vo_trace_path(self.h, x, y, frame, ctypes.addressof(param), _fp(out))
        return out

    # --- standalone pieces -----------------------------------------------------------------
    def bounds_of(self, vol, radius=0.05, brute_D=None):
        vol = np.ascontiguousarray(vol)
        nz, ny, nx = vol.shape
        out = np.empty((nz, ny, nx, 2), vol.dtype)
        u8 = vol.dtype == np.uint8
        if brute_D is None:
            (self.L.vo_bounds_u8 if u8 else self.L.vo_bounds_f32)(vol.ctypes.data, nx, ny, nz, radius, out.ctypes.data)
        else:
            (self.L.vo_bounds_brute_u8 if u8 else self.L.vo_bounds_brute_f32)(
                vol.ctypes.data, nx, ny, nz, brute_D, out.ctypes.data)
        return out

    def fbm_cloud(self, nx, ny, nz, seed=0):
        out = np.empty((nz, ny, nx), np.float32)
        self.L.vo_fbm_cloud_f32(nx, ny, nz, seed, _fp(out))
        return out

    def rng_sequence(self, x, y, frame, n):
        f = np.empty(n, np.float32)
        u = np.empty(n, np.uint32)
        self.L.vo_rng_sequence(x, y, frame, n, _fp(f), u.ctypes.data)
        return f, u

    def philox(self, ctr, key):
        c = np.asarray(ctr, np.uint32)
        k = np.asarray(key, np.uint32)
        o = np.empty(4, np.uint32)
        self.L.vo_philox4x32_10(c.ctypes.data, k.ctypes.data, o.ctypes.data)
        return o


class _RefBase:
    """Common part of the two builds of the reference (same driver, oracle/ref_driver.inc)."""

    def _bind(self, L):
        self.L = L
        L.ref_init_volume_host.argtypes = [c_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_fp, c_fp,
                                           ctypes.c_int]
        L.ref_set_filter.argtypes = [ctypes.c_int]
        L.ref_set_envmap.argtypes = [c_fp, ctypes.c_int, ctypes.c_int]
        L.ref_set_sun.argtypes = [c_fp, c_fp]
        L.ref_set_inv_view.argtypes = [c_fp]
        L.ref_precompute_opacity.argtypes = [c_fp]
        L.ref_render.argtypes = [c_vp, ctypes.c_int, ctypes.c_int, c_vp]
        L.ref_read_bounds.argtypes = [c_vp]
        L.ref_read_opacity.argtypes = [c_fp]
        L.ref_counters_read.argtypes = [c_vp]
        L.ref_bounds_u8.argtypes = [c_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float, c_vp]
        L.ref_bounds_f32.argtypes = [c_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float, c_vp]
        self.dims = None
        self.quantized = False

    def set_volume(self, vol, quantized, box=None, linear=True):
        vol = np.ascontiguousarray(vol)
        nz, ny, nx = vol.shape
        self.dims, self.quantized = (nx, ny, nz), bool(quantized)
        bmin = bmax = None
        if box is not None:
            lo, hi = np.asarray(box[0], np.float32), np.asarray(box[1], np.float32)
            bmin, bmax = _fp(lo), _fp(hi)
        rc = self.L.ref_init_volume_host(vol.ctypes.data, nx, ny, nz, int(quantized), bmin, bmax, int(linear))
        assert rc == 0

    def set_julia(self):
        # the no-OpenVDB build still wants a volume: the reference's main() passes a 32^3 extent
        # (volumeRender.cpp:1346); a zero volume gives box [-1,1]^3 and a zero opacity table
        vol = np.zeros((32, 32, 32), np.float32)
        self.set_volume(vol, False, None, linear=True)

    def set_filter(self, linear):
        self.L.ref_set_filter(int(linear))

    def set_envmap(self, env):
        env = np.ascontiguousarray(env, np.float32)
        self.L.ref_set_envmap(_fp(env), env.shape[1], env.shape[0])

    def set_sun(self, sun_dir, sun_power):
        d, p = np.ascontiguousarray(sun_dir, np.float32), np.ascontiguousarray(sun_power, np.float32)
        self.L.ref_set_sun(_fp(d), _fp(p))

    def set_inv_view(self, m12):
        m = np.ascontiguousarray(m12, np.float32)
        self.L.ref_set_inv_view(_fp(m))

    def precompute_opacity(self, sun_dir):
        d = np.ascontiguousarray(sun_dir, np.float32)
        assert self.L.ref_precompute_opacity(_fp(d)) == 0

    def bounds(self):
        nx, ny, nz = self.dims
        out = np.empty((nz, ny, nx, 2), np.uint8 if self.quantized else np.float32)
        assert self.L.ref_read_bounds(out.ctypes.data) == 0
        return out

    def opacity(self):
        nx, ny, nz = self.dims
        out = np.empty((nz, ny, nx), np.float32)
        assert self.L.ref_read_opacity(_fp(out)) == 0
        return out

    def counters(self):
        c = np.zeros(8, np.uint64)
        self.L.ref_counters_read(c.ctypes.data)
        return c

    def reset_counters(self):
        self.L.ref_counters_reset()

    def bounds_of(self, vol, radius=0.05):
        vol = np.ascontiguousarray(vol)
        nz, ny, nx = vol.shape
        out = np.empty((nz, ny, nx, 2), vol.dtype)
        fn = self.L.ref_bounds_u8 if vol.dtype == np.uint8 else self.L.ref_bounds_f32
        fn(vol.ctypes.data, nx, ny, nz, radius, out.ctypes.data)
        return out


class RefHost(_RefBase):
    def __init__(self, instrumented=False, julia=False, mis=False):
        name = "libvolpath_ref_host%s%s%s.so" % ("_julia" if julia else "", "_mis" if mis else "", "_instr" if instrumented else "")
        self._bind(ctypes.CDLL(os.path.join(REF_DIR, name)))

    def render(self, param, first_frame, n_frames, accum=None):
        if accum is None:
            accum = np.zeros((param.height, param.width, 4), np.float32)
        assert self.L.ref_render(accum.ctypes.data, first_frame, n_frames, ctypes.addressof(param)) == 0
        return accum


class RefCuda(_RefBase):
    """The reference kernel rebuilt for sm_100 -- needs a GPU."""

    def __init__(self, instrumented=False, julia=False, mis=False):
        name = "libvolpath_ref_cuda%s%s%s.so" % ("_julia" if julia else "", "_mis" if mis else "", "_instr" if instrumented else "")
        L = ctypes.CDLL(os.path.join(REF_DIR, name))
        self._bind(L)
        L.ref_dev_alloc.restype = c_vp
        L.ref_dev_alloc.argtypes = [ctypes.c_size_t]
        L.ref_dev_free.argtypes = [c_vp]
        L.ref_dev_zero.argtypes = [c_vp, ctypes.c_size_t]
        L.ref_dev_to_host.argtypes = [c_vp, c_vp, ctypes.c_size_t]
        L.ref_host_to_dev.argtypes = [c_vp, c_vp, ctypes.c_size_t]
        L.ref_render_timed.restype = ctypes.c_float
        L.ref_render_timed.argtypes = [c_vp, ctypes.c_int, ctypes.c_int, c_vp]
        L.ref_init_volume_device.argtypes = [c_vp, c_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_fp,
                                             c_fp, ctypes.c_int]

    def render(self, param, first_frame, n_frames, accum=None):
        n = param.width * param.height * 16
        d = self.L.ref_dev_alloc(n)
        assert d
        if accum is not None:
            self.L.ref_host_to_dev(d, accum.ctypes.data, n)
        else:
            accum = np.zeros((param.height, param.width, 4), np.float32)
        rc = self.L.ref_render(d, first_frame, n_frames, ctypes.addressof(param))
        assert rc == 0, rc
        assert self.L.ref_dev_to_host(accum.ctypes.data, d, n) == 0
        self.L.ref_dev_free(d)
        return accum
