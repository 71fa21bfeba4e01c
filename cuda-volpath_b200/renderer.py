"""Host-side mirror of the reference's render interface for the hot path.

`Renderer` keeps the call names and argument meaning of the reference's `extern "C"` surface
(src/volumeRender.cpp:117-128, 347-356: init_cuda, set_texture_filter_mode, init_envmap, set_sun,
copy_inv_view_matrix, precompute_opacity, render_kernel, scale, gamma_correct) on top of one vp_context.
Device memory for the float4 accumulator is a torch CUDA tensor (plumbing only); every computation is a
kernel of libvolpath_b200.so reached through the C ABI."""
import ctypes

import numpy as np

from . import lib as _l
from .param import Param


def _fp(a):
    return a.ctypes.data_as(_l.c_fp)


class Renderer:
    def __init__(self, device=0):
        self.L = _l.load()
        self.device = device
        h = _l.c_vp()
        _l.check(self.L.vp_create(device, ctypes.byref(h)))
        self.h = h
        self.dims = None

    def close(self):
        if getattr(self, "h", None):
            self.L.vp_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- scene (reference names) ---------------------------------------------------------------------
    def init_cuda(self, volume, quantized, box=None, store=None, bounds=_l.BOUNDS_VOXEL | _l.BOUNDS_CELL):
        """init_cuda (K.cu:354): dense host volume [nz, ny, nx] (x fastest), uint8 if quantized else float32."""
        vol = np.ascontiguousarray(volume)
        if vol.dtype != (np.uint8 if quantized else np.float32):
            raise TypeError("quantized volumes are uint8, others float32")
        nz, ny, nx = vol.shape
        src = _l.VOXEL_U8 if quantized else _l.VOXEL_F32
        store = src if store is None else store
        lo = hi = None
        if box is not None:
            self._lo, self._hi = np.asarray(box[0], np.float32), np.asarray(box[1], np.float32)
            lo, hi = _fp(self._lo), _fp(self._hi)
        _l.check(self.L.vp_upload_volume(self.h, vol.ctypes.data, nx, ny, nz, src, store, _l.MEM_HOST, lo, hi, bounds))
        self.dims = (nx, ny, nz)

    def generate_cloud(self, nx, ny, nz, seed=0, store=_l.VOXEL_F32, bounds=_l.BOUNDS_CELL, keep_dense=False, box=None):
        lo = hi = None
        if box is not None:
            self._lo, self._hi = np.asarray(box[0], np.float32), np.asarray(box[1], np.float32)
            lo, hi = _fp(self._lo), _fp(self._hi)
        _l.check(self.L.vp_generate_cloud(self.h, nx, ny, nz, seed, store, lo, hi, bounds, int(keep_dense)))
        self.dims = (nx, ny, nz)

    def set_julia(self):
        _l.check(self.L.vp_set_julia(self.h))
        self.dims = None

    def set_texture_filter_mode(self, linear):
        _l.check(self.L.vp_set_filter(self.h, int(bool(linear))))

    def init_envmap(self, env):
        env = np.ascontiguousarray(env, np.float32)
        assert env.ndim == 3 and env.shape[2] == 4
        _l.check(self.L.vp_set_envmap(self.h, _fp(env), env.shape[1], env.shape[0]))

    def set_sun(self, sun_dir, sun_power):
        d, p = np.ascontiguousarray(sun_dir, np.float32), np.ascontiguousarray(sun_power, np.float32)
        _l.check(self.L.vp_set_sun(self.h, _fp(d), _fp(p)))

    def set_env_sampling(self, enable):
        """The reference's PASSIVE_ENVMAP switch (K.cu:21): True = env-map importance sampling + one-sample MIS."""
        _l.check(self.L.vp_set_env_sampling(self.h, int(bool(enable))))

    def copy_inv_view_matrix(self, m12):
        m = np.ascontiguousarray(m12, np.float32)
        assert m.size == 12
        _l.check(self.L.vp_set_inv_view(self.h, _fp(m)))

    def precompute_opacity(self, sun_dir, sharded=False):
        """precompute_opacity (K.cu:526); sharded=True (after nccl_init, collective): every rank builds 1/G of the
        production table and an all-gather over NVLink completes it."""
        d = np.ascontiguousarray(sun_dir, np.float32)
        fn = self.L.vp_precompute_opacity_sharded if sharded else self.L.vp_precompute_opacity
        _l.check(fn(self.h, _fp(d)))

    def free_cuda_buffers(self):
        _l.check(self.L.vp_free_volume(self.h))

    # ---- render ----------------------------------------------------------------------------------------
    def render_kernel(self, d_sum_ptr, spp, param, mode=_l.MODE_PARITY, n_frames=1, frame_stride=1, stream=None):
        """render_kernel (K.cu:2364): add frame(s) into the device float4[W*H] sum at d_sum_ptr."""
        _l.check(self.L.vp_render(self.h, d_sum_ptr, spp, n_frames, frame_stride, ctypes.byref(param), mode, stream))

    def render(self, param, first_frame, n_frames, mode=_l.MODE_FAST, frame_stride=1, accum=None):
        """Host-buffer form: returns the float32 [H, W, 4] sum after adding the frames."""
        if accum is None:
            accum = np.zeros((param.height, param.width, 4), np.float32)
        assert accum.dtype == np.float32 and accum.flags["C_CONTIGUOUS"]
        _l.check(self.L.vp_render_to_host(self.h, accum.ctypes.data, first_frame, n_frames, frame_stride,
                                          ctypes.byref(param), mode))
        return accum

    def scale(self, dst_ptr, src_ptr, size, scale, stream=None):
        _l.check(self.L.vp_resolve(self.h, dst_ptr, src_ptr, size, scale, 0.0, stream))

    def gamma_correct(self, dst_ptr, src_ptr, size, scale, gamma, stream=None):
        _l.check(self.L.vp_resolve(self.h, dst_ptr, src_ptr, size, scale, gamma, stream))

    def sync(self):
        _l.check(self.L.vp_sync(self.h))

    # ---- multi-GPU: the library's own NCCL reduce (include/volpath.h, "combining the per-GPU sums") -------------
    def nccl_unique_id(self):
        """128 opaque bytes from ncclGetUniqueId: created on one rank, distributed by the host's own means."""
        buf = ctypes.create_string_buffer(128)
        _l.check(self.L.vp_nccl_unique_id(buf))
        return buf.raw

    def nccl_init(self, n_ranks, rank, unique_id):
        assert len(unique_id) == 128
        _l.check(self.L.vp_nccl_init(self.h, n_ranks, rank, unique_id))

    def reduce_nccl(self, d_send_ptr, d_recv_ptr, size, root=0, stream=None):
        """ncclReduce(sum) of `size` float4 onto rank `root` (d_recv_ptr may equal d_send_ptr; None off-root)."""
        _l.check(self.L.vp_reduce_nccl(self.h, d_send_ptr, d_recv_ptr, size, root, stream))

    def nccl_destroy(self):
        _l.check(self.L.vp_nccl_destroy(self.h))

    # ---- the same reduce over peer memory (CUDA IPC, one node): no communicator, no bootstrap ---------------------
    def dev_alloc(self, nbytes):
        """A plain cudaMalloc block on this context's device, zero-filled (IPC handles need base pointers)."""
        import torch  # only to make this context's device current for the allocation

        with torch.cuda.device(self.device):
            p = self.L.vp_dev_alloc(nbytes)
        if not p:
            raise _l.VolpathError("out of device memory")
        return p

    def ipc_export(self, d_base_ptr):
        buf = ctypes.create_string_buffer(64)
        _l.check(self.L.vp_ipc_export(self.h, d_base_ptr, buf))
        return buf.raw

    def reduce_ipc(self, d_sum_ptr, peer_handles, size, stream=None):
        """d_sum += every peer accumulator (list of 64-byte handles, rank order), read over NVLink through CUDA IPC."""
        blob = b"".join(peer_handles)
        assert len(blob) == 64 * len(peer_handles)
        _l.check(self.L.vp_reduce_ipc(self.h, d_sum_ptr, blob, len(peer_handles), size, stream))

    # ---- introspection -----------------------------------------------------------------------------------
    def bounds_voxel(self):
        nx, ny, nz = self.dims
        out = np.empty((nz, ny, nx, 2), np.float32)
        _l.check(self.L.vp_get_bounds_voxel(self.h, _fp(out)))
        return out

    def bounds_cell(self, raw_jumps=False):
        """(max, min) per bound cell of the fast renderer; raw_jumps=True keeps the encoded vacuum jump distances
        (vacuum cells hold -jump, in world units, in the max field)."""
        d = (ctypes.c_int * 3)()
        _l.check(self.L.vp_get_bounds_cell(self.h, None, d))
        out = np.empty((d[2], d[1], d[0], 2), np.float32)
        if raw_jumps:
            d[0] = -1
        _l.check(self.L.vp_get_bounds_cell(self.h, _fp(out), d))
        return out

    def bake_sunsky(self, state, width=1024, height=512, sunsky_scale=0.02, ground_albedo=0.01, gamma_scale=1.0):
        """update_sunsky(baked = true)'s per-texel loop + init_envmap in one device kernel (volumeRender.cpp:296-325);
        `state`: dict as sunsky.default_sky_state() / tests/golden/sunsky_states.npz (configs, radiances, ecf_sky,
        lambdas, weights, sun_dir, sun_power).  The sun itself still goes through set_sun."""
        from .sunsky import ground_radiance

        st = _l.SkyState()
        cfg = np.asarray(state["configs"], np.float64).reshape(11, 9)
        for w in range(11):
            for k in range(9):
                st.configs[w][k] = cfg[w, k]
            st.radiances[w] = float(state["radiances"][w])
            st.emission_correction_factor_sky[w] = float(state["ecf_sky"][w])
        for i in range(7):
            st.lambdas[i] = float(state["lambdas"][i])
            for ch in range(3):
                st.weights[i][ch] = float(state["weights"][i][ch])
        st.gamma_scale = gamma_scale
        g = ground_radiance(state["sun_dir"], state["sun_power"], ground_albedo)
        for ch in range(3):
            st.sun_dir[ch] = float(state["sun_dir"][ch])
            st.ground_rgb[ch] = float(g[ch])
        st.sunsky_scale = sunsky_scale
        _l.check(self.L.vp_bake_sunsky(self.h, ctypes.byref(st), width, height))

    def envmap(self):
        d = (ctypes.c_int * 2)()
        _l.check(self.L.vp_get_envmap(self.h, None, d))
        out = np.empty((d[1], d[0], 4), np.float32)
        _l.check(self.L.vp_get_envmap(self.h, _fp(out), d))
        return out

    def half_tables(self):
        """The half-precision per-cell tables of the production renderers (large volumes): ((max, min) as float16
        [cz][cy][cx][2] with the vacuum jumps still encoded, sun-clear float16 [cz][cy][cx], sun-clear float32), or None
        while the float tables are in use."""
        d = (ctypes.c_int * 3)()
        _l.check(self.L.vp_get_bounds_cell(self.h, None, d))
        mm = np.empty((d[2], d[1], d[0], 2), np.float16)
        cl = np.empty((d[2], d[1], d[0]), np.float16)
        cf = np.empty((d[2], d[1], d[0]), np.float32)
        present = ctypes.c_int(0)
        _l.check(self.L.vp_get_half_tables(self.h, mm.ctypes.data_as(ctypes.c_void_p), cl.ctypes.data_as(ctypes.c_void_p), _fp(cf),
                                           ctypes.byref(present)))
        return (mm, cl, cf) if present.value else None

    def opacity(self):
        nx, ny, nz = self.dims
        out = np.empty((nz, ny, nx), np.float32)
        _l.check(self.L.vp_get_opacity(self.h, _fp(out)))
        return out

    def opacity_fast(self):
        """The production renderers' table (swept build, fp16 octets) as a dense [nz, ny, nx] array."""
        nx, ny, nz = self.dims
        out = np.empty((nz, ny, nx), np.float32)
        _l.check(self.L.vp_get_opacity_fast(self.h, _fp(out)))
        return out

    def opacity_build_ms(self):
        ms = ctypes.c_float()
        _l.check(self.L.vp_opacity_build_ms(self.h, ctypes.byref(ms)))
        return float(ms.value)

    def fetch_density(self, pos, parity=True):
        pos = np.ascontiguousarray(pos, np.float32).reshape(-1, 3)
        out = np.empty(len(pos), np.float32)
        _l.check(self.L.vp_fetch_density(self.h, _fp(pos), len(pos), int(parity), _fp(out)))
        return out

    def dense_volume(self):
        """The device fp32 dense copy kept by generate_cloud(keep_dense=True), as a host array."""
        nx, ny, nz = self.dims
        p = self.L.vp_dense_volume(self.h)
        if not p:
            raise _l.VolpathError("no dense copy kept")
        out = np.empty((nz, ny, nx), np.float32)
        assert self.L.vp_dev_to_host(out.ctypes.data, p, out.nbytes) == 0
        return out

    def dense_volume_ptr(self):
        return self.L.vp_dense_volume(self.h)

    def volume_stats(self):
        s = (ctypes.c_ulonglong * 8)()
        _l.check(self.L.vp_volume_stats(self.h, s))
        keys = ["bricks", "nonempty_bricks", "octet_bytes", "bound_radius_voxels", "bounds_cell_bytes",
                "bounds_voxel_bytes", "opacity_bytes", "bound_cell_voxels"]
        return dict(zip(keys, [int(v) for v in s]))

    def rng_sequence(self, x, y, frame, n):
        f = np.empty(n, np.float32)
        u = np.empty(n, np.uint32)
        _l.check(self.L.vp_rng_sequence(self.h, x, y, frame, n, _fp(f), u.ctypes.data_as(ctypes.POINTER(ctypes.c_uint))))
        return f, u

    def philox2x32(self, c0, c1, key):
        o = (ctypes.c_uint * 2)()
        _l.check(self.L.vp_philox2x32(self.h, c0, c1, key, o))
        return int(o[0]), int(o[1])

    def set_stats(self, on):
        _l.check(self.L.vp_set_stats(self.h, int(on)))

    def counters(self, reset=True):
        s = (ctypes.c_ulonglong * 16)()
        _l.check(self.L.vp_render_counters(self.h, s, int(reset)))
        keys = ["track_fetches", "shadow_fetches", "segments", "opacity_fetches", "env_evals", "scatters"]
        d = dict(zip(keys, [int(v) for v in s][:6]))
        d["zero_track_fetches"], d["zero_shadow_fetches"] = int(s[6]), int(s[7])
        # binning efficiency of the megakernel: block executions per warp and active lanes, per block type
        for i, name in enumerate(["path", "scatter", "segment", "step"]):
            d["blocks_" + name], d["lanes_" + name] = int(s[8 + i]), int(s[12 + i])
        return d

    def last_kernel_ms(self):
        ms = ctypes.c_float()
        _l.check(self.L.vp_last_kernel_ms(self.h, ctypes.byref(ms)))
        return float(ms.value)

    def launch_count(self):
        n = ctypes.c_ulonglong()
        _l.check(self.L.vp_launch_count(self.h, ctypes.byref(n)))
        return int(n.value)
