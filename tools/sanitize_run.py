#!/usr/bin/env python3
"""GPU tool: a tiny run through every kernel of the library (build, opacity, parity / fast / wave renders, MIS variant,
resolve) -- the target of `compute-sanitizer --tool memcheck python tools/sanitize_run.py`."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import cuda_volpath_b200 as vp  # noqa: E402

r = vp.Renderer(0)
env, sd, sp = vp.default_sunsky()
env = np.ascontiguousarray(env[::8, ::8])
for store, quant in ((vp.VOXEL_F32, False), (vp.VOXEL_F16, False), (vp.VOXEL_U8, True)):
    r.generate_cloud(41, 27, 50, seed=3, store=vp.VOXEL_F32, bounds=vp.BOUNDS_CELL | vp.BOUNDS_VOXEL, keep_dense=True)
    vol = r.dense_volume()
    v = np.round(vol * 255).astype(np.uint8) if quant else vol
    r.init_cuda(v, quant, store=store)
    for linear in (False, True):
        r.set_texture_filter_mode(linear)
        r.init_envmap(env)
        r.set_sun(sd, sp)
        r.copy_inv_view_matrix(vp.inv_view_matrix())
        r.precompute_opacity(sd)
        P = vp.default_param(37, 21)
        P.density = 2000.0
        for mode in (vp.MODE_PARITY, vp.MODE_FAST, vp.MODE_WAVE):
            img = r.render(P, 9, 4, mode=mode)
            assert np.isfinite(img).all()
        Pc = vp.mat(P, *vp.MATERIALS[4])
        r.render(Pc, 9, 4, mode=vp.MODE_FAST)
        r.set_env_sampling(True)
        r.render(P, 9, 3, mode=vp.MODE_PARITY)
        r.render(P, 9, 3, mode=vp.MODE_FAST)
        r.set_env_sampling(False)
        r.opacity()
        r.bounds_cell()
# coarse bound cells + half-precision per-cell tables (what large volumes use), ragged row lengths for the staged sweep
os.environ["VOLPATH_FORCE_CELL_LOG2"] = "2"
os.environ["VOLPATH_HALF_TABLES"] = "1"
r.generate_cloud(203, 19, 23, seed=5, bounds=vp.BOUNDS_CELL | vp.BOUNDS_VOXEL)
r.set_texture_filter_mode(True)
r.set_sun(sd, sp)
r.precompute_opacity(sd)
assert r.half_tables() is not None
P = vp.default_param(37, 21)
for mode in (vp.MODE_FAST, vp.MODE_WAVE):
    assert np.isfinite(r.render(P, 12, 3, mode=mode)).all()
del os.environ["VOLPATH_FORCE_CELL_LOG2"], os.environ["VOLPATH_HALF_TABLES"]
# sun/sky bake on the device
r.bake_sunsky(vp.default_sky_state(), 64, 32)
assert np.isfinite(r.envmap()).all()
r.render(P, 0, 2, mode=vp.MODE_FAST)
# the render_kernel shim's overlap ring: consecutive one-frame launches on four internal streams
import ctypes  # noqa: E402
import torch  # noqa: E402

L = r.L
s = vp.Renderer.__new__(vp.Renderer)
s.L, s.device, s.dims, s.h = L, 0, None, ctypes.c_void_p(L.vp_shim_context())
s.generate_cloud(41, 27, 50, seed=3, bounds=vp.BOUNDS_CELL)
s.set_texture_filter_mode(True)
s.init_envmap(env)
s.set_sun(sd, sp)
s.copy_inv_view_matrix(vp.inv_view_matrix())
L.vp_shim_set_mode(vp.MODE_FAST)
acc = torch.zeros(21, 37, 4, device="cuda")
L.render_kernel.argtypes = [vp.lib.Dim3, vp.lib.Dim3, ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(vp.Param)]
L.render_kernel.restype = None
for f in range(9):
    L.render_kernel(vp.lib.Dim3(5, 3, 1), vp.lib.Dim3(8, 8, 1), acc.data_ptr(), f, ctypes.byref(P))
torch.cuda.synchronize()
assert torch.isfinite(acc).all()
r.set_julia()
r.render(vp.default_param(24, 16), 0, 2, mode=vp.MODE_FAST)
r.render(vp.default_param(24, 16), 0, 2, mode=vp.MODE_WAVE)
r.render(vp.default_param(24, 16), 0, 2, mode=vp.MODE_PARITY)
r.close()
print("sanitize_run ok")
