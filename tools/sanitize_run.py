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
r.set_julia()
r.render(vp.default_param(24, 16), 0, 2, mode=vp.MODE_FAST)
r.render(vp.default_param(24, 16), 0, 2, mode=vp.MODE_WAVE)
r.render(vp.default_param(24, 16), 0, 2, mode=vp.MODE_PARITY)
r.close()
print("sanitize_run ok")
