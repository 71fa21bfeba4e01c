#!/usr/bin/env python3
"""GPU diagnostic: fast vs parity image / scatter means at high spp for several storage types and densities."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

import cuda_volpath_b200 as vp  # noqa: E402
from oraclelib import Oracle  # noqa: E402

orc = Oracle()
env, sd, sp = vp.default_sunsky()
r = vp.Renderer(0)
spp = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
for dims, seed in (((64, 48, 80), 4),):
    vol = orc.fbm_cloud(*dims, seed=seed)
    for store in ("f32", "u8", "f16"):
        for density in (300.0, 3000.0):
            quant = store == "u8"
            v = np.round(vol * 255).astype(np.uint8) if quant else vol
            r.init_cuda(v, quant, store=vp.VOXEL_F16 if store == "f16" else None)
            r.set_texture_filter_mode(True)
            r.init_envmap(env)
            r.set_sun(sd, sp)
            r.copy_inv_view_matrix(vp.inv_view_matrix())
            P = vp.default_param(96, 64)
            P.density = density
            a = r.render(P, 0, spp, mode=vp.MODE_PARITY)
            b = r.render(P, spp, spp, mode=vp.MODE_PARITY)
            f = r.render(P, 0, spp, mode=vp.MODE_FAST)
            g = r.render(P, spp, spp, mode=vp.MODE_FAST)
            ma, mb, mf, mg = [x[..., :3].mean() for x in (a, b, f, g)]
            sa, sb, sf, sg = [x[..., 3].mean() for x in (a, b, f, g)]
            print("%s density %5.0f: radiance parity %.4f %.4f  fast %.4f %.4f  (fast/parity %.4f)   scatters parity %.3f %.3f fast %.3f %.3f (ratio %.4f)"
                  % (store, density, ma / spp, mb / spp, mf / spp, mg / spp, (mf + mg) / (ma + mb), sa / spp, sb / spp, sf / spp, sg / spp, (sf + sg) / (sa + sb)), flush=True)
