#!/usr/bin/env python3
"""GPU tool: per-path-sample work of the REFERENCE algorithm (L density fetches, S segments, O
opacity-table fetches, E env evaluations) counted by the instrumented build of the reference's own CUDA kernel
(oracle/_ref/libvolpath_ref_cuda_instr.so: atomicAdd probes next to the fetch sites, nothing else changed).
Writes profiles/ref_counters.json, which bench.py turns into algorithmic bytes per path-sample.

Scene: the C2 cloud family at 1/4 dims (the reference layout cannot hold the full grid next to ours and its CPU bound
sweep indexes with int), reference camera / Param / sun / sky, 1920x1080, frames 16..23 (all > 10, like 99 % of the
frames of a 1024-spp render)."""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import cuda_volpath_b200 as vp  # noqa: E402
from oraclelib import RefCuda  # noqa: E402


def cloud(dims, W, H, first, n, material=-1, albedo=1.0, density=800.0, label="C2"):
    nx, ny, nz = dims
    env, sd, sp = vp.default_sunsky()
    view = vp.inv_view_matrix()
    r = vp.Renderer(0)
    r.generate_cloud(nx, ny, nz, seed=0, store=vp.VOXEL_F32, bounds=vp.BOUNDS_CELL | vp.BOUNDS_VOXEL, keep_dense=True)
    bv = torch.from_numpy(r.bounds_voxel()).cuda()
    ref = RefCuda(instrumented=True)
    assert ref.L.ref_init_volume_device(r.dense_volume_ptr(), bv.data_ptr(), nx, ny, nz, 0, None, None, 1) == 0
    ref.dims = dims
    ref.set_envmap(env)
    ref.set_sun(sd, sp)
    ref.set_inv_view(view)
    ref.precompute_opacity(sd)
    P = vp.default_param(W, H)
    P.density = density
    P.albedo[:] = [albedo] * 3
    if material >= 0:
        P = vp.mat(P, *vp.MATERIALS[material])
    acc = torch.zeros(H, W, 4, device="cuda")
    ref.reset_counters()
    assert ref.L.ref_render(acc.data_ptr(), first, n, ctypes.addressof(P)) == 0
    c = ref.counters().astype(np.float64) / (W * H * n)
    # ours, same scene, for the record
    r.set_texture_filter_mode(True)
    r.init_envmap(env)
    r.set_sun(sd, sp)
    r.copy_inv_view_matrix(view)
    r.precompute_opacity(sd)
    r.set_stats(True)
    r.counters(reset=True)
    r.render(P, first, n, mode=vp.MODE_FAST)
    o = r.counters()
    out = {"workload": "%s on the C2 cloud family %dx%dx%d fp32, %dx%d, frames %d..%d" % (label, nx, ny, nz, W, H, first, first + n - 1),
           "L": c[0] + c[1], "L_track": c[0], "L_shadow": c[1], "S": c[2], "O": c[3], "E": c[4], "scatters": c[5],
           "ours": {k: v / (W * H * n) for k, v in o.items()}}
    r.close()
    return out


def julia(W, H, n):
    env, sd, sp = vp.default_sunsky()
    ref = RefCuda(instrumented=True, julia=True)
    ref.set_julia()
    ref.set_envmap(env)
    ref.set_sun(sd, sp)
    ref.set_inv_view(vp.inv_view_matrix())
    P = vp.default_param(W, H)
    acc = torch.zeros(H, W, 4, device="cuda")
    ref.reset_counters()
    assert ref.L.ref_render(acc.data_ptr(), 0, n, ctypes.addressof(P)) == 0
    c = ref.counters().astype(np.float64) / (W * H * n)
    return {"workload": "C1 Julia set %dx%d, frames 0..%d" % (W, H, n - 1), "L": c[0] + c[1], "L_track": c[0],
            "L_shadow": c[1], "S": c[2], "O": c[3], "E": c[4], "scatters": c[5]}


if __name__ == "__main__":
    import faulthandler

    faulthandler.enable()
    q = (497, 338, 612)
    out = {"cloud": cloud(q, 1920, 1080, 16, 8), "c1": julia(512, 512, 8),
           "c3a": cloud(q, 1920, 1080, 16, 8, material=8, label="C3 (Mat preset 8)"),
           "c3b": cloud(q, 1920, 1080, 16, 8, material=4, label="C3 (Mat preset 4)"),
           "c4": cloud(q, 1920, 1080, 16, 8, albedo=0.999, density=3000.0, label="C4 (albedo 0.999, density 3000)")}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    for d in ("gpurun_out", "profiles"):
        json.dump(out, open(os.path.join(ROOT, d, "ref_counters.json"), "w"), indent=1)
    print(json.dumps(out), flush=True)
    os._exit(0)
