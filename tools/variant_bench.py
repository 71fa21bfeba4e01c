#!/usr/bin/env python3
"""GPU tool: time VP_MODE_FAST launches of ONE library build (VOLPATH_B200_LIB selects a kernel-variant .so from
tools/build_variant.sh) without torch: device memory through the library's own helpers, CUDA-event times from
vp_last_kernel_ms.  One JSON line.
    VOLPATH_B200_LIB=gpurun_variants/lib_x.so python tools/variant_bench.py [--dims ..] [--image W H] [--frames F]
                     [--launches N] [--material M] [--albedo A] [--density D] [--mode fast|wave] [--sync-frames N]"""
import argparse
import ctypes
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cuda_volpath_b200 as vp  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dims", type=int, nargs=3, default=[1987, 1351, 2449])
    ap.add_argument("--image", type=int, nargs=2, default=[1920, 1080])
    ap.add_argument("--frames", type=int, default=64)
    ap.add_argument("--launches", type=int, default=4)
    ap.add_argument("--material", type=int, default=-1)
    ap.add_argument("--albedo", type=float, default=1.0)
    ap.add_argument("--density", type=float, default=800.0)
    ap.add_argument("--mode", default="fast")
    ap.add_argument("--julia", action="store_true")
    ap.add_argument("--sync-frames", type=int, default=0, help="also time N one-frame launches with a sync after each (the reference host's pattern)")
    ap.add_argument("--tag", default="")
    ap.add_argument("--store", default="f32", choices=["f32", "f16"])
    a = ap.parse_args()
    W, H = a.image
    r = vp.Renderer(0)
    env, sd, sp = vp.default_sunsky()
    t0 = time.perf_counter()
    if a.julia:
        r.set_julia()
    else:
        r.generate_cloud(*a.dims, seed=0, bounds=vp.BOUNDS_CELL, store=vp.VOXEL_F16 if a.store == "f16" else vp.VOXEL_F32)
    r.set_texture_filter_mode(True)
    r.init_envmap(env)
    r.set_sun(sd, sp)
    r.copy_inv_view_matrix(vp.inv_view_matrix())
    r.precompute_opacity(sd)
    r.sync()
    setup = time.perf_counter() - t0
    P = vp.default_param(W, H)
    P.density = a.density
    P.albedo[:] = [a.albedo] * 3
    if a.material >= 0:
        P = vp.mat(P, *vp.MATERIALS[a.material])
    mode = vp.MODE_WAVE if a.mode == "wave" else vp.MODE_FAST
    acc = r.L.vp_dev_alloc(W * H * 16)
    r.render_kernel(acc, 12, P, mode=mode, n_frames=8)  # warm-up
    r.sync()
    ms = []
    for i in range(a.launches):
        r.render_kernel(acc, 20 + i * a.frames, P, mode=mode, n_frames=a.frames)
        ms.append(r.last_kernel_ms())
    out = {"tag": a.tag or os.path.basename(os.environ.get("VOLPATH_B200_LIB", "HEAD")), "dims": a.dims, "image": [W, H], "frames": a.frames,
           "material": a.material, "ms": [round(m, 2) for m in ms], "Mps": [round(W * H * a.frames / m / 1e3, 1) for m in ms],
           "best_Mps": round(W * H * a.frames / min(ms) / 1e3, 1), "setup_s": round(setup, 2)}
    if a.sync_frames:
        r.sync()
        t0 = time.perf_counter()
        for f in range(a.sync_frames):
            r.render_kernel(acc, 1000 + f, P, mode=mode, n_frames=1)
            r.sync()
        dt = time.perf_counter() - t0
        out["sync_per_frame_Mps"] = round(W * H * a.sync_frames / dt / 1e6, 1)
    import numpy as np

    h = np.empty((H, W, 4), np.float32)
    r.L.vp_dev_to_host(h.ctypes.data, acc, h.nbytes)
    n = 8 + a.frames * a.launches + a.sync_frames
    out["mean_scatters"] = round(float(h[..., 3].sum() / (W * H * n)), 4)
    out["image_mean"] = round(float(h[..., :3].sum() / (3 * W * H * n)), 6)
    r.L.vp_dev_free(acc)
    r.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
