#!/usr/bin/env python3
"""GPU tool: the reference host's own calling pattern through the 14-name shims -- one render_kernel call per frame --
on the C2 cloud family, with and without a synchronisation after every frame.
    python tools/shim_frames.py [--dims nx ny nz] [--image W H] [--frames F]
VOLPATH_SHIM_OVERLAP=0 switches the overlap ring of the shim off (see vp_context::ring in csrc/volpath_api.cu)."""
import argparse
import ctypes
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dims", type=int, nargs=3, default=[497, 338, 612])
    ap.add_argument("--image", type=int, nargs=2, default=[1920, 1080])
    ap.add_argument("--frames", type=int, default=64)
    args = ap.parse_args()
    import torch

    import cuda_volpath_b200 as vp

    L = vp.lib.load()
    W, H = args.image
    r = vp.Renderer.__new__(vp.Renderer)  # a Renderer over the shims' implicit context
    r.L, r.device, r.dims = L, 0, None
    r.h = ctypes.c_void_p(L.vp_shim_context())
    env, sd, sp = vp.default_sunsky()
    r.generate_cloud(*args.dims, seed=0, bounds=vp.BOUNDS_CELL)
    r.set_texture_filter_mode(True)
    r.init_envmap(env)
    r.set_sun(sd, sp)
    r.copy_inv_view_matrix(vp.inv_view_matrix())
    r.precompute_opacity(sd)
    L.vp_shim_set_mode(vp.MODE_FAST)
    P = vp.default_param(W, H)
    acc = torch.zeros(H, W, 4, device="cuda")
    grid, block = vp.lib.Dim3((W + 7) // 8, (H + 7) // 8, 1), vp.lib.Dim3(8, 8, 1)
    L.render_kernel.argtypes = [vp.lib.Dim3, vp.lib.Dim3, ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(vp.Param)]
    L.render_kernel.restype = None
    for sync_each in (True, False):
        for rep in range(2):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for f in range(args.frames):
                L.render_kernel(grid, block, acc.data_ptr(), 16 + f, ctypes.byref(P))
                if sync_each:
                    torch.cuda.synchronize()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
        print("render_kernel x %d, %s: %.1f M path-samples/s" % (args.frames, "sync after every frame" if sync_each else "one sync at the end",
                                                                  W * H * args.frames / dt / 1e6))
    err = L.vp_last_error()
    if err:
        print("last error:", err.decode())


if __name__ == "__main__":
    main()
