#!/usr/bin/env python3
"""GPU tool: a short, fixed run of the hot path for ncu (one scene build, a few launches).
    python tools/profile_run.py [--dims nx ny nz] [--image W H] [--frames F] [--launches N] [--which fast|parity|ref]"""
import argparse
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dims", type=int, nargs=3, default=[497, 338, 612])
    ap.add_argument("--image", type=int, nargs=2, default=[960, 540])
    ap.add_argument("--frames", type=int, default=8)
    ap.add_argument("--first", type=int, default=16)
    ap.add_argument("--launches", type=int, default=3)
    ap.add_argument("--which", default="fast", choices=["fast", "wave", "parity", "ref", "julia"])
    ap.add_argument("--store", default="f32")
    args = ap.parse_args()
    import torch

    import cuda_volpath_b200 as vp

    nx, ny, nz = args.dims
    W, H = args.image
    env, sd, sp = vp.default_sunsky()
    view = vp.inv_view_matrix()
    r = vp.Renderer(0)
    P = vp.default_param(W, H)
    acc = torch.zeros(H, W, 4, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    if args.which == "julia":
        r.set_julia()
    else:
        need_voxel = args.which in ("parity", "ref")
        r.generate_cloud(nx, ny, nz, seed=0, store=vp.VOXEL_F32 if args.store == "f32" else vp.VOXEL_F16,
                         bounds=vp.BOUNDS_CELL | (vp.BOUNDS_VOXEL if need_voxel else 0), keep_dense=args.which == "ref")
    r.set_texture_filter_mode(True)
    r.init_envmap(env)
    r.set_sun(sd, sp)
    r.copy_inv_view_matrix(view)
    r.precompute_opacity(sd)
    if args.which == "ref":
        from oraclelib import RefCuda

        os.dup2(2, 1)
        bv = torch.from_numpy(r.bounds_voxel()).cuda()
        ref = RefCuda()
        assert ref.L.ref_init_volume_device(r.dense_volume_ptr(), bv.data_ptr(), nx, ny, nz, 0, None, None, 1) == 0
        ref.dims = (nx, ny, nz)
        ref.set_envmap(env)
        ref.set_sun(sd, sp)
        ref.set_inv_view(view)
        ref.precompute_opacity(sd)
        for i in range(args.launches):
            ms = ref.L.ref_render_timed(acc.data_ptr(), args.first + i * args.frames, args.frames, ctypes.addressof(P))
            print("ref launch batch %d: %.3f ms, %.1f M path-samples/s" % (i, ms, W * H * args.frames / ms / 1e3), file=sys.stderr)
        os._exit(0)
    mode = {"parity": vp.MODE_PARITY, "wave": vp.MODE_WAVE}.get(args.which, vp.MODE_FAST)
    for i in range(args.launches):
        r.render_kernel(acc.data_ptr(), args.first + i * args.frames, P, mode=mode, n_frames=args.frames, stream=stream)
        ms = r.last_kernel_ms()
        print("launch %d: %.3f ms, %.1f M path-samples/s" % (i, ms, W * H * args.frames / ms / 1e3), file=sys.stderr)
    torch.cuda.synchronize()
    r.close()


if __name__ == "__main__":
    main()
