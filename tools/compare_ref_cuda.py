#!/usr/bin/env python3
"""GPU tool (own process): the reference's own CUDA kernel rebuilt for sm_100 (oracle/_ref/libvolpath_ref_cuda.so)
beside VP_MODE_FAST on the SAME scene, same GPU: path-samples/s of both, image agreement, RMSE against a high-spp
reference image at equal spp ("matched image error").  Prints one JSON line.

The reference layout needs 12 B/voxel of cudaArrays (density + float2 bounds) plus a 4 B/voxel opacity array, and its
CPU bound sweep indexes with int, so the shared scene is the C2 cloud family at reduced dims (default 1/4: 497x338x612);
the reference gets our device volume and our (bit-identical, tests/test_gpu_parity.py) bound volume."""
import argparse
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dims", type=int, nargs=3, default=[497, 338, 612])
    ap.add_argument("--image", type=int, nargs=2, default=[1920, 1080])
    ap.add_argument("--frames", type=int, default=16)
    ap.add_argument("--truth-frames", type=int, default=0, help="extra reference frames for an RMSE ground truth")
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--mode", default="fast", choices=["fast", "wave", "parity"])
    ap.add_argument("--density", type=float, default=800.0)
    ap.add_argument("--albedo", type=float, default=1.0)
    ap.add_argument("--material", type=int, default=-1, help="index into the reference's Mat() table")
    ap.add_argument("--exact-bounds", action="store_true", help="VP_BOUNDS_EXACT: the fast renderer uses the reference's per-voxel windows")
    ap.add_argument("--julia", action="store_true", help="config C1: the no-OpenVDB build (procedural Julia set)")
    ap.add_argument("--rmse-truth-spp", type=int, default=0,
                    help="time-to-RMSE leg: spp of the reference-kernel ground truth (0 = skip); BASELINE.json metric, second half")
    ap.add_argument("--rmse-image", type=int, nargs=2, default=[1920, 1080])
    args = ap.parse_args()
    # the reference printf()s to stdout
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    import faulthandler

    faulthandler.enable()
    import numpy as np
    import torch

    import cuda_volpath_b200 as vp
    from oraclelib import RefCuda

    MODE = {"wave": vp.MODE_WAVE, "parity": vp.MODE_PARITY}.get(args.mode, vp.MODE_FAST)
    torch.cuda.set_device(args.device)
    nx, ny, nz = args.dims
    W, H = args.image
    env, sd, sp = vp.default_sunsky()
    view = vp.inv_view_matrix()
    r = vp.Renderer(args.device)
    if args.julia:
        r.set_julia()
        ref = RefCuda(julia=True)
        ref.set_julia()
    else:
        r.generate_cloud(nx, ny, nz, seed=0, store=vp.VOXEL_F32, bounds=vp.BOUNDS_CELL | vp.BOUNDS_VOXEL | (vp.BOUNDS_EXACT if args.exact_bounds else 0), keep_dense=True)
    r.set_texture_filter_mode(True)
    r.init_envmap(env)
    r.set_sun(sd, sp)
    r.copy_inv_view_matrix(view)
    r.precompute_opacity(sd)
    if not args.julia:
        bv = torch.from_numpy(r.bounds_voxel()).cuda()
        ref = RefCuda()
        rc = ref.L.ref_init_volume_device(r.dense_volume_ptr(), bv.data_ptr(), nx, ny, nz, 0, None, None, 1)
        assert rc == 0, rc
        del bv
        ref.dims, ref.quantized = (nx, ny, nz), False
    ref.set_envmap(env)
    ref.set_sun(sd, sp)
    ref.set_inv_view(view)
    ref.precompute_opacity(sd)
    P = vp.default_param(W, H)
    P.density = args.density
    P.albedo[:] = [args.albedo] * 3
    if args.material >= 0:
        P = vp.mat(P, *vp.MATERIALS[args.material])
    pa = ctypes.addressof(P)
    stream = torch.cuda.current_stream().cuda_stream
    warm = 12  # frames 0..11 warm both kernels up; the timed frames are all > 10 (opacity-table regime)
    acc_r = torch.zeros(H, W, 4, device="cuda")
    acc_o = torch.zeros(H, W, 4, device="cuda")
    assert ref.L.ref_render_timed(acc_r.data_ptr(), 0, warm, pa) > 0
    r.render_kernel(acc_o.data_ptr(), 0, P, mode=MODE, n_frames=warm, stream=stream)
    acc_r.zero_()
    acc_o.zero_()
    ms_ref = ref.L.ref_render_timed(acc_r.data_ptr(), warm, args.frames, pa)
    assert ms_ref > 0, ms_ref
    r.render_kernel(acc_o.data_ptr(), warm, P, mode=MODE, n_frames=args.frames, stream=stream)
    ms_ours = r.last_kernel_ms()
    n = W * H * args.frames
    a, b = acc_r.cpu().numpy() / args.frames, acc_o.cpu().numpy() / args.frames
    scene = "C1 Julia set (no-OpenVDB build)" if args.julia else "C2 cloud family %dx%dx%d fp32" % (nx, ny, nz)
    out = {"workload": "%s, %dx%d, frames %d..%d, density %g, albedo %g, material %d"
                       % (scene, W, H, warm, warm + args.frames - 1, args.density, args.albedo, args.material),
           "reference_kernel_path_samples_per_s": n / (ms_ref * 1e-3), "ours_path_samples_per_s": n / (ms_ours * 1e-3),
           "speedup": ms_ref / ms_ours, "ms_reference": ms_ref, "ms_ours": ms_ours,
           "image_mean_rel_diff": float(abs(a[..., :3].mean() - b[..., :3].mean()) / a[..., :3].mean()),
           "scatter_mean_reference": float(a[..., 3].mean()), "scatter_mean_ours": float(b[..., 3].mean())}
    if args.truth_frames > 0:
        t = torch.zeros(H, W, 4, device="cuda")
        first = warm + args.frames
        ref.L.ref_render_timed(t.data_ptr(), first, args.truth_frames, pa)
        truth = t.cpu().numpy() / args.truth_frames
        out["truth_frames"] = args.truth_frames
        out["rmse_reference_vs_truth"] = float(np.sqrt(np.mean((a[..., :3] - truth[..., :3]) ** 2)))
        out["rmse_ours_vs_truth"] = float(np.sqrt(np.mean((b[..., :3] - truth[..., :3]) ** 2)))
        out["mean_rel_ours_vs_truth"] = float(abs(b[..., :3].mean() - truth[..., :3].mean()) / truth[..., :3].mean())
        out["mean_rel_reference_vs_truth"] = float(abs(a[..., :3].mean() - truth[..., :3].mean()) / truth[..., :3].mean())
    if args.rmse_truth_spp > 0:
        out["rmse"] = time_to_rmse(args, vp, r, ref, MODE, P)
    r.close()
    sys.stdout.flush()
    os.dup2(saved, 1)
    print(json.dumps(out), flush=True)
    # two CUDA runtimes unload in one process: skip their destructors, but run the interpreter's exit hooks first (the
    # driver's loaded-library recorder is one of them), so this leg's libvolpath_ref_cuda.so shows up in its evidence
    import atexit

    try:
        atexit._run_exitfuncs()
    finally:
        os._exit(0)


def time_to_rmse(args, vp, r, ref, MODE, P0):
    """Time to a fixed per-pixel RMSE for both kernels (BASELINE.json metric: "time-to-RMSE vs reference kernel").
    Truth = the reference kernel's own image at --rmse-truth-spp on a disjoint frame range.  Each checkpoint is ONE batch
    of spp frames per kernel (device-timed); RMSE over rgb of image/spp - truth.  Target = the reference kernel's RMSE at
    64 spp; spp-to-target by log-log interpolation between checkpoints, time-to-target from the measured ms per spp."""
    import ctypes

    import numpy as np
    import torch

    W, H = args.rmse_image
    P = P0.copy()
    P.width, P.height = W, H
    pa = ctypes.addressof(P)
    stream = torch.cuda.current_stream().cuda_stream
    t = torch.zeros(H, W, 4, device="cuda")
    first_truth = 100000
    ms_truth = ref.L.ref_render_timed(t.data_ptr(), first_truth, args.rmse_truth_spp, pa)
    truth = (t[..., :3] / args.rmse_truth_spp).double()
    cps = [16, 32, 64, 128, 256]
    res = {"reference": {"rmse": [], "ms": []}, "ours": {"rmse": [], "ms": []}}
    first = 12
    for spp in cps:
        a = torch.zeros(H, W, 4, device="cuda")
        ms = ref.L.ref_render_timed(a.data_ptr(), first, spp, pa)
        res["reference"]["ms"].append(float(ms))
        res["reference"]["rmse"].append(float(torch.sqrt((((a[..., :3] / spp).double() - truth) ** 2).mean()).item()))
        b = torch.zeros(H, W, 4, device="cuda")
        r.render_kernel(b.data_ptr(), first, P, mode=MODE, n_frames=spp, stream=stream)
        res["ours"]["ms"].append(float(r.last_kernel_ms()))
        res["ours"]["rmse"].append(float(torch.sqrt((((b[..., :3] / spp).double() - truth) ** 2).mean()).item()))
        first += spp
    target = res["reference"]["rmse"][cps.index(64)]

    def to_target(which):
        rm, ms = np.array(res[which]["rmse"]), np.array(res[which]["ms"])
        lx, ly = np.log(np.array(cps, float)), np.log(rm)
        # rmse falls monotonically with spp: interpolate log(spp) as a function of log(rmse) (extrapolate with slope -1/2)
        if target >= rm[0]:
            spp_t = cps[0] * (rm[0] / target) ** 2
        elif target <= rm[-1]:
            spp_t = cps[-1] * (rm[-1] / target) ** 2
        else:
            spp_t = float(np.exp(np.interp(-np.log(target), -ly, lx)))
        ms_t = float(np.interp(spp_t, cps, ms)) if spp_t <= cps[-1] else float(ms[-1] * spp_t / cps[-1])
        return float(spp_t), ms_t

    sr, tr = to_target("reference")
    so, to = to_target("ours")
    return {"image": [W, H], "truth": "reference CUDA kernel, %d spp, frames %d.." % (args.rmse_truth_spp, first_truth),
            "truth_spp": args.rmse_truth_spp, "truth_ms": float(ms_truth), "checkpoints_spp": cps,
            "rmse_reference": res["reference"]["rmse"], "rmse_ours": res["ours"]["rmse"],
            "ms_reference": res["reference"]["ms"], "ms_ours": res["ours"]["ms"],
            "truth_noise_rmse_estimate": float(res["reference"]["rmse"][-1] * (cps[-1] / args.rmse_truth_spp) ** 0.5),
            "target_rmse": target, "target": "the reference kernel's RMSE at 64 spp",
            "spp_to_target_reference": sr, "spp_to_target_ours": so,
            "time_to_rmse_reference_ms": tr, "time_to_rmse_ours_ms": to, "time_to_rmse_ratio": tr / to}


if __name__ == "__main__":
    main()
