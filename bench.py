#!/usr/bin/env python3
"""bench.py -- path-samples/s of the render hot path (BASELINE.json metric), one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload c1|c2|c3a|c3b|c4|c5|c2q|c2e]

A "step" is one pass of the hot path over one batch: --frames-per-step frames (sample indices) of the whole image.

N = 1 (default workload c2 = config C2 of BASELINE.json: synthetic fBm cloud at WDAS dims 1987x1351x2449 fp32,
1920x1080, default sun/sky, camera and Param of the reference): one step = one launch of k_render_fast.

N > 1 (default workload c5 = config C5: the same volume at 3840x2160, STRONG scaling): the frames of a step are split
over the ranks by sample index (rank r renders frames r, r + N, ...; volume replicated), so the total work per step does
not depend on N.  Every rank renders into one of two accumulators; the library's own NCCL reduce (vp_reduce_nccl, C ABI)
runs on a side stream and overlaps the next step's render; the root adds the reduced step into the image.  All K reduces
are inside the timed region (the final event waits for the last one).

`value`   : device-timed (CUDA events on the launch stream), accumulators resident in HBM, max over ranks.
`e2e`     : the same steps through the host-buffer C-ABI call vp_render_to_host (pinned host float4 sum in and out).
`roofline`: algorithmic bytes per path-sample (SURVEY.md 8d formula, L/S/O/E counted by the instrumented reference
            kernel, profiles/ref_counters.json) x path-samples/s vs the measured HBM copy peak.
`check`   : the benchmarked kernel variant (rank directory + half tables + coarse bound cells on the full grid) against
            the reference-faithful renderer (VP_MODE_PARITY, per-voxel windows) at FULL dims on a 480x270 subset.
`cpu_baseline` / `--impl reference`: the reference's own kernel source compiled for the host (oracle/_ref, OpenMP over
            all cores) on a bounded sample of the same cloud family; `gpu_same_sample` = this repo's GPU arm on exactly
            that sample, so one ratio in the line is like for like.
`ref_cuda`: the reference's CUDA kernel rebuilt for sm_100 beside ours on the same GPU and scene, incl. time-to-RMSE.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

C2_DIMS = (1987, 1351, 2449)
WORKLOADS = {
    # name: (dims, image, material preset / overrides, description)
    "c2": (C2_DIMS, (1920, 1080), {}, "C2: synthetic fBm cloud 1987x1351x2449 fp32, 1920x1080, default sun/sky"),
    "c3a": (C2_DIMS, (1920, 1080), {"material": 8},
            "C3: chromatic medium Mat(0.74,0.88,1.01 / 0.032,0.17,0.48) on the C2 cloud, 1920x1080"),
    "c3b": (C2_DIMS, (1920, 1080), {"material": 4},
            "C3: strongly chromatic medium Mat(0.18,0.07,0.03 / 0.061,0.97,1.45) on the C2 cloud, 1920x1080"),
    "c4": (C2_DIMS, (1920, 1080), {"albedo": 0.999, "density": 3000.0},
           "C4: albedo 0.999, density 3000 (deep paths, opacity-table branch) on the C2 cloud, 1920x1080"),
    "c5": (C2_DIMS, (3840, 2160), {}, "C5: synthetic fBm cloud 1987x1351x2449 fp32, 3840x2160, sample-split"),
    "c2q": ((497, 338, 612), (1920, 1080), {}, "C2 cloud family at 1/4 dims 497x338x612 fp32, 1920x1080"),
    "c2e": ((248, 168, 306), (480, 270), {}, "C2 cloud family at 1/8 dims 248x168x306 fp32, 480x270"),
    "c1": (None, (512, 512), {}, "C1: procedural Julia set (no-OpenVDB build), 512x512"),
}
CLOUD_SEED = 0
CPU_SAMPLE_DIMS, CPU_SAMPLE_IMAGE = (248, 168, 306), (480, 270)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def algorithmic_bytes_per_path(workload, frames_per_rmw):
    """B = L*8*4 + S*8 + O*8*4 + E*16 + 32/F with L,S,O,E per path-sample from the instrumented
    REFERENCE kernel (profiles/ref_counters.json, written by tools/measure_ref_counters.py on a B200)."""
    p = os.path.join(ROOT, "profiles", "ref_counters.json")
    src = "survey probe (SURVEY.md 8d)"
    L, S, O, E = 118.0, 60.0, 5.6, 1.0
    if os.path.exists(p):
        j = json.load(open(p))
        key = workload if workload in j else ("c1" if workload == "c1" else "cloud")
        if key in j:
            c = j[key]
            L, S, O, E = c["L"], c["S"], c["O"], c["E"]
            src = "profiles/ref_counters.json (%s)" % c.get("workload", key)
    if workload == "c1":
        return S * 0 + E * 16 + 32.0 / frames_per_rmw, dict(L=L, S=S, O=O, E=E, source=src)
    B = L * 8 * 4 + S * 8 + O * 8 * 4 + E * 16 + 32.0 / frames_per_rmw
    return B, dict(L=L, S=S, O=O, E=E, source=src)


class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows if len(r) > 2 + i)]
        mx = [int(r[1]) for r in self.rows if r[1].isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx[0] if mx else None, "reasons": reasons,
                "samples": len(self.rows)}


def scene_inputs(vp):
    env, sun_dir, sun_power = vp.default_sunsky()
    return env, sun_dir, sun_power, vp.inv_view_matrix()


def workload_param(vp, W, H, over):
    P = vp.default_param(W, H)
    if "material" in over:
        P = vp.mat(P, *vp.MATERIALS[over["material"]])
    if "albedo" in over:
        P.albedo[:] = [over["albedo"]] * 3
    if "density" in over:
        P.density = over["density"]
    return P


def cpu_sample_desc(julia, fps):
    if julia:
        return "C1 Julia set (no-OpenVDB build), 256x256, frames 0..%d per step, host cores" % (fps - 1)
    return ("C2 cloud family at 1/8 dims (%dx%dx%d fp32, the dims of the reference's own wdas_cloud_eighth asset), %dx%d, "
            "frames 0..%d per step (<= 10: shadow walks, no opacity table), host cores"
            % (CPU_SAMPLE_DIMS + CPU_SAMPLE_IMAGE + (fps - 1,)))


def host_reference_run(args):
    """The reference's own kernel source compiled by g++ (oracle/_ref/libvolpath_ref_host*.so), OpenMP over all host
    cores, on a bounded sample: the C2 cloud family at 1/8 dims (248x168x306 fp32, the dims the reference's own asset
    wdas_cloud_eighth has), 480x270 -- the full 26 GB volume and its 52 GB CPU bound volume do not fit a host run."""
    julia = args.workload == "c1"
    name = "libvolpath_ref_host%s.so" % ("_julia" if julia else "")
    # the reference printf()s progress to stdout: keep fd 1 for the one JSON line
    sys.stdout.flush()
    saved_fd = os.dup(1)
    os.dup2(2, 1)
    try:
        return _host_reference_run(args, julia, name)
    finally:
        sys.stdout.flush()
        os.dup2(saved_fd, 1)
        os.close(saved_fd)


def _host_reference_run(args, julia, name):
    import numpy as np

    import cuda_volpath_b200 as vp
    from oraclelib import Oracle, RefHost, have_ref

    env, sun_dir, sun_power, view = scene_inputs(vp)
    if have_ref(name):
        ref, kind = RefHost(julia=julia), "reference"
    else:
        ref, kind = Oracle(), "port"
    over = WORKLOADS[args.workload][2]
    if julia:
        W, H = 256, 256
        ref.set_julia()
    else:
        W, H = CPU_SAMPLE_IMAGE
        vol = Oracle().fbm_cloud(*CPU_SAMPLE_DIMS, seed=CLOUD_SEED)
        ref.set_volume(vol, False, None, linear=True)
    ref.set_envmap(env)
    ref.set_sun(sun_dir, sun_power)
    ref.set_inv_view(view)
    P = workload_param(vp, W, H, over)
    cores = int(os.environ.get("OMP_NUM_THREADS", "0") or 0) or (os.cpu_count() or 1)
    acc = np.zeros((H, W, 4), np.float32)
    # frames 0..10 only: from frame 11 on the kernel reads the precomputed sun-opacity table (K.cu:2183), whose
    # construction (_precompute_opacity: ~10^10 emulated texture fetches at these dims) is not feasible on host cores
    fps = min(max(1, args.ref_frames), 11)
    for _ in range(args.warmup):
        ref.render(P, 0, 1, accum=acc)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ref.render(P, 0, fps, accum=acc)
    dt = time.perf_counter() - t0
    value = W * H * fps * args.steps / dt
    # the same loop on ONE host thread (north_star: "single-threaded plus OpenMP"), one frame
    single = None
    try:
        import ctypes as _c

        omp = _c.CDLL("libgomp.so.1")
        omp.omp_set_num_threads(1)
        t1 = time.perf_counter()
        ref.render(P, 0, 1, accum=acc)
        single = W * H / (time.perf_counter() - t1)
        omp.omp_set_num_threads(cores)
    except Exception:
        pass
    sample = cpu_sample_desc(julia, fps)
    return value, dt, sample, dict(value=value, unit="path-samples/s", cores=cores, kind=kind, sample=sample,
                                   single_thread_value=single, frames_per_step=fps,
                                   mean_scatters_per_path=float(acc[..., 3].sum() / (W * H * (fps * args.steps + args.warmup + (1 if single else 0)))))


def gpu_same_sample(vp, device, workload, fps, over):
    """This repo's GPU arm on EXACTLY the cpu_baseline sample (same volume, image, frames 0..fps-1, host buffers in and
    out through vp_render_to_host): the like-for-like ratio against the host build of the reference."""
    import numpy as np

    r = vp.Renderer(device)
    env, sun_dir, sun_power, view = scene_inputs(vp)
    if workload == "c1":
        W, H = 256, 256
        r.set_julia()
    else:
        W, H = CPU_SAMPLE_IMAGE
        r.generate_cloud(*CPU_SAMPLE_DIMS, seed=CLOUD_SEED, bounds=vp.BOUNDS_CELL)
    r.set_texture_filter_mode(True)
    r.init_envmap(env)
    r.set_sun(sun_dir, sun_power)
    r.copy_inv_view_matrix(view)
    P = workload_param(vp, W, H, over)
    acc = np.zeros((H, W, 4), np.float32)
    r.render(P, 0, fps, accum=acc)  # warm-up
    acc[:] = 0
    reps = 20
    t0 = time.perf_counter()
    for _ in range(reps):
        r.render(P, 0, fps, accum=acc)
    dt = time.perf_counter() - t0
    out = {"value": W * H * fps * reps / dt, "unit": "path-samples/s", "ms_per_call": 1e3 * dt / reps,
           "mean_scatters_per_path": float(acc[..., 3].sum() / (W * H * fps * reps)),
           "note": "vp_render_to_host, host float4 sum in and out; a %d-sample call is launch- and tail-bound on a B200" % (W * H * fps)}
    r.close()
    return out


def check_block(vp, device, dims, over, subset=(480, 270), frames=256):
    """the variant the bench times (k_render_fast on the FULL grid: rank directory, half-precision
    per-cell tables, bound cells of 8^3 voxels, swept fp16 opacity octets) against the reference-faithful renderer
    (k_render_parity: reference RNG, draw order and segmenting, the reference's own per-voxel +-50-voxel windows
    K.cu:1626-1661, the bit-faithful opacity table) on the same full-dims volume: 480x270, frames 12..12+frames-1."""
    import numpy as np

    r = vp.Renderer(device)
    t0 = time.perf_counter()
    env, sun_dir, sun_power, view = scene_inputs(vp)
    r.generate_cloud(*dims, seed=CLOUD_SEED, bounds=vp.BOUNDS_CELL | vp.BOUNDS_VOXEL)
    r.set_texture_filter_mode(True)
    r.init_envmap(env)
    r.set_sun(sun_dir, sun_power)
    r.copy_inv_view_matrix(view)
    r.precompute_opacity(sun_dir)
    r.sync()
    setup = time.perf_counter() - t0
    st = r.volume_stats()
    W, H = subset
    P = workload_param(vp, W, H, over)
    fast = r.render(P, 12, 4 * frames, mode=vp.MODE_FAST)      # 4x the samples: the fast estimate's noise is not the limit
    fast_half = r.render(P, 12, frames // 2, mode=vp.MODE_FAST)
    par_a = r.render(P, 12, frames // 2, mode=vp.MODE_PARITY)
    par_b = r.render(P, 12 + frames // 2, frames - frames // 2, mode=vp.MODE_PARITY)
    r.close()
    par = par_a + par_b
    n_f, n_p = W * H * 4 * frames, W * H * frames
    sf, sp = fast[..., 3].sum() / n_f, par[..., 3].sum() / n_p
    mf, mp = fast[..., :3].sum() / n_f, par[..., :3].sum() / n_p
    half = W * H * (frames // 2)
    # equal-spp per-pixel RMSE: fast vs parity against parity vs parity (two disjoint frame sets of frames/2 each); a ratio
    # above 1 means the fast estimate is noisier than the reference estimator's own Monte-Carlo noise
    fa = fast_half[..., :3] / (frames // 2)
    pa, pb = par_a[..., :3] / (frames // 2), par_b[..., :3] / (frames - frames // 2)
    rmse_fp = float(np.sqrt(np.mean((fa - pb) ** 2)))
    rmse_pp = float(np.sqrt(np.mean((pa - pb) ** 2)))
    noise_s = abs(par_a[..., 3].sum() / half - par_b[..., 3].sum() / (n_p - half)) / sp
    noise_m = abs(par_a[..., :3].sum() / half - par_b[..., :3].sum() / (n_p - half)) / mp
    return {"what": "k_render_fast (benchmarked layout: rank directory, half tables, %d^3-voxel bound cells, fp16 opacity octets) vs "
                    "k_render_parity (per-voxel windows) on the full %dx%dx%d grid, %dx%d, frames 12..%d (fast: 4x as many)"
                    % ((st["bound_cell_voxels"],) + tuple(dims) + (W, H, 12 + frames - 1)),
            "mean_scatters_fast": float(sf), "mean_scatters_parity": float(sp), "scatter_rel": float(abs(sf - sp) / sp),
            "image_mean_fast": float(mf), "image_mean_parity": float(mp), "mean_rel": float(abs(mf - mp) / mp),
            "parity_half_vs_half": {"scatter_rel": float(noise_s), "mean_rel": float(noise_m)},
            "rmse_equal_spp": {"spp": frames // 2, "fast_vs_parity": rmse_fp, "parity_vs_parity": rmse_pp, "ratio": rmse_fp / rmse_pp},
            "max_pixel_sample_mean": {"fast": float(np.nanmax(fast[..., :3]) / (4 * frames)), "parity": float(np.nanmax(par[..., :3]) / frames)},
            "nonfinite_pixels": {"fast": int((~np.isfinite(fast)).any(axis=-1).sum()), "parity": int((~np.isfinite(par)).any(axis=-1).sum())},
            "tolerance": {"scatter_rel": 0.01, "mean_rel": 0.005},
            "ok": bool(abs(sf - sp) / sp <= 0.01 and abs(mf - mp) / mp <= 0.005 + noise_m),
            "bounds_voxel_bytes": st["bounds_voxel_bytes"], "setup_s": round(setup, 2)}


def _timed_nccl_boot(vp, r, rank, world):
    t0 = time.perf_counter()
    vp.init_nccl_via_store(r, rank, world)
    return time.perf_counter() - t0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=None, choices=list(WORKLOADS))
    ap.add_argument("--frames-per-step", type=int, default=256,
                    help="frames (sample indices) of the whole image per step; N = 1: one launch; N > 1: split over the ranks")
    ap.add_argument("--scaling", default=None, choices=["strong", "weak"],
                    help="N > 1: strong (default) = a step is --frames-per-step frames in total; weak = per GPU")
    ap.add_argument("--store", default="f32", choices=["f32", "f16"])
    ap.add_argument("--ref-frames", type=int, default=8, help="frames per step of the host reference arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ref-cuda", action="store_true")
    ap.add_argument("--no-check", action="store_true")
    ap.add_argument("--shard-opacity", type=int, default=0,
                    help="N > 1: each rank sweeps 1/N of the sun-opacity table + ncclAllGather (needs the communicator first, so the NCCL "
                         "bootstrap no longer overlaps the volume build; default 0: every rank builds all of it, 0.55 s)")
    ap.add_argument("--reduce", default="end", choices=["end", "step"],
                    help="N > 1: 'end' = every rank accumulates all its frames, ONE vp_reduce_nccl at the end of the timed region "
                         "(the NCCL bootstrap, seconds at 8 ranks, overlaps setup and rendering); 'step' = one reduce per step on a "
                         "side stream, overlapped with the next step's render (progressive image on the root)")
    ap.add_argument("--transport", default="ipc", choices=["ipc", "nccl"],
                    help="--reduce end: 'ipc' = vp_reduce_ipc, ONE kernel on the root sums the peers' accumulators over NVLink through "
                         "CUDA IPC mappings (no communicator, no bootstrap); 'nccl' = vp_reduce_nccl")
    ap.add_argument("--render-streams", type=int, default=2, choices=[1, 2],
                    help="consecutive steps alternate between this many streams, so the tail of one launch overlaps the next launch")
    ap.add_argument("--truth-spp", type=int, default=4096, help="spp of the reference-kernel ground truth of the time-to-RMSE leg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.workload is None:
        args.workload = "c2" if max(world, args.gpus) == 1 else "c5"
    dims, (W, H), over, desc = WORKLOADS[args.workload]

    if args.impl == "reference":
        if rank != 0:
            return 0
        # torchrun pins OMP_NUM_THREADS=1; the reference arm uses every host core it can
        os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
        value, dt, sample, cb = host_reference_run(args)
        line = {"impl": "reference", "metric": "path-samples/s", "value": value, "unit": "path-samples/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": "bounded CPU sample of %s: %s" % (args.workload.upper(), sample), "b200_arm_workload": desc},
                "cpu_baseline": cb,
                "e2e": {"value": value, "unit": "path-samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    # keep fd 1 for the ONE JSON line: NCCL / library chatter (e.g. "NCCL version ...") goes to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(obj) + "\n").encode())

    import numpy as np
    import torch
    import torch.distributed as dist

    import cuda_volpath_b200 as vp

    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local)
    t_job = time.perf_counter()
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    strong = world > 1 and (args.scaling or "strong") == "strong"
    r = vp.Renderer(local)
    env, sun_dir, sun_power, view = scene_inputs(vp)
    t_setup = time.perf_counter()
    store = vp.VOXEL_F32 if args.store == "f32" else vp.VOXEL_F16
    # The library's own NCCL communicator (C ABI).  torch.distributed only carries the 128-byte id (through its TCP store)
    # and the barriers.  ncclCommInitRank plus NCCL's lazy ring set-up cost seconds at 8 ranks, so the bootstrap runs in a
    # background thread beside the volume build -- and, with --reduce end, beside the rendering.
    nccl_thread, nccl_info = None, {}
    use_ipc = world > 1 and args.reduce == "end" and args.transport == "ipc" and not args.shard_opacity
    if world > 1 and not use_ipc:
        def _nccl_boot():
            t0 = time.perf_counter()
            vp.init_nccl_via_store(r, rank, world)
            nccl_info["s"] = time.perf_counter() - t0

        nccl_thread = threading.Thread(target=_nccl_boot, daemon=True)
        nccl_thread.start()
    tb = [time.perf_counter()]
    if dims is None:
        r.set_julia()
    else:
        r.generate_cloud(*dims, seed=CLOUD_SEED, store=store, bounds=vp.BOUNDS_CELL)
    r.sync()
    tb.append(time.perf_counter())
    r.set_texture_filter_mode(True)
    r.init_envmap(env)
    r.set_sun(sun_dir, sun_power)
    r.copy_inv_view_matrix(view)
    r.sync()
    tb.append(time.perf_counter())
    tb.append(time.perf_counter())
    if world > 1 and args.shard_opacity:
        nccl_thread.join()  # the sharded sweep is a collective: it needs the communicator now
    r.precompute_opacity(sun_dir, sharded=world > 1 and args.shard_opacity != 0)
    r.sync()
    tb.append(time.perf_counter())
    t_setup = time.perf_counter() - t_setup
    setup_breakdown = {"volume_bounds_octets_s": round(tb[1] - tb[0], 3), "env_sun_tables_s": round(tb[2] - tb[1], 3),
                       "opacity_s": round(tb[4] - tb[3], 3)}
    opacity_ms = r.opacity_build_ms()
    stats = r.volume_stats() if dims is not None else {}
    P = workload_param(vp, W, H, over)
    fps = args.frames_per_step
    step_frames = fps if (world == 1 or strong) else fps * world  # frames of one step, all ranks together
    main_stream = torch.cuda.current_stream()
    stream = main_stream.cuda_stream
    # consecutive steps alternate between the render streams: the persistent grid of step k + 1 moves in as the CTAs of
    # step k retire, so the tail of a launch (its last, longest paths) is filled with the next launch's work
    rs = [torch.cuda.Stream() for _ in range(args.render_streams)]
    total = torch.zeros(H, W, 4, device="cuda", dtype=torch.float32)  # N = 1 / --reduce end: this rank's accumulator; the image on the root
    total_ptr = total.data_ptr()
    reduce_step = world > 1 and args.reduce == "step"
    ipc_fallback = None
    if use_ipc:
        # peer-memory reduce: the accumulator is a plain cudaMalloc block (IPC handles need base pointers); its 64-byte
        # handle travels to the root through torch.distributed's TCP store.  Where CUDA IPC is not permitted (some container
        # set-ups) every rank learns it here and the job falls back to the NCCL transport -- loudly, in the JSON line.
        ok, ipc_ptr, peer_handles = 1, None, []
        store = dist.distributed_c10d._get_default_store()
        try:
            if os.environ.get("VOLPATH_BENCH_NO_IPC"):
                raise RuntimeError("disabled by VOLPATH_BENCH_NO_IPC")
            ipc_ptr = r.dev_alloc(W * H * 16)
            store.set("volpath_ipc_%d" % rank, r.ipc_export(ipc_ptr))
        except Exception as e:
            ok, ipc_fallback = 0, repr(e)[:200]
            store.set("volpath_ipc_%d" % rank, b"")
        if rank == 0:
            try:
                peer_handles = [bytes(store.get("volpath_ipc_%d" % q)) for q in range(1, world)]
                if ok and all(len(h) == 64 for h in peer_handles):
                    r.reduce_ipc(ipc_ptr, peer_handles, 0, stream=stream)  # maps the peers now (cudaIpcOpenMemHandle + peer access), adds nothing
                else:
                    ok = 0
            except Exception as e:
                ok, ipc_fallback = 0, repr(e)[:200]
        flag = torch.tensor([ok], device="cuda", dtype=torch.int32)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 1:
            del total
            total_ptr = ipc_ptr
        else:
            use_ipc = False
            ipc_fallback = ipc_fallback or "CUDA IPC unavailable on another rank"
            nccl_thread = threading.Thread(target=lambda: nccl_info.update(s=_timed_nccl_boot(vp, r, rank, world)), daemon=True)
            nccl_thread.start()
    if reduce_step:
        nccl_thread.join()
        bufs = [torch.zeros(H, W, 4, device="cuda", dtype=torch.float32) for _ in range(2)]
        side = torch.cuda.Stream()
        ev_render = [torch.cuda.Event() for _ in range(2)]
        ev_free = [torch.cuda.Event() for _ in range(2)]
        used = [False, False]
    torch.cuda.synchronize()

    def step(k):
        # global frames of step k: [k * step_frames, (k + 1) * step_frames); this rank takes every world-th one
        first, count, stride = vp.frames_for_rank(k * step_frames, step_frames, rank, world)
        if not reduce_step:
            st = rs[k % len(rs)]
            r.render_kernel(total_ptr, first, P, mode=vp.MODE_FAST, n_frames=count, frame_stride=stride, stream=st.cuda_stream)
            return
        b = k & 1
        st = rs[b % len(rs)]
        if used[b]:
            st.wait_event(ev_free[b])  # its reduce (step k - 2) has drained and zeroed the buffer
        r.render_kernel(bufs[b].data_ptr(), first, P, mode=vp.MODE_FAST, n_frames=count, frame_stride=stride, stream=st.cuda_stream)
        ev_render[b].record(st)
        side.wait_event(ev_render[b])
        r.reduce_nccl(bufs[b].data_ptr(), bufs[b].data_ptr() if rank == 0 else None, W * H, root=0, stream=side.cuda_stream)
        if rank == 0:
            vp.lib.check(r.L.vp_accumulate(r.h, total.data_ptr(), bufs[b].data_ptr(), W * H, side.cuda_stream))
        with torch.cuda.stream(side):
            bufs[b].zero_()
        ev_free[b].record(side)
        used[b] = True

    def drain():
        # the main stream waits for everything the steps put on the other streams
        for st in rs:
            main_stream.wait_stream(st)
        if reduce_step:
            for b in range(2):
                if used[b]:
                    main_stream.wait_event(ev_free[b])

    for k in range(args.warmup):
        step(k)
    drain()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    sampler.start()
    n0 = r.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall = time.perf_counter()
    e0.record()
    for st in rs:
        st.wait_event(e0)
    for k in range(args.warmup, args.warmup + args.steps):
        step(k)
    drain()
    if use_ipc:
        # ONE reduce inside the timed region: every rank's render is complete (stream sync + barrier), then one kernel on
        # the root adds the peers' accumulators, read over NVLink through their IPC mappings, in rank order
        torch.cuda.synchronize()
        dist.barrier()
        if rank == 0:
            r.reduce_ipc(total_ptr, peer_handles, W * H, stream=stream)
    elif world > 1 and not reduce_step:
        # ONE reduce of the whole accumulator, inside the timed region (ncclReduce of W*H float4 over NVLink)
        nccl_thread.join()
        r.reduce_nccl(total_ptr, total_ptr if rank == 0 else None, W * H, root=0, stream=stream)
    e1.record()
    torch.cuda.synchronize()
    t_wall = time.perf_counter() - t_wall
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    launches = r.launch_count() - n0
    sampler.stop_flag = True
    sampler.join(timeout=2)
    if world > 1:
        t = torch.tensor([ms, t_setup], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, t_setup_max = float(t[0].item()), float(t[1].item())
    else:
        t_setup_max = t_setup
    paths = W * H * step_frames * args.steps
    value = paths / (ms * 1e-3)
    # what the image holds: scatter counts are integers, so the mean scatter count per path is exact evidence that every
    # (pixel, frame) item was rendered once (N > 1: on the root, after the reduces)
    img_stats = None
    if rank == 0:
        n_img = W * H * step_frames * (args.warmup + args.steps)
        if use_ipc:
            host = np.empty((H, W, 4), np.float32)
            assert r.L.vp_dev_to_host(host.ctypes.data, total_ptr, host.nbytes) == 0
            total = torch.from_numpy(host)
        finite = torch.isfinite(total).all(dim=-1)
        rgb = torch.where(finite.unsqueeze(-1), total[..., :3], torch.zeros_like(total[..., :3]))
        img_stats = {"mean_scatters_per_path": float(total[..., 3].double().sum().item() / n_img),
                     "image_mean": float(rgb.double().sum().item() / (3 * n_img)), "nonfinite_pixels": int((~finite).sum().item()),
                     "max_pixel_mean": float(rgb.max().item() / (step_frames * (args.warmup + args.steps))),
                     "frames_in_image": step_frames * (args.warmup + args.steps)}

    # kernel-only average launch duration for the roofline: this rank's render launch alone
    kms = []
    first, count, stride = vp.frames_for_rank(10 ** 6, step_frames, rank, world)
    scratch = torch.zeros(H, W, 4, device="cuda", dtype=torch.float32)
    for k in range(2):
        r.render_kernel(scratch.data_ptr(), first + k * step_frames, P, mode=vp.MODE_FAST, n_frames=count, frame_stride=stride, stream=stream)
        kms.append(r.last_kernel_ms())
    kernel_ms = sum(kms) / len(kms)  # an ISOLATED launch (CUDA events on its own stream); the timed region overlaps tails
    kernel_paths = W * H * count
    del scratch

    # e2e: host-buffer call, pinned float4 sum in and out
    h_sum = torch.zeros(H, W, 4, dtype=torch.float32).pin_memory()
    e2e_steps = max(2, min(args.steps, 4))
    # one untimed call: first-use costs of the host-buffer path (device accumulator allocation, copy engines)
    first, count, stride = vp.frames_for_rank(999 * step_frames, step_frames, rank, world)
    vp.lib.check(r.L.vp_render_to_host(r.h, h_sum.data_ptr(), first, count, stride, ctypes.byref(P), vp.MODE_FAST))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for k in range(e2e_steps):
        first, count, stride = vp.frames_for_rank((1000 + k) * step_frames, step_frames, rank, world)
        r.copy_inv_view_matrix(view)
        vp.lib.check(r.L.vp_render_to_host(r.h, h_sum.data_ptr(), first, count, stride, ctypes.byref(P), vp.MODE_FAST))
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dt], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    e2e = {"value": W * H * step_frames * e2e_steps / dt, "unit": "path-samples/s",
           "h2d_bytes_per_step": W * H * 16 + 44 + 48, "d2h_bytes_per_step": W * H * 16}
    if world > 1:
        e2e["note"] = "per rank: its own pinned host float4 sum in and out around its share of the step (bytes are per rank)"

    if rank != 0:
        r.close()
        if world > 1:
            dist.destroy_process_group()
        return 0

    peak, peak_src = peaks()
    B, cnt = algorithmic_bytes_per_path(args.workload, max(count, 1))
    per_launch_bytes = B * kernel_paths
    achieved = per_launch_bytes / (kernel_ms * 1e-3) / 1e9
    roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
            "kernel": "k_render_fast", "kernel_ms": kernel_ms, "bytes_per_path_sample": B, "counts": cnt, "peak_source": peak_src,
            "path_samples_per_launch": kernel_paths}
    if args.workload == "c1":
        roof["note"] = ("C1 is procedural (no density fetches): the kernel is issue-bound, the HBM figure only covers the env texel "
                        "and the accumulator; SURVEY.md 8d asks for instructions per path there (profiles/README.md)")
    prof = os.path.join(ROOT, "profiles", "dram_traffic.json")
    if os.path.exists(prof):
        t = json.load(open(prof)).get(args.workload)
        if t and count == t.get("frames", 64) and (W, H) == (1920, 1080):
            roof["traffic"] = t["dram_bytes_per_launch"]  # bytes per launch, same launch shape as `achieved`
            roof["traffic_source"] = t["source"]
    roof["algorithmic_bytes_per_launch"] = per_launch_bytes

    # wall-to-image of a job that renders exactly the timed steps (no warm-up): the NCCL bootstrap runs beside setup (and,
    # with --reduce end, beside the rendering); --reduce step needs the communicator before its first step
    nb = nccl_info.get("s", 0.0)
    wall_to_image = (max(t_setup_max, nb) + ms * 1e-3) if reduce_step else max(t_setup_max + ms * 1e-3, nb)
    how = ("vp_reduce_nccl per step on a side stream, double-buffered accumulators" if reduce_step else
           ("each rank accumulates its frames, ONE vp_reduce_ipc at the end of the timed region (a kernel on the root sums the peers' "
            "accumulators over NVLink through CUDA IPC mappings; no communicator)" if use_ipc else
            "each rank accumulates its frames, ONE vp_reduce_nccl at the end of the timed region (NCCL bootstrap in a background thread)"))
    par = ("one GPU" if world == 1 else
           "sample-index sharding x%d (%s scaling: %d frames per step %s), %s"
           % (world, "strong" if strong else "weak", step_frames, "in total" if strong else "= %d per GPU" % fps, how))
    par += "; consecutive steps alternate over %d render stream(s)" % len(rs)
    if ipc_fallback:
        par += "; CUDA IPC was NOT available (%s): fell back to the NCCL transport" % ipc_fallback
    line = {"metric": "path-samples/s", "value": value, "unit": "path-samples/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong" if (strong or world == 1) else "weak",
            "vs_baseline": None, "dtype": "f32" if args.store == "f32" else "f16", "data": "synthetic",
            "config": {"workload": desc, "frames_per_step": step_frames, "frames_per_step_per_gpu": step_frames // world if world > 1 else fps,
                       "mode": "fast (megakernel)", "store": args.store,
                       "l2": "inputs larger than L2 (octet store %.1f GB)" % (stats.get("octet_bytes", 0) / 1e9),
                       "parallelism": par, "volume": stats, "setup_s": round(t_setup_max, 2), "setup_breakdown_rank0": setup_breakdown, "opacity_build_s": round(opacity_ms * 1e-3, 3),
                       "wall_to_image_s": round(wall_to_image, 2), "nccl_bootstrap_s_rank0": round(nccl_info.get("s", 0.0), 2),
                       "wall_note": "max over ranks of setup (cloud, bricks, bounds, sun tables: per GPU, replicated) + the timed region; the NCCL "
                                    "bootstrap (background thread) counts where it is the longer pole"},
            "roofline": roof, "e2e": e2e, "gpu_launches": launches, "ms_per_step_region": ms / args.steps, "clocks": sampler.summary(), "image": img_stats}
    if world == 1:
        line["scaling"] = "weak"  # one GPU: per-GPU work is what it is

    if use_ipc:
        r.L.vp_dev_free(total_ptr)
    r.close()
    del total
    torch.cuda.empty_cache()
    if world == 1 and not args.no_check and dims is not None and dims == C2_DIMS:
        try:
            line["check"] = check_block(vp, local, dims, over)
        except Exception as e:  # out of memory on a smaller GPU, ...
            line["check"] = {"unavailable": repr(e)[:300]}
        torch.cuda.empty_cache()
    if world == 1 and not args.no_cpu_baseline:
        j = run_tool([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--workload", args.workload,
                      "--steps", "2", "--warmup", "1"], 240)
        cb = j.get("cpu_baseline", j)
        if "value" in cb:
            try:
                g = gpu_same_sample(vp, local, args.workload, cb.get("frames_per_step", 8), over)
                g["ratio_vs_cpu_baseline"] = g["value"] / cb["value"]
                cb["gpu_same_sample"] = g
            except Exception as e:
                cb["gpu_same_sample"] = {"unavailable": repr(e)[:300]}
        line["cpu_baseline"] = cb
    if world == 1 and not args.no_ref_cuda and dims is not None:
        line["ref_cuda"] = ref_cuda_compare(local, fps, over, args.truth_spp)
        if isinstance(line["ref_cuda"], dict) and "rmse" in line["ref_cuda"]:
            line["rmse"] = line["ref_cuda"]["rmse"]
    emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


def run_tool(cmd, timeout):
    """Run a measurement leg in its own process (the reference libraries export the same symbol names as ours) and
    return its one JSON line, or a reason."""
    try:
        p = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT)
        for ln in reversed(p.stdout.strip().splitlines()):
            if ln.startswith("{"):
                return json.loads(ln)
        return {"unavailable": "rc %d: %s" % (p.returncode, (p.stderr or "").strip()[-300:])}
    except Exception as e:  # timeout, missing file
        return {"unavailable": repr(e)[:300]}


def ref_cuda_compare(device, fps, over, truth_spp):
    """The reference's own CUDA kernel rebuilt for sm_100, beside ours on the same scene and GPU
    (tools/compare_ref_cuda.py, C2 cloud family at 1/4 dims -- what the reference's layout can hold), and the
    time-to-RMSE leg against a high-spp image of the reference kernel."""
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libvolpath_ref_cuda.so")):
        return {"unavailable": "oracle/_ref/libvolpath_ref_cuda.so not built"}
    cmd = [sys.executable, os.path.join(ROOT, "tools", "compare_ref_cuda.py"), "--frames", str(fps), "--device", str(device),
           "--rmse-truth-spp", str(truth_spp)]
    if "material" in over:
        cmd += ["--material", str(over["material"])]
    if "albedo" in over:
        cmd += ["--albedo", str(over["albedo"])]
    if "density" in over:
        cmd += ["--density", str(over["density"])]
    return run_tool(cmd, 420)


if __name__ == "__main__":
    sys.exit(main())
