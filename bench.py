#!/usr/bin/env python3
"""bench.py -- path-samples/s of the render hot path (BASELINE.json metric), one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload c2|c2q|c1]

A "step" is one pass of the hot path over one batch: --frames-per-step frames (sample indices) of the whole image.
Default workload (N = 1) is config C2 of BASELINE.json: synthetic fBm cloud at WDAS dims 1987x1351x2449 fp32,
1920x1080, default sun/sky, camera and Param of the reference.  With N > 1 every rank holds the whole
volume and renders its own strided subset of the frames (sample-index sharding); the float4 sums are combined by one
NCCL reduce per step inside the timed region.

`value`   : device-timed (CUDA events on the launch stream), accumulator resident in HBM.
`e2e`     : the same step through the host-buffer C-ABI call vp_render_to_host (pinned host float4 sum in and out).
`roofline`: algorithmic bytes per path-sample (SURVEY.md 8d formula, L/S/O/E counted by the instrumented reference
            kernel, profiles/ref_counters.json) x path-samples/s vs the measured HBM copy peak.
`cpu_baseline` / `--impl reference`: the reference's own kernel source compiled for the host (oracle/_ref, OpenMP over
            all cores) on a bounded sample of the same cloud family.
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
    # name: (dims, image, description)
    "c2": (C2_DIMS, (1920, 1080), "C2: synthetic fBm cloud 1987x1351x2449 fp32, 1920x1080, default sun/sky"),
    "c2q": ((497, 338, 612), (1920, 1080), "C2 cloud family at 1/4 dims 497x338x612 fp32, 1920x1080"),
    "c2e": ((248, 168, 306), (480, 270), "C2 cloud family at 1/8 dims 248x168x306 fp32, 480x270"),
    "c5": (C2_DIMS, (3840, 2160), "C5: synthetic fBm cloud 1987x1351x2449 fp32, 3840x2160, sample-split"),
    "c1": (None, (512, 512), "C1: procedural Julia set (no-OpenVDB build), 512x512"),
}
CLOUD_SEED = 0


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
        key = "c1" if workload == "c1" else "cloud"
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


def host_reference_run(args, workload_desc):
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
    if julia:
        W, H = 256, 256
        ref.set_julia()
        sample = "Julia set, 256x256, %d frame(s) per step"
    else:
        W, H = 480, 270
        vol = Oracle().fbm_cloud(248, 168, 306, seed=CLOUD_SEED)
        ref.set_volume(vol, False, None, linear=True)
        sample = "C2 cloud family at 1/8 dims (248x168x306 fp32), 480x270, frames 0..%d-1 per step (shadow walks, no opacity table)"
    ref.set_envmap(env)
    ref.set_sun(sun_dir, sun_power)
    ref.set_inv_view(view)
    P = vp.default_param(W, H)
    cores = int(os.environ.get("OMP_NUM_THREADS", "0") or 0) or (os.cpu_count() or 1)
    fps = max(1, args.ref_frames)
    acc = np.zeros((H, W, 4), np.float32)
    # frames 0..10 only: from frame 11 on the kernel reads the precomputed sun-opacity table (K.cu:2183), whose
    # construction (_precompute_opacity: ~10^10 emulated texture fetches at these dims) is not feasible on host cores
    fps = min(fps, 11)
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
    return value, dt, dict(value=value, unit="path-samples/s", cores=cores, kind=kind, sample=sample % fps,
                           single_thread_value=single)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=None, choices=list(WORKLOADS))
    ap.add_argument("--frames-per-step", type=int, default=256,
                    help="frames (sample indices) of the whole image per step = per launch; the default 4 steps x 256 are the 1024 spp of config C2")
    ap.add_argument("--store", default="f32", choices=["f32", "f16"])
    ap.add_argument("--ref-frames", type=int, default=8, help="frames per step of the host reference arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ref-cuda", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.workload is None:
        args.workload = "c2"
    dims, (W, H), desc = WORKLOADS[args.workload]

    if args.impl == "reference":
        if rank != 0:
            return 0
        # torchrun pins OMP_NUM_THREADS=1; the reference arm uses every host core it can
        os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
        value, dt, cb = host_reference_run(args, desc)
        line = {"impl": "reference", "metric": "path-samples/s", "value": value, "unit": "path-samples/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": desc}, "cpu_baseline": cb,
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
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    r = vp.Renderer(local)
    env, sun_dir, sun_power, view = scene_inputs(vp)
    t_setup = time.perf_counter()
    store = vp.VOXEL_F32 if args.store == "f32" else vp.VOXEL_F16
    if dims is None:
        r.set_julia()
    else:
        r.generate_cloud(*dims, seed=CLOUD_SEED, store=store, bounds=vp.BOUNDS_CELL)
    r.set_texture_filter_mode(True)
    r.init_envmap(env)
    r.set_sun(sun_dir, sun_power)
    r.copy_inv_view_matrix(view)
    r.precompute_opacity(sun_dir)
    r.sync()
    t_setup = time.perf_counter() - t_setup
    stats = r.volume_stats() if dims is not None else {}
    P = vp.default_param(W, H)
    fps = args.frames_per_step
    acc = torch.zeros(H, W, 4, device="cuda", dtype=torch.float32)
    stream = torch.cuda.current_stream().cuda_stream

    def step(k):
        # global frames of step k: [k*fps*world, (k+1)*fps*world); this rank takes every world-th one
        first, count, stride = vp.frames_for_rank(k * fps * world, fps * world, rank, world)
        r.render_kernel(acc.data_ptr(), first, P, mode=vp.MODE_FAST, n_frames=count, frame_stride=stride, stream=stream)
        if world > 1:
            vp.reduce_accumulators(acc, dst=0)

    for k in range(args.warmup):
        step(k)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    sampler.start()
    n0 = r.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kern_ms = []
    e0.record()
    for k in range(args.warmup, args.warmup + args.steps):
        step(k)
        kern_ms.append(None)
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    launches = r.launch_count() - n0
    sampler.stop_flag = True
    sampler.join(timeout=2)
    if world > 1:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    paths = W * H * fps * world * args.steps
    value = paths / (ms * 1e-3)

    # kernel-only average launch duration for the roofline (the step IS one launch of k_render_fast)
    kms = []
    for k in range(args.warmup + args.steps, args.warmup + args.steps + 2):
        first, count, stride = vp.frames_for_rank(k * fps * world, fps * world, rank, world)
        r.render_kernel(acc.data_ptr(), first, P, mode=vp.MODE_FAST, n_frames=count, frame_stride=stride, stream=stream)
        kms.append(r.last_kernel_ms())
    kernel_ms = sum(kms) / len(kms)
    if world == 1:
        kernel_ms = ms / args.steps  # the timed region is exactly `steps` launches of k_render_fast: CUDA events over it

    # e2e: host-buffer call, pinned float4 sum in and out
    e2e = None
    if rank == 0 or world > 1:
        h_sum = torch.zeros(H, W, 4, dtype=torch.float32).pin_memory()
        e2e_steps = max(2, min(args.steps, 4))
        # one untimed call: first-use costs of the host-buffer path (device accumulator allocation, copy engines)
        first, count, stride = vp.frames_for_rank(999 * fps * world, fps * world, rank, world)
        vp.lib.check(r.L.vp_render_to_host(r.h, h_sum.data_ptr(), first, count, stride, ctypes.byref(P), vp.MODE_FAST))
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for k in range(e2e_steps):
            first, count, stride = vp.frames_for_rank((1000 + k) * fps * world, fps * world, rank, world)
            r.copy_inv_view_matrix(view)
            vp.lib.check(r.L.vp_render_to_host(r.h, h_sum.data_ptr(), first, count, stride, ctypes.byref(P), vp.MODE_FAST))
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {"value": W * H * fps * world * e2e_steps / dt, "unit": "path-samples/s",
               "h2d_bytes_per_step": W * H * 16 + 44 + 48, "d2h_bytes_per_step": W * H * 16}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peak, peak_src = peaks()
    B, cnt = algorithmic_bytes_per_path(args.workload, fps)
    per_launch_bytes = B * W * H * fps
    achieved = per_launch_bytes / (kernel_ms * 1e-3) / 1e9
    roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
            "kernel": "k_render_fast", "kernel_ms": kernel_ms, "bytes_per_path_sample": B, "counts": cnt, "peak_source": peak_src}
    if args.workload == "c1":
        roof["note"] = ("C1 is procedural (no density fetches): the kernel is issue-bound, the HBM figure only covers the env texel "
                        "and the accumulator; SURVEY.md 8d asks for instructions per path there (profiles/README.md)")
    prof = os.path.join(ROOT, "profiles", "dram_traffic.json")
    if os.path.exists(prof):
        t = json.load(open(prof)).get(args.workload)
        if t and fps == t.get("frames", 64) and (W, H) == (1920, 1080):
            roof["traffic"] = t["dram_bytes_per_launch"]  # bytes per launch, same launch shape as `achieved`
            roof["traffic_source"] = t["source"]
        roof["algorithmic_bytes_per_launch"] = per_launch_bytes

    line = {"metric": "path-samples/s", "value": value, "unit": "path-samples/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32" if args.store == "f32" else "f16", "data": "synthetic",
            "config": {"workload": desc, "frames_per_step_per_gpu": fps, "mode": "fast (megakernel)", "store": args.store,
                       "l2": "inputs larger than L2 (octet store %.1f GB)" % (stats.get("octet_bytes", 0) / 1e9),
                       "parallelism": "sample-index sharding x%d, NCCL reduce per step" % world,
                       "volume": stats, "setup_s": round(t_setup, 2)},
            "roofline": roof, "e2e": e2e, "gpu_launches": launches, "clocks": sampler.summary()}

    r.close()
    del acc
    torch.cuda.empty_cache()
    if world == 1 and not args.no_cpu_baseline:
        j = run_tool([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--workload", args.workload,
                      "--steps", "2", "--warmup", "1"], 240)
        line["cpu_baseline"] = j.get("cpu_baseline", j)
    if world == 1 and not args.no_ref_cuda and dims is not None:
        line["ref_cuda"] = ref_cuda_compare(local, fps)
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


def ref_cuda_compare(device, fps):
    """The reference's own CUDA kernel rebuilt for sm_100, beside ours on the same scene and GPU
    (tools/compare_ref_cuda.py, C2 cloud family at 1/4 dims -- what the reference's layout can hold)."""
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libvolpath_ref_cuda.so")):
        return {"unavailable": "oracle/_ref/libvolpath_ref_cuda.so not built"}
    return run_tool([sys.executable, os.path.join(ROOT, "tools", "compare_ref_cuda.py"), "--frames", str(fps),
                     "--device", str(device)], 240)


if __name__ == "__main__":
    sys.exit(main())
