#!/usr/bin/env python3
"""GPU tool: work counters and binning efficiency of VP_MODE_FAST on a cloud scene (instrumented kernel variant)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cuda_volpath_b200 as vp  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--dims", type=int, nargs=3, default=[497, 338, 612])
ap.add_argument("--image", type=int, nargs=2, default=[1920, 1080])
ap.add_argument("--frames", type=int, default=16)
ap.add_argument("--mode", default="fast")
ap.add_argument("--material", type=int, default=-1)
a = ap.parse_args()
import torch  # noqa: E402

env, sd, sp = vp.default_sunsky()
r = vp.Renderer(0)
r.generate_cloud(*a.dims, seed=0, bounds=vp.BOUNDS_CELL)
r.set_texture_filter_mode(True)
r.init_envmap(env)
r.set_sun(sd, sp)
r.copy_inv_view_matrix(vp.inv_view_matrix())
r.precompute_opacity(sd)
W, H = a.image
P = vp.default_param(W, H)
if a.material >= 0:
    P = vp.mat(P, *vp.MATERIALS[a.material])
acc = torch.zeros(H, W, 4, device="cuda")
r.set_stats(True)
r.counters(reset=True)
r.render_kernel(acc.data_ptr(), 16, P, mode=vp.MODE_WAVE if a.mode == "wave" else vp.MODE_FAST, n_frames=a.frames, stream=torch.cuda.current_stream().cuda_stream)
ms = r.last_kernel_ms()
c = r.counters()
n = W * H * a.frames
out = {k: v / n for k, v in c.items() if not k.startswith(("blocks_", "lanes_"))}
for name in ("path", "scatter", "segment", "step"):
    b, l = c["blocks_" + name], c["lanes_" + name]
    out["block_" + name] = {"warp_execs_per_path": b / n, "avg_active_lanes": l / max(b, 1)}
tot_b = sum(c["blocks_" + k] for k in ("path", "scatter", "segment", "step"))
tot_l = sum(c["lanes_" + k] for k in ("path", "scatter", "segment", "step"))
out["avg_active_lanes_all_blocks"] = tot_l / max(tot_b, 1)
out["ms_instrumented"] = ms
print(json.dumps(out, indent=1))
