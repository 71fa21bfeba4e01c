"""CPU: host-side mirror of the reference interface (Param presets, camera matrix, sun/sky fixture, frame sharding)
and the multi-GPU reduce on gloo with world_size 2."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_default_param_and_material_presets(vp):
    p = vp.default_param()
    assert (p.width, p.height, p.density, p.brightness) == (960, 512, 800.0, 1.0)  # volumeRender.cpp:1286-1292
    assert abs(p.g - 0.877) < 1e-6 and list(p.albedo) == [1, 1, 1] and list(p.sigma_t) == [1, 1, 1]
    q = vp.mat(p, *vp.MATERIALS[8])  # Mat(P, 0.74,0.88,1.01, 0.032,0.17,0.48), volumeRender.cpp:44-57,1304
    st = np.array([0.74 + 0.032, 0.88 + 0.17, 1.01 + 0.48], np.float32)
    assert np.allclose(list(q.sigma_t), st / st.max(), rtol=1e-6) and max(q.sigma_t) == 1.0
    assert np.allclose(list(q.albedo), np.array([0.74, 0.88, 1.01]) / st, rtol=1e-6)
    assert len(vp.MATERIALS) == 13 and vp.MATERIALS[-1] == (1.0, 1.0, 1.0, 0.0, 0.0, 0.0)


def test_inv_view_matrix_is_a_rigid_camera_frame(vp):
    m = vp.inv_view_matrix().reshape(3, 4)
    R, eye = m[:, :3], m[:, 3]
    assert np.allclose(R.T @ R, np.eye(3), atol=1e-5)
    assert np.allclose(eye, [3.922986, -0.782739, 0.03], atol=1e-6)  # volumeRender.cpp:108
    assert np.allclose(-R[:, 2], [-0.978148, 0.207912, 0.0], atol=1e-5)  # camera looks along `forward`
    assert abs(np.linalg.det(R) - 1) < 1e-5


def test_default_sunsky_fixture(vp):
    env, sd, sp = vp.default_sunsky()
    assert env.shape == (512, 1024, 4) and env.dtype == np.float32
    assert np.allclose(sd, [0, 0.951057, -0.309017], atol=1e-5)
    assert np.allclose(sp, [51797.3, 42480.1, 32578.5], rtol=1e-4)
    assert env[:256, :, :3].mean() > env[256:, :, :3].mean() > 0  # sky brighter than the 1 % ground


@pytest.mark.parametrize("n,world", [(0, 2), (1, 2), (7, 2), (64, 8), (4096, 8), (5, 8)])
def test_frames_for_rank_partitions_the_frames(vp, n, world):
    seen = []
    for r in range(world):
        first, count, stride = vp.frames_for_rank(10, n, r, world)
        seen += [first + i * stride for i in range(count)]
    assert sorted(seen) == list(range(10, 10 + n))


def test_split_frames_strong_scaling(vp):
    for total, world in ((256, 8), (4096, 8), (7, 4), (3, 8)):
        parts = vp.split_frames(total, world)
        assert sum(parts) == total and max(parts) - min(parts) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist

    import cuda_volpath_b200 as vp
    from conftest import setup_scene
    from oraclelib import Oracle

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = np.load(os.path.join(ROOT, "tests", "golden", "ref_golden.npz"))
    orc = Oracle()
    setup_scene(orc, vp, g["vol_f32"], False, True)
    P = vp.default_param(16, 12)
    P.density = 60.0
    first, count, stride = vp.frames_for_rank(0, 5, rank, world)
    acc = np.zeros((12, 16, 4), np.float32)
    for i in range(count):  # the oracle stands in for the per-GPU renderer: same (x, y, frame) -> sample map
        orc.render(P, first + i * stride, 1, accum=acc)
    t = torch.from_numpy(acc)
    vp.reduce_accumulators(t, dst=0)
    if rank == 0:
        q.put(t.numpy().copy())
    dist.barrier()
    dist.destroy_process_group()


def test_sample_sharded_render_reduces_to_the_single_rank_image(vp, oracle, golden):
    import torch.multiprocessing as mp

    from conftest import setup_scene

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    setup_scene(oracle, vp, golden["vol_f32"], False, True)
    P = vp.default_param(16, 12)
    P.density = 60.0
    want = oracle.render(P, 0, 5)
    assert np.array_equal(got[..., 3], want[..., 3])  # scatter counts: integers, exact under any order
    assert np.allclose(got[..., :3], want[..., :3], rtol=1e-6, atol=1e-7)


def test_committed_bench_line_carries_the_contract_keys():
    """The JSON line bench.py printed on a B200 at the end of the round (profiles/r2_bench_c2.json): the keys the
    measurement contract names, with internally consistent numbers."""
    import json

    j = json.load(open(os.path.join(ROOT, "profiles", "r2_bench_c2.json")))
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks", "check", "rmse"):
        assert k in j, k
    assert j["metric"] == "path-samples/s" and j["n_gpus"] == 1 and j["gpu_launches"] == j["steps"] and "workload" in j["config"]
    roof = j["roofline"]
    assert roof["bound"] == "hbm" and abs(roof["frac"] - roof["achieved"] / roof["peak"]) < 1e-9 and roof["traffic"] > 0
    paths = 1920 * 1080 * j["config"]["frames_per_step"] * j["steps"]
    assert abs(j["value"] - paths / (j["ms_per_step"] * j["steps"] * 1e-3)) / j["value"] < 1e-6
    assert j["e2e"]["h2d_bytes_per_step"] > 0 and j["e2e"]["d2h_bytes_per_step"] > 0 and j["e2e"]["value"] < j["value"] * 1.02
    cb = j["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and "sample" in cb and "gpu_same_sample" in cb
    chk = j["check"]
    assert chk["ok"] and chk["scatter_rel"] <= 0.01 and chk["mean_rel"] <= 0.005
    assert j["rmse"]["time_to_rmse_ratio"] > 1 and len(j["rmse"]["rmse_ours"]) == len(j["rmse"]["checkpoints_spp"])
