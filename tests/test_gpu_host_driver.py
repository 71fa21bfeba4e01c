"""GPU (-m gpu): the headless C++ host (host/volpath_host.cpp) calls only the reference's 14 extern "C" entry points;
the SAME source is linked once against libvolpath_b200.so and once against the reference kernel rebuilt for sm_100.
Same inputs -> the dumped float4 sums must agree (point filter: 1e-5 relative on >= 98 % of pixels)."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "host", "volpath_host")
HOST_REF = os.path.join(ROOT, "host", "volpath_host_ref")


def run(binary, out, extra=()):
    # 11 samples = frames 0..10: from frame 11 on both sides read the sun-opacity TABLE through a linear filter, which
    # is the hardware texture unit on the reference side (undocumented 9-bit lerp) -- covered by the looser test below
    cmd = [binary, "--blob", "56", "--size", "96", "64", "--spp", "11", "--density", "300", "--dump", out] + list(extra)
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr[-2000:]
    assert "M samples / s" in p.stdout
    return np.fromfile(out, np.float32).reshape(64, 96, 4)


@pytest.mark.parametrize("quantized", [False, True])
def test_cpp_host_is_a_drop_in_for_the_reference_boundary(tmp_path, quantized):
    if not (os.path.exists(HOST) and os.path.exists(HOST_REF)):
        pytest.skip("host binaries not built (make -C host all ref)")
    extra = ["--point"] + (["--quantized"] if quantized else [])
    ours = run(HOST, str(tmp_path / "ours.f32"), extra)
    ref = run(HOST_REF, str(tmp_path / "ref.f32"), extra)
    assert ref[..., 3].sum() > 0
    rel = np.abs(ours - ref) / np.maximum(np.abs(ref), 1e-4)
    assert float((rel.max(axis=-1) <= 1e-5).mean()) >= 0.98


def test_cpp_host_with_opacity_table_frames(tmp_path):
    """Frames 11..15 take the precomputed-opacity branch for deep paths (K.cu:2183): values agree to the precision of
    the texture unit's filter (1e-3 relative), paths do not split (the table never feeds an accept/reject test)."""
    if not (os.path.exists(HOST) and os.path.exists(HOST_REF)):
        pytest.skip("host binaries not built (make -C host all ref)")
    extra = ["--point", "--spp", "16", "--density", "900"]
    ours = run(HOST, str(tmp_path / "ours.f32"), extra)
    ref = run(HOST_REF, str(tmp_path / "ref.f32"), extra)
    assert ref[..., 3].max() > 16 * 5
    rel = np.abs(ours - ref) / np.maximum(np.abs(ref), 1e-4)
    assert float((rel.max(axis=-1) <= 2e-3).mean()) >= 0.97
    assert float((ours[..., 3] == ref[..., 3]).mean()) >= 0.97  # same scatter counts: same paths


def test_cpp_host_fast_mode_matches_in_the_mean(tmp_path):
    if not (os.path.exists(HOST) and os.path.exists(HOST_REF)):
        pytest.skip("host binaries not built (make -C host all ref)")
    cmd_extra = ["--spp", "256"]
    fast = run(HOST, str(tmp_path / "fast.f32"), ["--fast"] + cmd_extra)
    ref = run(HOST_REF, str(tmp_path / "ref.f32"), cmd_extra)
    # 256 spp x 6144 pixels: the Monte-Carlo noise of these means is ~1 % (radiance, heavy-tailed) / ~0.3 % (scatters)
    assert abs(fast[..., :3].mean() - ref[..., :3].mean()) <= 0.04 * ref[..., :3].mean()
    assert abs(fast[..., 3].mean() - ref[..., 3].mean()) <= 0.015 * ref[..., 3].mean()


def test_cpp_host_batched_launches_render_the_same_samples(tmp_path):
    """--batch N (vp_render with n_frames = N from the C++ host) renders the same (pixel, frame) samples as one
    render_kernel call per frame: identical scatter counts, rgb up to the fp32 order of the per-pixel additions."""
    if not os.path.exists(HOST):
        pytest.skip("host binary not built (make -C host all)")
    extra = ["--fast", "--spp", "40", "--density", "600"]
    one = run(HOST, str(tmp_path / "one.f32"), extra)
    many = run(HOST, str(tmp_path / "many.f32"), extra + ["--batch", "16"])
    assert one[..., 3].sum() > 0
    assert np.array_equal(one[..., 3], many[..., 3])
    assert np.allclose(one[..., :3], many[..., :3], rtol=1e-5, atol=1e-6)


MULTI = os.path.join(ROOT, "host", "volpath_multi")


def _multi(out, devices, env=None):
    cmd = [MULTI, "--devices", devices, "--blob", "56", "--size", "96", "64", "--spp", "29", "--density", "300", "--dump", out]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=dict(os.environ, **(env or {})))
    assert p.returncode == 0, p.stderr[-2000:]
    assert "M samples / s" in p.stdout
    return np.fromfile(out, np.float32).reshape(64, 96, 4)


def test_cpp_multi_gpu_host_sample_sharding_and_reduce(tmp_path):
    """host/volpath_multi.cpp: one context per GPU, frames interleaved (vp_render with stride G), sums combined with
    vp_reduce.  Scatter counts (.w, integers) are identical to the one-context render for any G; rgb up to fp32 order.
    On a one-GPU box the G contexts share device 0 and vp_reduce takes its peer-copy path; with >= 2 GPUs it is NCCL."""
    import torch

    if not os.path.exists(MULTI):
        pytest.skip("host/volpath_multi not built (make -C host)")
    one = _multi(str(tmp_path / "one.f32"), "0")
    assert one[..., 3].sum() > 0
    three = _multi(str(tmp_path / "three.f32"), "0,0,0", env={"VOLPATH_REDUCE_P2P": "1"})
    assert np.array_equal(one[..., 3], three[..., 3])
    assert np.allclose(one[..., :3], three[..., :3], rtol=1e-5, atol=1e-6)
    if torch.cuda.device_count() >= 2:
        two = _multi(str(tmp_path / "two.f32"), "0,1")
        assert np.array_equal(one[..., 3], two[..., 3])
        assert np.allclose(one[..., :3], two[..., :3], rtol=1e-5, atol=1e-6)


def test_nccl_reduce_through_the_c_abi_single_rank():
    """vp_nccl_unique_id / vp_nccl_init / vp_reduce_nccl on a one-rank communicator: binds libnccl.so.2 at run time and
    runs ncclReduce on this GPU (the multi-rank case is bench.py --gpus N and tests under torchrun on >= 2 GPUs)."""
    import torch

    import cuda_volpath_b200 as vp

    r = vp.Renderer(0)
    if not r.L.vp_nccl_available():
        pytest.skip("libnccl.so.2 not loadable")
    r.nccl_init(1, 0, r.nccl_unique_id())
    # the sharded opacity build degenerates to the plain one on a one-rank communicator
    import numpy as np

    from oraclelib import Oracle

    vol = Oracle().fbm_cloud(48, 32, 56, seed=1)
    r.init_cuda(vol, False)
    r.set_texture_filter_mode(True)
    sun = np.array([0.0, 0.951057, -0.309017], np.float32)
    r.set_sun(sun, np.ones(3, np.float32))
    r.precompute_opacity(sun)
    plain = r.opacity_fast()
    r.precompute_opacity(sun, sharded=True)
    assert np.array_equal(plain, r.opacity_fast()) and plain.max() > 0
    a = torch.rand(1000, 4, device="cuda")
    b = torch.zeros_like(a)
    s = torch.cuda.current_stream().cuda_stream
    r.reduce_nccl(a.data_ptr(), b.data_ptr(), 1000, root=0, stream=s)
    torch.cuda.synchronize()
    assert torch.equal(a, b)
    with pytest.raises(vp.VolpathError):
        r.reduce_nccl(a.data_ptr(), b.data_ptr(), 1000, root=3, stream=s)
    r.nccl_destroy()
    r.close()


def _ipc_child(conn, root):
    import sys

    sys.path.insert(0, root)
    import numpy as np

    import cuda_volpath_b200 as vp

    r = vp.Renderer(0)
    n = 5000
    p = r.dev_alloc(n * 16)
    a = (np.arange(n * 4, dtype=np.float32) * 0.25).reshape(n, 4)
    assert r.L.vp_host_to_dev(p, a.ctypes.data, a.nbytes) == 0
    r.sync()
    conn.send(r.ipc_export(p))
    conn.recv()  # the parent has read the buffer
    r.L.vp_dev_free(p)
    r.close()


def test_peer_memory_reduce_through_cuda_ipc_two_processes():
    """vp_ipc_export / vp_reduce_ipc: another PROCESS's accumulator is mapped through its 64-byte IPC handle and added by
    one kernel in rank order (here both processes share GPU 0; across GPUs the same kernel reads over NVLink:
    tests/test_gpu_multi_rank.py, bench.py --gpus N)."""
    import multiprocessing as mp

    import cuda_volpath_b200 as vp

    ctx = mp.get_context("spawn")
    parent, child = ctx.Pipe()
    proc = ctx.Process(target=_ipc_child, args=(child, ROOT))
    proc.start()
    try:
        assert parent.poll(120), "child did not export a handle"
        handle = parent.recv()
        assert len(handle) == 64
        r = vp.Renderer(0)
        n = 5000
        mine = np.full((n, 4), 2.0, np.float32)
        p = r.dev_alloc(n * 16)
        assert r.L.vp_host_to_dev(p, mine.ctypes.data, mine.nbytes) == 0
        r.reduce_ipc(p, [handle], n)
        r.reduce_ipc(p, [handle], n)  # the mapping is cached: a second reduce adds the peer again
        r.sync()
        out = np.empty_like(mine)
        assert r.L.vp_dev_to_host(out.ctypes.data, p, out.nbytes) == 0
        want = mine + 2 * (np.arange(n * 4, dtype=np.float32) * 0.25).reshape(n, 4)
        assert np.array_equal(out, want)
        r.L.vp_ipc_close(r.h)
        r.L.vp_dev_free(p)
        r.close()
    finally:
        parent.send("done")
        proc.join(60)
    assert proc.exitcode == 0
