"""GPU (-m gpu): edge cases of the path against the reference's own CUDA kernel / the CPU oracle -- ragged and tiny
grids, a non-default box (anisotropic voxels), zero albedo, the 800-scatter cap, every storage type in the fast
renderer, odd image sizes."""
import numpy as np
import pytest

from conftest import SUN_DIR, SUN_POWER, setup_renderer, setup_scene
from test_gpu_parity import _ref_cuda, match_fraction, small_cloud

pytestmark = pytest.mark.gpu


def assert_same_in_the_mean(R, vp, P, spp, what, first=0):
    """fast vs parity means, judged against the noise of the statistic itself: two independent parity renders (disjoint
    frame ranges) give the Monte-Carlo scatter of the mean at this spp; fast must sit within 3 of those + 0.3 %."""
    a = R.render(P, first, spp, mode=vp.MODE_PARITY)
    b = R.render(P, first + spp, spp, mode=vp.MODE_PARITY)
    f = R.render(P, first, spp, mode=vp.MODE_FAST)
    assert np.isfinite(f).all()
    for name, sl in (("radiance", np.s_[..., :3]), ("scatters", np.s_[..., 3])):
        ma, mb, mf = a[sl].mean(), b[sl].mean(), f[sl].mean()
        ref = 0.5 * (ma + mb)
        noise = abs(ma - mb)
        assert abs(mf - ref) <= 3.0 * noise + 0.003 * ref, (what, name, ma, mb, mf)
    return a, f


@pytest.fixture(scope="module")
def R(vp):
    r = vp.Renderer(0)
    yield r
    r.close()


def test_custom_box_with_anisotropic_voxels_vs_reference_kernel(R, oracle, vp):
    ref = _ref_cuda()
    vol = small_cloud(oracle, (40, 56, 24), seed=8)
    box = ((-0.9, -0.5, -1.1), (1.0, 0.7, 0.6))
    P = vp.default_param(80, 56)
    P.density = 250.0
    ref.set_volume(vol, False, box, linear=False)
    ref.set_envmap(vp.constant_sky())
    ref.set_sun(SUN_DIR, SUN_POWER)
    ref.set_inv_view(vp.inv_view_matrix())
    R.init_cuda(vol, False, box=box)
    R.set_texture_filter_mode(False)
    R.init_envmap(vp.constant_sky())
    R.set_sun(SUN_DIR, SUN_POWER)
    R.copy_inv_view_matrix(vp.inv_view_matrix())
    want = ref.render(P, 0, 3)
    got = R.render(P, 0, 3, mode=vp.MODE_PARITY)
    assert want[..., 3].sum() > 0
    assert match_fraction(got, want, 1e-5) >= 0.98
    # and the production renderer agrees in the mean on the same box
    assert_same_in_the_mean(R, vp, P, 512, "custom box")


@pytest.mark.parametrize("dims", [(1, 1, 1), (2, 3, 1), (9, 1, 17)])
def test_tiny_and_degenerate_grids_vs_oracle(R, oracle, vp, dims):
    nx, ny, nz = dims
    vol = np.full((nz, ny, nx), 0.6, np.float32)
    vol.flat[0] = 0.0
    P = vp.default_param(48, 32)
    P.density = 20.0
    setup_scene(oracle, vp, vol, False, True)
    setup_renderer(R, vp, vol, False, True)
    want = oracle.render(P, 0, 2)
    got = R.render(P, 0, 2, mode=vp.MODE_PARITY)
    assert match_fraction(got, want, 1e-4) >= 0.95
    f = R.render(P, 0, 64, mode=vp.MODE_FAST)
    w = R.render(P, 0, 64, mode=vp.MODE_WAVE)
    assert np.isfinite(f).all() and np.array_equal(f[..., 3], w[..., 3])


def test_zero_albedo_and_the_scatter_cap(R, oracle, vp):
    """albedo 0: throughput becomes 0 after the first real collision and the weights go 0/0 -- the reference survives
    because max(NaN, 0) = 0 at the accumulate (Q9); the scatter cap (max_depth 800, K.cu:34) ends every path."""
    ref = _ref_cuda()
    vol = small_cloud(oracle, (32, 24, 40), seed=9)
    P = vp.default_param(48, 32)
    P.density = 400.0
    P.albedo[:] = [0.0, 0.0, 0.0]
    setup_scene(ref, vp, vol, False, False)
    setup_renderer(R, vp, vol, False, False)
    want = ref.render(P, 0, 2)
    got = R.render(P, 0, 2, mode=vp.MODE_PARITY)
    assert np.isfinite(got).all() and match_fraction(got, want, 1e-5) >= 0.97
    fast = R.render(P, 0, 2, mode=vp.MODE_FAST)
    assert np.isfinite(fast).all()
    # an extremely dense, conservative medium: a tail of paths runs into the 800-scatter cap
    P = vp.default_param(24, 16)
    P.density = 2.0e5
    want = ref.render(P, 0, 2)
    got = R.render(P, 0, 2, mode=vp.MODE_PARITY)
    assert want[..., 3].max() <= 2 * 800 and (want[..., 3] >= 800).any()
    assert np.array_equal(got[..., 3] >= 800, want[..., 3] >= 800) or match_fraction(got, want, 1e-4) >= 0.9
    a, f = assert_same_in_the_mean(R, vp, P, 128, "scatter cap")
    assert f[..., 3].max() <= 128 * 800 and (f[..., 3] >= 800).any()


@pytest.mark.parametrize("store", ["u8", "f16", "f32"])
def test_fast_renderer_on_every_storage_type(R, oracle, vp, store):
    vol = small_cloud(oracle, (64, 48, 80), seed=4)
    env, sd, sp = vp.default_sunsky()
    quant = store == "u8"
    v = np.round(vol * 255).astype(np.uint8) if quant else vol
    kw = {"store": vp.VOXEL_F16} if store == "f16" else {}
    setup_renderer(R, vp, v, quant, True, env=env, **kw)
    P = vp.default_param(96, 64)
    P.density = 300.0
    assert_same_in_the_mean(R, vp, P, 512, "store " + store)


@pytest.mark.parametrize("size", [(1, 1), (7, 3), (33, 5), (8, 4)])
def test_odd_image_sizes_cover_every_pixel_exactly_once(R, oracle, vp, size):
    vol = small_cloud(oracle, (32, 24, 40), seed=9)
    setup_renderer(R, vp, vol, False, True)
    W, H = size
    P = vp.default_param(W, H)
    P.density = 100.0
    p = R.render(P, 0, 6, mode=vp.MODE_PARITY)
    f = R.render(P, 0, 6, mode=vp.MODE_FAST)
    w = R.render(P, 0, 6, mode=vp.MODE_WAVE)
    assert p.shape == (H, W, 4) and np.isfinite(f).all()
    assert np.array_equal(f[..., 3], w[..., 3])
    # sky pixels (no scatter in either) carry exactly 6 environment samples in all renderers
    sky = (p[..., 3] == 0) & (f[..., 3] == 0)
    assert np.allclose(f[sky][:, :3], p[sky][:, :3], rtol=1e-4, atol=1e-6)


def test_volume_reupload_and_filter_switch_keep_the_context_consistent(R, oracle, vp):
    a = small_cloud(oracle, (32, 24, 40), seed=1)
    b = small_cloud(oracle, (48, 16, 24), seed=2)
    P = vp.default_param(40, 24)
    P.density = 150.0
    setup_renderer(R, vp, a, False, False)
    first = R.render(P, 0, 2, mode=vp.MODE_PARITY)
    setup_renderer(R, vp, b, False, True)
    R.render(P, 0, 2, mode=vp.MODE_FAST)
    setup_renderer(R, vp, a, False, False)
    again = R.render(P, 0, 2, mode=vp.MODE_PARITY)
    assert np.array_equal(first, again)  # the parity renderer is deterministic
    R.free_cuda_buffers()
    with pytest.raises(vp.VolpathError, match="no volume"):
        R.render(P, 0, 1, mode=vp.MODE_FAST)


def test_checkpoint_resume_on_the_gpu_path(tmp_path, R, oracle, vp):
    """SURVEY.md 8f row 3 on the CUDA path: a sample is addressed by (pixel, frame), so resuming a checkpointed float4 sum
    continues the SAME render.  Parity mode adds frames in order inside a thread: bit-identical to the uninterrupted
    render.  The production renderer accumulates with atomics: scatter counts identical, rgb up to fp32 order."""
    vol = small_cloud(oracle, (48, 32, 56))
    setup_renderer(R, vp, vol, False, True)
    P = vp.default_param(64, 48)
    P.density = 200.0
    for mode, exact in ((vp.MODE_PARITY, True), (vp.MODE_FAST, False)):
        full = R.render(P, 0, 9, mode=mode)
        part = R.render(P, 0, 4, mode=mode)
        p = str(tmp_path / ("ck%d.npz" % mode))
        vp.io.save_checkpoint(p, part, 4, P)
        acc, nxt, pb = vp.io.load_checkpoint(p)
        assert nxt == 4 and bytes(pb) == bytes(P)
        resumed = R.render(P, nxt, 5, mode=mode, accum=acc)
        assert np.array_equal(resumed[..., 3], full[..., 3])
        if exact:
            assert np.array_equal(resumed.view(np.uint32), full.view(np.uint32))
        else:
            assert np.allclose(resumed[..., :3], full[..., :3], rtol=1e-5, atol=1e-6)
