"""GPU (-m gpu): the variant bench.py runs on the full C2 volume -- rank directory instead of the flat brick table,
half-precision per-cell tables, coarse bound cells (k_render_fast<.., LY = 2>) -- pinned on volumes small enough to
test: VOLPATH_FORCE_RANK_DIR / VOLPATH_HALF_TABLES / VOLPATH_FORCE_CELL_LOG2 put a small volume on the large-volume
path.  Plus the production sun-opacity table (swept build, fp16 octets) against the bit-faithful per-voxel march, the
work-pool counter ring under many streams, and the conservative sun-clear clip on tiny grids.
Reference: the fetch is K.cu:682-695 (vol_sigma_t), the bound at segment start K.cu:1626-1661, the opacity table
K.cu:483-553 / 2183-2195."""
import ctypes

import numpy as np
import pytest

from conftest import SUN_DIR, SUN_POWER, setup_renderer, setup_scene
from test_gpu_parity import _ref_cuda, small_cloud

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def R(vp):
    r = vp.Renderer(0)
    yield r
    r.close()


def _sample_density(oracle, pos):
    import ctypes as C

    want = np.empty(len(pos), np.float32)
    oracle.L.vo_sample_density.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.c_int, C.POINTER(C.c_float)]
    oracle.L.vo_sample_density(oracle.h, pos.ctypes.data_as(C.POINTER(C.c_float)), len(pos), want.ctypes.data_as(C.POINTER(C.c_float)))
    return want


@pytest.mark.parametrize("store", ["u8", "f16", "f32"])
@pytest.mark.parametrize("linear", [False, True])
def test_rank_directory_fetch_bit_exact_vs_texture_emulation(R, oracle, vp, monkeypatch, store, linear):
    """vp_fetch_density through the RANK DIRECTORY (the slot lookup of every volume above 1 M bricks) equals the texture
    emulation bit for bit, and equals the flat-table fetch bit for bit."""
    vol = small_cloud(oracle, (75, 43, 91), seed=11)  # 10 x 6 x 12 bricks: directory words straddle rows and slices
    quant = store == "u8"
    src = np.round(vol * 255).astype(np.uint8) if quant else vol
    emu = src if store != "f16" else vol.astype(np.float16).astype(np.float32)
    rs = np.random.RandomState(3)
    lo, hi = np.array([-1, -43 / 75, -91 / 75], np.float32), np.array([1, 43 / 75, 91 / 75], np.float32)
    pos = (lo + (hi - lo) * (rs.rand(30000, 3).astype(np.float32) * 1.1 - 0.05)).astype(np.float32)
    inside = np.all((pos >= lo) & (pos <= hi), axis=1)
    kw = dict(store=vp.VOXEL_F16) if store == "f16" else {}
    R.init_cuda(src, quant, **kw)
    R.set_texture_filter_mode(linear)
    flat_parity = R.fetch_density(pos, parity=True)
    flat_fast = R.fetch_density(pos[inside], parity=False)
    monkeypatch.setenv("VOLPATH_FORCE_RANK_DIR", "1")
    R.init_cuda(src, quant, **kw)
    R.set_texture_filter_mode(linear)
    got = R.fetch_density(pos, parity=True)
    oracle.set_volume(emu, quant, None, linear=linear)
    want = _sample_density(oracle, pos)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    assert np.array_equal(got.view(np.uint32), flat_parity.view(np.uint32))
    # the production fetch (what k_render_fast calls) through the directory: the same bits as through the flat table
    assert np.array_equal(R.fetch_density(pos[inside], parity=False).view(np.uint32), flat_fast.view(np.uint32))


def _scene(R, vp, vol, env=None):
    setup_renderer(R, vp, vol, False, True, env=env)
    R.precompute_opacity(SUN_DIR)


def test_fast_render_rank_directory_and_layout_variants_render_the_same_samples(R, oracle, vp, monkeypatch):
    """The four storage layouts of the production kernel: {flat table, rank directory} x {float, half per-cell tables}.
    LY = 1 (flat + float) and LY = 2 (directory + half: what the full C2 bench runs) are compile-time specialisations;
    the mixed ones run the generic kernel.  The directory must not change a single path (.w identical); half tables move
    a majorant by its last half bit, so they are compared within their own pair."""
    vol = small_cloud(oracle, (96, 64, 80), seed=3)
    env, sd, sp = vp.default_sunsky()
    P = vp.default_param(128, 80)
    P.density = 500.0
    monkeypatch.setenv("VOLPATH_FORCE_CELL_LOG2", "1")
    out = {}
    for half in ("0", "1"):
        for rank in ("0", "1"):
            monkeypatch.setenv("VOLPATH_HALF_TABLES", half)
            monkeypatch.setenv("VOLPATH_FORCE_RANK_DIR", rank)
            _scene(R, vp, vol, env)
            assert (R.half_tables() is not None) == (half == "1")
            out[half, rank] = R.render(P, 5, 24, mode=vp.MODE_FAST)  # frames > 10: opacity-table branch included
    for half in ("0", "1"):
        a, b = out[half, "0"], out[half, "1"]
        assert a[..., 3].sum() > 0 and a[..., 3].max() / 24 > 20
        assert np.array_equal(a[..., 3], b[..., 3])
        assert np.allclose(a[..., :3], b[..., :3], rtol=2e-5, atol=1e-6)
    # and the wavefront form on the large-volume layout renders the megakernel's samples
    w = R.render(P, 5, 24, mode=vp.MODE_WAVE)
    assert np.array_equal(w[..., 3], out["1", "1"][..., 3])


@pytest.mark.parametrize("cell_log2,scat_tol,mean_tol", [(1, 0.01, 0.005), (3, 0.03, 0.005)])
def test_benchmarked_variant_vs_reference_cuda_kernel(R, vp, monkeypatch, cell_log2, scat_tol, mean_tol):
    """k_render_fast<LY = 2> (rank directory, half tables, coarse bound cells) against the reference's OWN CUDA kernel on
    the C2 cloud family at 1/4 dims (D = 13).  c = 2 is the production ratio (c / D = 0.15, as c = 8 at D = 50 on the full
    grid): north_star tolerances (scatter count 1 %, image mean 0.5 %).  c = 8 is the literal cell size of the bench at a
    4x coarser window ratio: the reference estimator's own window bias (-1.8 % scatters, 1.4e-3
    image mean, reproduced by the reference kernel itself when fed the same windows) bounds it."""
    import torch

    ref = _ref_cuda()
    nx, ny, nz = 497, 338, 612
    W, H, spp = 480, 270, 96
    env, sd, sp = vp.default_sunsky()
    monkeypatch.setenv("VOLPATH_FORCE_CELL_LOG2", str(cell_log2))
    monkeypatch.setenv("VOLPATH_HALF_TABLES", "1")
    monkeypatch.setenv("VOLPATH_FORCE_RANK_DIR", "1")
    R.generate_cloud(nx, ny, nz, seed=0, bounds=vp.BOUNDS_CELL | vp.BOUNDS_VOXEL, keep_dense=True)
    assert R.volume_stats()["bound_cell_voxels"] == 1 << cell_log2 and R.half_tables() is not None
    R.set_texture_filter_mode(True)
    R.init_envmap(env)
    R.set_sun(sd, sp)
    R.copy_inv_view_matrix(vp.inv_view_matrix())
    R.precompute_opacity(sd)
    bv = torch.from_numpy(R.bounds_voxel()).cuda()
    assert ref.L.ref_init_volume_device(R.dense_volume_ptr(), bv.data_ptr(), nx, ny, nz, 0, None, None, 1) == 0
    del bv
    ref.dims, ref.quantized = (nx, ny, nz), False
    ref.set_envmap(env)
    ref.set_sun(sd, sp)
    ref.set_inv_view(vp.inv_view_matrix())
    ref.precompute_opacity(sd)
    P = vp.default_param(W, H)
    a = ref.render(P, 12, spp)
    b = ref.render(P, 12 + spp, spp)
    f = R.render(P, 12, 2 * spp, mode=vp.MODE_FAST)
    R.free_cuda_buffers()
    want = a + b
    mean_rel = abs(f[..., :3].mean() - want[..., :3].mean()) / want[..., :3].mean()
    scat_rel = abs(f[..., 3].mean() - want[..., 3].mean()) / want[..., 3].mean()
    noise = abs(a[..., :3].mean() - b[..., :3].mean()) / want[..., :3].mean()
    print("LY=2, c=%d: scatter rel %.4f, image mean rel %.5f (ref-vs-ref %.5f)" % (1 << cell_log2, scat_rel, mean_rel, noise))
    assert scat_rel <= scat_tol and mean_rel <= mean_tol + noise


def _opacity_err(R):
    want, got = R.opacity(), R.opacity_fast()
    stored = want != 0
    assert stored.mean() > 0.2
    err = np.abs(got[stored] - want[stored])
    return float(err.max()), float(err.mean()), float(want.max())


def test_swept_opacity_octets_vs_bit_faithful_march(R, vp, monkeypatch):
    """The production table (checkpoint slabs every K voxels along the sun's dominant axis, fp16 octets) against the
    bit-faithful per-voxel march of K.cu:483-524 (itself pinned by the reference golden in test_opacity_table_vs_oracle).
    C2 cloud family at 1/4 dims with K = 16: 21 interpolation levels along y, as many as the full grid has at K = 64, on a
    field that is 4x rougher per voxel (the finest noise octave spans 10 voxels here, 41 there)."""
    env, sd, sp = vp.default_sunsky()
    monkeypatch.setenv("VOLPATH_OPACITY_FAITHFUL", "1")
    monkeypatch.setenv("VOLPATH_OPACITY_K", "16")
    R.generate_cloud(497, 338, 612, seed=0, bounds=vp.BOUNDS_CELL)
    R.set_texture_filter_mode(True)
    R.set_sun(sd, sp)
    R.precompute_opacity(sd)
    mx, mean, scale = _opacity_err(R)
    print("swept opacity, 497x338x612, K=16 (21 levels): max |err| %.3e, mean %.3e of range %.3f" % (mx, mean, scale))
    assert mx <= 1e-2 * scale and mean <= 1e-3 * scale
    # an oblique sun whose dominant axis is x, negative: the sweep runs the other way along another axis
    sun = np.array([-0.8, 0.5, 0.33166], np.float32)
    R.set_sun(sun, sp)
    R.precompute_opacity(sun)
    mx, mean, scale = _opacity_err(R)
    print("swept opacity, oblique sun (-x dominant): max |err| %.3e, mean %.3e of range %.3f" % (mx, mean, scale))
    assert mx <= 1e-2 * scale and mean <= 1e-3 * scale
    R.free_cuda_buffers()


@pytest.mark.parametrize("K", [4, 64])
def test_swept_opacity_on_a_rough_small_grid(R, oracle, vp, monkeypatch, K):
    """80 x 96 x 72: the finest noise octave is 1.7 voxels, so the table itself has kinks at every voxel and interpolating
    it is at its worst: K = 4 stacks 24 levels.  The mean error stays a fraction of a percent of the range."""
    vol = small_cloud(oracle, (80, 96, 72), seed=7)
    monkeypatch.setenv("VOLPATH_OPACITY_FAITHFUL", "1")
    monkeypatch.setenv("VOLPATH_OPACITY_K", str(K))
    setup_renderer(R, vp, vol, False, True)
    R.precompute_opacity(SUN_DIR)
    mx, mean, scale = _opacity_err(R)
    print("swept opacity, rough 80x96x72, K=%d: max |err| %.3e, mean %.3e of range %.3f" % (K, mx, mean, scale))
    assert mx <= 3e-2 * scale and mean <= 3e-3 * scale


def test_deep_paths_with_the_swept_table_match_the_parity_renderer(R, oracle, vp):
    """C4-like: albedo 0.999, density 3000, frames > 10 -- most sun contributions come from the opacity table."""
    vol = small_cloud(oracle, (64, 48, 80), seed=4)
    env, sd, sp = vp.default_sunsky()
    setup_renderer(R, vp, vol, False, True, env=env)
    R.precompute_opacity(sd)
    P = vp.default_param(96, 64)
    P.albedo[:] = [0.999, 0.999, 0.999]
    P.density = 3000.0
    a = R.render(P, 11, 256, mode=vp.MODE_PARITY)
    b = R.render(P, 11 + 256, 256, mode=vp.MODE_PARITY)
    f = R.render(P, 11, 512, mode=vp.MODE_FAST)
    assert a[..., 3].max() / 256 > 20
    for sl in (np.s_[..., :3], np.s_[..., 3]):
        ma, mb, mf = a[sl].mean(), b[sl].mean(), f[sl].mean() / 2
        assert abs(mf - 0.5 * (ma + mb)) <= 3 * abs(ma - mb) + 0.004 * ma, (ma, mb, mf)


def test_work_pool_counters_are_safe_under_many_streams(R, oracle, vp):
    """launches on more streams than counter slots must not share a live counter.  40 launches round-robin
    over 5 non-blocking streams, each into its own accumulator, against the same launches run one by one."""
    import torch

    vol = small_cloud(oracle, (48, 32, 56))
    setup_renderer(R, vp, vol, False, True)
    P = vp.default_param(160, 96)
    P.density = 300.0
    streams = [torch.cuda.Stream() for _ in range(5)]
    accs = [torch.zeros(96, 160, 4, device="cuda") for _ in range(40)]
    torch.cuda.synchronize()
    for i, acc in enumerate(accs):
        R.render_kernel(acc.data_ptr(), 3 * i, P, mode=vp.MODE_FAST, n_frames=3, stream=streams[i % 5].cuda_stream)
    torch.cuda.synchronize()
    for i in (0, 7, 16, 21, 39):
        one = torch.zeros(96, 160, 4, device="cuda")
        R.render_kernel(one.data_ptr(), 3 * i, P, mode=vp.MODE_FAST, n_frames=3, stream=torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        assert torch.equal(one[..., 3], accs[i][..., 3]), i
    # every item exactly once: the scatter-count totals of all 40 launches equal one 120-frame launch
    full = torch.zeros(96, 160, 4, device="cuda")
    R.render_kernel(full.data_ptr(), 0, P, mode=vp.MODE_FAST, n_frames=120, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert torch.equal(sum(a[..., 3] for a in accs), full[..., 3])


@pytest.mark.parametrize("dims", [(32, 32, 32), (24, 40, 16)])
def test_tiny_full_box_grid_sun_clear_clip_is_conservative(R, oracle, vp, dims):
    """grids with nx <= 40 have D = 1, where the vacuum proof of the sun-clear clip needs the neighbour cells
    too; medium touching the box wall, sun grazing the exit face.  fast vs parity in the mean, incl. shadow walks."""
    from test_gpu_edge_cases import assert_same_in_the_mean

    nx, ny, nz = dims
    vol = np.full((nz, ny, nx), 0.7, np.float32)       # medium up to every wall
    vol[:, ny // 2:ny // 2 + 2] = 0.0                    # a thin vacuum sheet the shadow walks cross
    P = vp.default_param(64, 48)
    P.density = 40.0
    setup_renderer(R, vp, vol, False, True)
    assert R.volume_stats()["bound_radius_voxels"] == 1
    for sun in (SUN_DIR, np.array([0.99875, 0.05, 0.0], np.float32)):  # default, and grazing the +x face
        R.set_sun(sun, SUN_POWER)
        assert_same_in_the_mean(R, vp, P, 384, "tiny grid %s sun %s" % (dims, sun))
    R.set_sun(SUN_DIR, SUN_POWER)
