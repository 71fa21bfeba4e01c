"""GPU (-m gpu): the CUDA path, called through the C ABI, against
  (1) the CPU oracle (bit-exact for integer / compare work; tolerance stated per test for floating point),
  (2) the committed golden vectors generated from the reference,
  (3) the reference's own kernel rebuilt for sm_100 (oracle/_ref/libvolpath_ref_cuda.so) on the same GPU.
Nothing here reads /root/reference."""
import ctypes

import os

import numpy as np
import pytest

from conftest import SUN_DIR, SUN_POWER, param_from_bytes, setup_renderer, setup_scene

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def R(vp):
    r = vp.Renderer(0)
    yield r
    r.close()


def small_cloud(oracle, dims=(56, 40, 64), seed=5):
    return oracle.fbm_cloud(*dims, seed=seed)


def match_fraction(a, b, rtol):
    rel = np.abs(a - b) / np.maximum(np.abs(b), 1e-4)
    return float((rel.max(axis=-1) <= rtol).mean())


# ---- integer / compare work: bit-exact -----------------------------------------------------------------
def test_reference_rng_stream_bit_exact(R, oracle, golden):
    for (x, y, fr), gf, gu in zip(golden["rng_triples"], golden["rng_float"], golden["rng_u32"]):
        f, u = R.rng_sequence(int(x), int(y), int(fr), 16)
        assert np.array_equal(u, gu) and np.array_equal(f.view(np.uint32), gf.view(np.uint32))
    f, u = R.rng_sequence(123, 456, 789, 4096)
    of, ou = oracle.rng_sequence(123, 456, 789, 4096)
    assert np.array_equal(u, ou) and np.array_equal(f, of)


def test_philox2x32_known_answers(R):
    # Random123 kat_vectors, philox2x32 10 rounds
    assert R.philox2x32(0, 0, 0) == (0xff1dae59, 0x6cd10df2)
    assert R.philox2x32(0xffffffff, 0xffffffff, 0xffffffff) == (0x2c3f628b, 0xab4fd7ad)
    assert R.philox2x32(0x243f6a88, 0x85a308d3, 0x13198a2e) == (0xdd7ce038, 0xf62a4c12)


@pytest.mark.parametrize("dims", [(56, 40, 64), (41, 17, 9), (8, 8, 8), (130, 7, 33)])
def test_device_cloud_generator_bit_exact(R, oracle, vp, dims):
    R.generate_cloud(*dims, seed=9, bounds=vp.BOUNDS_CELL, keep_dense=True)
    assert np.array_equal(R.dense_volume().view(np.uint32), oracle.fbm_cloud(*dims, seed=9).view(np.uint32))


@pytest.mark.parametrize("tag", ["a", "b"])
def test_voxel_bounds_bit_exact_vs_reference_routine(R, golden, vp, tag):
    vf = golden["bounds_in_f32_" + tag]
    R.init_cuda(vf, False)
    assert np.array_equal(R.bounds_voxel().view(np.uint32), golden["bounds_out_f32_" + tag].view(np.uint32))
    vu = golden["bounds_in_u8_" + tag]
    R.init_cuda(vu, True)
    want = golden["bounds_out_u8_" + tag].astype(np.float32) / np.float32(255.0)  # normalised-float read (K.cu:261)
    assert np.array_equal(R.bounds_voxel(), want)


@pytest.mark.parametrize("dims,cell", [((56, 40, 64), 1), ((100, 20, 31), 1), ((9, 9, 9), 1), ((500, 12, 10), 2), ((1001, 9, 11), 4)])
def test_bounds_ragged_grids_vs_oracle(R, oracle, vp, dims, cell, monkeypatch):
    vol = small_cloud(oracle, dims)
    if cell > 1:
        # small grids use the reference's own windows (c = 1); force the coarse rule (c = pow2 <= D/6) for this check
        monkeypatch.setenv("VOLPATH_FORCE_CELL_LOG2", str(cell.bit_length() - 1))
    R.init_cuda(vol, False)
    bv = R.bounds_voxel()
    assert np.array_equal(bv, oracle.bounds_of(vol))
    # fast-renderer grid: cells of c^3 voxels, c = pow2 <= max(1, D/6).  Values: the reference's own window of the cell's
    # CENTRE voxel (bit-exact copy of that voxel's per-voxel bound); vacuum classification: the union of the windows of
    # all its voxels (conservative), checked through the raw view below
    assert R.volume_stats()["bound_cell_voxels"] == cell
    bc = R.bounds_cell()
    nz, ny, nx = vol.shape
    assert bc.shape[:3] == (-(-nz // cell), -(-ny // cell), -(-nx // cell))
    c = cell
    mz = np.minimum(np.arange(bc.shape[0]) * c + c // 2, nz - 1)
    my = np.minimum(np.arange(bc.shape[1]) * c + c // 2, ny - 1)
    mx = np.minimum(np.arange(bc.shape[2]) * c + c // 2, nx - 1)
    want = bv[mz][:, my][:, :, mx]
    assert np.array_equal(bc, want)
    if cell > 1:
        raw = R.bounds_cell(raw_jumps=True)[..., 0]
        union = np.zeros(bc.shape[:3], np.float32)
        for cz in range(bc.shape[0]):
            for cy in range(bc.shape[1]):
                blk = bv[cz * c:cz * c + c, cy * c:cy * c + c, :, 0]
                pad = (-blk.shape[2]) % c
                union[cz, cy] = np.pad(blk, ((0, 0), (0, 0), (0, pad)), constant_values=-np.inf).reshape(blk.shape[0], blk.shape[1], -1, c).max(axis=(0, 1, 3))
        assert ((raw > 0) >= (union > 0)).all()   # a cell with medium in ANY voxel's window is never skippable vacuum
        assert ((raw <= 0) <= (union == 0)).all()


def test_vacuum_jump_distances_are_conservative(R, oracle, vp):
    """Vacuum cells of the fast bound grid carry -jump: 0.999 (k-1) cell edges with k the chessboard distance (in
    cells) to the nearest cell with medium -- checked against scipy's exact chessboard distance transform."""
    from scipy import ndimage

    vol = small_cloud(oracle, (72, 56, 88), seed=6)
    R.init_cuda(vol, False)
    raw = R.bounds_cell(raw_jumps=True)[..., 0]
    plain = R.bounds_cell()[..., 0]
    occupied = plain > 0
    assert occupied.any() and (~occupied).any()
    assert np.array_equal(raw[occupied], plain[occupied])
    k = ndimage.distance_transform_cdt(~occupied, metric="chessboard")  # exact distance in cells, 0 on occupied
    cw = np.float32(2.0 / 72)  # world size of a voxel (= cell: D = 2 here)
    # fringe margin: a 0.05 segment + 2 voxels of footprint / slack needs 0.1056, the +-2-voxel window covers 0.0556
    margin = int(np.ceil((0.05 + 2 * cw - 2 * cw) / cw))
    assert margin == 2
    fringe = (~occupied) & (k <= margin)
    vacuum = (~occupied) & (k > margin)
    assert fringe.any() and vacuum.any()
    assert np.all(raw[fringe] == np.float32(1e-30))  # tracked like the reference: d_max floored to 1e-4
    want = np.float32(0.999) * (np.minimum(k, 64) - 1 - margin).astype(np.float32) * cw
    assert np.allclose(-raw[vacuum], want[vacuum], rtol=1e-5, atol=1e-7) and (raw[vacuum] <= 0).all()


@pytest.mark.parametrize("store", ["u8", "f32"])
@pytest.mark.parametrize("linear", [False, True])
def test_octet_store_fetch_equals_texture_emulation(R, oracle, vp, store, linear):
    vol = small_cloud(oracle, (37, 23, 41))
    quant = store == "u8"
    if quant:
        vol = np.round(vol * 255).astype(np.uint8)
    R.init_cuda(vol, quant)
    R.set_texture_filter_mode(linear)
    oracle.set_volume(vol, quant, None, linear=linear)
    rs = np.random.RandomState(1)
    lo, hi = np.array([-1, -23 / 37, -41 / 37], np.float32), np.array([1, 23 / 37, 41 / 37], np.float32)
    pos = (lo + (hi - lo) * (rs.rand(20000, 3).astype(np.float32) * 1.1 - 0.05)).astype(np.float32)  # incl. outside
    got = R.fetch_density(pos, parity=True)
    # the oracle's vol_sigma_t with density 1 through a one-pixel render is overkill: use its sampler via trace
    import ctypes as C
    L = oracle.L
    want = np.empty(len(pos), np.float32)
    L.vo_sample_density.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.c_int, C.POINTER(C.c_float)]
    L.vo_sample_density(oracle.h, pos.ctypes.data_as(C.POINTER(C.c_float)), len(pos), want.ctypes.data_as(C.POINTER(C.c_float)))
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    # the production fetch (full-precision weights, no clamp addressing: it is only ever called inside the box)
    inside = np.all((pos >= lo) & (pos <= hi), axis=1)
    fast = R.fetch_density(pos[inside], parity=False)
    assert inside.sum() > 10000
    assert np.abs(fast - want[inside]).max() <= (2.5e-3 if linear else 1e-6)  # vs 1.8 fixed-point weights


def test_f16_store_is_the_rounded_volume(R, oracle, vp):
    vol = small_cloud(oracle, (37, 23, 41))
    R.init_cuda(vol, False, store=vp.VOXEL_F16)
    R.set_texture_filter_mode(False)
    oracle.set_volume(vol.astype(np.float16).astype(np.float32), False, None, linear=False)
    pos = (np.random.RandomState(2).rand(5000, 3).astype(np.float32) - 0.5)
    import ctypes as C
    want = np.empty(len(pos), np.float32)
    oracle.L.vo_sample_density.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.c_int, C.POINTER(C.c_float)]
    oracle.L.vo_sample_density(oracle.h, pos.ctypes.data_as(C.POINTER(C.c_float)), len(pos), want.ctypes.data_as(C.POINTER(C.c_float)))
    assert np.array_equal(R.fetch_density(pos, parity=True), want)


def test_empty_and_full_volumes(R, vp):
    z = np.zeros((9, 10, 11), np.float32)
    R.init_cuda(z, False)
    assert R.volume_stats()["nonempty_bricks"] == 0
    setup_renderer(R, vp, z, False, True)
    P = vp.default_param(32, 16)
    img = R.render(P, 0, 2, mode=vp.MODE_FAST)
    img2 = R.render(P, 0, 2, mode=vp.MODE_PARITY)
    assert np.all(img[..., 3] == 0) and np.all(img2[..., 3] == 0)  # nothing scatters
    assert np.allclose(img, img2, rtol=1e-5)                        # only sky
    o = np.ones((8, 8, 8), np.float32)
    R.init_cuda(o, False)
    st = R.volume_stats()
    assert st["nonempty_bricks"] == st["bricks"] == 8


def test_opacity_table_vs_oracle(R, oracle, golden, vp):
    vol = golden["vol_f32"]
    setup_renderer(R, vp, vol, False, True)
    R.precompute_opacity(SUN_DIR)
    got = R.opacity()
    want = golden["opacity_f32_linear"]
    stored = got != 0
    # only voxels a scatter point can read are stored (9^3 apron of non-empty bricks); compare those.
    # float tolerance: the GPU contracts o + d*t into an FMA like the reference's own CUDA build; the golden is
    # the reference's host build (no contraction): sums of ~1000 samples agree to 1e-4 relative.
    assert stored.mean() > 0.25
    assert np.allclose(got[stored], want[stored], rtol=2e-4, atol=1e-6)


def test_resolve_kernels_vs_oracle(R, oracle, vp):
    import torch

    src = torch.rand(1000, 4, device="cuda") * 5
    dst = torch.empty_like(src)
    R.scale(dst.data_ptr(), src.data_ptr(), 1000, 0.125)
    assert torch.equal(dst, src * 0.125)
    R.gamma_correct(dst.data_ptr(), src.data_ptr(), 1000, 0.125, 2.2)
    want = torch.pow(src * 0.125, 1 / 2.2)
    want[:, 3] = 1
    assert torch.allclose(dst, want, rtol=2e-6)  # powf: <= 2 ulp (K.cu:2353-2356)


# ---- the path itself --------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,quantized,linear,pkey", [
    ("f32_point", False, False, "param_default"), ("f32_linear", False, True, "param_default"),
    ("u8_point", True, False, "param_default"), ("u8_linear", True, True, "param_default"),
    ("f32_linear_chroma", False, True, "param_chroma")])
def test_parity_renderer_vs_reference_golden(R, golden, vp, name, quantized, linear, pkey):
    """Fixed-seed traces against the reference's host-compiled kernel.  Tolerance: 1e-5 relative per pixel
    (north_star); a path can legitimately split from the reference when an FMA-contracted float lands on the other
    side of an accept/reject threshold (the golden is a no-FMA host build), so >= 95 % of pixels must agree."""
    vol = golden["vol_f32"]
    if quantized:
        vol = np.round(vol * 255).astype(np.uint8)
    P = param_from_bytes(vp, golden[pkey])
    setup_renderer(R, vp, vol, quantized, linear)
    got = R.render(P, 0, 2, mode=vp.MODE_PARITY)
    want = golden["render_" + name + "_f0_2"]
    assert match_fraction(got, want, 1e-5) >= 0.95
    R.precompute_opacity(SUN_DIR)
    Pd = P.copy()
    Pd.density = 400.0
    got = R.render(Pd, 11, 2, mode=vp.MODE_PARITY)
    want = golden["render_" + name + "_f11_2_dense"]
    assert match_fraction(got, want, 1e-4) >= 0.85  # long paths (> 20 scatters): more thresholds to split on


def test_parity_renderer_julia_vs_reference_golden(R, golden, vp):
    P = param_from_bytes(vp, golden["param_julia"])
    R.set_julia()
    R.init_envmap(vp.constant_sky())
    R.set_sun(SUN_DIR, SUN_POWER)
    R.copy_inv_view_matrix(vp.inv_view_matrix())
    got = R.render(P, 0, 2, mode=vp.MODE_PARITY)
    assert match_fraction(got, golden["render_julia_f0_2"], 1e-5) >= 0.95


def _ref_cuda(julia=False):
    from oraclelib import RefCuda, have_ref

    name = "libvolpath_ref_cuda%s.so" % ("_julia" if julia else "")
    if not have_ref(name):
        pytest.skip("oracle/_ref/%s not built" % name)
    return RefCuda(julia=julia)


@pytest.mark.parametrize("quantized", [False, True])
def test_parity_renderer_vs_reference_cuda_kernel_point_filter(R, oracle, vp, quantized):
    """Trace parity with the reference's OWN kernel on the same GPU (point filter: texel values are exact, so the
    only differences left are the texture unit's coordinate rounding at texel borders).  1e-5 relative."""
    ref = _ref_cuda()
    vol = small_cloud(oracle, (56, 40, 64))
    if quantized:
        vol = np.round(vol * 255).astype(np.uint8)
    env, sd, sp = vp.default_sunsky()
    P = vp.default_param(96, 64)
    P.density = 300.0
    setup_scene(ref, vp, vol, quantized, False, env=env)
    setup_renderer(R, vp, vol, quantized, False, env=env)
    # the reference's own bound volume equals ours bit for bit
    rb = ref.bounds().astype(np.float32)
    if quantized:
        rb = rb / np.float32(255.0)
    assert np.array_equal(R.bounds_voxel(), rb)
    want = ref.render(P, 0, 4)
    got = R.render(P, 0, 4, mode=vp.MODE_PARITY)
    assert np.array_equal(got[..., 3] > 0, want[..., 3] > 0) or match_fraction(got, want, 1e-5) > 0.98
    assert match_fraction(got, want, 1e-5) >= 0.98
    # pixel / sample indexing: frame k of a 4-frame batch is the same sample as a single launch of frame k
    one = R.render(P, 3, 1, mode=vp.MODE_PARITY)
    three = R.render(P, 0, 3, mode=vp.MODE_PARITY)
    assert np.array_equal((three + one)[..., 3], got[..., 3])


def test_reference_named_shims_are_a_drop_in(R, oracle, vp):
    """The 14 extern "C" names, called exactly as the reference's host code calls them, on both libraries."""
    import torch

    ref = _ref_cuda()
    L = vp.lib.load()
    vol = small_cloud(oracle, (40, 24, 48))
    env = vp.constant_sky()
    P = vp.default_param(64, 40)
    P.density = 150.0
    view = vp.inv_view_matrix()
    sd, sp = SUN_DIR.copy(), SUN_POWER.copy()
    fp = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))
    # ours, through the reference-named entry points
    L.vp_shim_set_mode(vp.MODE_PARITY)
    L.init_cuda(vol.ctypes.data, vp.lib.Extent(40, 24, 48), False, None, None)
    L.set_texture_filter_mode(False)
    L.init_envmap(env.ctypes.data, env.shape[1], env.shape[0])
    L.set_sun(fp(sd), fp(sp))
    L.copy_inv_view_matrix(fp(view), 48)
    L.init_rng(vp.lib.Dim3(8, 5, 1), vp.lib.Dim3(8, 8, 1), 64, 40)
    acc = torch.zeros(40, 64, 4, device="cuda")
    for spp in range(3):
        L.render_kernel(vp.lib.Dim3(8, 5, 1), vp.lib.Dim3(8, 8, 1), acc.data_ptr(), spp, ctypes.byref(P))
    out = torch.empty_like(acc)
    L.scale(out.data_ptr(), acc.data_ptr(), 64 * 40, 1.0 / 3)
    torch.cuda.synchronize()
    got = acc.cpu().numpy()
    # the reference, same calls
    setup_scene(ref, vp, vol, False, False)
    want = ref.render(P, 0, 3)
    assert match_fraction(got, want, 1e-5) >= 0.98
    assert torch.allclose(out, acc * (1.0 / 3))
    L.free_rng()
    L.free_cuda_buffers()


def _stat_compare(got, want, spp, what):
    """Converged-image parity: mean relative error of the image mean <= tol; per-pixel RMSE of the difference
    consistent with Monte-Carlo noise (both images are independent estimates at equal spp)."""
    g, w = got[..., :3] / spp, want[..., :3] / spp
    mean_rel = abs(g.mean() - w.mean()) / w.mean()
    scat_rel = abs(got[..., 3].mean() - want[..., 3].mean()) / max(want[..., 3].mean(), 1e-9)
    return mean_rel, scat_rel


@pytest.mark.parametrize("linear", [False, True])
def test_fast_renderer_statistical_parity_vs_reference_cuda_kernel(R, oracle, vp, linear):
    """VP_MODE_FAST uses Philox and skips vacuum, so it is equal in distribution, not per trace.  Tolerances
    (north_star): image mean within 0.5 % of the reference kernel at high spp; mean scatter count (an integer-valued,
    weight-free statistic of the same random walk) within 1 %; per-pixel RMSE between the two images no larger
    than between two independent reference renders (its own Monte-Carlo noise at equal spp) x 1.15."""
    ref = _ref_cuda()
    vol = small_cloud(oracle, (96, 64, 112), seed=2)
    env, sd, sp = vp.default_sunsky()
    P = vp.default_param(160, 96)
    P.density = 400.0
    spp = 512
    setup_scene(ref, vp, vol, False, linear, env=env)
    ref.precompute_opacity(sd)
    setup_renderer(R, vp, vol, False, linear, env=env)
    R.precompute_opacity(sd)
    a = ref.render(P, 0, spp)
    b = ref.render(P, spp, spp)
    f = R.render(P, 0, spp, mode=vp.MODE_FAST)
    assert np.isfinite(f).all()
    mean_rel, scat_rel = _stat_compare(2 * f, a + b, 1.0, "fast")
    noise_rel, _ = _stat_compare(a, b, 1.0, "ref-vs-ref")
    print("fast vs ref: mean rel %.4f (ref-vs-ref %.4f), scatter rel %.4f" % (mean_rel, noise_rel, scat_rel))
    assert mean_rel <= 0.005 + noise_rel
    assert scat_rel <= 0.01
    rm_ref = np.sqrt(np.mean(((a - b)[..., :3] / spp) ** 2))
    rm_fast = np.sqrt(np.mean(((f - a)[..., :3] / spp) ** 2))
    assert rm_fast <= 1.15 * rm_ref


@pytest.mark.parametrize("material", [8, 4])
def test_fast_renderer_chromatic_vs_reference_cuda_kernel(R, oracle, vp, material):
    """Config C3 against the reference's OWN kernel (the chromatic test compared against this repo's parity
    renderer only): `Mat` presets 8 and 4 (volumeRender.cpp:1296-1308), spectral tracking with a 3-channel throughput.
    north_star tolerances per channel: image mean 0.5 % (+ the reference's own run-to-run noise), scatter count 1 %."""
    ref = _ref_cuda()
    vol = small_cloud(oracle, (96, 64, 112), seed=2)
    env, sd, sp = vp.default_sunsky()
    P = vp.mat(vp.default_param(160, 96), *vp.MATERIALS[material])
    P.density = 400.0
    spp = 512
    setup_scene(ref, vp, vol, False, True, env=env)
    ref.precompute_opacity(sd)
    setup_renderer(R, vp, vol, False, True, env=env)
    R.precompute_opacity(sd)
    a = ref.render(P, 0, spp)
    b = ref.render(P, spp, spp)
    f = R.render(P, 0, 2 * spp, mode=vp.MODE_FAST)
    assert np.isfinite(f).all()
    for ch in range(3):
        ma, mb, mf = a[..., ch].mean(), b[..., ch].mean(), f[..., ch].mean() / 2
        assert abs(mf - 0.5 * (ma + mb)) <= 0.005 * ma + 2 * abs(ma - mb), (material, ch, ma, mb, mf)
    sa, sb, sf = a[..., 3].mean(), b[..., 3].mean(), f[..., 3].mean() / 2
    assert abs(sf - 0.5 * (sa + sb)) <= 0.01 * sa, (material, sa, sb, sf)


def test_fast_renderer_chromatic_and_high_albedo(R, oracle, vp):
    """Configs C3 (chromatic sigma_t: spectral tracking) and C4 (albedo 0.999, deep paths) against the parity
    renderer (itself trace-checked against the reference): same tolerances as above, at smaller spp."""
    vol = small_cloud(oracle, (64, 48, 80), seed=4)
    env, sd, sp = vp.default_sunsky()
    setup_renderer(R, vp, vol, False, True, env=env)
    R.precompute_opacity(sd)
    base = vp.default_param(96, 64)
    for P in (vp.mat(base, *vp.MATERIALS[8]), vp.mat(base, *vp.MATERIALS[4])):
        P.density = 300.0
        a = R.render(P, 0, 256, mode=vp.MODE_PARITY)
        f = R.render(P, 0, 256, mode=vp.MODE_FAST)
        for ch in range(3):
            ma, mf = a[..., ch].mean(), f[..., ch].mean()
            assert abs(mf - ma) <= 0.015 * ma, (ch, ma, mf)
        assert abs(f[..., 3].mean() - a[..., 3].mean()) <= 0.015 * a[..., 3].mean()
    P = base.copy()
    P.albedo[:] = [0.999, 0.999, 0.999]
    P.density = 3000.0
    a = R.render(P, 11, 128, mode=vp.MODE_PARITY)
    f = R.render(P, 11, 128, mode=vp.MODE_FAST)
    assert a[..., 3].max() / 128 > 20  # deep paths exercised (opacity-table branch)
    print("C4: scatters/path parity %.3f fast %.3f; radiance parity %.5f fast %.5f"
          % (a[..., 3].mean() / 128, f[..., 3].mean() / 128, a[..., :3].mean() / 128, f[..., :3].mean() / 128))
    assert abs(f[..., 3].mean() - a[..., 3].mean()) <= 0.02 * a[..., 3].mean()
    assert abs(f[..., :3].mean() - a[..., :3].mean()) <= 0.02 * a[..., :3].mean()


def test_fast_renderer_julia_statistical_parity(R, vp):
    ref = _ref_cuda(julia=True)
    env, sd, sp = vp.default_sunsky()
    P = vp.default_param(128, 128)
    ref.set_julia()
    ref.set_envmap(env)
    ref.set_sun(sd, sp)
    ref.set_inv_view(vp.inv_view_matrix())
    R.set_julia()
    R.init_envmap(env)
    R.set_sun(sd, sp)
    R.copy_inv_view_matrix(vp.inv_view_matrix())
    a = ref.render(P, 0, 64)
    p = R.render(P, 0, 64, mode=vp.MODE_PARITY)
    f = R.render(P, 0, 64, mode=vp.MODE_FAST)
    assert match_fraction(p, a, 1e-5) >= 0.97
    assert abs(f[..., 3].mean() - a[..., 3].mean()) <= 0.02 * a[..., 3].mean()
    assert abs(f[..., :3].mean() - a[..., :3].mean()) <= 0.02 * a[..., :3].mean()


def test_half_precision_cell_tables_round_to_the_safe_side(R, oracle, vp, monkeypatch):
    """Large volumes keep the two per-cell tables of the production renderers in half precision so that they stay
    L2-resident: max (and the negative vacuum jumps) rounded up, min down, sun-clear up -- never the other way -- and
    the renders agree with the float tables (same paths except where a majorant moved by its last half bit)."""
    vol = small_cloud(oracle, (96, 64, 80), seed=3)
    monkeypatch.setenv("VOLPATH_FORCE_CELL_LOG2", "1")
    monkeypatch.setenv("VOLPATH_HALF_TABLES", "0")
    setup_renderer(R, vp, vol, False, True)
    assert R.half_tables() is None
    P = vp.default_param(96, 64)
    P.density = 300.0
    a = R.render(P, 0, 16, mode=vp.MODE_FAST)
    monkeypatch.setenv("VOLPATH_HALF_TABLES", "1")
    setup_renderer(R, vp, vol, False, True)
    mm, cl, cf = R.half_tables()
    raw = R.bounds_cell(raw_jumps=True)
    assert (mm[..., 0].astype(np.float32) >= raw[..., 0]).all()  # max up, jumps (negative) toward zero
    assert (mm[..., 1].astype(np.float32) <= raw[..., 1]).all() and (mm[..., 1] >= 0).all()
    assert ((mm[..., 0] > 0) == (raw[..., 0] > 0)).all()          # a cell stays vacuum / medium
    assert (cl.astype(np.float32) >= cf).all()
    assert np.allclose(mm.astype(np.float32), raw, rtol=1e-3, atol=1e-7) and np.allclose(cl.astype(np.float32), cf, rtol=1e-3, atol=1e-7)
    b = R.render(P, 0, 16, mode=vp.MODE_FAST)
    same = (a[..., 3] == b[..., 3]).mean()
    assert same > 0.9, same
    # the paths that did change are fresh samples: judge the means against the paired per-pixel noise (4 sigma)
    for x, y in ((a[..., 3], b[..., 3]), (a[..., :3].sum(-1), b[..., :3].sum(-1))):
        d = (x - y).astype(np.float64)
        assert abs(d.mean()) <= 4.0 * d.std() / np.sqrt(d.size) + 1e-3 * x.mean(), (d.mean(), d.std(), x.mean())


def test_fast_renderer_frame_sharding_is_exact_in_scatter_counts(R, oracle, vp):
    """Multi-GPU contract on one GPU: the union of the strided frame subsets is the full frame set; .w (integer
    scatter counts) is identical under any partition, rgb agrees up to fp32 summation order."""
    vol = small_cloud(oracle, (48, 32, 56))
    setup_renderer(R, vp, vol, False, True)
    P = vp.default_param(64, 48)
    P.density = 200.0
    full = R.render(P, 0, 12, mode=vp.MODE_FAST)
    parts = np.zeros_like(full)
    for rank in range(4):
        first, count, stride = vp.frames_for_rank(0, 12, rank, 4)
        parts += R.render(P, first, count, mode=vp.MODE_FAST, frame_stride=stride)
    assert np.array_equal(parts[..., 3], full[..., 3])
    assert np.allclose(parts[..., :3], full[..., :3], rtol=1e-5, atol=1e-6)


def test_accumulate_combines_sharded_sums(R, oracle, vp):
    import torch

    vol = small_cloud(oracle, (48, 32, 56))
    setup_renderer(R, vp, vol, False, True)
    P = vp.default_param(64, 48)
    P.density = 200.0
    stream = torch.cuda.current_stream().cuda_stream
    parts = [torch.zeros(48, 64, 4, device="cuda") for _ in range(2)]
    for rank in range(2):
        first, count, stride = vp.frames_for_rank(0, 7, rank, 2)
        R.render_kernel(parts[rank].data_ptr(), first, P, mode=vp.MODE_FAST, n_frames=count, frame_stride=stride, stream=stream)
    vp.lib.check(R.L.vp_accumulate(R.h, parts[0].data_ptr(), parts[1].data_ptr(), 64 * 48, stream))
    full = torch.zeros(48, 64, 4, device="cuda")
    R.render_kernel(full.data_ptr(), 0, P, mode=vp.MODE_FAST, n_frames=7, stream=stream)
    torch.cuda.synchronize()
    assert torch.equal(parts[0][..., 3], full[..., 3])
    assert torch.allclose(parts[0][..., :3], full[..., :3], rtol=1e-5, atol=1e-6)


def test_fast_counters_and_launch_count(R, oracle, vp):
    vol = small_cloud(oracle, (48, 32, 56))
    setup_renderer(R, vp, vol, False, True)
    P = vp.default_param(64, 48)
    P.density = 200.0
    n0 = R.launch_count()
    R.set_stats(True)
    R.counters(reset=True)
    img = R.render(P, 0, 4, mode=vp.MODE_FAST)
    c = R.counters()
    R.set_stats(False)
    assert R.launch_count() == n0 + 1
    assert c["scatters"] == int(img[..., 3].sum())
    assert c["env_evals"] <= 64 * 48 * 4 and c["track_fetches"] > c["scatters"] > 0


@pytest.mark.parametrize("chromatic", [False, True])
def test_wavefront_form_renders_the_same_samples_as_the_megakernel(R, oracle, vp, chromatic):
    """VP_MODE_WAVE keeps ray states in shared-memory pools and batches them by event type; the Philox streams are
    addressed by (pixel, frame, draw), so every path is THE SAME path as in VP_MODE_FAST: scatter counts are equal
    exactly, radiance up to the fp32 order of the atomic adds."""
    vol = small_cloud(oracle, (64, 48, 80), seed=4)
    env, sd, sp = vp.default_sunsky()
    setup_renderer(R, vp, vol, False, True, env=env)
    R.precompute_opacity(sd)
    P = vp.default_param(100, 60)  # not a multiple of the 8x4 item tiles
    P.density = 600.0
    if chromatic:
        P = vp.mat(P, *vp.MATERIALS[8])
    a = R.render(P, 5, 24, mode=vp.MODE_FAST)
    b = R.render(P, 5, 24, mode=vp.MODE_WAVE)
    assert a[..., 3].sum() > 0 and a[..., 3].max() / 24 > 20
    assert np.array_equal(a[..., 3], b[..., 3])
    assert np.allclose(a[..., :3], b[..., :3], rtol=2e-5, atol=1e-6)


def test_wavefront_form_julia(R, vp):
    env, sd, sp = vp.default_sunsky()
    R.set_julia()
    R.init_envmap(env)
    R.set_sun(sd, sp)
    R.copy_inv_view_matrix(vp.inv_view_matrix())
    P = vp.default_param(64, 64)
    a = R.render(P, 0, 8, mode=vp.MODE_FAST)
    b = R.render(P, 0, 8, mode=vp.MODE_WAVE)
    assert np.array_equal(a[..., 3], b[..., 3]) and np.allclose(a[..., :3], b[..., :3], rtol=2e-5, atol=1e-6)


# ---- env-map importance sampling + one-sample MIS (the reference's PASSIVE_ENVMAP 0 variant) ---------------------
def _mis_env():
    rs = np.random.RandomState(5)
    env = (rs.rand(8, 16, 4).astype(np.float32) ** 3 * 4).astype(np.float32)
    env[..., 3] = 1
    env[6:, :, :3] = 0
    return env


@pytest.mark.parametrize("name,pkey", [("gray", "param_default"), ("chroma", "param_chroma")])
def test_env_sampling_parity_vs_reference_golden(R, golden, vp, name, pkey):
    P = param_from_bytes(vp, golden[pkey])
    setup_renderer(R, vp, golden["vol_f32"], False, True, env=golden["mis_env"])
    R.set_env_sampling(True)
    try:
        got = R.render(P, 0, 3, mode=vp.MODE_PARITY)
    finally:
        R.set_env_sampling(False)
    assert match_fraction(got, golden["render_mis_" + name + "_f0_3"], 1e-5) >= 0.93


def test_env_sampling_parity_vs_reference_cuda_kernel(R, oracle, vp):
    """The reference's own kernel compiled with PASSIVE_ENVMAP 0, same GPU, point filter: fixed-seed traces, 1e-5."""
    ref = _ref_cuda_mis()
    vol = small_cloud(oracle, (56, 40, 64))
    env, sd, sp = vp.default_sunsky()
    env = np.ascontiguousarray(env[::8, ::8])  # 128 x 64 map: the CDF search depth of a real map, small enough for a test
    P = vp.default_param(96, 64)
    P.density = 300.0
    setup_scene(ref, vp, vol, False, False, env=env)
    setup_renderer(R, vp, vol, False, False, env=env)
    R.set_env_sampling(True)
    try:
        want = ref.render(P, 0, 4)
        got = R.render(P, 0, 4, mode=vp.MODE_PARITY)
        assert want[..., 3].sum() > 0
        assert match_fraction(got, want, 1e-5) >= 0.97
        # production renderer, same estimator in distribution
        a = R.render(P, 0, 512, mode=vp.MODE_PARITY)
        b = R.render(P, 512, 512, mode=vp.MODE_PARITY)
        f = R.render(P, 0, 512, mode=vp.MODE_FAST)
        for sl in (np.s_[..., :3], np.s_[..., 3]):
            ma, mb, mf = a[sl].mean(), b[sl].mean(), f[sl].mean()
            assert abs(mf - 0.5 * (ma + mb)) <= 3 * abs(ma - mb) + 0.004 * ma, (ma, mb, mf)
        # the wavefront form carries the same three-stage scatter event: the SAME samples as the megakernel
        w = R.render(P, 0, 64, mode=vp.MODE_WAVE)
        f64 = R.render(P, 0, 64, mode=vp.MODE_FAST)
        # (393 k paths: the two kernels are separate compilations of the same expressions, so once in ~10^5 paths a
        # differently contracted float lands on the other side of an accept / reject threshold)
        differ = w[..., 3] != f64[..., 3]
        assert differ.sum() <= 4, (int(differ.sum()), w[..., 3].sum(), f64[..., 3].sum())
        assert np.allclose(w[..., :3][~differ], f64[..., :3][~differ], rtol=2e-5, atol=1e-6)
    finally:
        R.set_env_sampling(False)
    # the two estimators (env picked up by escaping paths / sampled at every scatter event) agree in the mean
    p = R.render(P, 0, 512, mode=vp.MODE_FAST)
    assert abs(p[..., :3].mean() - f[..., :3].mean()) <= 0.03 * p[..., :3].mean()


def _ref_cuda_mis():
    from oraclelib import RefCuda, have_ref

    if not have_ref("libvolpath_ref_cuda_mis.so"):
        pytest.skip("oracle/_ref/libvolpath_ref_cuda_mis.so not built")
    return RefCuda(mis=True)


@pytest.mark.parametrize("tag", ["default", "low", "zenith"])
def test_sky_bake_on_device_vs_reference_bake(R, vp, tag):
    """vp_bake_sunsky: the device evaluates the map the reference's host loop bakes
    (volumeRender.cpp:296-322) from the same host-side Hosek state; <= 1e-5 relative to the reference's own output."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "sunsky_states.npz"))
    st = {k: g[tag + "_" + k] for k in ("configs", "radiances", "ecf_sky", "lambdas", "weights", "sun_dir", "sun_power")}
    want = g[tag + "_env"]
    h, w = want.shape[:2]
    R.bake_sunsky(st, w, h)
    got = R.envmap()
    assert got.shape == want.shape and np.all(np.isfinite(got))
    assert np.array_equal(got[h // 2:], want[h // 2:]) and np.array_equal(got[..., 3], want[..., 3])
    rel = np.abs(got - want) / np.maximum(np.abs(want), 1e-6)
    assert rel.max() <= 1e-5, rel.max()


def test_sky_bake_default_equals_the_shipped_fixture_and_renders_the_same(R, oracle, vp):
    """The default sun/sky baked on the device equals the fixture baked by the reference's host code, and a render
    with it equals a render with the uploaded fixture up to the 1e-5 of the map."""
    env, sd, sp = vp.default_sunsky()
    vol = small_cloud(oracle, (48, 32, 56))
    setup_renderer(R, vp, vol, False, True, env=env)
    R.set_sun(sd, sp)
    P = vp.default_param(64, 48)
    P.density = 200.0
    a = R.render(P, 0, 8, mode=vp.MODE_FAST)
    R.bake_sunsky(vp.default_sky_state())
    got = R.envmap()
    assert got.shape == env.shape
    rel = np.abs(got - env) / np.maximum(np.abs(env), 1e-6)
    # <= 1e-5 relative, except where a 1-ulp difference between device and host sinf / cosf / acosf is amplified: the
    # texels next to the sun (gamma = acos(dot -> 1)) and the last rows above the horizon (exp(c / (cos(theta) + 0.01))):
    # <= 2e-4 there, and those are < 0.05 % of the map
    worst = np.unravel_index(np.argmax(rel), rel.shape)
    assert (rel <= 1e-5).mean() >= 0.9995 and rel.max() <= 2e-4, ((rel <= 1e-5).mean(), rel.max(), worst, got[worst], env[worst])
    b = R.render(P, 0, 8, mode=vp.MODE_FAST)
    assert np.array_equal(a[..., 3], b[..., 3])
    assert np.allclose(a[..., :3], b[..., :3], rtol=1e-4, atol=1e-6)
