"""CPU: the oracle (oracle/volpath_oracle.cpp) against golden vectors generated from the REFERENCE ITSELF
(tests/golden/make_golden.py, reference kernel source compiled by g++).  Bit-exact: both sides are CPU builds with
no FMA contraction and the same libm."""
import os

import numpy as np
import pytest

from conftest import SUN_DIR, param_from_bytes, setup_scene

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_hash_and_rng_streams(oracle, golden):
    for s, h in zip(golden["hash_in"], golden["hash_out"]):
        assert oracle.L.vo_hash(int(s)) == int(h)
    for (x, y, fr), gf, gu in zip(golden["rng_triples"], golden["rng_float"], golden["rng_u32"]):
        f, u = oracle.rng_sequence(int(x), int(y), int(fr), 16)
        assert np.array_equal(u, gu)
        assert np.array_equal(f.view(np.uint32), gf.view(np.uint32))
        assert (f >= 0).all() and (f < 1).all()


@pytest.mark.parametrize("tag", ["a", "b"])
def test_local_bounds_match_reference_routine(oracle, golden, tag):
    for kind in ("f32", "u8"):
        vol = golden["bounds_in_%s_%s" % (kind, tag)]
        want = golden["bounds_out_%s_%s" % (kind, tag)]
        got = oracle.bounds_of(vol)
        assert np.array_equal(got, want)
        # and both equal the brute-force clamped cube
        D = oracle.L.vo_bound_radius_voxels(vol.shape[2], 0.05)
        assert np.array_equal(oracle.bounds_of(vol, brute_D=D), want)


def test_bound_radius():
    from oraclelib import Oracle

    L = Oracle().L
    assert L.vo_bound_radius_voxels(50, 0.05) == 2
    assert L.vo_bound_radius_voxels(13, 0.05) == 1
    assert L.vo_bound_radius_voxels(1987, 0.05) == 50  # config C2


def test_julia_density(oracle, golden):
    got = np.array([oracle.L.vo_julia_density(*[float(c) for c in p]) for p in golden["julia_pts"]], np.float32)
    assert np.array_equal(got, golden["julia_density"])
    assert 0 < got.sum() < len(got)


@pytest.mark.parametrize("name,quantized,linear,pkey", [
    ("f32_point", False, False, "param_default"), ("f32_linear", False, True, "param_default"),
    ("u8_point", True, False, "param_default"), ("u8_linear", True, True, "param_default"),
    ("f32_linear_chroma", False, True, "param_chroma")])
def test_render_matches_reference_bit_for_bit(oracle, golden, vp, name, quantized, linear, pkey):
    vol = golden["vol_f32"]
    if quantized:
        vol = np.round(vol * 255).astype(np.uint8)
    P = param_from_bytes(vp, golden[pkey])
    setup_scene(oracle, vp, vol, quantized, linear)
    got = oracle.render(P, 0, 2)
    want = golden["render_" + name + "_f0_2"]
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    assert want[..., 3].sum() > 0  # the scene scatters
    # deep paths + the precomputed sun-opacity branch (frames > 10, more than 20 scatters)
    oracle.precompute_opacity(SUN_DIR)
    Pd = P.copy()
    Pd.density = 400.0
    got = oracle.render(Pd, 11, 2)
    want = golden["render_" + name + "_f11_2_dense"]
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    assert want[..., 3].max() > 20
    if name == "f32_linear":
        assert np.array_equal(oracle.opacity().view(np.uint32), golden["opacity_f32_linear"].view(np.uint32))


def test_julia_render_matches_reference(oracle, golden, vp):
    from conftest import SUN_POWER

    P = param_from_bytes(vp, golden["param_julia"])
    oracle.set_julia()
    oracle.set_envmap(vp.constant_sky())
    oracle.set_sun(SUN_DIR, SUN_POWER)
    oracle.set_inv_view(vp.inv_view_matrix())
    got = oracle.render(P, 0, 2)
    assert np.array_equal(got.view(np.uint32), golden["render_julia_f0_2"].view(np.uint32))


def test_single_path_trace_equals_image_pixel(oracle, golden, vp):
    vol = golden["vol_f32"]
    P = param_from_bytes(vp, golden["param_default"])
    setup_scene(oracle, vp, vol, False, False)
    img = oracle.render(P, 5, 1)
    for (x, y) in [(0, 0), (8, 6), (15, 11), (7, 3)]:
        assert np.array_equal(oracle.trace_path(P, x, y, 5), img[y, x])


def test_philox4x32_known_answers(oracle):
    # Random123 kat_vectors (Salmon et al., SC'11): philox4x32 10 rounds
    assert [hex(v) for v in oracle.philox([0, 0, 0, 0], [0, 0])] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]
    assert [hex(v) for v in oracle.philox([0xffffffff] * 4, [0xffffffff] * 2)] == ["0x408f276d", "0x41c83b0e", "0xa20bc7c6", "0x6d5451fd"]


def test_fbm_cloud_is_deterministic_and_cloudlike(oracle):
    a = oracle.fbm_cloud(40, 28, 48, seed=3)
    b = oracle.fbm_cloud(40, 28, 48, seed=3)
    assert np.array_equal(a, b)
    assert a.min() == 0.0 and 0.5 < a.max() <= 1.0
    occ = float((a > 0).mean())
    assert 0.1 < occ < 0.5


def test_resolve_kernels(oracle):
    src = np.random.RandomState(0).rand(37, 4).astype(np.float32)
    dst = np.empty_like(src)
    oracle.L.vo_scale(dst.ctypes.data_as(oracle.L.vo_scale.argtypes[0]), src.ctypes.data_as(oracle.L.vo_scale.argtypes[1]), 37, 0.25)
    assert np.array_equal(dst, src * np.float32(0.25))


@pytest.mark.parametrize("name,pkey", [("gray", "param_default"), ("chroma", "param_chroma")])
def test_env_importance_sampling_mis_variant_bit_for_bit(golden, vp, name, pkey):
    """PASSIVE_ENVMAP 0 (compiled out in the reference as shipped, K.cu:21): CDF tables built like init_envmap does
    (K.cu:1144-1210), lower-bound sampling, one-sample MIS between phase and env-map sampling (K.cu:2220-2297)."""
    from oraclelib import Oracle

    o = Oracle()
    o.set_env_sampling(True)
    setup_scene(o, vp, golden["vol_f32"], False, True, env=golden["mis_env"])
    P = param_from_bytes(vp, golden[pkey])
    got = o.render(P, 0, 3)
    want = golden["render_mis_" + name + "_f0_3"]
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    # and it is a different estimator of the same image: the passive variant differs per sample
    o.set_env_sampling(False)
    assert not np.array_equal(o.render(P, 0, 3), want)
    o.close()


@pytest.mark.parametrize("tag", ["default", "low", "zenith"])
def test_sky_bake_restatement_vs_reference_bake(tag):
    """oracle/sky_oracle.py (the per-texel loop of update_sunsky + Skydome::skyColor + the Hosek spectral radiance)
    against maps baked by the reference's own code (tests/golden/make_sunsky.py): <= 5e-6 relative (float vs double
    libm differences), ground half and alpha exact."""
    import sys

    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import sky_oracle

    g = np.load(os.path.join(ROOT, "tests", "golden", "sunsky_states.npz"))
    st = {k: g[tag + "_" + k] for k in ("configs", "radiances", "ecf_sky", "lambdas", "weights", "sun_dir", "sun_power")}
    want = g[tag + "_env"]
    h, w = want.shape[:2]
    got = sky_oracle.bake_sunsky(st, w, h)
    assert np.array_equal(got[h // 2:], want[h // 2:]) and np.array_equal(got[..., 3], want[..., 3])
    assert np.all(np.isfinite(got))
    rel = np.abs(got - want) / np.maximum(np.abs(want), 1e-6)
    assert rel.max() <= 5e-6, rel.max()
