#!/usr/bin/env python3
"""Generates tests/golden/ref_golden.npz from the REFERENCE ITSELF (oracle/_ref/libvolpath_ref_host*.so = the
reference's kernel source compiled by g++, built by oracle/build_ref.py from /root/reference).  Run in the build
container only (needs oracle/_ref); the fixture it writes is committed and pins the CPU oracle on machines where
the reference does not exist.  The reference ships no golden vectors of its own."""
import ctypes
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cuda_volpath_b200 as vp  # noqa: E402  (host-side helpers only: Param, camera, sky)
from oraclelib import RefHost  # noqa: E402

RNG_TRIPLES = [(0, 0, 0), (1, 0, 0), (0, 1, 0), (511, 511, 63), (1919, 1079, 1023), (3839, 2159, 4095)]


def small_cloud(nx, ny, nz, seed):
    """A smooth blob + noise volume in plain numpy (float32), independent of this repo's generators."""
    rs = np.random.RandomState(seed)
    z, y, x = np.meshgrid(np.linspace(-1, 1, nz), np.linspace(-1, 1, ny), np.linspace(-1, 1, nx), indexing="ij")
    r2 = (x * x + y * y + z * z).astype(np.float32)
    v = np.clip(1.2 - 1.6 * r2 + 0.35 * rs.rand(nz, ny, nx).astype(np.float32) - 0.2, 0.0, 1.0).astype(np.float32)
    return np.ascontiguousarray(v)


def scene(ref, vol, quantized, linear):
    ref.set_volume(vol, quantized, None, linear=linear)
    ref.set_envmap(vp.constant_sky())
    sd = np.array([0.0, 0.951057, -0.309017], np.float32)
    ref.set_sun(sd, np.array([51797.3, 42480.1, 32578.5], np.float32))
    ref.set_inv_view(vp.inv_view_matrix())
    return sd


def main():
    out = {}
    ref = RefHost()
    L = ref.L
    L.ref_rng_sequence.argtypes = [ctypes.c_uint, ctypes.c_uint, ctypes.c_uint, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    L.ref_hash.restype = ctypes.c_uint
    L.ref_hash.argtypes = [ctypes.c_uint]
    # RNG streams and the hash
    f = np.empty((len(RNG_TRIPLES), 16), np.float32)
    u = np.empty((len(RNG_TRIPLES), 16), np.uint32)
    for i, (x, y, fr) in enumerate(RNG_TRIPLES):
        L.ref_rng_sequence(x, y, fr, 16, f[i].ctypes.data, u[i].ctypes.data)
    out["rng_triples"] = np.array(RNG_TRIPLES, np.uint32)
    out["rng_float"], out["rng_u32"] = f, u
    seeds = np.array([0, 1, 61, 0xdeadbeef, 0xffffffff, 12345, 1 << 16, (511 << 16) | 511], np.uint32)
    out["hash_in"] = seeds
    out["hash_out"] = np.array([L.ref_hash(int(s)) for s in seeds], np.uint32)
    # local bounds (the reference's CPU routine), f32 and u8, D = 2 (nx = 50) and D = 1 (nx = 13)
    rs = np.random.RandomState(7)
    for tag, shape in (("a", (11, 9, 50)), ("b", (7, 5, 13))):
        vf = rs.rand(*shape).astype(np.float32)
        vf[vf < 0.3] = 0.0
        vu = (rs.rand(*shape) * 255).astype(np.uint8)
        out["bounds_in_f32_" + tag], out["bounds_out_f32_" + tag] = vf, ref.bounds_of(vf)
        out["bounds_in_u8_" + tag], out["bounds_out_u8_" + tag] = vu, ref.bounds_of(vu)
    # renders: 16x12 pixels, a 24x16x28 volume, float and quantised, point and linear filter
    vol = small_cloud(24, 16, 28, 3)
    out["vol_f32"] = vol
    vol8 = np.round(vol * 255).astype(np.uint8)
    P = vp.default_param(16, 12)
    P.density = 60.0
    Pc = vp.mat(P, *vp.MATERIALS[8])  # chromatic preset (volumeRender.cpp:1304)
    for name, v, q, lin, par in (("f32_point", vol, False, False, P), ("f32_linear", vol, False, True, P),
                                 ("u8_point", vol8, True, False, P), ("u8_linear", vol8, True, True, P),
                                 ("f32_linear_chroma", vol, False, True, Pc)):
        sd = scene(ref, v, q, lin)
        out["render_" + name + "_f0_2"] = ref.render(par, 0, 2)
        ref.precompute_opacity(sd)
        Pd = par.copy()
        Pd.density = 400.0  # deep paths: reaches the n > 20 opacity-table branch at frames > 10
        out["render_" + name + "_f11_2_dense"] = ref.render(Pd, 11, 2)
        if name == "f32_linear":
            out["opacity_f32_linear"] = ref.opacity()
    out["param_default"] = np.frombuffer(bytes(P), np.uint8).copy()
    out["param_chroma"] = np.frombuffer(bytes(Pc), np.uint8).copy()
    # config C1: the no-OpenVDB (Julia) build
    rj = RefHost(julia=True)
    rj.L.ref_julia_density.restype = ctypes.c_float
    rj.L.ref_julia_density.argtypes = [ctypes.c_float] * 3
    pts = (np.random.RandomState(11).rand(256, 3).astype(np.float32) * 2 - 1) * np.float32(0.8)
    out["julia_pts"] = pts
    out["julia_density"] = np.array([rj.L.ref_julia_density(*[float(c) for c in p]) for p in pts], np.float32)
    rj.set_julia()
    rj.set_envmap(vp.constant_sky())
    rj.set_sun(np.array([0.0, 0.951057, -0.309017], np.float32), np.array([51797.3, 42480.1, 32578.5], np.float32))
    rj.set_inv_view(vp.inv_view_matrix())
    Pj = vp.default_param(24, 24)
    Pj.density = 100.0
    out["param_julia"] = np.frombuffer(bytes(Pj), np.uint8).copy()
    out["render_julia_f0_2"] = rj.render(Pj, 0, 2)
    # the PASSIVE_ENVMAP 0 build: env-map importance sampling + one-sample MIS (K.cu:904-1034, 2220-2297)
    rm = RefHost(mis=True)
    rs = np.random.RandomState(5)
    env = (rs.rand(8, 16, 4).astype(np.float32) ** 3 * 4).astype(np.float32)
    env[..., 3] = 1
    env[6:, :, :3] = 0  # rows the CDF can never pick
    out["mis_env"] = env
    for name, par in (("gray", P), ("chroma", Pc)):
        rm.set_volume(vol, False, None, linear=True)
        rm.set_envmap(env)
        rm.set_sun(np.array([0.0, 0.951057, -0.309017], np.float32), np.array([51797.3, 42480.1, 32578.5], np.float32))
        rm.set_inv_view(vp.inv_view_matrix())
        out["render_mis_" + name + "_f0_3"] = rm.render(par, 0, 3)
    np.savez_compressed(os.path.join(HERE, "ref_golden.npz"), **out)
    print("wrote ref_golden.npz with", len(out), "arrays")


if __name__ == "__main__":
    main()
