#!/usr/bin/env python3
"""Bake the reference's DEFAULT sun/sky (setup_sunsky(0.5, 0.2), volumeRender.cpp:1388-1390) with the
reference's own Hosek/Tungsten model (oracle/_ref/libvolpath_ref_sunsky.so, built by
oracle/build_ref.py from /root/reference/src/sunsky) and store what the reference hands to
init_envmap / set_sun (volumeRender.cpp:285-330) as a fixture.

Only the sky half (rows j < H/2) varies; the ground half is one constant colour (H.cpp:315-321), so
the fixture keeps rows 0..H/2-1 (rgb, fp32) + the ground colour.  Run in the build container only
(it needs /root/reference through the built .so); the fixture is what travels to the GPU box.
"""
import ctypes
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def sky_state(lib, x, y):
    """The host-side state behind one sun position (oracle/ref_sunsky_driver.cpp: ref_sky_state) -- the input of the
    GPU sky bake (vp_bake_sunsky)."""
    fp, dp = ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_double)
    cfg, rad, ecf = np.zeros((11, 9)), np.zeros(11), np.zeros(11)
    lam, wts = np.zeros(10, np.float32), np.zeros((10, 3), np.float32)
    sd, sp = np.zeros(3, np.float32), np.zeros(3, np.float32)
    lib.ref_sky_state.argtypes = [ctypes.c_float, ctypes.c_float, dp, dp, dp, fp, fp, fp, fp]
    lib.ref_sky_state(x, y, cfg.ctypes.data_as(dp), rad.ctypes.data_as(dp), ecf.ctypes.data_as(dp), lam.ctypes.data_as(fp),
                      wts.ctypes.data_as(fp), sd.ctypes.data_as(fp), sp.ctypes.data_as(fp))
    return dict(configs=cfg, radiances=rad, ecf_sky=ecf, lambdas=lam, weights=wts, sun_dir=sd, sun_power=sp)


def main():
    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libvolpath_ref_sunsky.so"))
    W, H = 1024, 512
    img = np.zeros((H, W, 4), np.float32)
    sd = np.zeros(3, np.float32)
    sp = np.zeros(3, np.float32)
    fp = ctypes.POINTER(ctypes.c_float)
    lib.ref_bake_sunsky.argtypes = [ctypes.c_float, ctypes.c_float, ctypes.c_int, ctypes.c_int, fp, fp, fp]
    lib.ref_bake_sunsky(0.5, 0.2, W, H, img.ctypes.data_as(fp), sd.ctypes.data_as(fp), sp.ctypes.data_as(fp))
    sky = img[: H // 2, :, :3].copy()
    ground = img[H // 2, 0, :3].copy()
    # alpha is make_float4(c, 1) * 0.02 in the sky half, 1 in the ground half; the kernel ignores it
    assert np.all(img[H // 2:, :, :3] == ground) and np.all(img[H // 2:, :, 3] == 1.0)
    assert np.all(img[: H // 2, :, 3] == np.float32(0.02))
    out = os.path.join(ROOT, "cuda-volpath_b200", "data", "sunsky_default.npz")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    st = sky_state(lib, 0.5, 0.2)
    assert np.array_equal(st["sun_dir"], sd) and np.array_equal(st["sun_power"], sp)
    np.savez_compressed(out, sky=sky, ground=ground, sun_dir=sd, sun_power=sp, width=W, height=H,
                        **{"state_" + k: v for k, v in st.items() if k not in ("sun_dir", "sun_power")})
    # golden vectors for the sky bake (oracle/sky_oracle.py, k_bake_sunsky): state in, the reference's own map out
    gold = {}
    for tag, (x, y, w, h) in {"default": (0.5, 0.2, 128, 64), "low": (0.3, 0.9, 96, 48), "zenith": (0.8, 0.02, 64, 32)}.items():
        im = np.zeros((h, w, 4), np.float32)
        a, b = np.zeros(3, np.float32), np.zeros(3, np.float32)
        lib.ref_bake_sunsky(x, y, w, h, im.ctypes.data_as(fp), a.ctypes.data_as(fp), b.ctypes.data_as(fp))
        for k, v in sky_state(lib, x, y).items():
            gold[tag + "_" + k] = v
        gold[tag + "_env"] = im
    g = os.path.join(HERE, "sunsky_states.npz")
    np.savez_compressed(g, **gold)
    print(g, os.path.getsize(g))
    print("sun_dir", sd, "sun_power", sp, "ground", ground, "sky range", sky.min(), sky.max())
    print(out, os.path.getsize(out))


if __name__ == "__main__":
    main()
