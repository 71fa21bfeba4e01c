"""Sun / sky input.  The reference bakes a 1024x512 lat-long float4 env map and a sun direction / disk
radiance on the HOST with its vendored Hosek-Wilkie model (src/volumeRender.cpp:276-333) and hands them to
init_envmap / set_sun; north_star keeps that input as is.  The default configuration
(setup_sunsky(0.5, 0.2), volumeRender.cpp:1388-1390) is shipped as a fixture baked by the reference's own
code (tests/golden/make_sunsky.py); arbitrary float4 maps (e.g. from an .hdr) go through the same call."""
import os

import numpy as np

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "sunsky_default.npz")


def default_sunsky():
    """-> (env rgba float32 [H, W, 4], sun_dir float32[3], sun_power float32[3]) exactly as the reference's
    update_sunsky(baked=true) produces them (alpha 0.02 in the sky half, 1 in the ground half)."""
    z = np.load(_DATA)
    W, H = int(z["width"]), int(z["height"])
    env = np.empty((H, W, 4), np.float32)
    env[: H // 2, :, :3] = z["sky"]
    env[: H // 2, :, 3] = np.float32(0.02)
    env[H // 2:, :, :3] = z["ground"]
    env[H // 2:, :, 3] = 1.0
    return env, z["sun_dir"].astype(np.float32), z["sun_power"].astype(np.float32)


def default_sky_state():
    """The host-side state behind default_sunsky() (what the reference's Skydome holds after prepareForRender() for
    setup_sunsky(0.5, 0.2)): the input of Renderer.bake_sunsky, which evaluates the map on the device."""
    z = np.load(_DATA)
    st = {k[len("state_"):]: z[k] for k in z.files if k.startswith("state_")}
    st["sun_dir"] = z["sun_dir"].astype(np.float32)
    st["sun_power"] = z["sun_power"].astype(np.float32)
    return st


def ground_radiance(sun_dir, sun_power, ground_albedo=0.01):
    """Lower half of the baked map: ground_albedo * sun_dir.y * sun_power * (pi (0.45 / 94)^2), volumeRender.cpp:315-320."""
    k = np.pi * (0.45 / np.float32(94.0) * 0.45 / np.float32(94.0))
    return ((np.float32(ground_albedo) * np.float32(sun_dir[1]) * np.asarray(sun_power, np.float32)).astype(np.float32) * k).astype(np.float32)


def constant_sky(rgb=(0.03, 0.07, 0.23), ground=(0.03, 0.03, 0.03), width=16, height=8):
    """The reference's tiny two-colour test map (volumeRender.cpp:1372-1385)."""
    env = np.empty((height, width, 4), np.float32)
    env[:5, :, :3] = rgb
    env[5:, :, :3] = ground
    env[..., 3] = 1.0
    return env
