"""oracle/sky_oracle.py -- TEST INFRASTRUCTURE (never imported by the product).

numpy restatement of the reference's per-texel sun/sky bake, i.e. of what update_sunsky(baked = true) computes for
every texel of the lat-long environment map (src/volumeRender.cpp:296-322) through Skydome::skyColor
(src/sunsky/sky_tungsten.cpp:400-419), arhosekskymodel_radiance and ArHosekSkyModel_GetRadianceInternal
(src/sunsky/hosek/ArHosekSkyModel.cpp:519-561, 291-304).  The INPUT is the host-side state the reference holds after
Skydome::prepareForRender() (cooked configurations of the 11 spectral bands etc.) -- the table interpolation that
produces it stays host code and is not restated.  Pinned by tests/golden/sunsky_states.npz, which was baked by the
reference's own code (tests/golden/make_sunsky.py)."""
import numpy as np

NUM_SAMPLES_VALID = 7  # sky_tungsten.cpp:367


def radiance_internal(cfg, theta, gamma):
    """ArHosekSkyModel_GetRadianceInternal (ArHosekSkyModel.cpp:291-304), double precision."""
    cg = np.cos(gamma)
    ct = np.cos(theta)
    exp_m = np.exp(cfg[4] * gamma)
    ray_m = cg * cg
    mie_m = (1.0 + cg * cg) / np.power(1.0 + cfg[8] * cfg[8] - 2.0 * cfg[8] * cg, 1.5)
    zenith = np.sqrt(ct)
    return (1.0 + cfg[0] * np.exp(cfg[1] / (ct + 0.01))) * (cfg[2] + cfg[3] * exp_m + cfg[5] * ray_m + cfg[6] * mie_m + cfg[7] * zenith)


def spectral_radiance(state, theta, gamma, wavelength):
    """arhosekskymodel_radiance (ArHosekSkyModel.cpp:519-561)."""
    low = int((wavelength - 320.0) / 40.0)
    if low < 0 or low >= 11:
        return np.zeros_like(theta)
    interp = np.fmod((wavelength - 320.0) / 40.0, 1.0)
    val = radiance_internal(state["configs"][low], theta, gamma) * state["radiances"][low] * state["ecf_sky"][low]
    if interp < 1e-6:
        return val
    res = (1.0 - interp) * val
    if low + 1 < 11:
        res = res + interp * radiance_internal(state["configs"][low + 1], theta, gamma) * state["radiances"][low + 1] * state["ecf_sky"][low + 1]
    return res


def xyz_to_rgb(x, y, z):
    """Spectral::xyzToRgb (sky_tungsten.cpp:318-323)."""
    f = np.float32
    return (f(3.240479) * x + f(-1.537150) * y + f(-0.498535) * z,
            f(-0.969256) * x + f(1.875991) * y + f(0.041556) * z,
            f(0.055648) * x + f(-0.204043) * y + f(1.057311) * z)


def bake_sunsky(state, width, height, sunsky_scale=np.float32(0.02), ground_albedo=np.float32(0.01)):
    """-> float32 [height][width][4]: the map update_sunsky(baked = true) hands to init_envmap (H.cpp:296-322)."""
    f = np.float32
    i = np.arange(width, dtype=np.float32)[None, :]
    j = np.arange(height // 2, dtype=np.float32)[:, None]
    phi = (i / f(width) * f(2) * np.pi).astype(np.float32)      # float(i) / hdrwidth * 2 * M_PI  -> float
    th = ((j / f(height)) * np.pi).astype(np.float32)
    d = (np.sin(th) * np.sin(phi), np.cos(th) + 0 * phi, np.sin(th) * -np.cos(phi))  # must match Envmap::uv_to_dir
    d = [c.astype(np.float32) for c in d]
    sun = np.asarray(state["sun_dir"], np.float32)
    theta = np.arccos(d[1])                                     # skyColor: acos(direction.y)
    dot = (d[0] * sun[0] + d[1] * sun[1] + d[2] * sun[2]).astype(np.float32)
    gamma = np.clip(np.arccos(np.clip(dot, f(-1), f(1))) * f(state.get("gamma_scale", 1.0)), f(0), f(np.pi)).astype(np.float32)
    xyz = [np.zeros_like(theta, dtype=np.float32) for _ in range(3)]
    for k in range(NUM_SAMPLES_VALID):
        r = spectral_radiance(state, theta.astype(np.float64), gamma.astype(np.float64), float(state["lambdas"][k])).astype(np.float32)
        for c in range(3):
            xyz[c] = (xyz[c] + f(state["weights"][k][c]) * r).astype(np.float32)
    rgb = xyz_to_rgb(*xyz)
    out = np.empty((height, width, 4), np.float32)
    for c in range(3):
        out[: height // 2, :, c] = rgb[c] * sunsky_scale
    out[: height // 2, :, 3] = sunsky_scale
    sp = np.asarray(state["sun_power"], np.float32)
    # ground half: ground_albedo * sun_dir.y * sun_power * (M_PI * (0.45 / 94.0f * 0.45 / 94.0f)) in double (H.cpp:315-320)
    k = np.pi * (0.45 / np.float32(94.0) * 0.45 / np.float32(94.0))
    out[height // 2:, :, :3] = ((ground_albedo * sun[1] * sp).astype(np.float32) * k).astype(np.float32)
    out[height // 2:, :, 3] = 1.0
    return out
