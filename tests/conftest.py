import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "ref_golden.npz"))


@pytest.fixture(scope="session")
def oracle():
    from oraclelib import Oracle

    o = Oracle()
    yield o
    o.close()


@pytest.fixture(scope="session")
def vp():
    import cuda_volpath_b200

    return cuda_volpath_b200


def param_from_bytes(vp, raw):
    import ctypes

    p = vp.Param()
    ctypes.memmove(ctypes.byref(p), bytes(raw.tobytes()), ctypes.sizeof(p))
    return p


SUN_DIR = np.array([0.0, 0.951057, -0.309017], np.float32)
SUN_POWER = np.array([51797.3, 42480.1, 32578.5], np.float32)


def setup_scene(target, vp, vol, quantized, linear, env=None):
    """Same scene calls on any of Oracle / RefHost / RefCuda (tests/oraclelib.py)."""
    target.set_volume(vol, quantized, None, linear=linear)
    target.set_envmap(vp.constant_sky() if env is None else env)
    target.set_sun(SUN_DIR, SUN_POWER)
    target.set_inv_view(vp.inv_view_matrix())


def setup_renderer(r, vp, vol, quantized, linear, env=None, **kw):
    r.init_cuda(vol, quantized, **kw)
    r.set_texture_filter_mode(linear)
    r.init_envmap(vp.constant_sky() if env is None else env)
    r.set_sun(SUN_DIR, SUN_POWER)
    r.copy_inv_view_matrix(vp.inv_view_matrix())
