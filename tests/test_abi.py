"""CPU: the C-ABI shared library loads and exports every symbol include/volpath.h declares; without a GPU
the product refuses to compute (no CPU fallback) and never touches oracle/."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "volpath.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b([a-z_][a-z0-9_]*)\s*\([^;{]*\)\s*;", src)
    return sorted(set(n for n in names if n not in ("defined", "visibility")))


def test_header_declares_the_reference_boundary():
    names = header_functions()
    # the 14 extern "C" entry points of the reference (volumeRender.cpp:117-128, 347-356)
    for ref_name in ["init_cuda", "set_texture_filter_mode", "free_cuda_buffers", "precompute_opacity", "init_envmap",
                     "free_envmap", "set_sun", "copy_inv_view_matrix", "copy_inv_model_matrix", "init_rng", "free_rng",
                     "scale", "gamma_correct", "render_kernel"]:
        assert ref_name in names
    assert "vp_render" in names and "vp_create" in names


def test_library_exports_every_declared_symbol(vp):
    lib = vp.lib.load()  # raises if the .so is missing or a bound symbol is absent
    raw = ctypes.CDLL(vp.lib.LIB_PATH)
    for name in header_functions():
        assert hasattr(raw, name), "libvolpath_b200.so does not export " + name
        assert name in vp.lib.SIGNATURES, "cuda-volpath_b200/lib.py does not bind " + name
    assert set(vp.lib.SIGNATURES) == set(header_functions())
    assert b"sm_100a" in lib.vp_version()


def test_param_is_the_reference_pod(vp):
    # src/param.h:4-12: uint w,h; float density, brightness; float3 albedo; float g; float3 sigma_t
    P = vp.Param
    assert ctypes.sizeof(P) == 44
    assert [(n, getattr(P, n).offset) for n, _ in P._fields_] == [
        ("width", 0), ("height", 4), ("density", 8), ("brightness", 12), ("albedo", 16), ("g", 28), ("sigma_t", 32)]
    assert ctypes.sizeof(vp.lib.Extent) == 24 and ctypes.sizeof(vp.lib.Dim3) == 12


def test_no_cpu_fallback(vp):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(vp.VolpathError, match="no CUDA device"):
        vp.Renderer(0)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "cuda-volpath_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")) or f == "Makefile":
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oraclelib" not in text and "libvolpath_oracle" not in text and "_ref/" not in text, f
