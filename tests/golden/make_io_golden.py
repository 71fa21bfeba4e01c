#!/usr/bin/env python3
"""Generates tests/golden/io_golden.npz from the reference's OWN file-format code (oracle/_ref/libvolpath_ref_io.so,
built by oracle/build_ref.py from src/image.cpp and the loader lines of src/volumeRender.cpp): the bytes its
Image::dump_ppm / dump_hdr write for a fixed float4 image, what its loadBinaryFile returns for a fixed .bin file, and
its two uchar quantisation rules.  Run here (the reference is not on the GPU box); the .npz is committed."""
import ctypes
import os
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
L = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libvolpath_ref_io.so"))
fp = ctypes.POINTER(ctypes.c_float)
L.ref_dump_ppm.argtypes = [fp, ctypes.c_int, ctypes.c_int, ctypes.c_char_p]
L.ref_dump_hdr.argtypes = [fp, ctypes.c_int, ctypes.c_int, ctypes.c_char_p]
L.ref_tonemap_gamma.argtypes = [fp, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_float]
L.ref_load_bin.restype = ctypes.c_longlong
L.ref_load_bin.argtypes = [ctypes.c_char_p, ctypes.POINTER(ctypes.c_int), ctypes.c_int, ctypes.c_void_p, ctypes.c_longlong]
L.ref_quantize_by_max.argtypes = [fp, ctypes.c_longlong, ctypes.c_float, ctypes.c_void_p]


def main():
    rs = np.random.RandomState(7)
    out = {}
    # an image wider than one 127-byte RLE run, with zeros, values > 1, tiny and huge values
    img = (rs.rand(6, 300, 4).astype(np.float32) ** 4 * 40).astype(np.float32)
    img[0, :5, :3] = 0
    img[1, 0, :3] = [1.0, 0.5, 0.25]
    img[2, 1, :3] = [1e-20, 3e-33, 0]
    img[3, 2, :3] = [7e4, 1.0, 1e-3]
    img[4, :, :3] = np.linspace(0, 1.2, 300, dtype=np.float32)[:, None]
    out["image"] = img
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "a.ppm").encode()
        L.ref_dump_ppm(img.ctypes.data_as(fp), 300, 6, p)
        out["ppm_bytes"] = np.frombuffer(open(p, "rb").read(), np.uint8)
        p = os.path.join(d, "a.hdr").encode()
        L.ref_dump_hdr(img.ctypes.data_as(fp), 300, 6, p)
        out["hdr_bytes"] = np.frombuffer(open(p, "rb").read(), np.uint8)
        g = img.copy()
        L.ref_tonemap_gamma(g.ctypes.data_as(fp), 300, 6, 0.03, 2.2)
        out["gamma_scale"] = np.array([0.03, 2.2], np.float32)
        out["gamma_image"] = g
        # .bin: int nx, ny, nz + floats (x fastest); values on both sides of [0, 1]
        vol = (rs.rand(5, 4, 7).astype(np.float32) * 1.6 - 0.3).astype(np.float32)
        vol.flat[:6] = [0.0, 1.0, 0.5, 0.999, 1.0 / 255, 254.5 / 255]
        p = os.path.join(d, "v.bin")
        with open(p, "wb") as f:
            f.write(np.array([7, 4, 5], np.int32).tobytes())
            f.write(vol.tobytes())
        out["bin_bytes"] = np.frombuffer(open(p, "rb").read(), np.uint8)
        dims = (ctypes.c_int * 3)()
        q = np.empty(vol.size, np.uint8)
        assert L.ref_load_bin(p.encode(), dims, 1, q.ctypes.data, q.nbytes) == vol.size and list(dims) == [7, 4, 5]
        out["bin_quantized"] = q.reshape(vol.shape)
        f32 = np.empty(vol.size, np.float32)
        assert L.ref_load_bin(p.encode(), dims, 0, f32.ctypes.data, f32.nbytes) == vol.size
        out["bin_float"] = f32.reshape(vol.shape)
    w = (rs.rand(1000).astype(np.float32) * 5 - 1).astype(np.float32)
    qm = np.empty(w.size, np.uint8)
    L.ref_quantize_by_max(w.ctypes.data_as(fp), w.size, float(w.max()), qm.ctypes.data)
    out["by_max_in"], out["by_max_out"] = w, qm
    np.savez_compressed(os.path.join(HERE, "io_golden.npz"), **out)
    print("wrote io_golden.npz:", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
