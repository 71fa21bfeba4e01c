"""CPU: the data formats either side of the path: .bin volume ingest + the reference's two uchar
quantisation rules, the ppm / Radiance-hdr sinks of the float4 image, accumulator checkpoint + exact resume."""
import numpy as np
import pytest

from conftest import setup_scene


def test_bin_round_trip_and_quantisation_rules(tmp_path, vp):
    io = vp.io
    rs = np.random.RandomState(0)
    v = (rs.rand(5, 4, 7).astype(np.float32) * 1.4 - 0.2).astype(np.float32)  # values outside [0, 1] on both sides
    p = str(tmp_path / "v.bin")
    io.save_bin(p, v)
    raw = open(p, "rb").read()
    assert np.frombuffer(raw[:12], np.int32).tolist() == [7, 4, 5] and len(raw) == 12 + 4 * v.size
    assert np.array_equal(io.load_bin(p, quantized=False), v)
    q = io.load_bin(p, quantized=True)
    assert q.dtype == np.uint8 and q.shape == v.shape
    # volumeRender.cpp:956 -- VolumeType(max(0, min(v, 1)) * 255.0f): truncation, 1.0 -> 255, negatives -> 0
    for val, want in [(-0.3, 0), (0.0, 0), (0.5, 127), (0.999, 254), (1.0, 255), (1.7, 255)]:
        assert io.quantize_clamp(np.float32(val)) == want
    # volumeRender.cpp:1008 -- max(0, v) / max_value * 255
    w = np.array([-1.0, 0.0, 2.0, 4.0], np.float32)
    assert io.quantize_by_max(w).tolist() == [0, 0, 127, 255]
    with pytest.raises(ValueError):
        open(p, "wb").write(np.array([-1, 2, 2], np.int32).tobytes())
        io.load_bin(p)


def test_ppm_is_the_reference_layout(tmp_path, vp):
    img = np.zeros((2, 3, 4), np.float32)
    img[0, 0, :3] = [0.0, 0.5, 1.0]
    img[1, 2, :3] = [2.0, 0.25, 0.999]
    p = str(tmp_path / "a.ppm")
    vp.io.dump_ppm(p, img)
    raw = open(p, "rb").read()
    assert raw.startswith(b"P6\n3 2\n255\n")
    px = np.frombuffer(raw[len(b"P6\n3 2\n255\n"):], np.uint8).reshape(2, 3, 3)
    assert px[1, 0].tolist() == [0, 127, 255]      # image row 0 is written LAST (bottom-up, image.cpp:31)
    assert px[0, 2].tolist() == [255, 63, 254]     # min(1, c) * 255, truncated


def test_hdr_round_trip_within_rgbe_precision(tmp_path, vp):
    rs = np.random.RandomState(1)
    img = (rs.rand(5, 300, 4).astype(np.float32) ** 4 * 50).astype(np.float32)  # wider than one 127-byte run
    img[0, 0, :3] = 0
    p = str(tmp_path / "a.hdr")
    vp.io.dump_hdr(p, img)
    raw = open(p, "rb").read()
    assert raw.startswith(b"#?RADIANCE\n") and b"FORMAT=32-bit_rle_rgbe" in raw and b"-Y 5 +X 300\n" in raw
    back = vp.io.load_hdr(p)
    mx = img[..., :3].max(axis=-1, keepdims=True)
    assert np.all(np.abs(back - img[..., :3]) <= mx / 128 + 1e-30)  # 8-bit mantissa, truncating
    assert np.all(back[0, 0] == 0)
    rgbe = vp.io.to_rgbe(np.array([[1.0, 0.5, 0.25]], np.float32))
    assert rgbe.tolist() == [[128, 64, 32, 129]]  # frexp(1) = 0.5 * 2^1 -> scale 128, exponent 1 + 128


def test_checkpoint_resume_is_exact(tmp_path, vp, oracle, golden):
    """(pixel, frame) addresses a sample, so sum(frames 0..5) == resume(checkpoint(frames 0..2), frames 3..5) bit for bit
    (checked on the CPU oracle: same additions in the same order)."""
    setup_scene(oracle, vp, golden["vol_f32"], False, True)
    P = vp.default_param(16, 12)
    P.density = 60.0
    full = oracle.render(P, 0, 6)
    part = oracle.render(P, 0, 3)
    p = str(tmp_path / "ck.npz")
    vp.io.save_checkpoint(p, part, 3, P)
    acc, nxt, pb = vp.io.load_checkpoint(p)
    assert nxt == 3 and bytes(pb) == bytes(P)
    resumed = oracle.render(P, nxt, 3, accum=acc)
    assert np.array_equal(resumed.view(np.uint32), full.view(np.uint32))
    assert np.allclose(vp.io.resolve(full, 6), full / 6)


def test_hdr_reader_handles_repeat_runs(tmp_path, vp):
    """A hand-built Radiance file with run-length runs (count byte > 128), as real .hdr environment maps use."""
    w, h = 10, 2
    rgbe = np.zeros((h, w, 4), np.uint8)
    rgbe[0, :, :] = [128, 64, 32, 129]          # constant row -> one repeat run per channel
    rgbe[1, :5, :] = [10, 20, 30, 130]
    rgbe[1, 5:, :] = [200, 100, 50, 127]
    body = b""
    for j in range(h):
        body += bytes([2, 2, 0, w])
        for k in range(4):
            row = rgbe[j, :, k]
            if j == 0:
                body += bytes([128 + w, int(row[0])])
            else:
                body += bytes([128 + 5, int(row[0]), 5]) + row[5:].tobytes()
    p = str(tmp_path / "rle.hdr")
    open(p, "wb").write(b"#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n-Y 2 +X 10\n" + body)
    img = vp.io.load_hdr(p)
    assert img.shape == (2, 10, 3)
    # file scanline 0 is image row h-1 (rows are stored top-down in the file, bottom-up in memory)
    assert np.allclose(img[1, 3], np.array([128, 64, 32]) * 2.0 ** (129 - 136))
    assert np.allclose(img[0, 2], np.array([10, 20, 30]) * 2.0 ** (130 - 136))
    assert np.allclose(img[0, 7], np.array([200, 100, 50]) * 2.0 ** (127 - 136))


# ---- pinned by the reference's own file-format code (tests/golden/io_golden.npz, made by tests/golden/make_io_golden.py
# ---- from src/image.cpp and the loader lines of src/volumeRender.cpp compiled by oracle/build_ref.py) -------------------
@pytest.fixture(scope="module")
def io_golden():
    import os

    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "io_golden.npz"))


def test_ppm_and_hdr_bytes_equal_the_reference_writers(tmp_path, vp, io_golden):
    """Image::dump_ppm / Image::dump_hdr (image.cpp:20-111): byte-identical files for the same float4 image."""
    img = io_golden["image"]
    p = str(tmp_path / "a.ppm")
    vp.io.dump_ppm(p, img)
    assert np.array_equal(np.frombuffer(open(p, "rb").read(), np.uint8), io_golden["ppm_bytes"])
    p = str(tmp_path / "a.hdr")
    vp.io.dump_hdr(p, img)
    assert np.array_equal(np.frombuffer(open(p, "rb").read(), np.uint8), io_golden["hdr_bytes"])
    # and our reader takes the reference's file back within RGBE precision
    ref_file = str(tmp_path / "ref.hdr")
    open(ref_file, "wb").write(io_golden["hdr_bytes"].tobytes())
    back = vp.io.load_hdr(ref_file)
    mx = img[..., :3].max(axis=-1, keepdims=True)
    assert np.all(np.abs(back - img[..., :3]) <= mx / 128 + 1e-30)


def test_bin_loader_and_quantisation_rules_equal_the_reference(tmp_path, vp, io_golden):
    """loadBinaryFile (volumeRender.cpp:915-965) and the loadVdbFile uchar rule (volumeRender.cpp:1003-1009)."""
    p = str(tmp_path / "v.bin")
    open(p, "wb").write(io_golden["bin_bytes"].tobytes())
    assert np.array_equal(vp.io.load_bin(p, quantized=True), io_golden["bin_quantized"])
    f = vp.io.load_bin(p, quantized=False)
    assert np.array_equal(f.view(np.uint32), io_golden["bin_float"].view(np.uint32))
    # our writer produces the file the reference's loader was given
    vp.io.save_bin(str(tmp_path / "w.bin"), io_golden["bin_float"])
    assert open(str(tmp_path / "w.bin"), "rb").read() == io_golden["bin_bytes"].tobytes()
    w = io_golden["by_max_in"]
    assert np.array_equal(vp.io.quantize_by_max(w), io_golden["by_max_out"])


def test_gamma_resolve_equals_the_reference_tonemap(vp, io_golden):
    """Image::scale + Image::tonemap_gamma (image.cpp:189-209) on the host; the device kernels (K.cu:2333-2362) are
    checked against the same formula in tests/test_gpu_parity.py::test_resolve_kernels_vs_oracle."""
    s, g = [float(v) for v in io_golden["gamma_scale"]]
    img = io_golden["image"]
    want = io_golden["gamma_image"]
    got = np.power(np.clip(img[..., :3] * np.float32(s), 0, 1), np.float32(1.0 / g), dtype=np.float32)
    assert np.allclose(got, want[..., :3], rtol=3e-6, atol=1e-7)
