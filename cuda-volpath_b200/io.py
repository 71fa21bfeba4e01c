"""Data formats on either side of the path, restated from the reference's host code:

  * volume ingest: the `.bin` format (int nx, ny, nz + nx*ny*nz floats, x fastest) written by
    vdbloader/load_vdb.cpp:52-69 and read by loadBinaryFile (src/volumeRender.cpp:915-965), and the two uchar
    quantisation rules: clamp(v, 0, 1) * 255 for `.bin` (:953-957) and max(v, 0) / max * 255 for VDB grids (:1003-1009),
    both truncating like the C cast;
  * the resolve sink: Image::dump_ppm / dump_hdr (src/image.cpp:20-111) for the float4 image the path produces;
  * an accumulator checkpoint (float4 sum + frame counter) -- the reference cannot resume (its capture() writes the
    resolved image only); since a sample is addressed by (pixel, frame), resuming is exact."""
import struct

import numpy as np


def save_bin(path, volume):
    """volume: float32 [nz, ny, nx] -> the reference's .bin (load_vdb.cpp:52-69)."""
    v = np.ascontiguousarray(volume, np.float32)
    nz, ny, nx = v.shape
    with open(path, "wb") as f:
        f.write(struct.pack("<iii", nx, ny, nz))
        f.write(v.tobytes())


def quantize_clamp(v):
    """loadBinaryFile's rule (volumeRender.cpp:956): uchar(max(0, min(v, 1)) * 255), truncating."""
    v = np.asarray(v, np.float32)
    return (np.clip(v, np.float32(0), np.float32(1)) * np.float32(255.0)).astype(np.uint8)


def quantize_by_max(v, max_value=None):
    """loadVdbFile's rule (volumeRender.cpp:1008): uchar(max(0, v) / max_value * 255), truncating."""
    v = np.asarray(v, np.float32)
    m = np.float32(v.max() if max_value is None else max_value)
    return (np.maximum(v, np.float32(0)) / m * np.float32(255.0)).astype(np.uint8)


def load_bin(path, quantized=True):
    """loadBinaryFile (volumeRender.cpp:915-965) -> [nz, ny, nx] uint8 (quantized) or float32."""
    with open(path, "rb") as f:
        hdr = f.read(12)
        if len(hdr) != 12:
            raise ValueError("short .bin header")
        nx, ny, nz = struct.unpack("<iii", hdr)
        if nx < 0 or ny < 0 or nz < 0:
            raise ValueError("Invalid resolution of file '%s'" % path)  # :929-934
        total = nx * ny * nz
        if total > 1 << 33:
            raise ValueError("Resolution too large of file '%s'" % path)  # :937-942
        data = np.frombuffer(f.read(total * 4), np.float32)
        if data.size != total:
            raise ValueError("short .bin payload")
    v = data.reshape(nz, ny, nx)
    return quantize_clamp(v) if quantized else v.copy()


def resolve(sum_image, spp):
    """sum * (1 / spp), the reference's `scale` (K.cu:2333-2346)."""
    return np.asarray(sum_image, np.float32) * np.float32(1.0 / spp)


def dump_ppm(path, image):
    """Image::dump_ppm (image.cpp:20-41): P6, rows bottom-up, uchar(min(1, c) * 255); no gamma (the caller applies
    gamma_correct first, volumeRender.cpp:477-495)."""
    img = np.asarray(image, np.float32)
    h, w = img.shape[:2]
    rgb = (np.minimum(np.float32(1.0), img[::-1, :, :3]) * np.float32(255)).astype(np.int32).astype(np.uint8)
    with open(path, "wb") as f:
        f.write(("P6\n%d %d\n255\n" % (w, h)).encode())
        f.write(rgb.tobytes())


def to_rgbe(rgb):
    """toRGBE (image.cpp:55-69): shared-exponent bytes, truncating."""
    rgb = np.asarray(rgb, np.float32)
    d = rgb.max(axis=-1)
    m, e = np.frexp(d)
    scale = np.where(d > 1e-32, m.astype(np.float32) * np.float32(256.0) / np.where(d > 1e-32, d, 1), 0).astype(np.float32)
    out = np.zeros(rgb.shape[:-1] + (4,), np.uint8)
    out[..., :3] = (rgb * scale[..., None]).astype(np.int32).astype(np.uint8)
    out[..., 3] = np.where(d > 1e-32, (e + 128).astype(np.uint8), 0)
    out[d <= 1e-32] = 0
    return out


def dump_hdr(path, image):
    """Image::dump_hdr (image.cpp:71-111): Radiance header, rows bottom-up, each row as the "new RLE" scanline
    (2, 2, width hi, width lo) with four channel planes cut into literal runs of at most 127 bytes."""
    img = np.asarray(image, np.float32)
    h, w = img.shape[:2]
    with open(path, "wb") as f:
        f.write(b"#?RADIANCE\n# Made with custom writer\nFORMAT=32-bit_rle_rgbe\nEXPOSURE=1.0\n\n")
        f.write(("-Y %d +X %d\n" % (h, w)).encode())
        for j in range(h - 1, -1, -1):
            line = to_rgbe(img[j, :, :3])
            f.write(bytes([2, 2, (w >> 8) & 0xFF, w & 0xFF]))
            for k in range(4):
                cursor = 0
                while cursor < w:
                    n = min(127, w - cursor)
                    f.write(bytes([n]))
                    f.write(line[cursor:cursor + n, k].tobytes())
                    cursor += n


def load_hdr(path):
    """Radiance .hdr reader ("new RLE" scanlines with literal AND repeat runs, "-Y h +X w" orientation) -> float32
    [h, w, 3] with row 0 = the LAST scanline of the file, i.e. the layout dump_hdr writes from (image.cpp:88) and the
    reference's HDRLoader hands to init_envmap (src/hdr/HDRloader.cpp, volumeRender.cpp:220-257)."""
    with open(path, "rb") as f:
        raw = f.read()
    head, _, rest = raw.partition(b"\n\n")
    if not head.startswith(b"#?RADIANCE"):
        raise ValueError("not a Radiance file")
    dims, _, body = rest.partition(b"\n")
    parts = dims.split()
    h, w = int(parts[1]), int(parts[3])
    out = np.zeros((h, w, 4), np.uint8)
    p = 0
    for j in range(h - 1, -1, -1):
        if body[p] != 2 or body[p + 1] != 2 or ((body[p + 2] << 8) | body[p + 3]) != w:
            raise ValueError("bad scanline header")
        p += 4
        for k in range(4):
            cursor = 0
            while cursor < w:
                n = body[p]
                if n > 128:  # repeat run: the next byte n - 128 times
                    n -= 128
                    out[j, cursor:cursor + n, k] = body[p + 1]
                    p += 2
                else:        # literal run
                    out[j, cursor:cursor + n, k] = np.frombuffer(body[p + 1:p + 1 + n], np.uint8)
                    p += 1 + n
                cursor += n
    e = out[..., 3].astype(np.int32)
    scale = np.where(e > 0, np.ldexp(np.float32(1.0), e - 128 - 8), 0).astype(np.float32)
    return out[..., :3].astype(np.float32) * scale[..., None]


def save_checkpoint(path, sum_image, next_frame, param=None):
    """float4 sum + the index of the next frame to render (+ the Param bytes it was rendered with)."""
    extra = {} if param is None else {"param": np.frombuffer(bytes(param), np.uint8)}
    np.savez(path, sum=np.asarray(sum_image, np.float32), next_frame=np.int64(next_frame), **extra)


def load_checkpoint(path):
    z = np.load(path)
    return z["sum"].copy(), int(z["next_frame"]), (z["param"].copy() if "param" in z.files else None)
