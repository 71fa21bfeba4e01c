"""Render parameters -- host-side mirror of the reference's `Param` POD (src/param.h:4-12) and of the
material presets `Mat()` (src/volumeRender.cpp:44-57, 1286-1308).  The struct is passed to the C ABI
byte-for-byte (44 bytes), exactly what the reference passes by value to its kernel."""
import ctypes


class Param(ctypes.Structure):
    _fields_ = [
        ("width", ctypes.c_uint),
        ("height", ctypes.c_uint),
        ("density", ctypes.c_float),
        ("brightness", ctypes.c_float),
        ("albedo", ctypes.c_float * 3),
        ("g", ctypes.c_float),
        ("sigma_t", ctypes.c_float * 3),
    ]

    def copy(self):
        p = Param()
        ctypes.memmove(ctypes.byref(p), ctypes.byref(self), ctypes.sizeof(Param))
        return p


assert ctypes.sizeof(Param) == 44


def default_param(width=960, height=512):
    """main()'s defaults (volumeRender.cpp:1286-1292); the 13 Mat() calls that follow end on the
    white preset (1308), which leaves albedo = sigma_t = (1,1,1)."""
    p = Param()
    p.width, p.height = width, height
    p.density, p.brightness = 800.0, 1.0
    p.albedo[:] = [1.0, 1.0, 1.0]
    p.g = 0.877
    p.sigma_t[:] = [1.0, 1.0, 1.0]
    return p


def mat(p, sx, sy, sz, ax, ay, az):
    """Mat(P, sigma_s.rgb, sigma_a.rgb) (volumeRender.cpp:44-57): sigma_t = s + a, albedo = s / sigma_t,
    sigma_t normalised by its max.  float32 arithmetic like the reference."""
    import numpy as np

    f32 = np.float32
    s = np.array([sx, sy, sz], f32)
    a = np.array([ax, ay, az], f32)
    st = (s + a).astype(f32)
    alb = (s / st).astype(f32)
    st = (st / st.max()).astype(f32)
    q = p.copy()
    q.sigma_t[:] = [float(v) for v in st]
    q.albedo[:] = [float(v) for v in alb]
    return q


# the reference's preset table (volumeRender.cpp:1296-1308), (sigma_s rgb, sigma_a rgb)
MATERIALS = [
    (2.29, 2.39, 1.97, 0.0030, 0.0034, 0.046),
    (0.15, 0.21, 0.38, 0.015, 0.077, 0.19),
    (0.19, 0.25, 0.32, 0.018, 0.088, 0.20),
    (7.38, 5.47, 3.15, 0.0002, 0.0028, 0.0163),
    (0.18, 0.07, 0.03, 0.061, 0.97, 1.45),
    (2.19, 2.62, 3.00, 0.0021, 0.0041, 0.0071),
    (0.68, 0.70, 0.55, 0.0024, 0.0090, 0.12),
    (0.70, 1.22, 1.90, 0.0014, 0.0025, 0.0142),
    (0.74, 0.88, 1.01, 0.032, 0.17, 0.48),
    (1.09, 1.59, 1.79, 0.013, 0.070, 0.145),
    (11.6, 20.4, 14.9, 0.0, 0.0, 0.0),
    (2.55, 3.21, 3.77, 0.0011, 0.0024, 0.014),
    (1.0, 1.0, 1.0, 0.0, 0.0, 0.0),
]
