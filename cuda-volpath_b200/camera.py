"""Camera -> the 12 floats the reference uploads with copy_inv_view_matrix (src/volumeRender.cpp:617-623):
rows 0..2 of the row-major inverse of glm::lookAt(eye, eye + forward * focus, up).  GLM is not vendored
by the reference; lookAt (right-handed) is restated analytically: with f = normalize(center -
eye), s = normalize(cross(f, up)), u = cross(s, f) the inverse view has columns (s, u, -f, eye)."""
import numpy as np

# reference defaults (volumeRender.cpp:108-112)
DEFAULT_POSITION = (3.922986, -0.782739, 0.030000)
DEFAULT_FORWARD = (-0.978148, 0.207912, 0.000000)
DEFAULT_UP = (0.207912, 0.978148, -0.000000)
DEFAULT_FOCUS = 4.0


def _normalize(v):
    return (v / np.sqrt(np.dot(v, v), dtype=np.float32)).astype(np.float32)


def inv_view_matrix(position=DEFAULT_POSITION, forward=DEFAULT_FORWARD, up=DEFAULT_UP, focus=DEFAULT_FOCUS):
    eye = np.asarray(position, np.float32)
    center = (eye + np.asarray(forward, np.float32) * np.float32(focus)).astype(np.float32)
    f = _normalize(center - eye)
    s = _normalize(np.cross(f, np.asarray(up, np.float32)).astype(np.float32))
    u = np.cross(s, f).astype(np.float32)
    m = np.empty((3, 4), np.float32)
    m[:, 0] = s
    m[:, 1] = u
    m[:, 2] = -f
    m[:, 3] = eye
    return np.ascontiguousarray(m.reshape(12))
