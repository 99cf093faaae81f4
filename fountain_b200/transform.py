"""Host-side mirror of fountain's `Transform` (src/geometry/transform.rs:6-160).

These matrices are computed ONCE on the host and cross the C ABI as 16 column-major floats
(cgmath `Matrix4` layout, transform.rs:34-42), so nothing here is on the hot path.  The
matrices are built in float64 and rounded to float32 once.
"""
import math
import numpy as np


class Transform:
    """`Transform { t, invt }` (transform.rs:6-10).  `m`/`minv` are 4x4 row-major numpy
    arrays in the usual maths convention (m @ column-vector); `.flat()` emits cgmath's
    column-major order for the ABI."""

    __slots__ = ("m", "minv")

    def __init__(self, m, minv=None):
        self.m = np.asarray(m, dtype=np.float64).reshape(4, 4)
        self.minv = np.linalg.inv(self.m) if minv is None else np.asarray(minv, dtype=np.float64).reshape(4, 4)

    # transform.rs:18, 117-119
    @staticmethod
    def identity():
        return Transform(np.eye(4), np.eye(4))

    # transform.rs:34-42: 16 floats, column-major
    @staticmethod
    def from_flat(flat):
        return Transform(np.asarray(flat, dtype=np.float64).reshape(4, 4).T)

    # transform.rs:62-66
    @staticmethod
    def translate(delta):
        m = np.eye(4); m[:3, 3] = delta
        mi = np.eye(4); mi[:3, 3] = -np.asarray(delta, dtype=np.float64)
        return Transform(m, mi)

    # transform.rs:68-72
    @staticmethod
    def scale(sx, sy, sz):
        return Transform(np.diag([sx, sy, sz, 1.0]), np.diag([1.0 / sx, 1.0 / sy, 1.0 / sz, 1.0]))

    # transform.rs:74-78 (angle in degrees like the pbrt `Rotate` directive, pbrt.rs:580-584)
    @staticmethod
    def rotate(theta_deg, axis):
        a = np.asarray(axis, dtype=np.float64); a = a / np.linalg.norm(a)
        th = math.radians(theta_deg); s, c = math.sin(th), math.cos(th)
        x, y, z = a
        r = np.array([[c + x * x * (1 - c), x * y * (1 - c) - z * s, x * z * (1 - c) + y * s, 0],
                      [y * x * (1 - c) + z * s, c + y * y * (1 - c), y * z * (1 - c) - x * s, 0],
                      [z * x * (1 - c) - y * s, z * y * (1 - c) + x * s, c + z * z * (1 - c), 0],
                      [0, 0, 0, 1.0]])
        return Transform(r, r.T)

    # transform.rs:44-56: returns WORLD-TO-CAMERA (t = inverse of the camera frame matrix)
    @staticmethod
    def look_at(pos, look, up):
        pos = np.asarray(pos, dtype=np.float64); look = np.asarray(look, dtype=np.float64); up = np.asarray(up, dtype=np.float64)
        d = look - pos; d = d / np.linalg.norm(d)
        right = np.cross(up / np.linalg.norm(up), d); right = right / np.linalg.norm(right)
        new_up = np.cross(d, right)
        cam = np.eye(4)
        cam[:3, 0] = right; cam[:3, 1] = new_up; cam[:3, 2] = d; cam[:3, 3] = pos
        return Transform(np.linalg.inv(cam), cam)

    # transform.rs:58-60
    @staticmethod
    def camera_look_at(pos, look, up):
        return Transform.look_at(pos, look, up).inverse()

    # transform.rs:105-115
    @staticmethod
    def perspective(fov_deg, near, far):
        persp = np.array([[1, 0, 0, 0], [0, 1, 0, 0],
                          [0, 0, far / (far - near), -far * near / (far - near)],
                          [0, 0, 1, 0]], dtype=np.float64)
        inv_tan = 1.0 / math.tan(math.radians(fov_deg) / 2.0)
        return Transform.scale(inv_tan, inv_tan, 1.0) * Transform(persp)

    def inverse(self):   # transform.rs:121-123
        return Transform(self.minv, self.m)

    def __mul__(self, rhs):   # transform.rs:164-170
        return Transform(self.m @ rhs.m, rhs.minv @ self.minv)

    def then(self, nxt):   # transform.rs:130-132
        return nxt * self

    def swaps_handedness(self):   # transform.rs:126-128
        return float(np.linalg.det(self.m[:3, :3])) < 0.0

    def flat(self):
        """16 float32, column-major, for the ABI."""
        return np.ascontiguousarray(self.m.T, dtype=np.float32).reshape(16)

    def flat_inv(self):
        return np.ascontiguousarray(self.minv.T, dtype=np.float32).reshape(16)

    def is_identity(self):
        return np.array_equal(self.m, np.eye(4))

    # f32 evaluation with cgmath's operation order (see oracle/ref_math.h): used to move mesh
    # vertices to world space exactly as TriangleMesh::new does (triangle.rs:42-51).
    def apply_points_f32(self, p):
        p = np.asarray(p, dtype=np.float32)
        if self.is_identity():
            return p.copy()
        m = self.m.astype(np.float32)
        x, y, z = p[:, 0], p[:, 1], p[:, 2]
        one = np.float32(1.0)
        out = np.empty_like(p)
        cols = []
        for r in range(4):
            cols.append(((m[r, 0] * x + m[r, 1] * y) + m[r, 2] * z) + m[r, 3] * one)
        iw = one / cols[3]
        for r in range(3):
            out[:, r] = cols[r] * iw
        return out

    def apply_normals_f32(self, n):   # transform.rs:134-140: transpose of the inverse
        n = np.asarray(n, dtype=np.float32)
        if self.is_identity():
            return n.copy()
        mi = self.minv.astype(np.float32)
        x, y, z = n[:, 0], n[:, 1], n[:, 2]
        out = np.empty_like(n)
        for r in range(3):
            out[:, r] = (mi[0, r] * x + mi[1, r] * y) + mi[2, r] * z
        return out
