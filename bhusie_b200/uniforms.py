"""Host-side mirrors of the three small uniforms the ray pass consumes.

Byte layouts are the reference's `#[repr(C)]` structs; the bytes cross the C-ABI verbatim
(include/bh_abi.h), exactly as the reference hands them to `queue.write_buffer`
(src/renderer/mod.rs:386-388).

  CameraUniform     32 B   src/scene/camera.rs:66-73      (defaults :12-16)
  BlackHoleUniform 132 B   src/scene/blackhole.rs:37-51   (defaults :17-27, update :70-97)
  RayDetails        32 B   src/renderer/pipelines/ray_pipeline.rs:3-14 (defaults mod.rs:116-121)
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field

import numpy as np

f32 = np.float32


@dataclass
class Camera:
    """src/scene/camera.rs:3-16."""
    position: tuple = (0.0, 0.0, -19.0)
    forward: tuple = (0.0, 0.0, 1.0)
    fov: float = 1.0

    def uniform(self) -> bytes:
        """CameraUniform::update (camera.rs:85-89): pos@0, pad@12, forward@16, fov@28."""
        return struct.pack("<3fI3ff", *map(float, self.position), 0, *map(float, self.forward), float(self.fov))


@dataclass
class RayDetails:
    """ray_pipeline.rs:5-14; defaults from mod.rs:116-121 (method 0 = Euler, 1 = Cash–Karp RK)."""
    material_count: int = 0
    model_count: int = 0
    time: float = 0.0
    integration_method: int = 0
    step_size: float = 0.15
    max_iterations: int = 2000
    angle_division_threshold: float = 0.02
    highlight_interpolation: int = 0

    def uniform(self) -> bytes:
        return struct.pack("<iifififi", self.material_count, self.model_count, self.time, self.integration_method,
                           self.step_size, self.max_iterations, self.angle_division_threshold,
                           self.highlight_interpolation)


def _cross(a, b):
    return np.array([a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]], dtype=f32)


def _quat_from_euler(x: f32, y: f32, z: f32):
    """cgmath 0.18 `Quaternion::from(Euler{x,y,z})` in f32 (crate source not vendored: recalled;
    the default-scene bytes it produces are frozen in tests/golden/default_uniforms.json)."""
    half = f32(0.5)
    sx, cx = f32(np.sin(f32(x * half))), f32(np.cos(f32(x * half)))
    sy, cy = f32(np.sin(f32(y * half))), f32(np.cos(f32(y * half)))
    sz, cz = f32(np.sin(f32(z * half))), f32(np.cos(f32(z * half)))
    s = f32(f32(f32(-sx * sy) * sz) + f32(f32(cx * cy) * cz))
    vx = f32(f32(f32(sx * cy) * cz) + f32(f32(sy * sz) * cx))
    vy = f32(f32(f32(-sx * sz) * cy) + f32(f32(sy * cx) * cz))
    vz = f32(f32(f32(sx * sy) * cz) + f32(f32(sz * cx) * cy))
    return s, np.array([vx, vy, vz], dtype=f32)


def _quat_rotate(s: f32, v: np.ndarray, vec: np.ndarray) -> np.ndarray:
    """cgmath `Quaternion * Vector3`: tmp = v×vec + vec*s; (v×tmp)*2 + vec."""
    tmp = (_cross(v, vec) + vec * s).astype(f32)
    return (_cross(v, tmp) * f32(2.0) + vec).astype(f32)


@dataclass
class BlackHole:
    """src/scene/blackhole.rs:3-27."""
    position: tuple = (0.0, 0.0, 0.0)
    accretion_disk_rotation: tuple = (0.15, 0.0, 0.25)
    accretion_disk_inner: float = 2.0
    accretion_disk_outer: float = 10.0
    rotation_speed: float = 1.0
    relativity_sphere_radius: float = 20.0
    show_disk_texture: int = 1
    show_red_shift: int = 1
    feather_amount: float = 0.3

    def frame(self):
        """blackhole.rs:80-88: up = normalize(q·(0,-1,0)); right = (0,0,1)×up (not normalised, Q21);
        forward = right×up."""
        s, v = _quat_from_euler(*(f32(a) for a in self.accretion_disk_rotation))
        up = _quat_rotate(s, v, np.array([0.0, -1.0, 0.0], dtype=f32))
        mag = f32(np.sqrt(f32(f32(f32(up[0] * up[0]) + f32(up[1] * up[1])) + f32(up[2] * up[2]))))
        up = (up * f32(f32(1.0) / mag)).astype(f32)
        right = _cross(np.array([0.0, 0.0, 1.0], dtype=f32), up)
        forward = _cross(right, up)
        return right, up, forward

    def uniform(self) -> bytes:
        """BlackHoleUniform::update (blackhole.rs:70-97); 132 B incl. the 32 B Rust-side pad."""
        right, up, forward = self.frame()
        mat = [*right, 0.0, *up, 0.0, *forward, 0.0]
        return struct.pack("<4f3fi3fi12ff8i", self.accretion_disk_inner, self.accretion_disk_outer, self.rotation_speed,
                           self.relativity_sphere_radius, *map(float, self.position), self.show_disk_texture,
                           *map(float, up), self.show_red_shift, *map(float, mat), self.feather_amount, *([0] * 8))


CAMERA_UNIFORM_SIZE = 32
BLACK_HOLE_UNIFORM_SIZE = 132
RAY_DETAILS_SIZE = 32

# ModelUniform (src/renderer/triangle.rs:268-285), SURVEY.md App. B
MAX_MODEL_VERTICES = 524288
MAX_MODELS = 1
MODEL_UNIFORM_SIZE = 48234572
MU_POINTS, MU_NORMALS, MU_TRIANGLES, MU_NODES, MU_LOOKUP = 48, 8388656, 16777264, 29360176, 46137392


def pyramid_levels(base=(72, 41), multiplier=3, iters=4):
    """mod.rs:177-206: level sizes S_n = 3*S_{n-1} - 2 (f32 arithmetic, cast to u32 at use)."""
    cur = (f32(base[0]), f32(base[1]))
    out = []
    m = f32(multiplier)
    for i in range(iters):
        out.append((int(cur[0]), int(cur[1])))
        if i < iters - 1:
            cur = (f32(f32(cur[0] * m) - f32(m - f32(1.0))), f32(f32(cur[1] * m) - f32(m - f32(1.0))))
    return out
