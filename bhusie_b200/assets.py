"""Scene inputs for the ray pass: the three RGBA8 textures and the OBJ mesh.

The reference embeds color.png / disk.png / sky.png with `include_bytes!`
(src/renderer/pipelines/ray_pipeline.rs:63-70), decodes them with the `image` crate to RGBA8
(src/renderer/texture.rs:15-16) and loads lucy.obj at start-up (src/scene/mod.rs:23-26).  Here
the decoded RGBA8 bytes are what crosses the C-ABI.  The real assets are staged by
tools/stage_assets.py into assets/_ref/ (git-ignored, shipped to the GPU box); if they are
missing, seeded synthetic stand-ins of the same sizes are generated and every caller reports
`source == "synthetic"`.
"""
from __future__ import annotations

import os

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ASSET_DIR = os.path.join(_ROOT, "assets", "_ref")

TEXTURE_SIZES = {"color": (256, 256), "disk": (1000, 1000), "sky": (6000, 3000)}   # (w, h)


def have_reference_assets() -> bool:
    return all(os.path.exists(os.path.join(ASSET_DIR, n)) for n in ("color.png", "disk.png", "sky.png"))


def have_lucy() -> bool:
    return os.path.exists(os.path.join(ASSET_DIR, "lucy.obj"))


def lucy_path() -> str:
    return os.path.join(ASSET_DIR, "lucy.obj")


def _decode_png(path: str) -> np.ndarray:
    from PIL import Image
    Image.MAX_IMAGE_PIXELS = None
    with Image.open(path) as im:
        return np.ascontiguousarray(np.asarray(im.convert("RGBA"), dtype=np.uint8))   # image::to_rgba8


def synthetic_sky(w: int = 6000, h: int = 3000, stars: int = 20000, seed: int = 0xB1AC) -> np.ndarray:
    """Black sky with `stars` small gaussian blobs (SURVEY §8d fallback)."""
    rng = np.random.default_rng(seed)
    img = np.zeros((h, w, 4), np.float32)
    xs = rng.integers(2, w - 2, stars)
    ys = rng.integers(2, h - 2, stars)
    amp = rng.uniform(0.4, 1.0, stars).astype(np.float32)
    tint = rng.uniform(0.7, 1.0, (stars, 3)).astype(np.float32)
    for dy in (-1, 0, 1):
        for dx in (-1, 0, 1):
            wgt = np.float32(np.exp(-(dx * dx + dy * dy) / 1.2))
            np.maximum.at(img[..., :3], (ys + dy, xs + dx), (amp[:, None] * tint) * wgt)
    img[..., 3] = 1.0
    return np.ascontiguousarray((img * 255.0 + 0.5).astype(np.uint8))


def synthetic_disk(n: int = 1000, seed: int = 0xD15C) -> np.ndarray:
    """Smooth value-noise stand-in for disk.png (r=g=b=a like the reference asset)."""
    rng = np.random.default_rng(seed)
    acc = np.zeros((n, n), np.float32)
    amp_total = 0.0
    for octave, cells in enumerate((8, 16, 32, 64)):
        g = rng.random((cells + 1, cells + 1)).astype(np.float32)
        t = np.linspace(0, cells, n, endpoint=False, dtype=np.float32)
        i = t.astype(np.int32)
        f = t - i
        f = f * f * (3 - 2 * f)
        row = g[i][:, i] * (1 - f)[None, :] + g[i][:, i + 1] * f[None, :]
        row1 = g[i + 1][:, i] * (1 - f)[None, :] + g[i + 1][:, i + 1] * f[None, :]
        amp = 0.5 ** octave
        acc += amp * (row * (1 - f)[:, None] + row1 * f[:, None])
        amp_total += amp
    v = np.clip(acc / amp_total, 0, 1)
    u8 = (v * 255.0 + 0.5).astype(np.uint8)
    return np.ascontiguousarray(np.repeat(u8[..., None], 4, axis=2))


def synthetic_color_lut(n: int = 256) -> np.ndarray:
    """Blackbody-like ramp: u (shift) from deep red to blue-white; constant in v."""
    u = np.linspace(0, 1, n, dtype=np.float32)
    r = np.clip(1.6 - u, 0, 1)
    g = np.clip(0.2 + 0.9 * u, 0, 1)
    b = np.clip(u * u * 1.2, 0, 1)
    row = np.stack([r, g, b, np.ones_like(u)], axis=1)
    img = np.repeat(row[None, :, :], n, axis=0)
    return np.ascontiguousarray((img * 255.0 + 0.5).astype(np.uint8))


_CACHE: dict = {}


def load_textures(prefer_reference: bool = True) -> tuple[dict, str]:
    """Returns ({'color','disk','sky'} -> (h,w,4) uint8, source) with source in {'reference','synthetic'}."""
    key = ("tex", prefer_reference and have_reference_assets())
    if key not in _CACHE:
        if key[1]:
            tex = {n: _decode_png(os.path.join(ASSET_DIR, f"{n}.png")) for n in ("color", "disk", "sky")}
            src = "reference"
        else:
            tex = {"color": synthetic_color_lut(), "disk": synthetic_disk(), "sky": synthetic_sky()}
            src = "synthetic"
        for n, (w, h) in TEXTURE_SIZES.items():
            assert tex[n].shape == (h, w, 4), (n, tex[n].shape)
        _CACHE[key] = (tex, src)
    return _CACHE[key]


def small_textures(seed: int = 7) -> dict:
    """Tiny seeded textures for fast CPU tests (64x64 LUT, 96x96 disk, 256x128 sky)."""
    rng = np.random.default_rng(seed)
    return {
        "color": np.ascontiguousarray(synthetic_color_lut(64)),
        "disk": np.ascontiguousarray(synthetic_disk(96, seed)),
        "sky": np.ascontiguousarray(rng.integers(0, 256, (128, 256, 4), dtype=np.uint8)),
    }


def uv_sphere(n_lat: int = 224, n_lon: int = 224, radius: float = 8.0, jitter: float = 0.02, seed: int = 0x5EED):
    """Fallback mesh: UV sphere, n_lat*n_lon*2 triangles (100 352 at 224x224), seeded
    vertex jitter.  Un-indexed like lucy.obj: 3 fresh points + 3 normals per triangle.
    Returns (points (n,3) f32, normals (n,3) f32, triangles (m,6) i32) in the reference's model space
    (i.e. after the loader's (0.5,-0.5,0.5) scale)."""
    rng = np.random.default_rng(seed)
    th = np.linspace(0, np.pi, n_lat + 1)
    ph = np.linspace(0, 2 * np.pi, n_lon + 1)
    T, P = np.meshgrid(th, ph, indexing="ij")
    R = radius * (1 + jitter * rng.standard_normal(T.shape))
    R[:, -1] = R[:, 0]
    R[0, :] = R[0, 0]
    R[-1, :] = R[-1, 0]
    grid = np.stack([R * np.sin(T) * np.cos(P), R * np.cos(T), R * np.sin(T) * np.sin(P)], axis=-1)
    a = grid[:-1, :-1].reshape(-1, 3)
    b = grid[1:, :-1].reshape(-1, 3)
    c = grid[1:, 1:].reshape(-1, 3)
    d = grid[:-1, 1:].reshape(-1, 3)
    tri_pts = np.concatenate([np.stack([a, b, c], axis=1), np.stack([a, c, d], axis=1)], axis=0)   # (m,3,3)
    pts = tri_pts.reshape(-1, 3).astype(np.float32)
    nrm = pts / np.maximum(np.linalg.norm(pts, axis=1, keepdims=True), 1e-20)
    m = tri_pts.shape[0]
    idx = np.arange(3 * m, dtype=np.int32).reshape(m, 3)
    tris = np.concatenate([idx, idx], axis=1).astype(np.int32)
    return np.ascontiguousarray(pts), np.ascontiguousarray(nrm.astype(np.float32)), np.ascontiguousarray(tris)
