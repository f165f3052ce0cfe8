import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLD = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


def has_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def small_scene(oracle):
    """Seeded small textures + a 384-triangle mesh, built with the ORACLE's BVH builder."""
    from bhusie_b200 import assets
    tex = assets.small_textures()
    pts, nrm, tris = assets.uv_sphere(12, 16, radius=4.0)
    blob = oracle.new_model_blob()
    v = oracle.blob_views(blob)
    v["points"][: len(pts), :3] = pts
    v["normals"][: len(nrm), :3] = nrm
    v["triangles"][: len(tris)] = tris
    v["position"][:] = (-10.0, 0.0, 30.0)
    v["visible"][0] = 1
    hdr = blob[:48].view(np.int32)
    hdr[8], hdr[10] = len(pts), len(tris)        # point_count @32, triangle_count @40 (normal_count stays 0, Q17)
    oracle.build_bvh(blob, len(tris))
    return tex, blob, (pts, nrm, tris)


@pytest.fixture(scope="session")
def small_oracle_scene(oracle, small_scene):
    tex, blob, _ = small_scene
    return oracle.OracleScene(tex["color"], tex["disk"], tex["sky"], blob)


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)
