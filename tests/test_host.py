"""CPU tests of the product's host side: byte layouts, the C-ABI surface, the C++ BVH/OBJ code
(checked against the oracle's independent restatement), band layout, error behaviour."""
import ctypes as C
import json
import os
import re

import numpy as np
import pytest

from conftest import GOLD, ROOT, has_gpu
from bhusie_b200 import _lib, assets, pipelines as P, uniforms as U
from bhusie_b200.multi import BandLayout


def test_uniform_layouts():
    cam = U.Camera().uniform()
    assert len(cam) == 32
    f = np.frombuffer(cam, np.float32)
    assert tuple(f[0:3]) == (0.0, 0.0, -19.0) and tuple(f[4:7]) == (0.0, 0.0, 1.0) and f[7] == 1.0     # camera.rs:12-16,66-73
    det = U.RayDetails().uniform()
    assert len(det) == 32
    i = np.frombuffer(det, np.int32); g = np.frombuffer(det, np.float32)
    assert i[3] == 0 and g[4] == np.float32(0.15) and i[5] == 2000 and g[6] == np.float32(0.02)          # mod.rs:116-121
    hole = U.BlackHole().uniform()
    assert len(hole) == 132
    h = np.frombuffer(hole, np.float32); hi = np.frombuffer(hole, np.int32)
    assert tuple(h[0:4]) == (2.0, 10.0, 1.0, 20.0) and hi[7] == 1 and hi[11] == 1 and h[24] == np.float32(0.3)
    up = h[8:11]; right = h[12:15]; upm = h[16:19]; fwd = h[20:23]
    assert np.array_equal(up, upm)                                   # normal == matrix column 1 (blackhole.rs:90-96)
    assert np.isclose(np.linalg.norm(up), 1.0, atol=1e-6)
    assert np.allclose(right, np.cross([0, 0, 1], up), atol=1e-7)    # right = (0,0,1) x up, not normalised
    assert np.allclose(fwd, np.cross(right, up), atol=1e-7)
    assert np.all(hi[25:33] == 0)
    gold = json.load(open(os.path.join(GOLD, "default_uniforms.json")))
    assert cam.hex() == gold["camera_default"] and hole.hex() == gold["black_hole_default"] and det.hex() == gold["ray_details_default"]


def test_pyramid_sizes():
    assert U.pyramid_levels() == [(72, 41), (214, 121), (640, 361), (1918, 1081)]      # mod.rs:177-206
    assert U.pyramid_levels((9, 5), 3, 3) == [(9, 5), (25, 13), (73, 37)]


def test_abi_exports_every_declared_symbol():
    """The shared library loads on a CPU-only box and exports exactly what include/bh_abi.h declares."""
    hdr = open(os.path.join(ROOT, "include", "bh_abi.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(bh_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.bh_abi_version() == _lib.ABI_VERSION == 2
    # struct sizes the ABI promises
    assert C.sizeof(_lib.PassStats) == 72 and C.sizeof(_lib.ModelInfo) == 28


@pytest.mark.skipif(has_gpu(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    with pytest.raises(_lib.BhError) as e:
        P.Context(0)
    assert e.value.code == -19 and "no CPU fallback" in str(e.value)


def test_host_error_codes(tmp_path):
    lib = _lib.load()
    blob = np.zeros(U.MODEL_UNIFORM_SIZE, np.uint8)
    info = _lib.ModelInfo()
    assert lib.bh_model_load_obj(str(tmp_path / "nope.obj").encode(), blob.ctypes.data_as(C.c_void_p), C.byref(info)) == -2
    assert b"cannot open" in lib.bh_last_error()
    assert lib.bh_model_build_bvh(None, 3, C.byref(info)) == -22
    assert lib.bh_model_build_bvh(blob.ctypes.data_as(C.c_void_p), U.MAX_MODEL_VERTICES + 1, C.byref(info)) == -22
    quad = tmp_path / "quad.obj"
    quad.write_text("v 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nf 1 2 3 4\n")
    assert lib.bh_model_load_obj(str(quad).encode(), blob.ctypes.data_as(C.c_void_p), C.byref(info)) == -22
    with pytest.raises(_lib.BhError):
        P.model_from_arrays(np.zeros((3, 3)), np.zeros((1, 3)), np.array([[0, 1, 5, 0, 0, 0]]))     # index out of range
    assert lib.bh_ray_pipeline_local_rows(None) == 0 and lib.bh_ray_pipeline_output(None) is None


def test_model_builder_matches_oracle_small(oracle, small_scene):
    """Product C++ builder (iterative) vs oracle C builder (recursive, literal triangle.rs): same bytes."""
    _, oracle_blob, (pts, nrm, tris) = small_scene
    blob, info = P.model_from_arrays(pts, nrm, tris)
    assert np.array_equal(blob, oracle_blob)
    assert info["triangle_count"] == len(tris) and info["nodes_used"] > len(tris) // 2


def test_model_builder_matches_oracle_random(oracle):
    rng = np.random.default_rng(11)
    for n_tris in (0, 1, 2, 3, 17, 500):
        pts = rng.normal(size=(3 * n_tris, 3)).astype(np.float32) * 0.3 + rng.normal(size=(n_tris, 1, 3)).astype(np.float32).repeat(3, 1).reshape(-1, 3) * 4
        nrm = rng.normal(size=(max(n_tris, 1), 3)).astype(np.float32)
        tris = np.zeros((n_tris, 6), np.int32)
        tris[:, :3] = np.arange(3 * n_tris).reshape(-1, 3)
        tris[:, 3:] = np.arange(n_tris)[:, None]
        blob, info = P.model_from_arrays(pts.reshape(-1, 3) if n_tris else np.zeros((0, 3), np.float32), nrm, tris, position=(1, 2, 3), visible=0)
        ob = oracle.new_model_blob()
        v = oracle.blob_views(ob)
        v["points"][: 3 * n_tris, :3] = pts.reshape(-1, 3) if n_tris else 0
        v["normals"][: len(nrm), :3] = nrm
        v["triangles"][:n_tris] = tris
        used, depth = oracle.build_bvh(ob, n_tris)
        pv = oracle.blob_views(blob)
        assert info["nodes_used"] == used and info["max_depth"] == depth
        assert np.array_equal(pv["nodes"][:used], v["nodes"][:used]) and np.array_equal(pv["lookup"][:n_tris], v["lookup"][:n_tris])
        assert tuple(pv["position"]) == (1.0, 2.0, 3.0) and pv["visible"][0] == 0


def test_obj_loader_matches_oracle(oracle, tmp_path):
    rng = np.random.default_rng(12)
    lines = ["o a"]
    n = 40
    for _ in range(3 * n):
        lines.append("v %.17g %.17g %.17g" % tuple(rng.normal(size=3) * 3))
    for _ in range(n):
        lines.append("vn %.9f %.9f %.9f" % tuple(rng.normal(size=3)))
    for i in range(n):
        a, b, c = 3 * i + 1, 3 * i + 2, 3 * i + 3
        lines.append(f"f {a}//{i + 1} {b}//{i + 1} {c}//{(i + 7) % n + 1}")
    p = tmp_path / "r.obj"
    p.write_text("\n".join(lines) + "\n")
    blob, info = P.load_obj_model(str(p))
    ob, oinfo = oracle.load_obj(str(p))
    assert np.array_equal(blob, ob)
    assert info["triangle_count"] == oinfo["triangles"] == n and info["point_count"] == 3 * n
    # shared vertices + no normals + negative indices
    p.write_text("v 0 0 0\nv 2 0 0\nv 0 2 0\nv 2 2 1\nf 1 2 3\nf -3 -1 -2\n")
    blob, info = P.load_obj_model(str(p))
    ob, oinfo = oracle.load_obj(str(p))
    assert np.array_equal(blob, ob) and info["point_count"] == 4 and info["normal_count"] == 2


@pytest.mark.skipif(not assets.have_lucy(), reason="lucy.obj not staged")
def test_lucy_through_product_loader(oracle):
    blob, info = P.load_obj_model(assets.lucy_path())
    assert (info["triangle_count"], info["point_count"], info["nodes_used"], info["max_depth"], info["leaf_count"], info["max_leaf_size"]) == \
        (99970, 299910, 117343, 24, 58672, 60)
    ob, _ = oracle.load_obj(assets.lucy_path())
    assert np.array_equal(blob, ob)


def test_band_layout():
    for h, band, world in ((2160, 8, 8), (1081, 8, 4), (37, 5, 3), (16, 16, 2), (9, 4, 8)):
        lay = BandLayout(h, band, world)
        rows = [lay.rows_of(r) for r in range(world)]
        allrows = np.sort(np.concatenate(rows))
        assert np.array_equal(allrows, np.arange(h))                  # a partition of the frame
        for r in range(world):
            assert np.all((rows[r] // band) % world == r)
            assert lay.local_rows(r) == len(rows[r])
            # local (band-major) order is increasing global row order
            assert np.all(np.diff(rows[r]) > 0)
        assert lay.max_local_rows == max(len(x) for x in rows)
    assert BandLayout(2160, 10, 8).uniform and not BandLayout(1081, 8, 4).uniform
    with pytest.raises(ValueError):
        BandLayout(10, 0, 2)


def test_bench_accounting():
    import bench
    st = {"ray_steps": 1000, "node_visits": 10, "tri_tests": 5, "tex_samples": 7}
    assert bench.algorithmic_bytes(st, 100) == 256 * 1000 + 16 * 100 + 64 * 10 + 124 * 5 + 16 * 7
    peak, src, _ = bench.peaks()
    assert peak > 1000 and ("measured" in src or "fallback" in src)


def test_rust_shim_declares_every_abi_symbol():
    """rust_shim/  cannot be compiled here (no Rust toolchain); at least keep its raw bindings in step with
    the header: every function include/bh_abi.h declares appears as `pub fn <name>(` in rust_shim/src/ffi.rs, and nothing else."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    header = open(os.path.join(root, "include", "bh_abi.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(bh_[a-z0-9_]+)\s*\(", header))
    rust = open(os.path.join(root, "rust_shim", "src", "ffi.rs")).read()
    bound = set(re.findall(r"pub fn (bh_[a-z0-9_]+)\s*\(", rust))
    assert declared == bound, (sorted(declared - bound), sorted(bound - declared))
    assert declared == set(_lib.SYMBOLS)


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` is a CPU arm (the oracle's strict flavour with OpenMP on a bounded sample): it must run
    without a GPU and print ONE JSON line with the contract's keys, its own value echoed in cpu_baseline and e2e."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.strip().startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "Mray-steps/sec" and d["unit"] == "Mray-steps/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_kernel_constants_are_refused_when_stale(tmp_path, monkeypatch):
    """bench.py's issue-rate roofline uses per-warp-step counts from a committed ncu capture; they only count when the capture
    was taken from the kernel source this build is made of (bhusie_b200.build.kernel_source_hash)."""
    import bench
    from bhusie_b200 import build as B
    h = B.kernel_source_hash()
    assert len(h) == 16 and h == B.kernel_source_hash()
    prof = tmp_path / "profiles"
    prof.mkdir()
    monkeypatch.setattr(bench, "ROOT", str(tmp_path))
    (prof / "trace_kernel_dram.json").write_text(json.dumps({"fused": {"source_hash": h, "warp_inst_per_warp_step": 190.0}}))
    pj, state = bench.kernel_constants("fused")
    assert state == "current" and pj["warp_inst_per_warp_step"] == 190.0
    (prof / "trace_kernel_dram.json").write_text(json.dumps({"fused": {"source_hash": "0" * 16, "warp_inst_per_warp_step": 190.0}}))
    pj, state = bench.kernel_constants("fused")
    assert pj == {} and state.startswith("stale")
    pj, state = bench.kernel_constants("literal")
    assert pj == {}


def test_model_validate_on_the_host():
    """bh_model_validate (pure host code): a blob built by the library passes; corrupt indices, a root that is its own child and
    lookup entries beyond the triangle array are refused; the empty model the reference starts with is accepted."""
    from bhusie_b200 import assets
    blob, info = P.model_from_arrays(*assets.uv_sphere(6, 8, radius=2.0))
    P.validate_model(blob)
    tri0 = int(blob[U.MU_LOOKUP:U.MU_LOOKUP + 4].view(np.int32)[0])
    for off, val in ((U.MU_NODES + 12, 10 ** 6), (U.MU_NODES + 12, 0), (U.MU_TRIANGLES + 24 * tri0, 600000),
                     (U.MU_TRIANGLES + 24 * tri0 + 12, -5), (U.MU_LOOKUP, 1 << 20)):
        bad = blob.copy()
        bad[off:off + 4].view(np.int32)[0] = val
        with pytest.raises(_lib.BhError) as e:
            P.validate_model(bad)
        assert e.value.code == -22
    empty, _ = P.model_from_arrays(np.zeros((0, 3), np.float32), np.zeros((0, 3), np.float32), np.zeros((0, 6), np.int32))
    P.validate_model(empty)


def test_host_frame_name_falls_back_to_temp_dir():
    """HostTiledFrame's rank 0 tries the POSIX shared-memory name first and a file under the temp directory second, and hands the
    other ranks either the name that worked or every reason why none did (bhusie_b200/multi.py::_create_host_frame)."""
    import tempfile
    from bhusie_b200.multi import _create_host_frame

    class OnlyFiles:
        def __init__(self, ctx, name, nbytes, create):
            if name.count("/") < 2:
                raise OSError("No space left on device")
            self.name = name

    class Nothing:
        def __init__(self, ctx, name, nbytes, create):
            raise OSError(f"cannot map {name}")

    name, err, hf = _create_host_frame(OnlyFiles, None, "/bhframe_test", 1 << 20)
    assert err is None and name == os.path.join(tempfile.gettempdir(), "bhframe_test") and hf.name == name
    name, err, hf = _create_host_frame(Nothing, None, "/bhframe_test", 1 << 20)
    assert name is None and hf is None and "/bhframe_test" in err and tempfile.gettempdir() in err
