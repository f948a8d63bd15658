"""The C-ABI library loads on a machine without a GPU, exports every symbol include/pcr_b200.h
declares, and fails loudly (no CPU fallback) when asked to compute without a device."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "pcr_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pcr_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    from pointclouds_rs_b200 import _ffi

    lib = _ffi.load()
    names = declared_symbols()
    assert len(names) >= 35
    for n in names:
        assert hasattr(lib, n), f"{n} declared in pcr_b200.h but not exported"
    assert set(names) == set(_ffi.SIGNATURES), set(names) ^ set(_ffi.SIGNATURES)
    assert lib.pcr_version() == 122


def test_no_torch_in_library_dependencies():
    import subprocess

    from pointclouds_rs_b200 import _ffi

    out = subprocess.run(["ldd", _ffi.LIB_PATH], capture_output=True, text=True).stdout
    assert "torch" not in out and "libc10" not in out and "nccl" not in out  # NCCL is dlopen'ed on demand


def test_fails_loudly_without_device():
    from pointclouds_rs_b200 import _ffi

    lib = _ffi.load()
    if lib.pcr_device_count() > 0:
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    st = lib.pcr_ctx_create(0, C.byref(h))
    assert st == _ffi.PCR_ERR_NO_DEVICE and not h.value
    assert b"no CPU fallback" in lib.pcr_last_error(None)
    import pointclouds_rs_b200 as pcr

    with pytest.raises(pcr.PcrError):
        pcr.statistical_outlier_removal(pcr.PointCloud.from_numpy(__import__("numpy").zeros((4, 3), "float32")), 2, 1.0)


def test_null_arguments_are_rejected_not_crashing():
    from pointclouds_rs_b200 import _ffi

    lib = _ffi.load()
    assert lib.pcr_ctx_create(0, None) == _ffi.PCR_ERR_INVALID_ARG
    assert lib.pcr_ctx_synchronize(None) == _ffi.PCR_ERR_INVALID_ARG
    assert lib.pcr_index_len(None) == 0
    lib.pcr_index_free(None)
    lib.pcr_ctx_destroy(None)
    assert lib.pcr_comm_unique_id(None) == _ffi.PCR_ERR_INVALID_ARG


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "pointclouds_rs_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", "Makefile")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "pcr_oracle" not in text and "from oracle" not in text and "import oracle" not in text, f


def test_tools_do_not_import_the_oracle():
    """tools/ holds profiling helpers for the product path; anything that checks against the oracle lives in tests/."""
    tdir = os.path.join(ROOT, "tools")
    for f in os.listdir(tdir):
        if f.endswith(".py"):
            text = open(os.path.join(tdir, f), errors="ignore").read()
            assert "from oracle" not in text and "import oracle" not in text, f


def test_python_surface_mirrors_reference_names():
    import pointclouds_rs_b200 as pcr

    for name in ("PointCloud", "statistical_outlier_removal", "radius_outlier_removal", "estimate_normals", "IcpResult",
                 "icp_point_to_point", "icp_point_to_plane", "apply_transform", "voxel_downsample", "euclidean_cluster", "ransac_plane",
                 "PlaneResult"):
        assert hasattr(pcr, name)
    pc = pcr.PointCloud.from_numpy(__import__("numpy").arange(12, dtype="float64").reshape(4, 3))
    assert pc.len() == 4 and len(pc) == 4 and not pc.is_empty() and repr(pc) == "PointCloud(n=4)"
    assert pc.select([0, 2]).to_numpy().tolist() == [[0, 1, 2], [6, 7, 8]]
    assert pc.select_inverse([0, 2]).to_numpy().tolist() == [[3, 4, 5], [9, 10, 11]]
    with pytest.raises(IndexError):
        pc.select([7])
    with pytest.raises(TypeError):
        pcr.PointCloud.from_numpy(__import__("numpy").zeros((3, 3), "int32"))
    with pytest.raises(ValueError):
        pcr.PointCloud.from_numpy(__import__("numpy").zeros((3, 4), "float32"))
    with pytest.raises(ValueError):
        pcr.PointCloud.from_numpy(__import__("numpy").asfortranarray(__import__("numpy").zeros((3, 3), "float32")))


def test_normals_are_split_once_per_array():
    """The reference's Normals are SoA (crates/core/src/cloud.rs:13-18); the mirror keeps an (N, 3) array and its split
    form together: the split is redone only when another array is attached (a 1 M-point split costs more than the rest
    of an ICP call's setup)."""
    import numpy as np

    import pointclouds_rs_b200 as pcr

    pc = pcr.PointCloud.from_numpy(np.arange(12, dtype=np.float32).reshape(4, 3))
    pc.normals = np.arange(12, 24, dtype=np.float32).reshape(4, 3)
    a = pc._normals_soa()
    assert [x.tolist() for x in a] == [[12, 15, 18, 21], [13, 16, 19, 22], [14, 17, 20, 23]]
    assert all(x.flags["C_CONTIGUOUS"] and x.dtype == np.float32 for x in a)
    assert pc._normals_soa()[0] is a[0]                     # same array attached: nothing recomputed
    pc.normals = np.zeros((4, 3), np.float32)               # another array: split again
    assert pc._normals_soa()[0] is not a[0] and pc._normals_soa()[0].tolist() == [0, 0, 0, 0]
    sub = pc.select([1, 3])                                 # select carries the normals (cloud.rs:103-140) with a split of its own
    assert sub.normals.shape == (2, 3) and sub._normals_soa()[1].shape == (2,)


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` needs no GPU: one bounded step of the CPU port, one JSON line with the contract's keys."""
    import json
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=300, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "sor_normals_points_per_sec" and line["unit"] == "points/s"
    assert line["higher_is_better"] is True and line["value"] > 0 and line["steps"] == 1
    assert line["e2e"]["value"] == line["value"] and line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["cpu_baseline"]["kind"] in ("port", "reference") and line["cpu_baseline"]["cores"] >= 1
    assert "workload" in line["config"]
