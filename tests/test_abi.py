"""The C-ABI library loads without a GPU and exports every symbol include/nixis_b200.h declares.
No kernel is launched here; host-only entry points (init_perm) and host-side logic are checked."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from nixis_b200 import build, _lib
    build.build()
    return _lib.load()


def test_header_symbols_exported(lib):
    from nixis_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "nixis_b200.h")).read()
    declared = set(re.findall(r"\b(nxb_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"nxb_status"}
    assert len(declared) >= 25
    raw = ctypes.CDLL(_lib.SO_PATH)
    for name in sorted(declared):
        assert hasattr(raw, name), f"{name} declared in the header but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature"
    assert set(_lib.SIGNATURES) <= declared
    assert lib.nxb_version() == 1


def test_missing_library_fails_loudly(monkeypatch):
    from nixis_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "SO_PATH", "/nonexistent/libnixis_b200.so")
    with pytest.raises(ImportError):
        _lib.load()


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "nixis_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), f
                assert "nixis_oracle" not in src, f


def test_init_perm_host_exact(lib, golden):
    from nixis_b200 import opensimplex
    g = golden("init")
    for i, s in enumerate(g["seeds"]):
        perm, pgi = opensimplex.init(int(s))
        assert perm.dtype == np.int32 and pgi.dtype == np.int32
        assert np.array_equal(perm, g["perm"][i]) and np.array_equal(pgi, g["pgi"][i]), int(s)


def test_no_gpu_raises_not_falls_back():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from nixis_b200 import terrain, opensimplex
    perm, pgi = opensimplex.init(1)
    with pytest.raises(RuntimeError):
        terrain.sample_octaves(np.zeros((4, 3)), None, perm, pgi, verbose=False)


def test_power_summary_combine_matches_sequential_scan(oracle):
    """Host-side ordered combine (runtime.combine_power_summaries / power_bounds) against the
    oracle's literal sequential scan, on shards of random data incl. the if/elif corner case."""
    from nixis_b200 import runtime as rt

    def shard_summary(x, sel):
        has, F, U, M = 0.0, 0.0, float("-inf"), float("inf")
        for v, s in zip(x, sel):
            if not s:
                continue
            if not has:
                has, F, M = 1.0, v, v
            elif v >= M:
                U = max(U, v)
            else:
                M = v
        return has, F, U, M

    rng = np.random.default_rng(3)
    cases = [(np.array([3.0, 1.0, 2.0, 10.0, 1.5]), np.array([1, 1, 1, 0, 1], bool))]
    for _ in range(200):
        n = int(rng.integers(1, 40))
        x = np.round(rng.normal(size=n), 1)          # many ties
        cases.append((x, rng.random(n) < 0.6))
        cases.append((np.sort(x)[::-1].copy(), np.ones(n, bool)))   # strictly falling: all records
    for x, sel in cases:
        for nshard in (1, 2, 3, 7):
            cuts = np.linspace(0, len(x), nshard + 1).astype(int)
            summ = [shard_summary(x[a:b], sel[a:b]) for a, b in zip(cuts[:-1], cuts[1:])]
            lo, hi = rt.power_bounds(rt.combine_power_summaries(summ), x.min(), x.max())
            _, stats = oracle.power_rescale(x, sel, 1, 1.0, return_stats=True)
            assert (lo, hi) == (stats[2], stats[3]), (x, sel, nshard)


def test_swap_in_runner_reaches_our_create_mesh():
    """tools/run_nixis.py pre-seeds sys.modules and executes the unmodified reference CLI: without a
    GPU it must get as far as OUR create_mesh and fail loudly there (no CPU fallback).  Needs the
    reference checkout (present in the build container only)."""
    import subprocess
    import sys
    import torch
    ref = os.environ.get("NIXIS_REF", "/root/reference")
    if not os.path.exists(os.path.join(ref, "nixis.py")):
        pytest.skip("reference checkout not present")
    if torch.cuda.is_available():
        pytest.skip("GPU present: the run would go through")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "run_nixis.py"), "--ref", ref, "--",
                        "-d", "4", "-s", "12345", "--novis"], capture_output=True, text=True, timeout=300, cwd="/tmp")
    assert "hot path swapped onto nixis_b200" in r.stdout
    assert r.returncode != 0 and "nixis_b200 needs a CUDA device" in r.stderr
    assert "nixis_b200/util.py" in r.stderr and "create_mesh" in r.stderr
