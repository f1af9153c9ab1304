"""The swap-in mechanism on real hardware: tools/run_nixis.py pre-seeds sys.modules with the
nixis_b200 modules under the reference's module names and executes a CLI script that imports them the
way nixis.py does.  /root/reference does not exist on the GPU box, so the script is
tests/fixtures/mini_nixis/nixis.py, which replays nixis.py:247-417's call order; what the run leaves
behind is checked against the CPU oracle."""
import glob
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def relerr(a, ref):
    return np.abs(np.asarray(a, np.float64) - ref).max() / (ref.max() - ref.min())


def test_swap_in_runner_drives_the_whole_path(tmp_path, oracle):
    out = str(tmp_path / "run.npz")
    cmd = [sys.executable, os.path.join(ROOT, "tools", "run_nixis.py"), "--ref", os.path.join(ROOT, "tests", "fixtures", "mini_nixis"),
           "--erode", "--", "-d", "24", "-s", "12345", "--novis", "--save_img", "--snapshot", "--erode_iters", "3", "--out", out]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=str(tmp_path))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "hot path swapped onto nixis_b200" in r.stdout and "mini_nixis finished" in r.stdout
    assert "Starting terrain erosion..." in r.stdout and "Erosion pass: 3 of 3" in r.stdout      # erosion.py:173,181
    g = np.load(out)
    from oracle import icosphere
    pts, cells = icosphere.icosa_sphere(24)
    assert np.array_equal(g["points"], pts) and np.array_equal(g["cells"], cells)
    perm, pgi = oracle.init(12345)
    assert np.array_equal(g["perm"], perm) and np.array_equal(g["pgi"], pgi)
    adj = oracle.build_adjacency(cells)
    oracle.sort_adjacency(adj)
    assert np.array_equal(g["neighbors"], adj)
    h = oracle.sample_octaves(pts, None, perm, pgi, 7, 1.5, 0.4, 2.5, 0.5, 1.0)
    ref_asm, ref_ocean, ref_level = oracle.height_assembly(h)
    assert abs(float(g["ocean_level"]) - ref_level) < 1e-2
    assert relerr(g["assembled"], ref_asm) <= 2e-5
    ref_h = g["assembled"].copy()                      # erosion from the run's own (FP32-rounded) assembled heights
    oracle.erode_terrain3(pts, adj, ref_h, 3)
    assert bool(g["eroded"]) and relerr(g["height"], ref_h) <= 1e-4
    # export (nixis.py:516-540) + snapshots (erosion.py:186-192)
    assert g["img_height"].shape == (128, 256) and g["img_height"].dtype == np.uint8 and g["img_ocean"].dtype == np.uint8
    assert len(glob.glob(str(tmp_path / "mini_*.png"))) == 2
    snaps = sorted(glob.glob(str(tmp_path / "erosion_snapshot_*.png")))
    assert [os.path.basename(s) for s in snaps] == ["erosion_snapshot_001.png", "erosion_snapshot_002.png", "erosion_snapshot_003.png"]
    from PIL import Image
    last = np.asarray(Image.open(snaps[-1]))
    assert last.shape == (128, 256) and np.abs(last.astype(int) - g["img_height"].astype(int)).max() <= 1


def test_rgb_maps_and_dtype_policy(tmp_path):
    """util.py:310-341 make_rgb_array / :369-429 build_image_data dtype policy: an RGB map is the gray
    blend in three channels; uint16 sources stay uint16, everything else is exported as uint8."""
    from nixis_b200 import util
    pts, _ = util.create_mesh(8, verbose=False)
    util.build_KDTree(pts)
    ll = util.make_ll_arr(64, 32, 1.0)
    util.cfg.IMG_QUERY_DATA = util.cfg.KDT.query(ll, k=3)
    rng = np.random.default_rng(2)
    f = rng.normal(size=len(pts))
    u16 = rng.integers(0, 65535, len(pts)).astype(np.uint16)
    res = util.build_image_data({"g": [f.copy(), "gray"], "c": [f.copy(), "rgb"], "w": [u16, "gray"], "m": [f > 0, "GRAY"]})
    assert res["g"].shape == (32, 64) and res["g"].dtype == np.uint8
    assert res["c"].shape == (32, 64, 3) and all(np.array_equal(res["c"][:, :, i], res["g"]) for i in range(3))
    assert res["w"].dtype == np.uint16 and res["m"].dtype == np.uint8
    d, n = util.cfg.IMG_QUERY_DATA
    rgb = util.make_rgb_array(64, 32, d, n, util.rescale(f, 0, 255))
    assert rgb.shape == (32, 64, 3) and np.array_equal(rgb[:, :, 0].astype(np.uint8), res["g"])
