"""Host-side file I/O helpers (util.py:59-98, 430-553): CPU only."""
import json
import os

import numpy as np
import pytest


@pytest.fixture(scope="module")
def util():
    from nixis_b200 import build
    build.build()
    from nixis_b200 import util
    return util


def test_settings_round_trip(util, tmp_path):
    data = {"world_name": "x", "seed": 12345, "divisions": 320}
    util.save_settings(data, str(tmp_path), "w_config", fmt="json")
    assert util.load_settings(str(tmp_path / "w_config.json")) == data
    assert open(tmp_path / "w_config.json").read() == json.dumps(data, indent=4)
    with pytest.raises(SystemExit):
        util.load_settings(str(tmp_path / "missing.json"))


def test_image_round_trip_16_and_8_bit(util, tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    rng = np.random.default_rng(0)
    a16 = rng.integers(0, 65536, (32, 64)).astype(np.uint16)
    a8 = rng.integers(0, 256, (32, 64)).astype(np.uint8)
    util.save_image({"height_absolute": a16, "ocean": a8}, str(tmp_path), "world")
    assert os.path.exists(tmp_path / "world_height_absolute.png") and os.path.exists(tmp_path / "world_ocean.png")
    assert np.array_equal(util.image_to_array(str(tmp_path / "world_height_absolute.png")), a16 / 65536)
    assert np.array_equal(util.image_to_array(str(tmp_path / "world_ocean.png")), np.float64(a8) / 256)


def test_save_mesh_obj(util, tmp_path, monkeypatch):
    from oracle import icosphere
    monkeypatch.chdir(tmp_path)
    pts, cells = icosphere.icosa_sphere(3)
    util.save_mesh(pts, cells, str(tmp_path), "planet")
    lines = open(tmp_path / "planet.obj").read().splitlines()
    v = np.array([[float(x) for x in ln.split()[1:]] for ln in lines if ln.startswith("v ")])
    f = np.array([[int(x) for x in ln.split()[1:]] for ln in lines if ln.startswith("f ")])
    assert np.array_equal(v, pts) and np.array_equal(f - 1, cells)


def test_latlon_helpers(util):
    lat, lon = util.xyz2latlon(*util.latlon2xyz(20, 15, 6378100.0), 6378100.0)
    assert abs(lat - 20) < 1e-9 and abs(lon - 15) < 1e-9
    assert util.xyz2latlon(0.0, 0.0, 2.0, 1.0)[0] == 90.0          # clamped, no NaN (util.py:69-70)
    assert util.kelvin_to_c(util.c_to_kelvin(12.5)) == pytest.approx(12.5)
