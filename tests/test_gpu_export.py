"""GPU parity of the equirectangular export row: make_ll_arr, the 3-nearest-vertex query
that replaces scipy's KD-tree, make_gray_array -- against golden vectors computed by the reference
(+ scipy) and, at larger sizes, against scipy.spatial.KDTree directly."""
import numpy as np
import pytest

from oracle import icosphere

pytestmark = pytest.mark.gpu
EARTH_R = 6378100.0


@pytest.fixture(scope="module")
def nx():
    import torch
    assert torch.cuda.is_available()
    from nixis_b200 import util, runtime
    return util, runtime, torch


def _same_neighbours(dists, nbrs, rd, rn, scale, query, pts):
    """The answer is valid iff the distances equal the reference's and the ids really are at those
    distances; where several vertices are equidistant (4th tied with 3rd on symmetric pixels, poles)
    scipy and the analytic search may name different ones, so ids are compared only off ties."""
    assert np.allclose(dists, rd, rtol=1e-12, atol=1e-12 * scale)
    true_d = np.linalg.norm(query[..., None, :] - pts[nbrs], axis=-1)
    assert np.allclose(true_d, dists, rtol=1e-12, atol=1e-12 * scale)
    assert (np.diff(dists, axis=-1) >= 0).all()
    same = (np.sort(nbrs, axis=-1) == np.sort(rn, axis=-1)).all(axis=-1)
    assert same.mean() > 0.9, same.mean()       # the rest are exact distance ties (validated above)


@pytest.mark.parametrize("k,R,W,H", [(16, EARTH_R, 64, 32), (8, 1.0, 36, 19)])
def test_export_chain_vs_reference_golden(nx, golden, k, R, W, H):
    util, rt, torch = nx
    g = golden("export")
    t = f"k{k}_{W}x{H}"
    ll = util.make_ll_arr(W, H, R)
    assert ll.shape == (H, W, 3) and np.allclose(ll, g[f"{t}_ll"], rtol=0, atol=4e-16 * R)
    pts, _ = icosphere.icosa_sphere(k)
    util.build_KDTree(pts * R)
    dists, nbrs = util.cfg.KDT.query(g[f"{t}_ll"], k=3, workers=-1)       # nixis.py:283
    assert dists.dtype == np.float64 and nbrs.dtype == np.int64 and nbrs.shape == (H, W, 3)
    _same_neighbours(dists, nbrs, g[f"{t}_dists"], g[f"{t}_nbrs"], R, g[f"{t}_ll"], pts * R)
    # the blend, fed with the reference's own query result: bit-exact integers
    for src, ref in (("abs_src", "abs"), ("rel_src", "rel")):
        got = util.make_gray_array(W, H, g[f"{t}_dists"], g[f"{t}_nbrs"], g[f"{t}_{src}"])
        assert got.dtype == np.uint16 and np.array_equal(got, g[f"{t}_{ref}"])
    # build_image_data: dtype handling of util.py:393-429 on float heights
    util.cfg.IMG_QUERY_DATA = (g[f"{t}_dists"], g[f"{t}_nbrs"])
    res = util.build_image_data({"height": [g[f"{t}_height"].copy(), "gray"], "abs": [g[f"{t}_abs_src"], "gray"]})
    assert res["abs"].dtype == np.uint16 and np.array_equal(res["abs"], g[f"{t}_abs"])
    assert res["height"].dtype == np.uint8
    assert np.abs(res["height"].astype(int) - g[f"{t}_height255"].astype(int)).max() <= 1     # FP32 rescale feeding int()


def test_nearest3_vs_scipy_kdtree_large(nx):
    """k=200 (400 002 vertices), 1024x512 pixels: the analytic search returns what scipy's KD-tree returns."""
    util, rt, torch = nx
    from scipy.spatial import KDTree
    k, R, W, H = 200, EARTH_R, 1024, 512
    mesh = util.create_mesh(k, device=True, radius=R, verbose=False, with_cells=False)
    pts = mesh.points_numpy()
    ll = util.make_ll_arr(W, H, R)
    rd, rn = KDTree(pts, leafsize=10).query(ll, k=3, workers=-1)
    util.build_KDTree(pts)
    d, n = util.cfg.KDT.query(ll, k=3)
    _same_neighbours(d, n, rd, rn, R, ll, pts)
    # random directions as well (not only the lat/lon grid), incl. exact vertex positions
    rng = np.random.default_rng(3)
    q = rng.normal(size=(200000, 3))
    q = q / np.linalg.norm(q, axis=1, keepdims=True) * R
    q[:1000] = pts[rng.integers(0, len(pts), 1000)]
    rd, rn = KDTree(pts, leafsize=10).query(q, k=3, workers=-1)
    d, n = util.cfg.KDT.query(q, k=3)
    _same_neighbours(d, n, rd, rn, R, q, pts)
    assert (d[:1000, 0] < 1e-6).all()


def test_export_config1_4096x2048_d1000_properties(nx):
    """BASELINE configs[1]: d=1000 heights exported to a 4096x2048 equirectangular map, device resident."""
    util, rt, torch = nx
    from nixis_b200.pipeline import TerrainPipeline
    k, W, H = 1000, 4096, 2048
    pipe = TerrainPipeline(k, seed=12345, n_octaves=8, radius=1.0)
    pipe.build_mesh(with_adjacency=False)
    h, ocean, level = pipe.heights()
    ll = rt.ll_grid(W, H, 1.0)
    d, ids = rt.ico_nearest3(k, 1.0, ll)
    assert int(ids.min()) >= 0 and int(ids.max()) < pipe.V
    assert bool((d[..., 0] <= d[..., 1]).all()) and bool((d[..., 1] <= d[..., 2]).all())
    assert float(d.max()) < 3.0 / k                                   # nearest vertices are within a few edge lengths
    # the three ids of a pixel are mutually adjacent or identical-distance corners: all distinct
    assert bool((ids[..., 0] != ids[..., 1]).all()) and bool((ids[..., 1] != ids[..., 2]).all())
    img = rt.idw_gray(d, ids, (h.to(torch.float64) + 4000.0) * (65535.0 / 12850.0))
    assert img.shape == (H, W) and int(img.min()) >= 0 and int(img.max()) <= 65535
    # a blend of three values lies between their min and max
    v = ((h.to(torch.float64) + 4000.0) * (65535.0 / 12850.0))[ids]
    assert bool((img.to(torch.float64) <= v.max(dim=-1).values + 1e-6).all())
    assert bool((img.to(torch.float64) >= v.min(dim=-1).values.floor() - 1).all())


def test_device_export_maps_match_dropin_chain(nx):
    """the fused device-resident maps (TerrainPipeline.export_maps) equal what the
    drop-in chain of nixis.py:349,386-389 + build_image_data produces from the same FP32 heights --
    exactly for the uint16 maps when fed identical per-vertex values, and within 1 level of the
    float64 numpy chain (FP32 heights feeding int())."""
    util, rt, torch = nx
    from nixis_b200.pipeline import TerrainPipeline, find_percent_val
    k, W, H = 64, 512, 256
    pipe = TerrainPipeline(k, seed=12345, n_octaves=8, radius=1.0)
    pipe.build_mesh(with_adjacency=False)
    h, ocean, _ = pipe.heights()
    q = pipe.image_query(W, H)
    maps = pipe.export_maps(h, ocean, W, H, query=q)
    assert maps["height_absolute"].dtype == torch.uint16 and maps["ocean"].dtype == torch.uint8
    dists, nbrs = q[0].cpu().numpy(), q[1].cpu().numpy()
    h64 = h.cpu().numpy().astype(np.float64)

    def np_rescale(x, lo, hi):                      # util.py:143
        return ((x - x.min()) / (x.max() - x.min())) * (hi - lo) + lo

    def np_blend(colors):                           # util.py:343-367
        c = colors.astype(np.float64)[nbrs]
        sd = dists[..., 0] + dists[..., 1] + dists[..., 2]
        ws = 1 / ((dists + 0.00001) / sd[..., None])
        t = ws[..., 0] + ws[..., 1] + ws[..., 2]
        iw = ws / t[..., None]
        return (c[..., 0] * iw[..., 0] + c[..., 1] * iw[..., 1] + c[..., 2] * iw[..., 2]).astype(np.int64)

    absolute = (np_rescale(h64, -4000, 8850) + (32768 - find_percent_val(-4000, 8850, 55.0))).astype('uint16')
    relative = np_rescale(h64, 0, 65535).astype('uint16')
    assert np.array_equal(maps["height_absolute"].cpu().numpy(), np_blend(absolute).astype('uint16'))
    assert np.array_equal(maps["height_relative"].cpu().numpy(), np_blend(relative).astype('uint16'))
    mask = ocean.cpu().numpy().astype(bool)
    assert np.array_equal(maps["ocean"].cpu().numpy(), np_blend(np_rescale(mask.astype(np.float64), 0, 255)).astype('uint8'))
    # the same maps through the drop-in functions (host numpy in / out)
    util.cfg.IMG_QUERY_DATA = (dists, nbrs)
    res = util.build_image_data({"ocean": [mask, 'gray'], "height_absolute": [absolute, 'gray'],
                                 "height_relative": [relative, 'gray']})
    for key in ("ocean", "height_absolute", "height_relative"):
        assert res[key].dtype == maps[key].cpu().numpy().dtype
        assert np.abs(res[key].astype(int) - maps[key].cpu().numpy().astype(int)).max() <= (1 if key == "ocean" else 0), key
    # after erosion: a single uint8 `height` map (nixis.py:417)
    e = pipe.export_maps(h, None, W, H, eroded=True, query=q)["height"]
    assert np.array_equal(e.cpu().numpy(), np_blend(np_rescale(h64, 0, 255)).astype('uint8'))
