"""Generate the golden fixtures in this directory by RUNNING THE UNMODIFIED REFERENCE.

Run in the build container only (needs /root/reference + numba):

    python tests/golden/gen_golden.py

The reference is imported from $NIXIS_REF (default /root/reference) with the two
absent third-party names stubbed (`util.py:6-7` import meshzoo / meshio); nothing
of the reference is copied -- the fixtures are data it computed.  The mesh fed
to it comes from oracle/icosphere.py (meshzoo itself is unavailable: mesh parity
is unpinned, everything downstream of the mesh is pinned by these files).

The GPU box has no /root/reference; tests read only the .npz files written here.
"""
import hashlib
import os
import sys
import types
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("NIXIS_REF", "/root/reference")
os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/nixis_numba_cache")
sys.path.insert(0, REF)
sys.path.insert(1, ROOT)
for _name in ("meshzoo", "meshio"):
    sys.modules.setdefault(_name, types.ModuleType(_name))
warnings.filterwarnings("ignore")

import numpy as np  # noqa: E402

_cwd = os.getcwd()
os.chdir(REF)  # options.json is opened cwd-relative (util.py:381)
import opensimplex as osi  # noqa: E402
import terrain  # noqa: E402
import util  # noqa: E402
import erosion  # noqa: E402
os.chdir(_cwd)

from oracle import icosphere  # noqa: E402

EARTH_R = 6378100.0
SEEDS = [0, 1, 42, 12345, -7, 2 ** 31 - 1, 2 ** 31, 2 ** 32 + 5, -2 ** 40, 987654321987]


def save(name, **arrays):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **arrays)
    print(f"  {name}: {os.path.getsize(path) / 1024:.1f} KiB")


def gen_init():
    perms = np.stack([osi.init(s)[0] for s in SEEDS])
    pgis = np.stack([osi.init(s)[1] for s in SEEDS])
    save("init.npz", seeds=np.array(SEEDS, dtype=np.int64), perm=perms, pgi=pgis)


def noise_points(rng, n, dim):
    """Uniform points, points hugging lattice-cell / region boundaries, negatives, large."""
    parts = [rng.uniform(-4, 4, (n, dim)), rng.uniform(-300, 300, (n // 4, dim)),
             rng.uniform(-1e4, 1e4, (n // 8, dim)),
             np.round(rng.uniform(-8, 8, (n // 8, dim)) * 4) / 4,        # exact quarter-lattice points
             np.zeros((1, dim)), np.ones((1, dim)), -np.ones((1, dim))]
    return np.ascontiguousarray(np.concatenate(parts))


def gen_noise():
    rng = np.random.default_rng(20260117)
    out = {}
    for seed in (0, 12345):
        perm, pgi = osi.init(seed)
        p2, p3, p4 = noise_points(rng, 4000, 2), noise_points(rng, 6000, 3), noise_points(rng, 6000, 4)
        out[f"p2_{seed}"] = p2
        out[f"p3_{seed}"] = p3
        out[f"p4_{seed}"] = p4
        out[f"v2_{seed}"] = osi.noisearr2d(p2[:, 0].copy(), p2[:, 1].copy(), perm)
        out[f"v3_{seed}"] = osi.noisearr3d(p3[:, 0].copy(), p3[:, 1].copy(), p3[:, 2].copy(), perm, pgi)
        out[f"v4_{seed}"] = osi.noisearr4d(p4[:, 0].copy(), p4[:, 1].copy(), p4[:, 2].copy(), p4[:, 3].copy(), perm)
    save("noise.npz", **out)


def quiet(fn, *a, **k):
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def gen_fbm():
    out = {}
    cases = [(2, 12345, 8, 1.0), (8, 12345, 8, 1.0), (8, 0, 7, EARTH_R), (32, 12345, 8, 1.0),
             (32, 0, 8, 1.0), (32, 12345, 7, EARTH_R), (32, 42, 12, 1.0)]
    for k, seed, n_oct, R in cases:
        pts, _ = icosphere.icosa_sphere(k)
        perm, pgi = osi.init(seed)
        h = quiet(terrain.sample_octaves, pts * R, None, perm, pgi, n_oct, 1.5, 0.4, 2.5, 0.5, R)
        tag = f"k{k}_s{seed}_o{n_oct}_{'earth' if R != 1.0 else 'unit'}"
        out[tag] = h
    # accumulate-into-existing-array + non-default lacunarity/persistence
    pts, _ = icosphere.icosa_sphere(8)
    perm, pgi = osi.init(7)
    e = np.linspace(-1, 1, len(pts))
    out["k8_s7_accum"] = quiet(terrain.sample_octaves, pts, e.copy(), perm, pgi, 3, 2.0, 0.7, 2.0, 0.45, 1.0)
    out["cases"] = np.array([[k, s, o, R] for k, s, o, R in cases], dtype=np.float64)
    save("fbm.npz", **out)


def gen_adjacency():
    out = {}
    for k in (1, 2, 3, 4, 8, 17, 32):
        _, cells = icosphere.icosa_sphere(k)
        adj = util.build_adjacency(cells)
        out[f"unsorted_k{k}"] = adj.copy()
        if k >= 2:
            util.sort_adjacency.py_func(adj)     # sequential semantics = the race-free parity target
            out[f"sorted_k{k}"] = adj
    save("adjacency.npz", **out)


def assembly(height):
    """nixis.py:332-364, calling the reference's own functions in its order."""
    height = util.rescale(height, -4000, 8850)
    minval, maxval = np.amin(height), np.amax(height)
    ocean_level = util.find_percent_val(minval, maxval, 55.0)
    ocean = terrain.make_bool_elevation_mask(height, ocean_level)
    h1 = quiet(util.power_rescale, height, mask=ocean, mode=1, power=0.5)
    h2 = quiet(util.power_rescale, h1, mask=ocean, mode=0, power=2.0)
    h3 = h2 - ocean_level
    h4 = util.rescale(h3, -4000, 8850, mid=0)
    return height, ocean_level, ocean, h1, h2, h4


def gen_assembly():
    out = {}
    for k, seed, R in ((16, 12345, 1.0), (32, 0, EARTH_R)):
        pts, _ = icosphere.icosa_sphere(k)
        perm, pgi = osi.init(seed)
        raw = quiet(terrain.sample_octaves, pts * R, None, perm, pgi, 7, 1.5, 0.4, 2.5, 0.5, R)
        h0, level, ocean, h1, h2, h4 = assembly(raw)
        t = f"k{k}_s{seed}"
        out.update({f"{t}_raw": raw, f"{t}_rescaled": h0, f"{t}_level": np.float64(level),
                    f"{t}_ocean": ocean, f"{t}_pow1": h1, f"{t}_pow0": h2, f"{t}_final": h4})
    # rescale modes on a fixed vector
    rng = np.random.default_rng(5)
    x = rng.normal(size=2000)
    out["rs_x"] = x
    out["rs_plain"] = util.rescale(x, -1.0, 3.0)
    out["rs_mid"] = util.rescale(x, -2.0, 5.0, mid=0.25)
    out["rs_lower"] = util.rescale(x, -2.0, 5.0, mid=0.25, mode='lower')
    out["rs_upper"] = util.rescale(x, -2.0, 5.0, mid=0.25, mode='upper')
    out["rs_umin"] = util.rescale(x, 0.0, 1.0, u_min=-10.0, u_max=0.5)
    m = x > 0.1
    out["pr_mask"] = m
    out["pr_m1"] = quiet(util.power_rescale, x, mask=m, mode=1, power=1.7)
    out["pr_m0"] = quiet(util.power_rescale, x, mask=m, mode=0, power=0.3)
    # the if/elif quirk: first selected element is the selected maximum
    y = np.array([5.0, 1.0, 2.0, 3.0, -1.0, 4.0])
    my = np.array([True, True, True, True, False, False])
    out["pq_x"], out["pq_mask"] = y, my
    out["pq_m1"] = quiet(util.power_rescale, y, mask=my, mode=1, power=2.0)
    # ... and the case where it bites: the selected maximum (3) is a running-minimum record,
    # so the reference ends with mask_upper = 2 (util.py:203-207)
    y2 = np.array([3.0, 1.0, 2.0, 10.0, 1.5])
    my2 = np.array([True, True, True, False, True])
    out["pq2_x"], out["pq2_mask"] = y2, my2
    out["pq2_m1"] = quiet(util.power_rescale, y2, mask=my2, mode=1, power=2.0)
    out["pq2_m0"] = quiet(util.power_rescale, y2, mask=~my2, mode=0, power=0.5)
    save("assembly.npz", **out)


def gen_erosion():
    out = {}
    for k, seed, R, steps in ((8, 12345, 1.0, (1, 5, 11, 50)), (32, 12345, 1.0, (1, 5, 11, 50)),
                              (32, 0, EARTH_R, (1, 3, 5))):
        pts, cells = icosphere.icosa_sphere(k)
        nodes = pts * R
        perm, pgi = osi.init(seed)
        raw = quiet(terrain.sample_octaves, nodes, None, perm, pgi, 8, 1.5, 0.4, 2.5, 0.5, R)
        h_start = assembly(raw)[-1]
        adj = util.build_adjacency(cells)
        util.sort_adjacency.py_func(adj)
        t = f"k{k}_s{seed}_{'earth' if R != 1.0 else 'unit'}"
        out[f"{t}_h0"] = h_start
        # erode_terrain3 semantics (erosion.py:172-184) with the state kept
        h = h_start.copy()
        wat = np.zeros_like(h)
        sed = np.zeros_like(h)
        for it in range(1, max(steps) + 1):
            wat += 0.3 / 320
            erosion.erosion_iteration3(nodes, adj, h, wat, sed)
            if it in steps:
                out[f"{t}_it3_n{it}_h"] = h.copy()
                out[f"{t}_it3_n{it}_w"] = wat.copy()
                out[f"{t}_it3_n{it}_s"] = sed.copy()
        # the driver itself, 11 passes as nixis.py:410 calls it
        if R == 1.0:
            h = h_start.copy()
            quiet(erosion.erode_terrain3, nodes, adj, h, num_iter=11, snapshot=False)
            out[f"{t}_driver11_h"] = h
        # erosion_iteration1 / erode_terrain1
        h = h_start.copy()
        w = np.ones_like(h)
        erosion.erosion_iteration1(adj, h, w)
        out[f"{t}_it1_n1"] = w.copy()
        h = h_start.copy()
        r = quiet(erosion.erode_terrain1, pts, adj, h, num_iter=100)
        out[f"{t}_it1_n100"] = np.asarray(r).copy()
    save("erosion.npz", **out)


def gen_export():
    """make_ll_arr -> scipy KDTree query (k=3) -> make_gray_array, reference code."""
    from scipy.spatial import KDTree
    out = {}
    for k, R, W, H in ((16, EARTH_R, 64, 32), (8, 1.0, 36, 19)):
        pts, _ = icosphere.icosa_sphere(k)
        points = pts * R
        ll = util.make_ll_arr(W, H, R)
        dists, nbrs = KDTree(points, leafsize=10).query(ll, k=3)
        perm, pgi = osi.init(12345)
        raw = quiet(terrain.sample_octaves, points, None, perm, pgi, 7, 1.5, 0.4, 2.5, 0.5, R)
        height = assembly(raw)[-1]
        ocean = terrain.make_bool_elevation_mask(height, 0.0)
        t = f"k{k}_{W}x{H}"
        out[f"{t}_ll"], out[f"{t}_dists"], out[f"{t}_nbrs"], out[f"{t}_height"] = ll, dists, nbrs, height
        # nixis.py:382-389 and util.py:393-404 dtype handling, then util.py:343-367
        absolute = (util.rescale(height, -4000, 8850) + (32768 - util.find_percent_val(-4000, 8850, 55.0))).astype('uint16')
        relative = (util.rescale(height, 0, 65535)).astype('uint16')
        out[f"{t}_abs_src"], out[f"{t}_rel_src"] = absolute, relative
        out[f"{t}_abs"] = util.make_gray_array(W, H, dists, nbrs, absolute)
        out[f"{t}_rel"] = util.make_gray_array(W, H, dists, nbrs, relative)
        out[f"{t}_height255"] = util.make_gray_array(W, H, dists, nbrs, util.rescale(height, 0, 255)).astype('uint8')
        out[f"{t}_ocean"] = util.make_gray_array(W, H, dists, nbrs, util.rescale(ocean.astype(np.float64), 0, 255)).astype(ocean.dtype)
    save("export.npz", **out)


def gen_climate():
    """the per-vertex climate kernels of climate.py:345-597, reference code."""
    os.chdir(REF)
    import climate
    os.chdir(_cwd)
    out = {}
    rng = np.random.default_rng(777)
    for k, R in ((8, EARTH_R), (12, 1.0)):
        pts, _ = icosphere.icosa_sphere(k)
        P = np.ascontiguousarray(pts * R)
        h = rng.uniform(-4000.0, 8850.0, len(P))
        t = f"k{k}"
        out[f"{t}_R"] = np.array([R])
        out[f"{t}_height"] = h
        tilts = [climate.calculate_seasonal_tilt(23.44, 36), 0.0, -23.44, 80.0]
        out[f"{t}_tilts"] = np.array(tilts)
        for i, tilt in enumerate(tilts):
            out[f"{t}_temp_{i}"] = climate.assign_surface_temp(P, h, R, tilt)
            out[f"{t}_slice_{i}"] = climate.calc_insolation_slice(R, tilt)
            out[f"{t}_daily_{i}"] = climate.calc_daily_insolation(P, h, R, tilt)
            for j, rot in enumerate((0, 37.5, -180.0)):
                out[f"{t}_instant_{i}_{j}"] = climate.calc_instant_insolation(P, h, R, rot, tilt)
        out[f"{t}_brute_0"] = climate.brute_daily_insolation(P, h, R, tilts[0])
        out[f"{t}_yearly"] = climate.calc_yearly_insolation(P, h, R, 23.44)
    out["seasonal_tilt"] = np.array([climate.calculate_seasonal_tilt(23.44, d) for d in range(360)])
    save("climate.npz", **out)


def main():
    print("reference:", REF)
    gen_init()
    gen_noise()
    gen_fbm()
    gen_adjacency()
    gen_assembly()
    gen_erosion()
    gen_export()
    gen_climate()
    # provenance
    with open(os.path.join(HERE, "PROVENANCE.txt"), "w") as f:
        import numba
        f.write("Generated by tests/golden/gen_golden.py from the unmodified reference at /root/reference\n")
        f.write(f"numpy {np.__version__}, numba {numba.__version__}, python {sys.version.split()[0]}\n")
        for fn in sorted(os.listdir(HERE)):
            if fn.endswith(".npz"):
                sha = hashlib.sha1(open(os.path.join(HERE, fn), "rb").read()).hexdigest()[:16]
                f.write(f"{fn} sha1={sha}\n")
        for src in ("opensimplex.py", "terrain.py", "util.py", "erosion.py", "climate.py"):
            sha = hashlib.sha1(open(os.path.join(REF, src), "rb").read()).hexdigest()[:16]
            f.write(f"reference/{src} sha1={sha}\n")


if __name__ == "__main__":
    main()
