"""Drop-in for the hot-path half of the reference's `util` module (util.py) on B200.

Kept: `create_mesh`, `rescale`, `power_rescale`, `find_percent_val`, `build_adjacency`,
`sort_adjacency` -- same names, argument order, defaults, return / in-place conventions -- and,
from the "next" rows of SURVEY section 8f, the equirectangular export chain `make_ll_arr`,
`build_KDTree` (+ `cfg.KDT.query`), `make_gray_array`, `build_image_data`.
File I/O (save_image / save_mesh / settings), lat/lon helpers and memory pretty-printers are
outside the path and not provided here.

numpy in -> numpy out (float64 / int32, like the reference); `DeviceMesh` and CUDA tensors
in -> CUDA tensors out (the resident path bench.py and the multi-GPU driver use).
"""
import os
import time

import numpy as np
import torch

from . import runtime as rt
from . import shard

try:                    # inside the reference tree: share the reference's globals (cfg.py:3-14)
    import cfg
    if not hasattr(cfg, "IMG_QUERY_DATA"):
        raise ImportError
except ImportError:
    from . import cfg


class DeviceMesh:
    """Icosphere resident in HBM: unit-sphere float4 positions (+ radius), optional int32 cells
    and neighbour table.  Vertex order = meshzoo order = the reference's `points` order."""

    def __init__(self, k, xyz, radius=1.0, cells=None, adj=None, v_begin=0):
        self.k = int(k)
        self.xyz = xyz                  # float32 [n,4] unit sphere
        self.radius = float(radius)
        self.cells = cells              # int32 [T,3] or None
        self.adj = adj                  # int32 [V,6] or None
        self.v_begin = int(v_begin)     # first global vertex id held (multi-GPU shards)

    @property
    def n_vertices(self):
        return self.xyz.shape[0]

    def points64(self):
        """float64 [n,3] unit-sphere positions on the device (closed form, generated once and kept):
        what the fBm kernel reads -- the reference evaluates noise on float64 vertices, and the lattice
        cell / candidate selection must see the same numbers to make the same decisions."""
        if getattr(self, "_p64", None) is None:
            self._p64 = rt.mesh_points(self.k, self.v_begin, self.v_begin + self.n_vertices, f32=False, f64=True,
                                       device=self.xyz.device)[1]
        return self._p64

    def points_numpy(self):
        """float64 [V,3] radius-scaled positions, like `points` after nixis.py:249."""
        p64 = rt.mesh_points(self.k, self.v_begin, self.v_begin + self.n_vertices, f32=False, f64=True,
                             device=self.xyz.device)[1]
        return (p64 * self.radius).cpu().numpy()


def _device_xyz(verts, mult):
    """-> (float32 [n,4] CUDA positions, factor the caller must apply to frequencies)."""
    if isinstance(verts, DeviceMesh):
        return verts.xyz, verts.radius * mult
    if isinstance(verts, torch.Tensor):
        assert verts.is_cuda and verts.dtype == torch.float32 and verts.shape[1] == 4
        return verts, mult
    v = np.ascontiguousarray(verts, dtype=np.float64)
    assert v.ndim == 2 and v.shape[1] == 3, "verts must be [V,3]"
    return rt.xyz_from_f64(rt.upload(v), mult), 1.0


def create_mesh(divisions, device=False, radius=1.0, with_cells=True, verbose=True):
    """Icosphere in meshzoo.icosa_sphere order (util.py:17-50).

    device=False: (points float64 [V,3] on the unit sphere, cells int64 [T,3]) numpy arrays,
    the reference's return value.  device=True: a `DeviceMesh` (nothing leaves the GPU).
    Vertex arithmetic is FP64 on the device, so the numpy points are full double precision."""
    rt.require_cuda()
    k = int(divisions)
    if verbose:
        print("Generating the mesh...")
        print(f"k is {k}")
    t0 = time.perf_counter()
    if device:
        xyz, _ = rt.mesh_points(k)
        cells = rt.mesh_cells(k) if with_cells else None
        mesh = DeviceMesh(k, xyz, radius, cells)
        if verbose:
            torch.cuda.synchronize()
            print(f"Number of vertices: {xyz.shape[0]:,}")
            print(f"Number of triangles: {20 * k * k:,}")
            print(f"Mesh generated in {time.perf_counter() - t0 :.5f} sec")
        return mesh
    _, p64 = rt.mesh_points(k, f32=False, f64=True)
    cells = rt.mesh_cells(k)
    points = p64.cpu().numpy()
    cells = cells.cpu().numpy().astype(np.int64)       # meshzoo: dtype=int
    if verbose:
        print(f"Number of vertices: {points.shape[0]:,}")
        print(f"Number of triangles: {cells.shape[0]:,}")
        print(f"Mesh generated in {time.perf_counter() - t0 :.5f} sec")
    return points, cells


_MODES = {None: 0, "lower": 1, "upper": 2}


def rescale(x, lower, upper, mid=None, mode=None, u_min=None, u_max=None):
    """Re-scale (normalize) an array to a given lower and upper bound (util.py:110-175).
    Device arithmetic is FP32: float64 numpy input is rounded to float32 first (min / max and the
    mapping are then exact for those values); the float64 result agrees with the reference's to 1e-6
    of the output range."""
    if mode is not None and mid is None:
        print("ERROR: Must supply a middle value to use rescale modes.")
        print("Continuing with unmodified data.")
        return x
    if mode not in _MODES:
        return None          # the reference falls off the end of the function (util.py:161-175)
    dev_io = isinstance(x, torch.Tensor)
    xd = x if dev_io else rt.upload_f32(x)
    x_min, x_max = shard.collective().minmax(rt.minmax(xd))      # whole-planet min / max under a shard context
    if u_min is not None and u_min < x_min:
        x_min = u_min
    if u_max is not None and u_max > x_max:
        x_max = u_max
    out = rt.rescale(xd, x_min, x_max, lower, upper, mid, _MODES[mode])
    return out if dev_io else rt.download_f64(out)


def power_rescale(x, mask=None, mode=None, power=1.0, verbose=True, shift=0.0):
    """Rescale values using a power function (util.py:178-254).  `shift` (extension, default 0)
    is subtracted from the result, fusing nixis.py:359.  Device arithmetic is FP32 (float64 numpy input
    is rounded first; the order-dependent if / elif scan of util.py:203-214 is reproduced exactly on the
    rounded values)."""
    dev_io = isinstance(x, torch.Tensor)
    xd = x if dev_io else rt.upload_f32(x)
    coll = shard.collective()
    x_min, x_max = coll.minmax(rt.minmax(xd))
    if mode in (0, 1) and mask is not None:
        md = mask if isinstance(mask, torch.Tensor) else rt.upload(np.ascontiguousarray(mask).view(np.uint8))
        # the sequential if / elif scan (util.py:203-214) over the whole planet = the per-shard ordered
        # summaries combined in rank order
        summary = rt.combine_power_summaries(coll.gather_summaries(rt.power_summary(xd, md, int(mode))))
        lo, hi = rt.power_bounds(summary, x_min, x_max)
        sel = int(mode)
    else:
        md, lo, hi, sel = None, x_max, x_min, -1
    if verbose:
        print("x min:", x_min)
        print("x max:", x_max)
        print("mask min:", lo)
        print("mask max:", hi)
    out = rt.power_apply(xd, md, sel, lo, hi, power, shift)
    return out if dev_io else rt.download_f64(out)


def find_percent_val(minval, maxval, percent):
    """Value that is `percent` of the way from minval to maxval (util.py:556-566). Host scalar."""
    if not 0.0 < percent < 100.0:
        print("\n" + "    ERROR: Percent must be between 0 and 100.")
        print("    Defaulting to 50 percent." + "\n")
        percent = 50.0
    return minval + ((maxval - minval) * percent / 100.0)


def build_adjacency(triangles):
    """int32 [V,6] neighbour table, -1 padded, rows in the reference's slot order (util.py:591-613)."""
    if isinstance(triangles, DeviceMesh):
        triangles.adj = rt.adj_build(triangles.cells, 10 * triangles.k ** 2 + 2)
        return triangles.adj
    if isinstance(triangles, torch.Tensor):
        return rt.adj_build(triangles, (triangles.shape[0] + 4) // 2)
    tri = np.ascontiguousarray(triangles)
    V = int((len(tri) + 4) / 2)
    cells = rt.upload(tri.astype(np.int32, copy=False))
    return rt.adj_build(cells, V).cpu().numpy()


def sort_adjacency(adj):
    """Order every row as a ring walk (util.py:623-662).  numpy: in place, returns None, like the
    reference.  DeviceMesh / CUDA tensor: returns the sorted table (DeviceMesh.adj is replaced)."""
    if isinstance(adj, DeviceMesh):
        adj.adj = rt.adj_sort(adj.adj)
        return adj.adj
    if isinstance(adj, torch.Tensor):
        return rt.adj_sort(adj)
    a = np.ascontiguousarray(adj, dtype=np.int32)
    adj[...] = rt.adj_sort(rt.upload(a)).cpu().numpy()
    return None


# ---------------------------------------------------------------------------------------------
# Equirectangular export: util.py:275-429, nixis.py:270-302, 382-389
def make_ll_arr(width, height, radius):
    """XYZ of each pixel's latitude / longitude (util.py:290-308): float64 [height, width, 3]."""
    return rt._to_host(rt.ll_grid(width, height, radius))


class IcoNearest:
    """Stands in for scipy.spatial.KDTree over the icosphere's vertices (util.py:275-285,
    nixis.py:279-283): `.query(x, k=3)` returns (distances, indices) of the 3 nearest vertices, found
    analytically on the closed-form mesh instead of through a tree (26 s to build at k=2500)."""

    def __init__(self, points):
        if isinstance(points, DeviceMesh):
            self.k, self.radius = points.k, points.radius
            return
        pts = np.asarray(points)
        V = len(pts)
        k = int(round(((V - 2) / 10.0) ** 0.5))
        if 10 * k * k + 2 != V:
            raise ValueError("build_KDTree: the analytic nearest-vertex search needs the k-division icosphere "
                             f"(10k^2+2 vertices), got {V} points")
        self.k = k
        self.radius = float(np.linalg.norm(pts[0]))
        # the points must BE the meshzoo-ordered icosphere: compare a sample with the closed form
        probe = np.unique(np.linspace(0, V - 1, 64).astype(np.int64))
        ref = rt.mesh_points(k, f32=False, f64=True)[1][torch.from_numpy(probe).cuda()].cpu().numpy() * self.radius
        if not np.allclose(pts[probe], ref, rtol=1e-9, atol=1e-9 * self.radius):
            raise ValueError("build_KDTree: points are not the meshzoo-ordered icosphere this library generates")

    def query(self, x, k=3, workers=1, **_):
        if k != 3:
            raise NotImplementedError("only the k=3 query of nixis.py:283 is provided")
        dev_io = isinstance(x, torch.Tensor)
        q = x if dev_io else rt.upload(np.ascontiguousarray(x, dtype=np.float64))
        d, i = rt.ico_nearest3(self.k, self.radius, q)
        return (d, i) if dev_io else (rt._to_host(d), rt._to_host(i))


def build_KDTree(points, lf=10):
    """util.py:275-285: sets cfg.KDT (here an `IcoNearest`, no tree is built)."""
    print("Building KD Tree...")
    t0 = time.perf_counter()
    cfg.KDT = IcoNearest(points)
    print(f"KD built in {time.perf_counter() - t0 :.5f} sec")


def make_gray_array(width, height, dists, nbrs, colors):
    """Sample vertices and build a grayscale map (util.py:343-367): inverse-distance blend of the 3
    nearest vertices, int() truncation, result cast back to the dtype of `colors`."""
    dev_io = isinstance(dists, torch.Tensor)
    d = dists if dev_io else rt.upload(np.ascontiguousarray(dists, dtype=np.float64))
    i = nbrs if isinstance(nbrs, torch.Tensor) else rt.upload(np.ascontiguousarray(nbrs, dtype=np.int64))
    if isinstance(colors, torch.Tensor):
        c64, orig = colors.to(torch.float64), None
    else:
        colors = np.asarray(colors)
        c64, orig = rt.upload(np.ascontiguousarray(colors.astype(np.float64))), colors.dtype
    out = rt.idw_gray(d.reshape(height, width, 3), i.reshape(height, width, 3), c64)
    if dev_io:
        return out
    return rt._to_host(out).astype(orig)


def make_rgb_array(width, height, dists, nbrs, colors):
    """Sample vertices and build an RGB map (util.py:310-341): the same blended value in all three
    channels, int32 [height, width, 3] cast back to the dtype of `colors`."""
    gray = make_gray_array(width, height, dists, nbrs, colors)
    if isinstance(gray, torch.Tensor):
        return gray.unsqueeze(-1).expand(-1, -1, 3).contiguous()
    return np.repeat(gray[:, :, None], 3, axis=2)


# dtype policy of build_image_data (util.py:394-402, 421-427), as data.  Arrays are pre-scaled to
# 0..255 (as float64) before blending unless they are uint16 / uint32; the exported pixels are uint8
# except for uint16 sources, which stay uint16 (a pre-scaled array is float64 by then, and float64 /
# uint32 are on the reference's narrow-to-uint8 list).
_KEEP_UNSCALED = {"u2": np.uint16, "u4": np.uint8}      # source kind -> exported dtype; everything else: pre-scale, uint8


def _dtype_key(dt):
    dt = np.dtype(dt)
    return "b" if dt == np.bool_ else f"{dt.kind}{dt.itemsize}"


def build_image_data(colors=None, width=None, height=None):
    """Use the nearest-vertex query results to build the maps for export (util.py:369-429).
    colors: dict name -> [array, mode] with mode 'gray' / 'rgb'.  Width / height default to the shape
    of cfg.IMG_QUERY_DATA (the reference re-reads options.json, util.py:381-383)."""
    print("Sampling verts for texture...")
    dists, nbrs = cfg.IMG_QUERY_DATA[0], cfg.IMG_QUERY_DATA[1]
    height = height or dists.shape[0]
    width = width or dists.shape[1]
    if not isinstance(colors, dict):
        print("ERROR: Must pass a dict when saving out texture maps.")
    d_dev = dists if isinstance(dists, torch.Tensor) else rt.upload(np.ascontiguousarray(dists, dtype=np.float64))
    i_dev = nbrs if isinstance(nbrs, torch.Tensor) else rt.upload(np.ascontiguousarray(nbrs, dtype=np.int64))
    d_dev, i_dev = d_dev.reshape(height, width, 3), i_dev.reshape(height, width, 3)
    result = {}
    for key in list(colors):
        array, mode = colors[key][0], str(colors[key][1]).lower()
        src = np.asarray(array)
        kind = _dtype_key(src.dtype)
        values = src
        if kind not in _KEEP_UNSCALED:
            values = rescale(src.astype(np.float64), 0, 255)
        t0 = time.perf_counter()
        if mode not in ("gray", "grey", "rgb"):
            raise ValueError(f"build_image_data: unknown mode {colors[key][1]!r} for map {key!r}")
        c64 = rt.upload(np.ascontiguousarray(np.asarray(values).astype(np.float64)))
        pixels = rt._to_host(rt.idw_gray(d_dev, i_dev, c64))
        if mode == "rgb":
            pixels = np.repeat(pixels[:, :, None], 3, axis=2)
        colors[key] = None
        print(f"  {key} pixels built in   {time.perf_counter() - t0 :.5f} sec")
        # make_gray_array / make_rgb_array cast back to the dtype of the colours they were given
        # (util.py:341, 367), then the export dtype of the table above applies (util.py:424-427)
        pixels = pixels.astype(np.asarray(values).dtype)
        result[key] = pixels.astype(_KEEP_UNSCALED.get(kind, np.uint8))
    return result


# ---------------------------------------------------------------------------------------------
# File I/O and scalar helpers: host-side, same names as util.py:59-98, 430-553.
# Nothing here touches the GPU; options default to the reference's options.json values.
DEFAULT_OPTIONS = {"img_format": "png", "mesh_format": "obj", "point_format": "ply", "settings_format": "json"}


def xyz2latlon(x, y, z, r):
    """util.py:59-75."""
    lat = np.degrees(np.arcsin(min(max((z / r), -1), 1)))
    lon = np.degrees(np.arctan2(y, x))
    return (lat, lon)


def latlon2xyz(lat, lon, r):
    """util.py:79-88."""
    x = r * np.cos(lat * (np.pi / 180)) * np.cos(lon * (np.pi / 180))
    y = r * np.cos(lat * (np.pi / 180)) * np.sin(lon * (np.pi / 180))
    z = r * np.sin(lat * (np.pi / 180))
    return (x, y, z)


def kelvin_to_c(k):
    return k - 273.15


def c_to_kelvin(c):
    return c + 273.15


def load_settings(path):
    """util.py:531-541: JSON file -> dict; a missing file prints and exits like the reference."""
    import json
    import sys
    if os.path.exists(path):
        with open(path, "rt") as f:
            return json.loads(f.read())
    print("Path does not exist:", path)
    sys.exit(0)


def save_settings(data, path, name, fmt=None):
    """util.py:524-529."""
    import json
    with open(os.path.join(path, f"{name}.{fmt}"), "w") as f:
        json.dump(data, f, indent=4)


def _options():
    return load_settings("options.json") if os.path.exists("options.json") else dict(DEFAULT_OPTIONS)


def save_image(data, path, name):
    """util.py:430-445: one image file per key, `name_key.fmt` (uint8 -> 8-bit, uint16 -> 16-bit gray)."""
    from PIL import Image
    fmt = _options()["img_format"]
    for key, array in data.items():
        if isinstance(array, torch.Tensor):
            array = rt._to_host(array)
        Image.fromarray(array).save(os.path.join(path, f"{name}_{key}.{fmt}"))


def image_to_array(image_file):
    """util.py:447-470: image -> float64 heights in [0, 1)."""
    import sys
    from PIL import Image
    img = Image.open(image_file)
    if img.mode == "L":
        return np.float64(np.asarray(img)) / 256
    if img.mode in ("I", "I;16"):
        return np.asarray(img) / 65536
    if img.mode in ("RGB", "RGBA"):
        return np.float64(np.asarray(img.convert("L"))) / 256
    print("ERROR. Unsupported image format.")
    sys.exit(-1)


def save_mesh(verts, tris, path, name, confirm=True):
    """util.py:472-497.  The reference hands the arrays to meshio (absent here); the configured
    format is Wavefront OBJ, written directly: `v x y z` per vertex, `f a b c` (1-based) per triangle."""
    if len(tris) > 3000000 and confirm:
        print("\n" + f"ATTENTION. This mesh has {len(tris):,} triangles. "
              "Your 3D modeling software may not be able to open/edit it." + "\n")
        if input("Continue anyway? Y/N: ").lower() not in ('y', 'yes'):
            print("A smaller division setting will reduce the number of tris.")
            return
    fmt = _options()["mesh_format"]
    if fmt != "obj":
        raise NotImplementedError(f"mesh_format {fmt!r}: only obj is written without meshio")
    print("Saving mesh to disk...")
    t0 = time.perf_counter()
    verts = np.asarray(verts.cpu() if isinstance(verts, torch.Tensor) else verts, dtype=np.float64)[:, :3]
    tris = np.asarray(tris.cpu() if isinstance(tris, torch.Tensor) else tris).astype(np.int64) + 1
    with open(os.path.join(path, f"{name}.{fmt}"), "w") as f:
        np.savetxt(f, verts, fmt="v %.17g %.17g %.17g")
        np.savetxt(f, tris, fmt="f %d %d %d")
    print(f"Mesh saved in {time.perf_counter() - t0 :.5f} sec")


def save_point_cloud(verts, path, name):
    """util.py:499-522."""
    print("Not implemented yet.")


def save_log(path, name, fmt=None):
    print("Not implemented yet.")


def export_planet(data, path, name):
    print("Not implemented yet.")
