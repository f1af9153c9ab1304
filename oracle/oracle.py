"""ctypes front-end of the CPU oracle (oracle/nixis_oracle.c) -- TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
leg may import this module.  The product package (nixis_b200) never does.

Function names and argument order mirror the reference's Python surface
(opensimplex.py / terrain.py / util.py / erosion.py) so parity tests read like
calls into the reference.  Everything is float64, like the reference.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libnixis_oracle.so")
_lib = None

_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")
_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")


def build(force=False):
    """Compile the oracle with gcc (seconds).  Building the checker is not using it."""
    srcs = [os.path.join(_HERE, s) for s in ("nixis_oracle.c", "nixis_oracle4.c", "nixis_oracle_climate.c")]
    srcs = [s for s in srcs if os.path.exists(s)]
    if (not force and os.path.exists(_SO)
            and all(os.path.getmtime(_SO) >= os.path.getmtime(s) for s in srcs)):
        return _SO
    cmd = ["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fopenmp", "-fPIC",
           "-fvisibility=hidden", "-shared", "-o", _SO] + srcs + ["-lm"]
    subprocess.check_call(cmd)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.nxo_init.argtypes = [C.c_int64, _i32p, _i32p]
        L.nxo_noise2.restype = C.c_double
        L.nxo_noise2.argtypes = [C.c_double, C.c_double, _i32p]
        L.nxo_noise3.restype = C.c_double
        L.nxo_noise3.argtypes = [C.c_double] * 3 + [_i32p, _i32p]
        L.nxo_noise3_array.argtypes = [C.c_int64, _f64p, _f64p, _f64p, _i32p, _i32p, _f64p]
        L.nxo_noise2_array.argtypes = [C.c_int64, _f64p, _f64p, _i32p, _f64p]
        if hasattr(L, "nxo_noise4"):
            L.nxo_noise4.restype = C.c_double
            L.nxo_noise4.argtypes = [C.c_double] * 4 + [_i32p]
            L.nxo_noise4_array.argtypes = [C.c_int64, _f64p, _f64p, _f64p, _f64p, _i32p, _f64p]
            L.nxo_sample_octaves4.argtypes = [C.c_int64, _f64p, _f64p, _i32p, C.c_int] + [C.c_double] * 6 + [C.c_int]
        L.nxo_sample_octaves.argtypes = [C.c_int64, _f64p, _f64p, _i32p, _i32p, C.c_int] + [C.c_double] * 5 + [C.c_int]
        L.nxo_mask_le.argtypes = [C.c_int64, _f64p, C.c_double, _u8p]
        L.nxo_rescale.restype = C.c_int
        L.nxo_rescale.argtypes = [C.c_int64, _f64p, _f64p, C.c_double, C.c_double, C.c_int, C.c_double,
                                  C.c_int, C.c_int, C.c_double, C.c_int, C.c_double]
        L.nxo_power_rescale.argtypes = [C.c_int64, _f64p, _u8p, C.c_int, C.c_double, _f64p, _f64p]
        L.nxo_find_percent_val.restype = C.c_double
        L.nxo_find_percent_val.argtypes = [C.c_double] * 3
        L.nxo_build_adjacency.restype = C.c_int
        L.nxo_build_adjacency.argtypes = [C.c_int64, _i64p, _i32p]
        L.nxo_sort_adjacency.argtypes = [C.c_int64, _i32p, _i32p]
        L.nxo_erosion_iteration1.argtypes = [C.c_int64, _i32p, _f64p, _f64p]
        L.nxo_erode_terrain1.argtypes = [C.c_int64, _i32p, _f64p, C.c_int]
        L.nxo_erosion_iteration3.argtypes = [C.c_int64, _f64p, _i32p, _f64p, _f64p, _f64p]
        L.nxo_erode_terrain3.argtypes = [C.c_int64, _f64p, _i32p, _f64p, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        L.nxo_num_threads.restype = C.c_int
        L.nxo_set_num_threads.argtypes = [C.c_int]
        if hasattr(L, "nxo_sample_insolation"):
            L.nxo_seasonal_tilt.restype = C.c_double
            L.nxo_seasonal_tilt.argtypes = [C.c_double, C.c_double]
            L.nxo_assign_surface_temp.argtypes = [C.c_int64, _f64p, _f64p, C.c_double, C.c_double, _f32p]
            L.nxo_sample_insolation.argtypes = [C.c_int64, _f32p, _f64p, C.c_double, C.c_double, C.c_double, C.c_int, C.c_double]
            L.nxo_insolation_slice.argtypes = [C.c_double, C.c_double, _f32p]
            L.nxo_interpolate_insolation.argtypes = [C.c_int64, _f64p, _f32p, _f32p, C.c_double]
            L.nxo_daily_insolation.argtypes = [C.c_int64, _f64p, C.c_double, C.c_double, _f32p]
            L.nxo_yearly_insolation.argtypes = [C.c_int64, _f64p, C.c_double, C.c_double, _f32p]
        _lib = L
    return _lib


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


# ---- opensimplex.py -------------------------------------------------------
def init(seed=0):
    perm = np.zeros(256, np.int32)
    pgi = np.zeros(256, np.int32)
    seed = int(seed)
    seed = (seed + 2 ** 63) % 2 ** 64 - 2 ** 63   # numba int64 wrap
    lib().nxo_init(seed, perm, pgi)
    return perm, pgi


def noise2d(x, y, perm):
    return lib().nxo_noise2(x, y, perm)


def noise3d(x, y, z, perm, pgi):
    return lib().nxo_noise3(x, y, z, perm, pgi)


def noise4d(x, y, z, w, perm):
    return lib().nxo_noise4(x, y, z, w, perm)


def noisearr2d(x, y, perm):
    x, y = _f64(x), _f64(y)
    out = np.empty(x.size)
    lib().nxo_noise2_array(x.size, x, y, perm, out)
    return out


def noisearr3d(x, y, z, perm, pgi):
    x, y, z = _f64(x), _f64(y), _f64(z)
    out = np.empty(x.size)
    lib().nxo_noise3_array(x.size, x, y, z, perm, pgi, out)
    return out


def noisearr4d(x, y, z, w, perm):
    x, y, z, w = _f64(x), _f64(y), _f64(z), _f64(w)
    out = np.empty(x.size)
    lib().nxo_noise4_array(x.size, x, y, z, w, perm, out)
    return out


# ---- terrain.py -----------------------------------------------------------
def sample_octaves(verts, elevations, perm, pgi, n_octaves=1, n_init_roughness=1.5,
                   n_init_strength=0.4, n_roughness=2.0, n_persistence=0.5, world_radius=1.0,
                   nthreads=0):
    verts = _f64(verts)
    if elevations is None:
        elevations = np.zeros(len(verts), dtype=np.float64)
    lib().nxo_sample_octaves(len(verts), verts, elevations, perm, pgi, n_octaves,
                             n_init_roughness, n_init_strength, n_roughness, n_persistence,
                             world_radius, nthreads)
    return elevations


def sample_octaves4(verts, elevations, perm, n_octaves=1, n_init_roughness=1.5,
                    n_init_strength=0.4, n_roughness=2.0, n_persistence=0.5, world_radius=1.0,
                    w_scale=0.5, nthreads=0):
    """4-D fBm driver (builder-defined, the reference has none): w = w_scale * freq."""
    verts = _f64(verts)
    if elevations is None:
        elevations = np.zeros(len(verts), dtype=np.float64)
    lib().nxo_sample_octaves4(len(verts), verts, elevations, perm, n_octaves,
                              n_init_roughness, n_init_strength, n_roughness, n_persistence,
                              world_radius, w_scale, nthreads)
    return elevations


def make_bool_elevation_mask(height, mask_elevation):
    height = _f64(height)
    m = np.zeros(len(height), np.uint8)
    lib().nxo_mask_le(len(height), height, mask_elevation, m)
    return m.view(np.bool_)


# ---- util.py --------------------------------------------------------------
_MODES = {None: 0, "lower": 1, "upper": 2}


def rescale(x, lower, upper, mid=None, mode=None, u_min=None, u_max=None):
    x = _f64(x)
    out = np.empty_like(x)
    rc = lib().nxo_rescale(len(x), x, out, lower, upper, mid is not None, mid or 0.0, _MODES[mode],
                           u_min is not None, u_min or 0.0, u_max is not None, u_max or 0.0)
    return x if rc else out


def power_rescale(x, mask=None, mode=None, power=1.0, return_stats=False):
    x = _f64(x)
    out = np.empty_like(x)
    stats = np.zeros(4)
    m = np.zeros(len(x), np.uint8) if mask is None else np.ascontiguousarray(mask).view(np.uint8)
    lib().nxo_power_rescale(len(x), x, m, -1 if mode is None else int(mode), power, out, stats)
    return (out, stats) if return_stats else out


def find_percent_val(minval, maxval, percent):
    return lib().nxo_find_percent_val(minval, maxval, percent)


def build_adjacency(triangles):
    tri = np.ascontiguousarray(triangles, dtype=np.int64)
    adj = np.empty(((len(tri) + 4) // 2, 6), np.int32)
    rc = lib().nxo_build_adjacency(len(tri), tri, adj)
    if rc:
        raise ValueError("adjacency row overflow (inconsistent winding?)")
    return adj


def sort_adjacency(adj):
    """In place, like the reference (util.py:638-662)."""
    src = adj.copy()
    lib().nxo_sort_adjacency(len(adj), src, adj)


# ---- erosion.py -----------------------------------------------------------
def erosion_iteration1(neighbors, r_buff, w_buff):
    lib().nxo_erosion_iteration1(len(neighbors), neighbors, r_buff, w_buff)
    return w_buff


def erode_terrain1(nodes, neighbors, heights, num_iter=1, snapshot=None):
    lib().nxo_erode_terrain1(len(neighbors), neighbors, heights, num_iter)
    return heights


def erosion_iteration3(verts, neighbors, r_buff, wat, sed):
    lib().nxo_erosion_iteration3(len(neighbors), _f64(verts), neighbors, r_buff, wat, sed)


def erode_terrain3(nodes, neighbors, heights, num_iter=1, snapshot=False, return_state=False,
                   nthreads=0):
    nodes = _f64(nodes)
    if return_state:
        wat = np.zeros_like(heights)
        sed = np.zeros_like(heights)
        lib().nxo_erode_terrain3(len(neighbors), nodes, neighbors, heights, num_iter,
                                 wat.ctypes.data, sed.ctypes.data, nthreads)
        return wat, sed
    lib().nxo_erode_terrain3(len(neighbors), nodes, neighbors, heights, num_iter, None, None, nthreads)


def height_assembly(height, min_alt=-4000, max_alt=8850, ocean_percent=55.0):
    """nixis.py:332-364 call sequence.  Returns (height, ocean mask, ocean level)."""
    height = rescale(height, min_alt, max_alt)
    minval, maxval = np.amin(height), np.amax(height)
    ocean_level = find_percent_val(minval, maxval, ocean_percent)
    ocean = make_bool_elevation_mask(height, ocean_level)
    height = power_rescale(height, mask=ocean, mode=1, power=0.5)
    height = power_rescale(height, mask=ocean, mode=0, power=2.0)
    height -= ocean_level
    height = rescale(height, min_alt, max_alt, mid=0)
    return height, ocean, ocean_level


def set_num_threads(n):
    """Override OMP_NUM_THREADS for the oracle's OpenMP loops (torchrun sets it to 1 in its workers)."""
    lib().nxo_set_num_threads(int(n))


def num_threads():
    return lib().nxo_num_threads()


# ---- climate.py -------------------------------------------
def calculate_seasonal_tilt(axial_tilt, degrees):
    return lib().nxo_seasonal_tilt(float(axial_tilt), float(degrees))


def assign_surface_temp(verts, altitudes, radius, tilt):
    v = _f64(verts)
    out = np.zeros(len(v), dtype=np.float32)
    lib().nxo_assign_surface_temp(len(v), v, _f64(altitudes), float(radius), float(tilt), out)
    return out


def sample_insolation(arr, verts, radius, rotation, tilt):
    """In place on the float32 array `arr` (climate.py:415-448)."""
    v = _f64(verts)
    assert arr.dtype == np.float32 and arr.flags.c_contiguous
    lib().nxo_sample_insolation(len(v), arr, v, float(radius), float(rotation), 0.0, 1, float(tilt))


def brute_daily_insolation(verts, altitudes, radius, tilt, snapshot=False):
    v = _f64(verts)
    out = np.zeros(len(v), dtype=np.float32)
    lib().nxo_sample_insolation(len(v), out, v, float(radius), -180.0, 360.0 / 360, 360, float(tilt))
    return out


def calc_instant_insolation(verts, altitudes, radius, rotation, tilt):
    out = np.zeros(len(verts), dtype=np.float32)
    sample_insolation(out, verts, radius, rotation, tilt)
    return out


def calc_insolation_slice(radius, tilt):
    out = np.zeros(181, dtype=np.float32)
    lib().nxo_insolation_slice(float(radius), float(tilt), out)
    return out


def interpolate_insolation(verts, lookup_table, insolation, radius):
    v = _f64(verts)
    lib().nxo_interpolate_insolation(len(v), v, np.ascontiguousarray(lookup_table, dtype=np.float32), insolation, float(radius))


def calc_daily_insolation(verts, altitudes, radius, tilt):
    v = _f64(verts)
    out = np.zeros(len(v), dtype=np.float32)
    lib().nxo_daily_insolation(len(v), v, float(radius), float(tilt), out)
    return out


def calc_yearly_insolation(points, height, radius, axial_tilt, snapshot=False):
    v = _f64(points)
    out = np.zeros(len(v), dtype=np.float32)
    lib().nxo_yearly_insolation(len(v), v, float(radius), float(axial_tilt), out)
    return out
