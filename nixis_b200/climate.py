"""Per-vertex climate kernels with the reference's names and argument order (climate.py:167-201,
345-597).  numpy float64 `verts` in -> numpy float32 out, like the reference; a
float64 CUDA tensor for `verts` keeps everything on the device and returns CUDA tensors.

Every driver that loops over rotations or days in the reference (360 full passes over the vertices)
is one kernel launch here (csrc/nxb_climate.cu).  Snapshots (`snapshot=True`: one PNG per step,
climate.py:471-480, 589-595) are file I/O and not provided.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from . import runtime as rt

SBC = 5.670374419 * 10**-8  # climate.py:9


def day2deg(year_length, day):
    return (day / year_length) * 360.0


def deg2day(year_length, degrees):
    return degrees * (year_length / 360.0)


def calculate_seasonal_tilt(axial_tilt, degrees):
    """climate.py:193-201 (host scalar)."""
    return _lib.load().nxb_climate_seasonal_tilt(float(axial_tilt), float(degrees))


def calculate_tsi(star_radius, star_temp, orbital_distance):
    """climate.py:203-210 (host scalar)."""
    energy_at_sun = SBC * star_temp**4 * (4 * np.pi * star_radius**2)
    return energy_at_sun / (4 * np.pi * orbital_distance**2)


def _verts(verts):
    """-> (float64 CUDA tensor [n,3], came_from_device)."""
    if isinstance(verts, torch.Tensor):
        assert verts.is_cuda and verts.dtype == torch.float64 and verts.is_contiguous()
        return verts, True
    return rt.upload(np.ascontiguousarray(verts, dtype=np.float64)), False


def _ret(x, dev):
    return x if dev else rt._to_host(x)


def _dbl(vals):
    return (C.c_double * len(vals))(*[float(v) for v in vals])


def _insolation(verts_dev, radius, rotations, tilts, arr=None):
    n = verts_dev.shape[0]
    if arr is None:
        arr = torch.zeros((len(tilts), n), dtype=torch.float32, device=verts_dev.device)
    scratch = torch.empty(2 * len(tilts), dtype=torch.float64, device=verts_dev.device)
    _lib.call("nxb_climate_insolation_f32", rt._ptr(verts_dev), n, C.c_double(radius), _dbl(rotations), len(rotations),
              _dbl(tilts), len(tilts), rt._ptr(scratch), rt._ptr(arr), rt._stream())
    return arr


def _rotation_sweep():
    """rotation = -180; rotation += 360/360, 360 times (climate.py:463-470, 518-523)."""
    rot, out = -180.0, []
    for _ in range(360):
        out.append(rot)
        rot += 360.0 / 360
    return out


def assign_surface_temp(verts, altitudes, radius, tilt):
    """climate.py:345-372.  `altitudes` only enters through rescale(altitudes, 0, alt_intensity) with the
    literal alt_intensity = 0, i.e. not at all."""
    v, dev = _verts(verts)
    out = torch.empty(v.shape[0], dtype=torch.float32, device=v.device)
    _lib.call("nxb_climate_surface_temp_f32", rt._ptr(v), v.shape[0], C.c_double(radius), C.c_double(tilt), rt._ptr(out), rt._stream())
    return _ret(out, dev)


def sample_insolation(arr, verts, radius, rotation, tilt):
    """climate.py:415-448: accumulates one rotation's insolation into the float32 array `arr` in place."""
    v, _ = _verts(verts)
    if isinstance(arr, torch.Tensor):
        _insolation(v, radius, [rotation], [tilt], arr.view(1, -1))
        return
    assert arr.dtype == np.float32
    a = rt.upload(np.ascontiguousarray(arr)).view(1, -1)
    _insolation(v, radius, [rotation], [tilt], a)
    arr[...] = rt._to_host(a.view(-1))


def brute_daily_insolation(verts, altitudes, radius, tilt, snapshot=False):
    """climate.py:450-490: 360 rotations of sample_insolation, one launch."""
    if snapshot:
        raise NotImplementedError("snapshots are file I/O (climate.py:471-480)")
    v, dev = _verts(verts)
    return _ret(_insolation(v, radius, _rotation_sweep(), [tilt]).view(-1), dev)


def calc_instant_insolation(verts, altitudes, radius, rotation, tilt):
    """climate.py:492-501."""
    v, dev = _verts(verts)
    return _ret(_insolation(v, radius, [rotation], [tilt]).view(-1), dev)


def _slice_tables(radius, tilts, device):
    """calc_insolation_slice for every tilt at once: float32 CUDA [len(tilts), 181]."""
    host = (C.c_double * (181 * 3))()
    _lib.call("nxb_climate_slice_verts", C.c_double(radius), host)
    sv = rt.upload(np.frombuffer(host, dtype=np.float64).reshape(181, 3).copy())
    return _insolation(sv, radius, _rotation_sweep(), list(tilts))


def calc_insolation_slice(radius, tilt):
    """climate.py:503-536: float32[181] lookup table (index = integer latitude, negatives from the end)."""
    rt.require_cuda()
    return rt._to_host(_slice_tables(radius, [tilt], None).view(-1))


def interpolate_insolation(verts, lookup_table, insolation, radius):
    """climate.py:551-577: fills `insolation` (float32) in place."""
    v, _ = _verts(verts)
    tab = lookup_table if isinstance(lookup_table, torch.Tensor) else rt.upload(np.ascontiguousarray(lookup_table, dtype=np.float32))
    out = insolation if isinstance(insolation, torch.Tensor) else torch.empty(v.shape[0], dtype=torch.float32, device=v.device)
    _lib.call("nxb_climate_interpolate_f32", rt._ptr(v), v.shape[0], C.c_double(radius), rt._ptr(tab), 1, rt._ptr(out), rt._stream())
    if not isinstance(insolation, torch.Tensor):
        insolation[...] = rt._to_host(out)


def calc_daily_insolation(verts, altitudes, radius, tilt):
    """climate.py:539-548."""
    v, dev = _verts(verts)
    tab = _slice_tables(radius, [tilt], v.device)
    out = torch.empty(v.shape[0], dtype=torch.float32, device=v.device)
    _lib.call("nxb_climate_interpolate_f32", rt._ptr(v), v.shape[0], C.c_double(radius), rt._ptr(tab), 1, rt._ptr(out), rt._stream())
    return _ret(out, dev)


def calc_yearly_insolation(points, height, radius, axial_tilt, snapshot=False):
    """climate.py:579-597: 360 days x (slice table + interpolation) in two launches."""
    if snapshot:
        raise NotImplementedError("snapshots are file I/O (climate.py:589-595)")
    v, dev = _verts(points)
    tilts = [calculate_seasonal_tilt(axial_tilt, x) for x in range(360)]
    tabs = _slice_tables(radius, tilts, v.device)
    out = torch.empty(v.shape[0], dtype=torch.float32, device=v.device)
    _lib.call("nxb_climate_interpolate_f32", rt._ptr(v), v.shape[0], C.c_double(radius), rt._ptr(tabs), 360, rt._ptr(out), rt._stream())
    return _ret(out, dev)
