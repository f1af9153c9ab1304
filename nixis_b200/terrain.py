"""Drop-in for the reference's `terrain` module (terrain.py) on B200.

`sample_octaves` keeps the reference signature and semantics (terrain.py:32-59): octave o
uses frequency f0*roughness^o and amplitude a0*persistence^o (advanced in float64 on the
host), each octave adds ((noise3d+1)*0.5)*amplitude, `elevations` is accumulated in place
when given, and the same progress lines are printed when `verbose`.  All octaves run in ONE
fused kernel (nxb_fbm3_f32) instead of one numba prange per octave.

Inputs may be numpy (float64, radius-scaled vertices like nixis.py:249 -- copied to the GPU,
result returned as float64 numpy) or a `util.DeviceMesh` / CUDA tensors (resident fast path,
CUDA float32 tensor returned).
"""
import time

import numpy as np
import torch

from . import runtime as rt
from .util import DeviceMesh, _device_xyz          # _device_xyz: float4 positions for the 4-D kernel


def _fbm3(verts, tables, nr, amp, init=None, out=None, minmax=None):
    """The fused fBm kernel on whatever the caller holds.  nr[o] multiplies the vertices AS GIVEN
    (terrain.py:17 `verts * n_roughness`); numpy / float64 tensors / DeviceMesh go through the
    float64-position entry point, float4 tensors are promoted inside the kernel."""
    if isinstance(verts, torch.Tensor) and verts.dtype == torch.float32:
        assert verts.is_cuda and verts.shape[1] == 4
        return rt.fbm3(tables, verts, nr, amp, init=init, out=out, minmax=minmax)
    if isinstance(verts, DeviceMesh):
        v64, scale = verts.points64(), verts.radius            # nixis.py:249 points *= world_radius
    elif isinstance(verts, torch.Tensor):
        assert verts.is_cuda and verts.dtype == torch.float64 and verts.shape[1] == 3
        v64, scale = verts, 1.0
    else:
        v = np.ascontiguousarray(verts, dtype=np.float64)
        assert v.ndim == 2 and v.shape[1] == 3, "verts must be [V,3]"
        v64, scale = rt.upload(v), 1.0
    return rt.fbm3_pos64(tables, v64, scale, nr, amp, init=init, out=out, minmax=minmax)


def sample_noise(verts, perm, pgi, n_roughness=1, n_strength=0.2, radius=1):
    """One octave (terrain.py:12-29): ((noise3d(verts*n_roughness)+1)*0.5)*n_strength*radius."""
    tables = rt.tables_for(perm, pgi)
    out = _fbm3(verts, tables, [float(n_roughness)], [float(n_strength) * float(radius)])
    return out if _is_device(verts) else rt.download_f64(out)


def _is_device(v):
    return isinstance(v, (DeviceMesh, torch.Tensor))


def sample_octaves(verts, elevations, perm, pgi, n_octaves=1, n_init_roughness=1.5, n_init_strength=0.4,
                   n_roughness=2.0, n_persistence=0.5, world_radius=1.0, verbose=True, minmax=None, exact=False):
    """Sample octaves of noise and combine them together (terrain.py:32-59).

    exact=False (default): throughput kernel -- float64 lattice coordinates and candidate selection
    (the reference's own decisions), FP32 contributions: within 1e-5 of the range of the reference's result.
    exact=True: float64 kernel without FMA contraction in the reference's operation order -- the
    result is BIT-IDENTICAL to the reference's for the same float64 vertices (about 4x slower)."""
    t0 = time.perf_counter()
    if exact:
        return _sample_octaves_exact(verts, elevations, perm, pgi, n_octaves, n_init_roughness, n_init_strength,
                                     n_roughness, n_persistence, world_radius)
    tables = rt.tables_for(perm, pgi)
    freq, amp = rt.octave_schedule(n_octaves, n_init_roughness, n_init_strength, n_roughness, n_persistence)
    device_io = _is_device(verts)
    init = None
    if elevations is not None:
        init = elevations if isinstance(elevations, torch.Tensor) else rt.upload_f32(elevations)
    out_t = init if isinstance(elevations, torch.Tensor) else None
    mm = minmax if minmax is not None else rt.new_minmax(torch.device("cuda", torch.cuda.current_device()))
    # terrain.py:43: every octave samples verts * (n_freq / world_radius)
    out = _fbm3(verts, tables, [f / float(world_radius) for f in freq], amp, init=init, out=out_t, minmax=mm)
    if verbose:
        torch.cuda.synchronize()
        print(f"  Octaves 1..{n_octaves} (fused): {time.perf_counter() - t0:.5f} sec")
        lo, hi = mm.tolist()
        print("  Combined octaves min:", lo)
        print("  Combined octaves max:", hi)
    if device_io or isinstance(elevations, torch.Tensor):
        return out
    if elevations is not None:               # in place, like `elevations +=` (terrain.py:43)
        return rt.download_f64(out, out=elevations)
    return rt.download_f64(out)


def _sample_octaves_exact(verts, elevations, perm, pgi, n_octaves, f0, a0, roughness, persistence, world_radius):
    tables = rt.tables_for(perm, pgi)
    scale = 1.0
    if isinstance(verts, DeviceMesh):
        # the mesh holds unit-sphere positions; nixis.py:249 scales them by the radius first
        v64 = rt.mesh_points(verts.k, verts.v_begin, verts.v_begin + verts.n_vertices, f32=False, f64=True,
                             device=verts.xyz.device)[1]
        scale = verts.radius
    elif isinstance(verts, torch.Tensor):
        assert verts.dtype == torch.float64 and verts.shape[1] == 3, "exact mode needs float64 [n,3] positions"
        v64 = verts
    else:
        v64 = rt.upload(np.ascontiguousarray(verts, dtype=np.float64))
    init = None
    if elevations is not None:
        init = elevations if isinstance(elevations, torch.Tensor) else rt.upload(np.ascontiguousarray(elevations, dtype=np.float64))
    out = rt.fbm3_exact(tables, v64, n_octaves, f0, a0, roughness, persistence, float(world_radius), scale, init)
    if _is_device(verts) or isinstance(elevations, torch.Tensor):
        return out
    host = rt._to_host(out)
    if elevations is not None:
        elevations[...] = host
        return elevations
    return host


def sample_octaves4(verts, elevations, perm, n_octaves=1, n_init_roughness=1.5, n_init_strength=0.4,
                    n_roughness=2.0, n_persistence=0.5, world_radius=1.0, w_scale=0.5, verbose=False):
    """4-D fBm.  NOT a reference function (nothing in nixis calls noise4d): defined
    by analogy with sample_octaves, the 4th coordinate of octave o is w_scale * freq_o."""
    xyz, fscale = _device_xyz(verts, 1.0 / float(world_radius))
    tables = rt.tables_for(perm, None)
    freq, amp = rt.octave_schedule(n_octaves, n_init_roughness, n_init_strength, n_roughness, n_persistence)
    w = [w_scale * f for f in freq]
    freq = [f * fscale for f in freq]
    init = None
    if elevations is not None:
        init = elevations if isinstance(elevations, torch.Tensor) else rt.upload_f32(elevations)
    out = rt.fbm4(tables, xyz, freq, amp, w, init=init)
    if _is_device(verts) or isinstance(elevations, torch.Tensor):
        return out
    if elevations is not None:
        return rt.download_f64(out, out=elevations)
    return rt.download_f64(out)


def make_bool_elevation_mask(height, mask_elevation):
    """mask[h] = height[h] <= mask_elevation (terrain.py:61-72).  Device arithmetic is FP32: float64 numpy
    input (and the level) is rounded to float32 before the comparison, so a vertex within one float32
    ulp of the level may land on the other side than in the reference."""
    if isinstance(height, torch.Tensor):
        return rt.mask_le(height, float(mask_elevation))
    h = rt.upload_f32(height)
    return rt._to_host(rt.mask_le(h, float(mask_elevation))).view(np.bool_)
