"""Device-side plumbing: torch owns the buffers and streams, the C-ABI does the work.

Every function here takes / returns torch CUDA tensors (FP32 state, int32 tables) and
launches hand-written sm_100a kernels through `_lib`.  Nothing in this module computes
on the host; without a GPU or without the built library it raises.
"""
import ctypes as C
import os
import weakref

import numpy as np
import torch

from . import _lib

F32 = torch.float32
EARTH_RADIUS = 6378100.0


def require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("nixis_b200 needs a CUDA device: the hot path has no CPU fallback")
    _lib.load()


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous(), "device tensors must be contiguous CUDA tensors"
    return C.c_void_p(t.data_ptr())


def _dbl(values):
    return (C.c_double * len(values))(*values)


# ---------------------------------------------------------------------------------
# opensimplex tables
def init_perm(seed):
    """opensimplex.py:90-112 (host): returns (perm, pgi) int32[256] numpy arrays."""
    perm = np.zeros(256, np.int32)
    pgi = np.zeros(256, np.int32)
    seed = (int(seed) + 2 ** 63) % 2 ** 64 - 2 ** 63        # numba int64 wrap-around
    _lib.call("nxb_init_perm", C.c_int64(seed), perm.ctypes.data_as(C.c_void_p), pgi.ctypes.data_as(C.c_void_p))
    return perm, pgi


class _Tables:
    def __init__(self, handle):
        self.handle = handle

    def __del__(self):
        try:
            if self.handle:
                _lib.load().nxb_tables_destroy(C.c_void_p(self.handle))
        except Exception:
            pass


_tables_cache = {}


def tables_for(perm, pgi):
    """Device lookup tables for a (perm, pgi) pair; cached per device."""
    require_cuda()
    perm = np.ascontiguousarray(perm, dtype=np.int32)
    pgi = np.ascontiguousarray(pgi if pgi is not None else (perm % 24) * 3, dtype=np.int32)
    key = (torch.cuda.current_device(), perm.tobytes(), pgi.tobytes())
    t = _tables_cache.get(key)
    if t is None:
        h = C.c_void_p()
        _lib.call("nxb_tables_create", perm.ctypes.data_as(C.c_void_p), pgi.ctypes.data_as(C.c_void_p), C.byref(h))
        t = _Tables(h.value)
        if len(_tables_cache) > 64:
            _tables_cache.clear()
        _tables_cache[key] = t
    return t


def octave_schedule(n_octaves, n_init_roughness, n_init_strength, n_roughness, n_persistence):
    """terrain.py:36-45: per-octave frequency / amplitude, advanced in Python float64."""
    freq, amp = [], []
    f, a = float(n_init_roughness), float(n_init_strength)
    for _ in range(int(n_octaves)):
        freq.append(f)
        amp.append(a)
        f *= n_roughness
        a *= n_persistence
    return freq, amp


# ---------------------------------------------------------------------------------
# kernels on torch tensors
def fbm3(tables, xyz, freq, amp, init=None, out=None, minmax=None):
    """float4 unit-sphere positions (promoted to float64 inside); freq / amp per octave (unit-sphere terms)."""
    n = xyz.shape[0]
    if out is None:
        out = torch.empty(n, dtype=F32, device=xyz.device)
    _lib.call("nxb_fbm3_f32", C.c_void_p(tables.handle), _ptr(xyz), n, len(freq), _dbl(freq), _dbl(amp),
              _ptr(init), _ptr(out), _ptr(minmax), _stream())
    return out


def fbm3_pos64(tables, verts64, scale, nr, amp, init=None, out=None, minmax=None):
    """The throughput fBm on the reference's own float64 vertices (device float64 [n,3], times `scale`):
    lattice coordinate (verts * scale) * nr[o] and the candidate selection in float64 exactly as the
    reference evaluates them, FP32 contributions; amp[o] in output units."""
    n = verts64.shape[0]
    assert verts64.dtype == torch.float64 and verts64.shape[1] == 3
    if out is None:
        out = torch.empty(n, dtype=F32, device=verts64.device)
    _lib.call("nxb_fbm3_pos64_f32", C.c_void_p(tables.handle), _ptr(verts64), C.c_double(scale), n, len(nr),
              _dbl(nr), _dbl(amp), _ptr(init), _ptr(out), _ptr(minmax), _stream())
    return out


def fbm3_exact(tables, verts64, n_octaves, n_init_roughness, n_init_strength, n_roughness, n_persistence,
               world_radius, scale=1.0, init=None):
    """Reference-exact float64 fBm (nxb_fbm3_f64): verts64 float64 [n,3] CUDA, returns float64 [n] CUDA."""
    n = verts64.shape[0]
    nr, ns = [], []
    f, a = n_init_roughness, n_init_strength          # Python floats, advanced like terrain.py:44-45
    for _ in range(int(n_octaves)):
        nr.append(f / world_radius)
        ns.append(a / world_radius)
        f *= n_roughness
        a *= n_persistence
    out = torch.empty(n, dtype=torch.float64, device=verts64.device)
    _lib.call("nxb_fbm3_f64", C.c_void_p(tables.handle), _ptr(verts64), n, len(nr), _dbl(nr), _dbl(ns),
              C.c_double(world_radius), C.c_double(scale), _ptr(init), _ptr(out), _stream())
    return out


def fbm4(tables, xyz, freq, amp, w, init=None, out=None, minmax=None):
    n = xyz.shape[0]
    if out is None:
        out = torch.empty(n, dtype=F32, device=xyz.device)
    _lib.call("nxb_fbm4_f32", C.c_void_p(tables.handle), _ptr(xyz), n, len(freq), _dbl(freq), _dbl(amp), _dbl(w),
              _ptr(init), _ptr(out), _ptr(minmax), _stream())
    return out


def noise_array(tables, coords):
    """coords: list of 2, 3 or 4 float32 CUDA vectors."""
    n = coords[0].numel()
    out = torch.empty(n, dtype=F32, device=coords[0].device)
    name = {2: "nxb_noise2_f32", 3: "nxb_noise3_f32", 4: "nxb_noise4_f32"}[len(coords)]
    _lib.call(name, C.c_void_p(tables.handle), *[_ptr(c) for c in coords], n, _ptr(out), _stream())
    return out


def new_minmax(device):
    mm = torch.empty(2, dtype=F32, device=device)
    _lib.call("nxb_minmax_reset", _ptr(mm), _stream())
    return mm


def minmax(x, mm=None):
    """Device min/max of x as a float32[2] CUDA tensor (merged into mm if given)."""
    if mm is None:
        mm = new_minmax(x.device)
    _lib.call("nxb_minmax_f32", _ptr(x), x.numel(), _ptr(mm), _stream())
    return mm


def rescale(x, x_min, x_max, lower, upper, mid=None, mode=0, out=None):
    if out is None:
        out = torch.empty_like(x)
    _lib.call("nxb_rescale_f32", _ptr(x), x.numel(), C.c_float(x_min), C.c_float(x_max), C.c_float(lower),
              C.c_float(upper), int(mid is not None), C.c_float(0.0 if mid is None else mid), int(mode),
              _ptr(out), _stream())
    return out


def mask_le(h, level, out=None):
    if out is None:
        out = torch.empty(h.numel(), dtype=torch.uint8, device=h.device)
    _lib.call("nxb_mask_le_f32", _ptr(h), h.numel(), C.c_float(level), _ptr(out), _stream())
    return out


def power_summary(x, mask, sel_mode):
    """(has, F, U, M) of the selected elements as a float32[4] CUDA tensor (include/nixis_b200.h)."""
    s = torch.empty(4, dtype=F32, device=x.device)
    _lib.call("nxb_power_summary_f32", _ptr(x), _ptr(mask), x.numel(), int(sel_mode), _ptr(s), _stream())
    return s


def combine_power_summaries(summaries):
    """Ordered combine of per-shard (has, F, U, M) tuples (host floats)."""
    tot = None
    for has, F, U, M in summaries:
        if not has:
            continue
        if tot is None:
            tot = [1.0, F, U, M]
            continue
        tot[2] = max(tot[2], U, F if F >= tot[3] else float("-inf"))
        tot[3] = min(tot[3], M)
    return tot if tot is not None else [0.0, 0.0, float("-inf"), float("inf")]


def power_bounds(summary, x_min, x_max):
    """util.py:198-214 result (mask_lower, mask_upper) from the ordered summary."""
    has, F, U, M = summary
    if not has:
        return x_max, x_min
    lower = min(x_max, M)
    upper = max(x_min, U, F if F >= x_max else float("-inf"))
    return lower, upper


def power_apply(x, mask, sel_mode, lo, hi, power, shift=0.0, out=None):
    if out is None:
        out = torch.empty_like(x)
    _lib.call("nxb_power_apply_f32", _ptr(x), _ptr(mask), x.numel(), int(sel_mode), C.c_float(lo), C.c_float(hi),
              C.c_float(power), C.c_float(shift), _ptr(out), _stream())
    return out


def mesh_points(k, v_begin=0, v_end=None, f32=True, f64=False, device=None):
    V = 10 * k * k + 2
    v_end = V if v_end is None else v_end
    device = device or torch.device("cuda", torch.cuda.current_device())
    n = v_end - v_begin
    p32 = torch.empty((n, 4), dtype=F32, device=device) if f32 else None
    p64 = torch.empty((n, 3), dtype=torch.float64, device=device) if f64 else None
    _lib.call("nxb_mesh_icosa_points", int(k), v_begin, v_end, _ptr(p32), _ptr(p64), _stream())
    return p32, p64


def mesh_cells(k, t_begin=0, t_end=None, device=None):
    T = 20 * k * k
    t_end = T if t_end is None else t_end
    device = device or torch.device("cuda", torch.cuda.current_device())
    cells = torch.empty((t_end - t_begin, 3), dtype=torch.int32, device=device)
    _lib.call("nxb_mesh_icosa_cells", int(k), t_begin, t_end, _ptr(cells), _stream())
    return cells


def xyz_from_f64(verts64, scale):
    """float64 [n,3] CUDA tensor * scale -> float32 [n,4]."""
    n = verts64.shape[0]
    out = torch.empty((n, 4), dtype=F32, device=verts64.device)
    _lib.call("nxb_xyz_f64_to_f32", _ptr(verts64), n, C.c_double(scale), _ptr(out), _stream())
    return out


def adj_build(cells, V):
    T = cells.shape[0]
    adj = torch.empty((V, 6), dtype=torch.int32, device=cells.device)
    ws_bytes = _lib.load().nxb_adj_build_workspace(V)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=cells.device)
    _lib.call("nxb_adj_build", _ptr(cells), T, V, _ptr(adj), _ptr(ws), _stream())
    return adj


def icosa_adj_rows(k, v_begin, v_end, device=None, with_unsorted=False):
    """Sorted neighbour rows [v_begin, v_end) of the closed-form icosphere (global ids), built from the
    triangle generator with 52 B of scratch per ROW -- no cell array, no whole-mesh table."""
    device = device or torch.device("cuda", torch.cuda.current_device())
    n = v_end - v_begin
    adj = torch.empty((n, 6), dtype=torch.int32, device=device)
    uns = torch.empty((n, 6), dtype=torch.int32, device=device) if with_unsorted else None
    ws = torch.empty(max(16, _lib.load().nxb_mesh_icosa_adj_rows_workspace(n)), dtype=torch.uint8, device=device)
    _lib.call("nxb_mesh_icosa_adj_rows", int(k), int(v_begin), int(v_end), _ptr(adj), _ptr(uns), _ptr(ws), _stream())
    return (adj, uns) if with_unsorted else adj


def adj_sort(adj):
    out = torch.empty_like(adj)
    _lib.call("nxb_adj_sort", _ptr(adj), _ptr(out), adj.shape[0], _stream())
    return out


ERO_TILE = 256
ERO_DESC_BYTES, ERO_DESC_WORDS = 256, 64
ERO_DW_SEND = 28           # csrc/nxb_erosion_plan.cuh: word index of (send0, send1) in a tile descriptor


def round_up(n, m):
    return (n + m - 1) // m * m


def edge_lengths(nodes64, adj, n_own=None):
    """float64 [.,3] CUDA positions + int32 [n_own,6] table -> float32 [round_up(n_own,256)*6] edge lengths."""
    n_own = adj.shape[0] if n_own is None else n_own
    dist = torch.zeros(round_up(n_own, ERO_TILE) * 6, dtype=F32, device=adj.device)
    _lib.call("nxb_edge_lengths_f64", _ptr(nodes64), _ptr(adj), n_own, _ptr(dist), _stream())
    return dist


def icosa_edge_lengths(k, adj_rows, v_begin, v_end, radius):
    """Edge lengths of the closed-form icosphere for rows [v_begin, v_end) (global ids in adj_rows)."""
    n = v_end - v_begin
    dist = torch.zeros(round_up(n, ERO_TILE) * 6, dtype=F32, device=adj_rows.device)
    _lib.call("nxb_mesh_icosa_edge_lengths", int(k), _ptr(adj_rows), v_begin, v_end, C.c_double(radius), _ptr(dist), _stream())
    return dist


class ErosionPlan:
    """Tile plan (halo segments, 16-bit tile-local adjacency, implicit-adjacency constants) of an
    int32 [n_own,6] neighbour table whose indices address buffers of `capacity` elements."""

    def __init__(self, adj, capacity=None):
        self.adj = adj
        self.n_own = adj.shape[0]
        self.capacity = round_up(self.n_own, ERO_TILE) if capacity is None else int(capacity)
        nbytes = _lib.load().nxb_erode_plan_bytes(self.n_own)
        self.mem = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=adj.device)
        stats = (C.c_int32 * 8)()
        _lib.call("nxb_erode_plan_build", _ptr(adj), self.n_own, self.capacity, _ptr(self.mem), stats, _stream())
        self.n_tiles, self.n_irregular, self.max_halo, self.n_affine, self.n_affine3, self.n_two = (stats[i] for i in range(6))
        # descriptor size as the LIBRARY lays it out (plan = descriptors, then 16-bit codes of whole tiles)
        self.desc_bytes = (nbytes // self.n_tiles - ERO_TILE * 6 * 2) if self.n_tiles else ERO_DESC_BYTES
        self._dist3 = {}

    def descriptors(self):
        """int32 [n_tiles, 64] view of the tile descriptors (nxb_erosion_plan.cuh EroTileDesc)."""
        return self.mem[: self.n_tiles * self.desc_bytes].view(torch.int32).view(self.n_tiles, self.desc_bytes // 4)

    def dist3_for(self, dist):
        """One-length-per-edge table derived from the full [n,6] table `dist` (built once per table,
        cached).  The sweep reads it on the tiles the plan marks (affine tiles with constant owner
        rows: 36 B per vertex-sweep instead of 48).  NXB_ERO_DIST3=0 streams the full table everywhere."""
        if os.environ.get("NXB_ERO_DIST3", "1") != "1" or self.n_affine3 == 0:
            return None
        key = dist.data_ptr()
        if key not in self._dist3:
            d3 = torch.empty(_lib.load().nxb_erode_dist3_floats(self.n_own), dtype=F32, device=dist.device)
            _lib.call("nxb_erode_dist3_build", _ptr(self.adj), _ptr(dist), self.n_own, _ptr(d3), _stream())
            self._dist3 = {key: d3}
        return self._dist3[key]


def erode3_step(plan, dist, src, dst, rain):
    """src/dst: (hw, s) pairs: hw float32 [capacity, 2] = {height, water} per vertex, s float32 [capacity]."""
    d3 = plan.dist3_for(dist)
    _lib.call("nxb_erode3_plan_step_f32", _ptr(plan.mem), _ptr(plan.adj), _ptr(dist), None if d3 is None else _ptr(d3),
              _ptr(src[0]), _ptr(src[1]), _ptr(dst[0]), _ptr(dst[1]), plan.n_own, C.c_float(rain), _stream())


def erode3_run(plan, dist, a, b, rain, n_sweeps):
    """erosion.py:180-184 loop in C: n_sweeps sweeps ping-ponging between the (hw, s) sets a and b
    (sweep 0 reads a).  Returns the set that holds the result."""
    d3 = plan.dist3_for(dist)
    _lib.call("nxb_erode3_run_f32", _ptr(plan.mem), _ptr(plan.adj), _ptr(dist), None if d3 is None else _ptr(d3),
              _ptr(a[0]), _ptr(a[1]), _ptr(b[0]), _ptr(b[1]),
              plan.n_own, C.c_float(rain), int(n_sweeps), _stream(), launches=int(n_sweeps))
    return a if n_sweeps % 2 == 0 else b


def erode1_step(adj, h_in, h_out, v_begin, v_end):
    _lib.call("nxb_erode1_step_f32", _ptr(adj), _ptr(h_in), _ptr(h_out), v_begin, v_end, _stream())


def gather(src, idx, out=None):
    if out is None:
        out = torch.empty(idx.numel(), dtype=F32, device=src.device)
    _lib.call("nxb_gather_f32", _ptr(src), _ptr(idx), idx.numel(), _ptr(out), _stream())
    return out


def scatter(src, idx, dst):
    _lib.call("nxb_scatter_f32", _ptr(src), _ptr(idx), idx.numel(), _ptr(dst), _stream())


def ll_grid(width, height, radius):
    """util.py:290-308: float64 [height, width, 3] CUDA tensor of pixel positions."""
    require_cuda()
    out = torch.empty((height, width, 3), dtype=torch.float64, device="cuda")
    _lib.call("nxb_ll_grid_f64", int(width), int(height), C.c_double(radius), _ptr(out), _stream())
    return out


def ico_nearest3(k, radius, query64):
    """query64: float64 [..., 3] CUDA -> (dists float64 [..., 3] ascending, ids int64 [..., 3])."""
    n = query64.numel() // 3
    dists = torch.empty(query64.shape, dtype=torch.float64, device=query64.device)
    ids = torch.empty(query64.shape, dtype=torch.int64, device=query64.device)
    _lib.call("nxb_ico_nearest3_f64", int(k), C.c_double(radius), _ptr(query64), n, _ptr(dists), _ptr(ids), _stream())
    return dists, ids


def idw_gray(dists, ids, colors64):
    """util.py:343-367: int32 [...] blend of colors64 (float64 [V]) at the 3 nearest vertices."""
    n = dists.numel() // 3
    out = torch.empty(dists.shape[:-1], dtype=torch.int32, device=dists.device)
    _lib.call("nxb_idw_gray_f64", _ptr(dists), _ptr(ids), _ptr(colors64), n, _ptr(out), _stream())
    return out


def idw_map(dists, ids, field, x_min, x_max, lower, upper, add=0.0, quantize_u16=False, out_bits=8):
    """One export map (nixis.py:349,386-389,417 -> util.py:393-429) from a device-resident field
    (float32 heights or uint8 mask): rescale + optional uint16 truncation + 3-vertex blend, fused."""
    assert field.dtype in (torch.float32, torch.uint8) and field.is_contiguous()
    n = dists.numel() // 3
    out = torch.empty(dists.shape[:-1], dtype=torch.uint8 if out_bits == 8 else torch.uint16, device=dists.device)
    _lib.call("nxb_idw_map", _ptr(dists), _ptr(ids), _ptr(field), 0 if field.dtype == torch.float32 else 1, n,
              C.c_double(x_min), C.c_double(x_max), C.c_double(lower), C.c_double(upper), C.c_double(add),
              int(bool(quantize_u16)), int(out_bits), _ptr(out), _stream())
    return out


def to_f64(x32):
    out = torch.empty(x32.shape, dtype=torch.float64, device=x32.device)
    _lib.call("nxb_f32_to_f64", _ptr(x32), x32.numel(), _ptr(out), _stream())
    return out


def to_f32(x64):
    out = torch.empty(x64.shape, dtype=F32, device=x64.device)
    _lib.call("nxb_f64_to_f32", _ptr(x64), x64.numel(), _ptr(out), _stream())
    return out


# ---------------------------------------------------------------------------------
# host <-> device for the numpy-facing API (the reference's arrays are float64 numpy)
def upload(a, dtype=None):
    require_cuda()
    t = torch.from_numpy(np.ascontiguousarray(a))
    t = t.cuda(non_blocking=True)
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t


def upload_f32(a):
    """float64 numpy -> float32 CUDA vector (conversion on the device)."""
    a = np.ascontiguousarray(a)
    if a.dtype == np.float64:
        return to_f32(upload(a))
    return upload(a.astype(np.float32, copy=False))


def _to_host(t):
    """CUDA tensor -> numpy array backed by PINNED host memory (torch's caching host allocator keeps
    the blocks, so steady-state calls pay no cudaHostAlloc).  Full-speed D2H now, and a full-speed H2D
    if the caller passes the array back in (nixis.py feeds every result into the next call)."""
    host = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    host.copy_(t, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    return host.numpy()


def download_f64(x32, out=None):
    """float32 CUDA vector -> float64 numpy (conversion on the device, like the reference's dtype).
    With `out` the D2H copy lands directly in the caller's array (pinned or pageable)."""
    t64 = to_f64(x32)
    if out is not None and out.dtype == np.float64 and out.flags.c_contiguous and out.shape == tuple(t64.shape):
        torch.from_numpy(out).copy_(t64)
        return out
    t = _to_host(t64)
    if out is not None:
        out[...] = t
        return out
    return t


def ffma_peak_tflops(iters=2000):
    require_cuda()
    v = C.c_double()
    _lib.call("nxb_ffma_peak", int(iters), C.byref(v))
    return v.value
