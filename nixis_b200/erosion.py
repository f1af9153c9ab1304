"""Drop-in for the live half of the reference's `erosion` module (erosion.py) on B200.

Kept: `erode_terrain3` / `erosion_iteration3` (the variant nixis.py:410 calls) and
`erode_terrain1` / `erosion_iteration1` (the numerically stable variant).  Variants 2, 4, 5, 6
are dead, buggy or non-deterministic experiments in the reference and are
not provided.

State on the device is FP32 and ping-pongs between two buffer sets; there is no copy-back pass
(erosion.py:199-201, 274-277) and `water += rain` (erosion.py:182-183) is fused into the sweep.
numpy arguments are updated in place exactly where the reference updates them.
"""
import numpy as np
import torch

from . import runtime as rt
from .util import DeviceMesh

RAIN_AMOUNT = 0.3 / 320         # erosion.py:182


def _positions(nodes):
    """-> (float32 [V,4] positions / scale, scale).  Distances are scale * |delta|."""
    if isinstance(nodes, DeviceMesh):
        return nodes.xyz, nodes.radius
    if isinstance(nodes, torch.Tensor):
        return nodes, 1.0
    v = np.ascontiguousarray(nodes, dtype=np.float64)
    scale = float(np.abs(v[: min(len(v), 4096)]).max()) or 1.0
    return rt.xyz_from_f64(rt.upload(v), 1.0 / scale), scale


def _neighbors(neighbors):
    if isinstance(neighbors, torch.Tensor):
        return neighbors
    return rt.upload(np.ascontiguousarray(neighbors, dtype=np.int32))


class Erosion3State:
    """Device-resident state of erode_terrain3: (h, water, sediment) x ping-pong."""

    def __init__(self, xyz, scale, adj, heights32):
        self.xyz, self.scale, self.adj = xyz, float(scale), adj
        n = heights32.numel()
        self.cur = (heights32, torch.zeros(n, dtype=rt.F32, device=heights32.device),
                    torch.zeros(n, dtype=rt.F32, device=heights32.device))
        self.nxt = tuple(torch.empty_like(t) for t in self.cur)
        self.iterations = 0

    def step(self, rain=RAIN_AMOUNT):
        n = self.cur[0].numel()
        rt.erode3_step(self.xyz, self.adj, self.cur, self.nxt, 0, n, rain, self.scale)
        self.cur, self.nxt = self.nxt, self.cur
        self.iterations += 1

    def run(self, num_iter, rain=RAIN_AMOUNT):
        for _ in range(num_iter):
            self.step(rain)

    @property
    def heights(self):
        return self.cur[0]

    @property
    def water(self):
        """Water AFTER the last sweep (the reference's `water` array at that point)."""
        return self.cur[1]

    @property
    def sediment(self):
        return self.cur[2]


def erode_terrain3(nodes, neighbors, heights, num_iter=1, snapshot=False, verbose=True, return_state=False):
    """erosion.py:172-192.  `heights` (numpy float64) is eroded IN PLACE and None is returned;
    with CUDA tensors the new height tensor is returned.  water / sediment start at zero and are
    discarded unless return_state=True.  `snapshot` (per-iteration PNG export) is outside the hot
    path and not supported."""
    if snapshot:
        raise NotImplementedError("erosion snapshots use the image-export path, which is out of scope")
    if verbose:
        print("Starting terrain erosion...")
    if num_iter <= 0:
        num_iter = 1
    xyz, scale = _positions(nodes)
    adj = _neighbors(neighbors)
    dev_io = isinstance(heights, torch.Tensor)
    h32 = heights.clone() if dev_io else rt.upload_f32(heights)
    st = Erosion3State(xyz, scale, adj, h32)
    for i in range(num_iter):
        if verbose:
            print("  Erosion pass:", i + 1, "of", num_iter)
        st.step()
    if dev_io:
        return st if return_state else st.heights
    rt.download_f64(st.heights, out=heights)
    if return_state:
        return rt.download_f64(st.water), rt.download_f64(st.sediment)
    return None


def erosion_iteration3(verts, neighbors, r_buff, wat, sed):
    """One sweep (erosion.py:197-279): r_buff, wat, sed (numpy float64) are updated in place.
    `wat` must already contain this iteration's rain, as in erode_terrain3."""
    xyz, scale = _positions(verts)
    adj = _neighbors(neighbors)
    src = (rt.upload_f32(r_buff), rt.upload_f32(wat), rt.upload_f32(sed))
    dst = tuple(torch.empty_like(t) for t in src)
    rt.erode3_step(xyz, adj, src, dst, 0, src[0].numel(), 0.0, scale)
    rt.download_f64(dst[0], out=r_buff)
    rt.download_f64(dst[1], out=wat)
    rt.download_f64(dst[2], out=sed)


def erosion_iteration1(neighbors, r_buff, w_buff):
    """erosion.py:76-99: w = r + 0.0005 * (#higher - #lower neighbours); returns w_buff."""
    adj = _neighbors(neighbors)
    if isinstance(r_buff, torch.Tensor):
        rt.erode1_step(adj, r_buff, w_buff, 0, r_buff.numel())
        return w_buff
    src = rt.upload_f32(r_buff)
    dst = torch.empty_like(src)
    rt.erode1_step(adj, src, dst, 0, src.numel())
    return rt.download_f64(dst, out=w_buff)


def erode_terrain1(nodes, neighbors, heights, num_iter=1, snapshot=None, verbose=True):
    """erosion.py:42-73: heights updated in place after every pass and returned."""
    if verbose:
        print("Starting terrain erosion...")
    if num_iter <= 0:
        num_iter = 1
    adj = _neighbors(neighbors)
    dev_io = isinstance(heights, torch.Tensor)
    a = heights if dev_io else rt.upload_f32(heights)
    b = torch.empty_like(a)
    for i in range(num_iter):
        if verbose:
            print("  Erosion pass:", i + 1, "of", num_iter)
        rt.erode1_step(adj, a, b, 0, a.numel())
        a, b = b, a
    if dev_io:
        if a.data_ptr() != heights.data_ptr():
            heights.copy_(a)
        return heights
    return rt.download_f64(a, out=heights)
