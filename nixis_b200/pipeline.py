"""Device-resident driver of the terrain hot path (nixis.py:247-417 call order) for 1..N GPUs.

    mesh (closed form, on device) -> adjacency (build + ring sort) -> fBm (all octaves fused)
    -> height assembly (nixis.py:332-364) -> erosion sweeps (erosion.py:172-279)

Single GPU: everything stays in HBM, the host only sees a few scalars (min/max, ocean level).
Multi GPU (one process per GPU, torch.distributed): the vertex index range is split into
contiguous shards (see partition.py); fBm and the elementwise passes need no communication,
the min/max / power_rescale statistics need a few scalars all-gathered, and erosion exchanges
the boundary heights / water once per sweep (halo.py).
"""
import numpy as np
import torch

from . import runtime as rt
from .util import DeviceMesh

# nixis.py:312-320
MIN_ALT, MAX_ALT, OCEAN_PERCENT = -4000.0, 8850.0, 55.0
N_INIT_ROUGH, N_INIT_STRENGTH, N_ROUGHNESS, N_PERSISTENCE = 1.5, 0.4, 2.5, 0.5


def find_percent_val(minval, maxval, percent):
    return minval + ((maxval - minval) * percent / 100.0)      # util.py:556-566


class Collective:
    """The handful of scalar exchanges the sharded assembly needs.  world=1: identity."""

    def __init__(self, group=None, distributed=False):
        """distributed=False (default): single-domain, never touches torch.distributed even if a
        process group exists (a rank may run a whole-planet reference next to the sharded run)."""
        import torch.distributed as dist
        self.dist = dist if (distributed and dist.is_available() and dist.is_initialized()) else None
        self.group = group
        self.world = self.dist.get_world_size(group) if self.dist else 1
        self.rank = self.dist.get_rank(group) if self.dist else 0

    def _host_side(self):
        """gloo moves CPU tensors (CPU tests, several ranks sharing one GPU); NCCL device tensors."""
        return self.dist.get_backend(self.group) == "gloo"

    def minmax(self, mm):
        """mm: float32[2] CUDA tensor (local min, max) -> global (min, max) host floats."""
        if self.dist and self.world > 1:
            lo = mm[0:1].clone()
            hi = mm[1:2].clone()
            if self._host_side():
                lo, hi = lo.cpu(), hi.cpu()
            self.dist.all_reduce(lo, op=self.dist.ReduceOp.MIN, group=self.group)
            self.dist.all_reduce(hi, op=self.dist.ReduceOp.MAX, group=self.group)
            return float(lo.item()), float(hi.item())
        lo, hi = mm.tolist()
        return lo, hi

    def gather_summaries(self, s):
        """s: float32[4] CUDA tensor -> list of per-rank (has, F, U, M) host tuples, rank order."""
        if self.dist and self.world > 1:
            if self._host_side():
                s = s.cpu()
            buf = [torch.empty_like(s) for _ in range(self.world)]
            self.dist.all_gather(buf, s, group=self.group)
            return [b.tolist() for b in buf]
        return [s.tolist()]


def assemble_heights(h, coll=None, min_alt=MIN_ALT, max_alt=MAX_ALT, ocean_percent=OCEAN_PERCENT, mm=None):
    """nixis.py:332-364 on a (shard of a) float32 CUDA height vector, in place where possible.
    Returns (height, ocean mask uint8, ocean_level)."""
    coll = coll or Collective()
    lo, hi = coll.minmax(mm if mm is not None else rt.minmax(h))
    h = rt.rescale(h, lo, hi, min_alt, max_alt, out=h)                       # nixis.py:332
    minval, maxval = coll.minmax(rt.minmax(h))                                # :337-338
    ocean_level = find_percent_val(minval, maxval, ocean_percent)             # :343
    ocean = rt.mask_le(h, ocean_level)                                        # :346
    for mode, power, shift in ((1, 0.5, 0.0), (0, 2.0, ocean_level)):         # :352-359
        x_min, x_max = (minval, maxval) if mode == 1 else coll.minmax(rt.minmax(h))
        summ = rt.combine_power_summaries(coll.gather_summaries(rt.power_summary(h, ocean, mode)))
        p_lo, p_hi = rt.power_bounds(summ, x_min, x_max)
        h = rt.power_apply(h, ocean, mode, p_lo, p_hi, power, shift, out=h)
    lo, hi = coll.minmax(rt.minmax(h))
    h = rt.rescale(h, lo, hi, min_alt, max_alt, mid=0.0, out=h)               # :361
    return h, ocean, ocean_level


class TerrainPipeline:
    """One GPU's share of the planet.  world=1 -> the whole planet."""

    def __init__(self, k, seed=0, n_octaves=8, radius=1.0, rank=0, world=1, device=None, noise_dim=3, w_scale=0.5):
        """noise_dim 4: BASELINE configs[4]'s 4-D fBm (builder-defined driver, the reference has none):
        octave o samples noise4d(x f_o, y f_o, z f_o, w_scale f_o)."""
        rt.require_cuda()
        assert noise_dim in (3, 4)
        self.noise_dim, self.w_scale = int(noise_dim), float(w_scale)
        self.k, self.seed, self.n_octaves, self.radius = int(k), seed, int(n_octaves), float(radius)
        self.rank, self.world = rank, world
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self.V = 10 * self.k ** 2 + 2
        self.T = 20 * self.k ** 2
        self.perm, self.pgi = rt.init_perm(seed)
        self.tables = rt.tables_for(self.perm, self.pgi)
        self.freq, self.amp = rt.octave_schedule(n_octaves, N_INIT_ROUGH, N_INIT_STRENGTH, N_ROUGHNESS, N_PERSISTENCE)
        self.mesh = None
        self.adj = None

    # ---- single-GPU stages -------------------------------------------------------
    def build_mesh(self, with_adjacency=True):
        xyz, _ = rt.mesh_points(self.k, device=self.device)
        self.mesh = DeviceMesh(self.k, xyz, self.radius)
        if with_adjacency:
            cells = rt.mesh_cells(self.k, device=self.device)
            unsorted = rt.adj_build(cells, self.V)
            del cells
            self.adj = rt.adj_sort(unsorted)
            del unsorted
            self.mesh.adj = self.adj
        return self.mesh

    def fbm(self, out=None, minmax=None):
        if self.noise_dim == 4:
            return rt.fbm4(self.tables, self.mesh.xyz, self.freq, self.amp, [self.w_scale * f for f in self.freq],
                           out=out, minmax=minmax)
        nr = [f / self.radius for f in self.freq]                     # terrain.py:43 n_freq / world_radius
        return rt.fbm3_pos64(self.tables, self.mesh.points64(), self.radius, nr, self.amp, out=out, minmax=minmax)

    def heights(self):
        mm = rt.new_minmax(self.device)
        h = self.fbm(minmax=mm)
        return assemble_heights(h, mm=mm)

    def erosion_state(self, heights32):
        """Plan + edge lengths are built once per mesh and reused by every state."""
        from .erosion import Erosion3State, _edge_lengths
        if getattr(self, "_plan", None) is None:
            self._plan = rt.ErosionPlan(self.adj)
            self._dist = _edge_lengths(self.mesh, self.adj)
        return Erosion3State(self.mesh, self.adj, heights32, plan=self._plan, dist=self._dist)

    # ---- equirectangular export, everything stays on the device ------------
    def image_query(self, width, height):
        """nixis.py:270-283: per-pixel sphere position and its 3 nearest vertices (dists, ids)."""
        ll = rt.ll_grid(width, height, self.radius)
        return rt.ico_nearest3(self.k, self.radius, ll)

    def export_maps(self, heights32, ocean=None, width=4096, height=2048, eroded=False, query=None,
                    min_alt=MIN_ALT, max_alt=MAX_ALT, ocean_percent=OCEAN_PERCENT):
        """The maps nixis.py puts in `export_list` and sends through build_image_data, as device
        images: before erosion `ocean` (uint8), `height_absolute`, `height_relative` (uint16)
        (nixis.py:349, 386-389); after erosion `height` (uint8, nixis.py:417)."""
        dists, ids = query if query is not None else self.image_query(width, height)
        lo, hi = (float(v) for v in rt.minmax(heights32).tolist())
        maps = {}
        if eroded:
            maps["height"] = rt.idw_map(dists, ids, heights32, lo, hi, 0.0, 255.0, out_bits=8)
            return maps
        if ocean is not None:
            # bool mask -> rescale(mask.astype(float64), 0, 255) (util.py:395-396): min 0, max 1
            m_lo, m_hi = float(ocean.min()), float(ocean.max())
            maps["ocean"] = rt.idw_map(dists, ids, ocean, m_lo, m_hi, 0.0, 255.0, out_bits=8)
        add = 32768 - find_percent_val(min_alt, max_alt, ocean_percent)
        maps["height_absolute"] = rt.idw_map(dists, ids, heights32, lo, hi, min_alt, max_alt, add=add,
                                             quantize_u16=True, out_bits=16)
        maps["height_relative"] = rt.idw_map(dists, ids, heights32, lo, hi, 0.0, 65535.0,
                                             quantize_u16=True, out_bits=16)
        return maps
