"""Shard context of the reference-named numpy API (one process per GPU).

The reference is a single-process script; its functions see whole-planet arrays.  With a shard
context set, every rank calls the SAME functions (`terrain.sample_octaves`, `util.rescale`,
`util.power_rescale`, `terrain.make_bool_elevation_mask`, `erosion.erode_terrain3`) on its own
contiguous slice of the vertex arrays -- `points[begin:end]`, `neighbors[begin:end]` (rows keep GLOBAL
vertex ids), `height[begin:end]` -- and gets the slice of the whole-planet result:
  * per-vertex functions need nothing;
  * `rescale` / `power_rescale` all-reduce their min / max and combine the ordered power summaries in
    rank order (pipeline.Collective), so the scalars are the whole planet's;
  * `erode_terrain3` plans its halo from the rows it was given (partition.build_rank_plan_local),
    fetches the halo vertices' positions once for the edge lengths, and runs the sharded sweep loop
    with the fused NVLink exchange (multigpu.ShardedErosion).
The one call a sharded driver needs beyond the reference's own is `amin_amax` (nixis.py:337-338 takes
np.amin / np.amax of the whole array).
"""
import numpy as np

_ctx = None


class ShardContext:
    def __init__(self, ranges, group=None):
        import torch.distributed as dist
        from .pipeline import Collective
        assert dist.is_initialized(), "set_shard needs an initialised torch.distributed process group"
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.ranges = [tuple(int(v) for v in r) for r in ranges]
        assert len(self.ranges) == self.world
        self.begin, self.end = self.ranges[self.rank]
        self.coll = Collective(group, distributed=True)


def set_shard(ranges, group=None):
    """ranges: [(begin, end)] per rank, contiguous, tile-aligned (partition.vertex_ranges)."""
    global _ctx
    _ctx = ShardContext(ranges, group)
    return _ctx


def clear_shard():
    global _ctx
    _ctx = None


def current():
    return _ctx


def collective():
    """The scalar exchanges of the active shard context (identity without one)."""
    from .pipeline import Collective
    return _ctx.coll if _ctx is not None else Collective()


def amin_amax(x):
    """Whole-planet (np.amin(x), np.amax(x)) of a sharded float array (nixis.py:337-338)."""
    import torch
    from . import runtime as rt
    xd = x if isinstance(x, torch.Tensor) else rt.upload_f32(x)
    lo, hi = collective().minmax(rt.minmax(xd))
    return (lo, hi) if isinstance(x, torch.Tensor) else (np.float64(lo), np.float64(hi))
