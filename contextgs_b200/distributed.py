"""Multi-GPU partitioning of the hot path (the reference is single-process, so this is
new design).  One process per GPU, `torch.distributed` (NCCL over NVLink on the GPUs, gloo in the CPU
tests) only for the plumbing:

  * rendering        -- replicas of the model, camera batches sharded round-robin, NO collective;
  * entropy scoring  -- the 3-level coding plan is split by dependency root (the coarsest-level
                        representative a row's context chain ends in), so every shard's rows depend
                        only on rows of the same shard; one scalar all-reduce of the bit sums;
  * training         -- one all-reduce (sum) of a flat fp32 gradient bucket per step.
"""
from types import SimpleNamespace

import torch


# ----------------------------------------------------------------------------- cameras

def shard_cameras(n_cameras, rank, world):
    """Indices of the cameras rank `rank` renders: rank, rank + world, ... (weak scaling: with
    n_cameras = world * k every rank renders k frames)."""
    return list(range(rank, n_cameras, world))


# ----------------------------------------------------------------------------- entropy plan

def plan_roots(plan):
    """For every coded row of every level, the index (0 .. n_coarsest-1) of the coarsest-level row its
    context chain ends in.  Level order in `plan.levels` is coarse -> fine (level 2, 1, 0):
      level-2 rows are their own roots; a finer row inherits the root of the row that CODED its
      context source anchor `ctx_src` (scene/gaussian_model.py:1711-1724 gathers the context from
      already coded anchors only)."""
    dev = plan.levels[0].orig.device
    root_of_anchor = torch.full((plan.N,), -1, dtype=torch.long, device=dev)
    roots = []
    for li, lv in enumerate(plan.levels):
        if li == 0:
            r = torch.arange(lv.n, device=dev)
        else:
            r = root_of_anchor[lv.ctx_src.long()]
            if lv.n and int(r.min()) < 0:
                raise RuntimeError("level plan: a context source was not coded by a coarser level")
        root_of_anchor[lv.orig.long()] = r
        roots.append(r)
    return roots


def shard_level_plan(plan, rank, world):
    """Rows of `plan` owned by `rank`: contiguous blocks of coarsest-level roots, with every finer row
    following its root.  The shards partition the rows of every level and are closed under the context
    dependency, so each rank can run its three levels without any exchange."""
    roots = plan_roots(plan)
    n_root = plan.levels[0].n
    per = (n_root + world - 1) // world if n_root else 0
    lo, hi = rank * per, min(n_root, (rank + 1) * per)
    levels = []
    for lv, r in zip(plan.levels, roots):
        sel = torch.nonzero((r >= lo) & (r < hi))[:, 0]
        levels.append(SimpleNamespace(
            level=lv.level, orig=lv.orig[sel].contiguous(),
            ctx_src=None if lv.ctx_src is None else lv.ctx_src[sel].contiguous(),
            level_anchor=None if lv.level_anchor is None else lv.level_anchor[sel].contiguous(),
            n=int(sel.numel()), rows=sel))
    return SimpleNamespace(N=plan.N, levels=levels, rank=rank, world=world)


# ----------------------------------------------------------------------------- collectives

def all_reduce_sums(t, group=None):
    """In-place SUM all-reduce of a small tensor of bit sums (no-op without a process group)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


class GradientBucket:
    """One flat fp32 bucket for the per-step gradient all-reduce (111 floats per anchor + ~84 k MLP /
    codec weights; 888 MB at 2 M anchors).  `views[i]` aliases the slice of parameter i, so kernels
    can write gradients straight into the bucket and a single collective moves everything."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device if self.params else "cpu"
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        self.views, o = [], 0
        for p in self.params:
            self.views.append(self.flat[o:o + p.numel()].view(p.shape))
            o += p.numel()

    def attach(self):
        """Make every parameter's .grad a view into the bucket."""
        for p, v in zip(self.params, self.views):
            p.grad = v
        return self

    def zero(self):
        self.flat.zero_()

    def all_reduce(self, group=None, average=True):
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()):
            return self.flat
        world = dist.get_world_size(group)
        if world > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
            if average:
                self.flat.div_(world)
        return self.flat


def adjust_anchor_data_parallel(pc, group=None, **kw):
    """`adjust_anchor` for data-parallel training: every rank has accumulated `training_statis` over
    its own cameras, so the four accumulators are SUM all-reduced first; the random thinning of anchor_growing
    (scene/gaussian_model.py:769) is drawn on rank 0 and broadcast.  All ranks then grow and prune identically
    (the device path is deterministic: sorted cells, maxima), so parameters and Adam state stay replicated
    without any further exchange."""
    import torch.distributed as dist
    dev = pc._anchor.device
    n_slots = pc._anchor.shape[0] * pc.n_offsets
    rand = [torch.rand(n_slots, device=dev) for _ in range(pc.update_depth)]
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        for name in ("opacity_accum", "anchor_demon", "offset_gradient_accum", "offset_denom"):
            dist.all_reduce(getattr(pc, name), op=dist.ReduceOp.SUM, group=group)
        for r in rand:
            dist.broadcast(r, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    pc.adjust_anchor(rand=rand, **kw)
