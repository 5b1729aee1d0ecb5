"""CPU tests of the multi-GPU host logic (contextgs_b200/distributed.py): gloo, world_size 2, plus
single-process "fake world" checks of the shard structure."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from contextgs_b200 import synthetic
from contextgs_b200.context_model import build_level_plan, find_divide_scale
from contextgs_b200.distributed import GradientBucket, all_reduce_sums, plan_roots, shard_cameras, shard_level_plan
from contextgs_b200.gaussian_model import GaussianModel


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _cpu_model(N=4000, seed=3):
    scene = synthetic.make_scene("chair", N, seed=seed)
    torch.manual_seed(6)
    m = GaussianModel.from_tensors(scene, device="cpu")   # parameter container only: no kernels are called
    anchor = scene["anchor"]
    m.level_scale = find_divide_scale(m, anchor, m.target_ratio, m.level_num)
    return m, anchor


def test_shard_cameras_partition():
    for world in (1, 2, 4, 8):
        got = sorted(i for r in range(world) for i in shard_cameras(16, r, world))
        assert got == list(range(16))
        assert all(len(shard_cameras(16, r, world)) == 16 // world for r in range(world))


def test_level_plan_shards_partition_rows_and_are_dependency_closed():
    m, anchor = _cpu_model()
    plan = build_level_plan(m, anchor, None)
    assert sum(lv.n for lv in plan.levels) == plan.N
    roots = plan_roots(plan)
    for world in (2, 4, 7):
        shards = [shard_level_plan(plan, r, world) for r in range(world)]
        for li, lv in enumerate(plan.levels):
            rows = torch.cat([s.levels[li].rows for s in shards])
            assert sorted(rows.tolist()) == list(range(lv.n))              # partition of the level's rows
        for s in shards:
            coded = torch.zeros(plan.N, dtype=torch.bool)
            for li, lv in enumerate(s.levels):
                if lv.ctx_src is not None and lv.n:
                    assert bool(coded[lv.ctx_src.long()].all())            # context is coded by THIS shard, earlier
                coded[lv.orig.long()] = True
        sizes = [sum(lv.n for lv in s.levels) for s in shards]
        assert max(sizes) < 2.0 * plan.N / world                           # reasonably balanced
    assert int(roots[0].max()) == plan.levels[0].n - 1


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        m, anchor = _cpu_model()
        plan = build_level_plan(m, anchor, None)
        shard = shard_level_plan(plan, rank, world)
        # every rank "scores" its rows with a deterministic per-anchor weight; the all-reduced sums must
        # equal the whole-scene totals
        w = torch.arange(plan.N, dtype=torch.float64) * 0.5 + 1.0
        sums = torch.zeros(16, dtype=torch.float64)
        for li, lv in enumerate(shard.levels):
            sums[4 * li] = w[lv.orig.long()].sum()
            sums[4 * li + 3] = lv.n
        all_reduce_sums(sums)
        total = [float(w[lv.orig.long()].sum()) for lv in plan.levels]
        ok = all(abs(float(sums[4 * i]) - total[i]) < 1e-6 for i in range(3))
        ok = ok and int(sums[3] + sums[7] + sums[11]) == plan.N
        # gradient bucket: grads differ per rank, the averaged bucket is identical on all ranks
        torch.manual_seed(100 + rank)
        params = [torch.nn.Parameter(torch.zeros(5, 3)), torch.nn.Parameter(torch.zeros(7))]
        bucket = GradientBucket(params).attach()
        for p in params:
            p.grad.copy_(torch.randn(p.shape))
        mine = bucket.flat.clone()
        gathered = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(gathered, mine)
        bucket.all_reduce()
        ok = ok and torch.allclose(bucket.flat, sum(gathered) / world)
        ok = ok and params[1].grad.data_ptr() == bucket.views[1].data_ptr()
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_gloo_world2_bit_sums_and_gradient_bucket():
    world = 2
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert all(ret.get(r) for r in range(world)), dict(ret)


def _densify_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from contextgs_b200.distributed import adjust_anchor_data_parallel
        from tests.helpers import load_npz
        from tests.test_growing_cpu import model_from_golden, oracle_grow_cells
        g = load_npz("growing.npz")
        m = model_from_golden(g, 1, "cpu")
        GaussianModel.grow_cells = oracle_grow_cells          # host logic only: the device call is the oracle's
        # each rank saw different cameras: split the fixture's accumulators unevenly between the ranks
        share = 0.25 if rank == 0 else 0.75
        for k in ("opacity_accum", "anchor_demon", "offset_gradient_accum", "offset_denom"):
            setattr(m, k, getattr(m, k) * share)
        torch.manual_seed(rank)                               # different local generators: the draw must come from rank 0
        adjust_anchor_data_parallel(m, check_interval=100, success_threshold=0.8, grad_threshold=0.0002, min_opacity=0.005)
        flat = torch.cat([getattr(m, "_" + k).detach().reshape(-1) for k in ("anchor", "anchor_feat", "scaling", "offset")])
        sizes = torch.tensor([m._anchor.shape[0], m.offset_denom.shape[0]])
        gathered = [torch.zeros_like(sizes) for _ in range(world)]
        dist.all_gather(gathered, sizes)
        ok = all(torch.equal(gathered[0], t) for t in gathered)
        if ok:
            both = [torch.zeros_like(flat) for _ in range(world)]
            dist.all_gather(both, flat)
            ok = all(torch.equal(both[0], t) for t in both)
        ok = ok and m._anchor.shape[0] != g["c1_before_anchor"].shape[0]
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_gloo_world2_adjust_anchor_stays_replicated():
    world = 2
    ret = mp.Manager().dict()
    mp.spawn(_densify_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert all(ret.get(r) for r in range(world)), dict(ret)
