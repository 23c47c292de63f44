"""N>1 path on CPU: world_size-2 gloo processes shard a batch, render their share with the oracle standing
in for the CUDA op, gather, and must reproduce the single-process render bit-for-bit."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT
from helpers import rand_faces
from pytorch3d_b200.parallel import ShardPlan, lpt_partition


def test_lpt_partition_balances_and_is_deterministic():
    costs = [100, 5, 60, 40, 30, 30, 10, 80]
    a = lpt_partition(costs, 3)
    assert sorted(i for b in a for i in b) == list(range(8))
    loads = [sum(costs[i] for i in b) for b in a]
    assert max(loads) - min(loads) <= 30
    assert a == lpt_partition(costs, 3)
    assert lpt_partition([1, 1], 4) == [[0], [1], [], []]


def test_shard_plan_rebase_and_local_inputs():
    fv, _, _ = rand_faces(90, 1, seed=1)
    first, num = [0, 10, 40, 45], [10, 30, 5, 45]
    plan = ShardPlan.build(first, num, 2)
    assert sorted(plan.assignment[0] + plan.assignment[1]) == [0, 1, 2, 3]
    for r in range(2):
        loc = plan.local_inputs(fv, r)
        assert loc.face_verts.shape[0] == sum(num[i] for i in plan.assignment[r])
        off = 0
        for j, i in enumerate(plan.assignment[r]):
            assert torch.equal(loc.face_verts[off: off + num[i]], fv[first[i]: first[i] + num[i]])
            assert int(loc.first[j]) == off
            off += num[i]
        p2f = torch.full((len(plan.assignment[r]), 2, 2, 1), -1, dtype=torch.int64)
        for j in range(p2f.shape[0]):
            p2f[j, 0, 0, 0] = int(loc.first[j])  # first local face of each local mesh
        g = plan.rebase(p2f, r)
        for j, i in enumerate(plan.assignment[r]):
            assert int(g[j, 0, 0, 0]) == first[i] and int(g[j, 1, 1, 0]) == -1


def _oracle_raster(fv, first, num, nb, size, blur, K, bs, mf, persp, clip, cull):
    import oracle
    out = oracle.rasterize_meshes(fv.numpy(), first.numpy(), num.numpy(), size, blur, K, persp, clip, cull, nthreads=1)
    return tuple(torch.from_numpy(o) for o in out)


def _worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from pytorch3d_b200.parallel import rasterize_meshes_sharded
        fv, _, _ = rand_faces(330, 1, seed=4)
        first = torch.tensor([0, 100, 130, 250, 250])
        num = torch.tensor([100, 30, 120, 0, 80])
        out = rasterize_meshes_sharded(fv, first, num, (20, 24), 1e-3, 3, raster_fn=_oracle_raster)
        np.savez(os.path.join(tmp, "rank%d.npz" % rank), *[o.numpy() for o in out[:4]])
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_equals_single_process(tmp_path):
    import oracle
    oracle.build()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    fv, _, _ = rand_faces(330, 1, seed=4)
    first = torch.tensor([0, 100, 130, 250, 250])
    num = torch.tensor([100, 30, 120, 0, 80])
    want = oracle.rasterize_meshes(fv.numpy(), first.numpy(), num.numpy(), (20, 24), 1e-3, 3)
    for r in range(2):
        got = np.load(os.path.join(str(tmp_path), "rank%d.npz" % r))
        for i in range(4):
            assert np.array_equal(got["arr_%d" % i], want[i]), "rank %d output %d" % (r, i)
    assert (want[0] >= 0).sum() > 50
