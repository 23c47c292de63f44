"""Packed frame exchange (csrc/peer_exchange.cu, pytorch3d_b200/peer.py): pack -> (peer) memory -> unpack is lossless
and reproduces a single-GPU render of the whole batch bit-for-bit; on a box with >= 2 GPUs the same through CUDA-IPC
peer memory and NCCL, one process per GPU."""
import ctypes
import os
import sys

import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda:0")


def _batch(dev, K, blur, size, counts=(900, 2500, 400, 1600, 3000)):
    from pytorch3d_b200 import _C, synthetic
    m = synthetic.torus_batch_hetero(list(counts), seed=2)
    fv = synthetic.face_verts_of(m).to(dev)
    first, num = m.mesh_to_faces_packed_first_idx(), m.num_faces_per_mesh()
    nb = torch.full((fv.shape[0],), -1, dtype=torch.int64, device=dev)
    nb._b200_all_minus_one = True
    full = _C.rasterize_meshes(fv, first.to(dev), num.to(dev), nb, size, blur, K, 0, 0, False, False, False)
    return fv, first, num, full


@pytest.mark.parametrize("K,blur,size", [(8, 0.0, (64, 64)), (4, 1e-3, (40, 56)), (6, 1e-3, (33, 47)), (16, 1e-2, (32, 32))])
def test_pack_unpack_round_trip_of_a_sharded_batch(built_lib, dev, K, blur, size):
    """Three 'ranks' on one GPU: each shard is rendered in its local packing, packed into its region of an arena,
    and the regions are expanded into the full batch: equal to the single render, including re-based face ids."""
    from pytorch3d_b200 import _C, _lib, parallel
    lib = _lib.load()
    fv, first, num, full = _batch(dev, K, blur, size)
    H, W = size
    world = 3
    plan = parallel.ShardPlan.build(first.tolist(), num.tolist(), world)
    n_layout = plan.max_local
    rb = int(lib.b200r_packed_frames_bytes(n_layout, H, W, K))
    arena = torch.zeros(world * rb + 64, dtype=torch.uint8, device=dev)
    base = (arena.data_ptr() + 15) // 16 * 16
    cursor = torch.zeros(1, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream
    outs = [torch.full((len(num), H, W, K), 7, dtype=torch.int64, device=dev),
            torch.full((len(num), H, W, K), 7.0, device=dev), torch.full((len(num), H, W, K, 3), 7.0, device=dev),
            torch.full((len(num), H, W, K), 7.0, device=dev)]
    for r in range(world):
        loc = plan.local_inputs(fv, r)
        nb = torch.full((loc.face_verts.shape[0],), -1, dtype=torch.int64, device=dev)
        nb._b200_all_minus_one = True
        part = _C.rasterize_meshes(loc.face_verts, loc.first, loc.num, nb, size, blur, K, 0, 0, False, False, False)
        dst = (ctypes.c_void_p * 1)(base + r * rb)
        _lib.check(lib.b200r_fragments_pack_push(part[0].data_ptr(), part[1].data_ptr(), part[2].data_ptr(),
                                                 part[3].data_ptr(), len(loc.mesh_ids), H, W, K, n_layout, dst, 1,
                                                 cursor.data_ptr(), stream))
        used = int(cursor.item())
        assert used == int((part[0] >= 0).sum())
    for r in range(world):
        ids = plan.assignment[r]
        idx = torch.tensor(ids, dtype=torch.int32, device=dev)
        shift = torch.tensor(plan.local_shifts(r), dtype=torch.int64, device=dev)
        _lib.check(lib.b200r_fragments_unpack(base + r * rb, len(ids), H, W, K, n_layout, idx.data_ptr(),
                                              shift.data_ptr(), outs[0].data_ptr(), outs[1].data_ptr(),
                                              outs[2].data_ptr(), outs[3].data_ptr(), stream))
    torch.cuda.synchronize()
    for got, want in zip(outs, full):
        assert torch.equal(got, want)


def test_exchange_class_single_rank(built_lib, dev):
    from pytorch3d_b200 import parallel, peer
    fv, first, num, full = _batch(dev, 8, 0.0, (48, 48))
    plan = parallel.ShardPlan.build(first.tolist(), num.tolist(), 1)
    ex = peer.PackedFrameExchange(plan, 0, (48, 48), 8, device=dev)
    try:
        for _ in range(3):  # both arena halves, reuse
            got = ex.start(full).wait()
            torch.cuda.synchronize()
            for a, b in zip(got, full):
                assert torch.equal(a, b)
    finally:
        ex.close()


# ---------------------------------------------------------------------------------------------- two GPUs

def _two_rank_worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from pytorch3d_b200 import _C, parallel, peer, synthetic
        counts = [900, 2500, 400, 1600, 3000, 700, 1200]
        m = synthetic.torus_batch_hetero(counts, seed=2)
        fv = synthetic.face_verts_of(m).to(dev)
        first, num = m.mesh_to_faces_packed_first_idx(), m.num_faces_per_mesh()
        size, K, blur = (64, 80), 8, 1e-4
        nb = torch.full((fv.shape[0],), -1, dtype=torch.int64, device=dev)
        nb._b200_all_minus_one = True
        want = _C.rasterize_meshes(fv, first.to(dev), num.to(dev), nb, size, blur, K, 0, 0, False, False, False)
        # (1) dense NCCL gather through the public sharded entry point, with gradients for the local meshes
        fvg = fv.clone().requires_grad_(True)
        out = parallel.rasterize_meshes_sharded(fvg, first, num, size, blur, K)
        for a, b in zip(out[:4], want):
            assert torch.equal(a.detach(), b), "dense gather differs from the single-GPU render"
        plan = out[4]
        g = torch.Generator(device=dev).manual_seed(231)
        gz = torch.randn(want[1].shape, generator=g, device=dev)
        gb = torch.randn(want[2].shape, generator=g, device=dev)
        gd = torch.randn(want[3].shape, generator=g, device=dev)
        ((out[1] * gz).sum() + (out[2] * gb).sum() + (out[3] * gd).sum()).backward()
        gref = _C.rasterize_meshes_backward(fv, want[0], gz, gb, gd, False, False)
        mine = torch.zeros(fv.shape[0], dtype=torch.bool, device=dev)
        for i in plan.assignment[rank]:
            mine[plan.first[i]: plan.first[i] + plan.num[i]] = True
        scale = float(gref.abs().max())
        assert (fvg.grad[mine] - gref[mine]).abs().max() <= 2e-3 * scale, "local gradients differ"
        assert (fvg.grad[~mine] == 0).all(), "meshes of other ranks must not receive gradients here"
        # (2) packed exchange through peer memory, several steps (both arena halves, reuse)
        loc = plan.local_inputs(fv, rank)
        nbl = torch.full((loc.face_verts.shape[0],), -1, dtype=torch.int64, device=dev)
        nbl._b200_all_minus_one = True
        ex = peer.PackedFrameExchange(plan, rank, size, K)
        try:
            handles = []
            for _ in range(5):
                part = _C.rasterize_meshes(loc.face_verts, loc.first, loc.num, nbl, size, blur, K, 0, 0, False, False,
                                           False)
                handles.append(ex.start(part))
                if len(handles) > 1:
                    got = handles.pop(0).wait()
                    for a, b in zip(got, want):
                        assert torch.equal(a, b), "packed exchange differs from the single-GPU render"
            got = handles.pop(0).wait()
            torch.cuda.synchronize()
            for a, b in zip(got, want):
                assert torch.equal(a, b), "packed exchange differs from the single-GPU render"
        finally:
            ex.close()
        with open(os.path.join(tmp, "ok%d" % rank), "w") as fh:
            fh.write("ok")
    finally:
        dist.destroy_process_group()


def test_two_gpu_gather_and_peer_exchange(built_lib, tmp_path):
    """One process per GPU over NCCL: the gathered Fragments of both transports equal a one-GPU render bit-for-bit,
    and the sharded render back-propagates into the local meshes only."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run with gpurun --gpus 2)")
    import torch.multiprocessing as mp
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_two_rank_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(os.path.join(str(tmp_path), "ok0")) and os.path.exists(os.path.join(str(tmp_path), "ok1"))
