"""Component timings of the two frame-exchange transports (run under torchrun on >= 2 GPUs; development aid)."""
import ctypes
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pytorch3d_b200 import _C, _lib, parallel, peer, synthetic  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
lib = _lib.load()
nm, H, W, K = 8, 512, 512, 8
m = synthetic.torus_batch(nm, 187, 187, seed=0)
fv = synthetic.face_verts_of(m).to(dev)
first, num = m.mesh_to_faces_packed_first_idx().to(dev), m.num_faces_per_mesh().to(dev)
nb = torch.full((fv.shape[0],), -1, dtype=torch.int64, device=dev)
nb._b200_all_minus_one = True
f = _C.rasterize_meshes(fv, first, num, nb, (H, W), 0.0, K, 0, 0, False, False, False)
N = nm * world
plan = parallel.ShardPlan([list(range(r * nm, (r + 1) * nm)) for r in range(world)], [0] * N, [0] * N)


def gpu_ms(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    host = (time.perf_counter() - t0) / n * 1e3
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, host


out = {}
# ---- packed path components
ex = peer.PackedFrameExchange(plan, rank, (H, W), K)
cur = torch.cuda.current_stream(dev)


def region(holder, parity, source):
    return ex._arena[holder] + parity * ex.half_bytes + source * ex.region_bytes


cursor = torch.zeros(1, dtype=torch.int32, device=dev)
image_index = [torch.tensor(plan.assignment[r], dtype=torch.int32, device=dev) for r in range(world)]
face_shift = [torch.zeros(nm, dtype=torch.int64, device=dev) for r in range(world)]
dst_self = (ctypes.c_void_p * 1)(region(rank, 0, rank))
dst_all = (ctypes.c_void_p * world)(*[region(r, 0, rank) for r in range(world)])
dst_peer = (ctypes.c_void_p * 1)(region((rank + 1) % world, 0, rank))


def pack(dst, n):
    _lib.check(lib.b200r_fragments_pack_push(f[0].data_ptr(), f[1].data_ptr(), f[2].data_ptr(), f[3].data_ptr(), nm, H, W,
                                             K, ex.n_layout, dst, n, cursor.data_ptr(), cur.cuda_stream))


out["pack_to_self_ms"] = gpu_ms(lambda: pack(dst_self, 1))
out["pack_to_peer_ms"] = gpu_ms(lambda: pack(dst_peer, 1))
out["pack_to_all_ms"] = gpu_ms(lambda: pack(dst_all, world))
full = [torch.empty((N, H, W, K), dtype=torch.int64, device=dev), torch.empty((N, H, W, K), device=dev),
        torch.empty((N, H, W, K, 3), device=dev), torch.empty((N, H, W, K), device=dev)]
pack(dst_all, world)
torch.cuda.synchronize()
dist.barrier()


def unpack_all():
    for r in range(world):
        _lib.check(lib.b200r_fragments_unpack(region(rank, 0, r), nm, H, W, K, ex.n_layout,
                                              image_index[r].data_ptr(), face_shift[r].data_ptr(),
                                              full[0].data_ptr(), full[1].data_ptr(), full[2].data_ptr(),
                                              full[3].data_ptr(), cur.cuda_stream))


out["unpack_%d_sources_ms" % world] = gpu_ms(unpack_all)
tok = torch.zeros(1, dtype=torch.int32, device=dev)
out["allreduce_token_ms"] = gpu_ms(lambda: dist.all_reduce(tok))
out["exchange_sync_ms"] = gpu_ms(lambda: ex.start(f).wait())
# ---- dense path components
fg = parallel.FrameGather(plan, rank)
out["dense_sync_ms"] = gpu_ms(lambda: fg.start(f).wait())
t32 = f[0].to(torch.int32)
buf = t32.new_empty((world,) + tuple(t32.shape))
out["allgather_p2f_int32_ms"] = gpu_ms(lambda: dist.all_gather_into_tensor(buf, t32))
bufb = f[2].new_empty((world,) + tuple(f[2].shape))
out["allgather_bary_ms"] = gpu_ms(lambda: dist.all_gather_into_tensor(bufb, f[2]))
out["narrow_ms"] = gpu_ms(lambda: f[0].to(torch.int32))
out["widen_ms"] = gpu_ms(lambda: buf.to(torch.int64))
# ---- with compute overlapped (as bench.py does)
gz, gb, gd = torch.randn_like(f[1]), torch.randn_like(f[2]), torch.randn_like(f[3])


def step(start):
    ff = _C.rasterize_meshes(fv, first, num, nb, (H, W), 0.0, K, 0, 0, False, False, False)
    h = start(ff) if start else None
    _C.rasterize_meshes_backward(fv, ff[0], gz, gb, gd, False, False)
    return h


def loop(start):
    prev = None

    def one():
        nonlocal prev
        h = step(start)
        if prev is not None:
            prev.wait()
        prev = h
    return one


out["step_no_exchange_ms"] = gpu_ms(lambda: step(None), n=20)
out["step_packed_ms"] = gpu_ms(loop(ex.start), n=20)
out["step_dense_ms"] = gpu_ms(loop(fg.start), n=20)
ex.close()
if rank == 0:
    for k, v in out.items():
        print("%-28s gpu %.3f ms   host %.3f ms" % (k, v[0], v[1]), flush=True)
dist.destroy_process_group()
