"""Packed frame exchange over NVLink peer memory (csrc/peer_exchange.cu).

    ex = PackedFrameExchange(plan, rank, (H, W), K)          # collective: every rank constructs it once
    h  = ex.start(local_fragments)                           # local = (pix_to_face, zbuf, bary, dists), LOCAL face ids
    ...                                                      # backward of this batch, forward of the next ...
    p2f, zbuf, bary, dists = h.wait()                        # the whole batch, batch order, global face ids

Per step and rank: ONE kernel packs the rank's frames (1 B per pixel + 24 B per valid slot; the -1 padding does not
travel) and stores the stream into the arena of every rank -- its own and, through CUDA-IPC-mapped pointers, the
peers' -- then a 4-byte NCCL all-reduce on a side stream orders "all streams written" before "streams read", and one
kernel per source expands the streams into dense buffers.  Arenas are double-buffered by step parity; a rank reuses
a half only after the all-reduce of the NEXT step, which every rank enqueues behind its own unpack of this one.
Everything but the pack runs on a side stream: the exchange overlaps the backward pass and the next forward pass.
"""
import ctypes
from typing import List, Sequence

import torch
import torch.distributed as dist

from . import _lib


class _ExchangeHandle:
    def __init__(self, outs, event):
        self._outs, self._event = outs, event

    def wait(self) -> List[torch.Tensor]:
        cur = torch.cuda.current_stream()
        cur.wait_event(self._event)
        for t in self._outs:
            t.record_stream(cur)
        return self._outs


class PackedFrameExchange:
    """All-gather of rendered frames in packed form, pushed into peer memory by the packing kernel itself."""

    MAX_K = 32

    def __init__(self, plan, rank: int, image_size, faces_per_pixel: int, group=None, device=None):
        self.plan, self.rank, self.group = plan, rank, group
        self.H, self.W = (image_size, image_size) if isinstance(image_size, int) else image_size
        self.K = int(faces_per_pixel)
        if self.K > self.MAX_K:
            raise ValueError("PackedFrameExchange supports faces_per_pixel <= %d (use parallel.FrameGather)" % self.MAX_K)
        self.world = plan.world_size
        self.dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.lib = _lib.load()
        self.n_layout = max(plan.max_local, 1)
        with torch.cuda.device(self.dev):
            self.region_bytes = int(self.lib.b200r_packed_frames_bytes(self.n_layout, self.H, self.W, self.K))
            self.half_bytes = self.region_bytes * self.world
            base = ctypes.c_void_p()
            handle = ctypes.create_string_buffer(64)
            _lib.check(self.lib.b200r_peer_alloc(2 * self.half_bytes, ctypes.byref(base), handle))
            self._own = base.value
            # exchange the IPC handles (64 bytes per rank) and map every peer's arena
            mine = torch.tensor(list(handle.raw), dtype=torch.uint8, device=self.dev)
            if self.world > 1:
                allh = torch.empty((self.world, 64), dtype=torch.uint8, device=self.dev)
                dist.all_gather_into_tensor(allh, mine, group=group)
                allh = allh.cpu()
            else:
                allh = mine.cpu().view(1, 64)
            self._arena = []
            for r in range(self.world):
                if r == rank:
                    self._arena.append(self._own)
                    continue
                p = ctypes.c_void_p()
                _lib.check(self.lib.b200r_peer_open(bytes(allh[r].tolist()), ctypes.byref(p)))
                self._arena.append(p.value)
            self._cursor = torch.zeros(1, dtype=torch.int32, device=self.dev)
            self._token = torch.zeros(1, dtype=torch.int32, device=self.dev)
            self._side = torch.cuda.Stream(device=self.dev)
            # per source rank: where its frames go in the batch and how its local face ids become global ones
            self._image_index, self._face_shift = [], []
            for r in range(self.world):
                ids = plan.assignment[r]
                self._image_index.append(torch.tensor(ids if ids else [0], dtype=torch.int32, device=self.dev))
                sh = plan.local_shifts(r)
                self._face_shift.append(torch.tensor(sh if sh else [0], dtype=torch.int64, device=self.dev))
            self._free = [None, None]  # event after which a parity half may be overwritten by the peers
            self._step = 0
        if self.world > 1:
            dist.barrier(group=group)  # every arena is mapped before anyone pushes

    def _region(self, holder: int, parity: int, source: int) -> int:
        return self._arena[holder] + parity * self.half_bytes + source * self.region_bytes

    def start(self, fragments: Sequence[torch.Tensor]) -> _ExchangeHandle:
        p2f, zbuf, bary, dists = (t.contiguous() for t in fragments)
        n_local = int(p2f.shape[0])
        assert n_local == len(self.plan.assignment[self.rank]) and tuple(p2f.shape[1:]) == (self.H, self.W, self.K)
        parity = self._step & 1
        cur = torch.cuda.current_stream(self.dev)
        side = self._side
        lib = self.lib
        with torch.cuda.device(self.dev):
            if self._free[parity] is not None:
                cur.wait_event(self._free[parity])
            # ---- pack + push (compute stream: right behind the forward pass that produced the fragments)
            dst = (ctypes.c_void_p * self.world)(*[self._region(r, parity, self.rank) for r in range(self.world)])
            _lib.check(lib.b200r_fragments_pack_push(
                p2f.data_ptr(), zbuf.data_ptr(), bary.data_ptr(), dists.data_ptr(), n_local, self.H, self.W, self.K,
                self.n_layout, dst, self.world, self._cursor.data_ptr(), cur.cuda_stream))
            packed = torch.cuda.Event()
            packed.record(cur)
            # ---- side stream: all ranks packed -> expand every source's stream
            side.wait_event(packed)
            with torch.cuda.stream(side):
                if self.world > 1:
                    dist.all_reduce(self._token, group=self.group)
                # the all-reduce of step i also proves that every rank has finished UNPACKING step i-1
                prev_done = torch.cuda.Event()
                prev_done.record(side)
                self._free[parity ^ 1] = prev_done
                N = self.plan.n_meshes
                full = [torch.empty((N, self.H, self.W, self.K), dtype=torch.int64, device=self.dev),
                        torch.empty((N, self.H, self.W, self.K), dtype=torch.float32, device=self.dev),
                        torch.empty((N, self.H, self.W, self.K, 3), dtype=torch.float32, device=self.dev),
                        torch.empty((N, self.H, self.W, self.K), dtype=torch.float32, device=self.dev)]
                for r in range(self.world):
                    n_r = len(self.plan.assignment[r])
                    if n_r == 0:
                        continue
                    _lib.check(lib.b200r_fragments_unpack(
                        self._region(self.rank, parity, r), n_r, self.H, self.W, self.K, self.n_layout,
                        self._image_index[r].data_ptr(), self._face_shift[r].data_ptr(), full[0].data_ptr(),
                        full[1].data_ptr(), full[2].data_ptr(), full[3].data_ptr(), side.cuda_stream))
                done = torch.cuda.Event()
                done.record(side)
        self._step += 1
        return _ExchangeHandle(full, done)

    def close(self):
        torch.cuda.synchronize(self.dev)
        if self.world > 1 and dist.is_initialized():
            dist.barrier(group=self.group)
        for r, p in enumerate(self._arena):
            if r != self.rank and p:
                self.lib.b200r_peer_close(p)
        if self._own:
            self.lib.b200r_peer_free(self._own)
        self._arena, self._own = [], None
