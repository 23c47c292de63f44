"""Packed frame exchange over NVLink peer memory (csrc/peer_exchange.cu).

    ex = PackedFrameExchange(plan, rank, (H, W), K)          # collective: every rank constructs it once
    h  = ex.start(local_fragments)                           # local = (pix_to_face, zbuf, bary, dists), LOCAL face ids
    ...                                                      # backward of this batch, forward of the next ...
    p2f, zbuf, bary, dists = h.wait()                        # the whole batch, batch order, global face ids

Per step and rank: ONE kernel packs the rank's frames (1 B per pixel + 24 B per valid slot; the -1 padding does not
travel) and stores the stream into the arena of every rank -- its own and, through CUDA-IPC-mapped pointers, the
peers' -- then a 4-byte NCCL all-reduce on a side stream orders "all streams written" before "streams read", and one
kernel per source expands the streams into dense buffers.  Arenas and result buffers are double-buffered by step
parity; a rank reuses an arena half only after the all-reduce of the NEXT step, which every rank enqueues behind its
own expansion of this one.  Everything but the pack runs on a side stream: the exchange overlaps the backward pass and
the next forward pass.  The step's bookkeeping (events, launches) lives in the library: three C calls per step.
"""
import ctypes
from typing import List, Sequence

import torch
import torch.distributed as dist

from . import _lib


class _ExchangeHandle:
    def __init__(self, owner, outs, parity):
        self._owner, self._outs, self._parity = owner, outs, parity

    def wait(self) -> List[torch.Tensor]:
        """The current stream waits for the exchange; returns (pix_to_face, zbuf, bary, dists) of the whole batch.
        The tensors are rewritten by the exchange after the next one: consume them (on the stream that waited) before
        starting that."""
        o = self._owner
        cur = torch.cuda.current_stream(o.dev)
        _lib.check(o.lib.b200r_exchange_wait(o._handle, self._parity, cur.cuda_stream))
        o._reader = cur
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
        self.lib = lib = _lib.load()
        self.n_layout = max(plan.max_local, 1)
        self._handle, self._own, self._arena = None, None, []
        with torch.cuda.device(self.dev):
            self.region_bytes = int(lib.b200r_packed_frames_bytes(self.n_layout, self.H, self.W, self.K))
            self.half_bytes = self.region_bytes * self.world
            base = ctypes.c_void_p()
            handle = ctypes.create_string_buffer(64)
            _lib.check(lib.b200r_peer_alloc(2 * self.half_bytes, ctypes.byref(base), handle))
            self._own = base.value
            # exchange the IPC handles (64 bytes per rank) and map every peer's arena
            mine = torch.tensor(list(handle.raw), dtype=torch.uint8, device=self.dev)
            if self.world > 1:
                allh = torch.empty((self.world, 64), dtype=torch.uint8, device=self.dev)
                dist.all_gather_into_tensor(allh, mine, group=group)
                allh = allh.cpu()
            else:
                allh = mine.cpu().view(1, 64)
            for r in range(self.world):
                if r == rank:
                    self._arena.append(self._own)
                    continue
                p = ctypes.c_void_p()
                _lib.check(lib.b200r_peer_open(bytes(allh[r].tolist()), ctypes.byref(p)))
                self._arena.append(p.value)
            # per source rank: where its frames go in the batch and how its local face ids become global ones
            counts, index, shift = [], [], []
            for r in range(self.world):
                ids = plan.assignment[r]
                counts.append(len(ids))
                index += list(ids)
                shift += plan.local_shifts(r)
            n = max(len(index), 1)
            c_counts = (ctypes.c_int32 * self.world)(*counts)
            c_index = (ctypes.c_int32 * n)(*(index or [0]))
            c_shift = (ctypes.c_int64 * n)(*(shift or [0]))
            c_arenas = (ctypes.c_void_p * self.world)(*self._arena)
            h = ctypes.c_void_p()
            _lib.check(lib.b200r_exchange_create(self.world, rank, self.H, self.W, self.K, self.n_layout, c_counts,
                                                 c_index, c_shift, c_arenas, ctypes.byref(h)))
            self._handle = h.value
            self._token = torch.zeros(1, dtype=torch.int32, device=self.dev)
            self._side = torch.cuda.Stream(device=self.dev)
            self._out = [None, None]
            self._reader = None
        if self.world > 1:
            dist.barrier(group=group)  # every arena is mapped before anyone pushes

    def _outputs(self, parity: int):
        """The full-batch result buffers, double-buffered like the arenas."""
        if self._out[parity] is None:
            N, shape = self.plan.n_meshes, (self.H, self.W, self.K)
            self._out[parity] = [torch.empty((N,) + shape, dtype=torch.int64, device=self.dev),
                                 torch.empty((N,) + shape, dtype=torch.float32, device=self.dev),
                                 torch.empty((N,) + shape + (3,), dtype=torch.float32, device=self.dev),
                                 torch.empty((N,) + shape, dtype=torch.float32, device=self.dev)]
        return self._out[parity]

    def start(self, fragments: Sequence[torch.Tensor]) -> _ExchangeHandle:
        p2f, zbuf, bary, dists = (t if t.is_contiguous() else t.contiguous() for t in fragments)
        if p2f.shape[0] != len(self.plan.assignment[self.rank]) or tuple(p2f.shape[1:]) != (self.H, self.W, self.K):
            raise ValueError("fragments do not match the exchange's plan / image size / faces_per_pixel")
        lib, h = self.lib, self._handle
        cur = torch.cuda.current_stream(self.dev)
        side = self._side
        reader = self._reader if self._reader is not None else cur
        # the caller's tensors are read by the pack on `cur` only; nothing to record on the side stream
        _lib.check(lib.b200r_exchange_push(h, p2f.data_ptr(), zbuf.data_ptr(), bary.data_ptr(), dists.data_ptr(),
                                           cur.cuda_stream, side.cuda_stream, reader.cuda_stream))
        if self.world > 1:
            with torch.cuda.stream(side):
                dist.all_reduce(self._token, group=self.group)  # "every rank has pushed" (and expanded the step before)
        parity = ctypes.c_int32()
        # (parity of this step = number of completed steps & 1; the library returns it)
        outs = self._outputs(self._peek_parity())
        _lib.check(lib.b200r_exchange_expand(h, outs[0].data_ptr(), outs[1].data_ptr(), outs[2].data_ptr(),
                                             outs[3].data_ptr(), side.cuda_stream, ctypes.byref(parity)))
        self._steps = getattr(self, "_steps", 0) + 1
        return _ExchangeHandle(self, outs, parity.value)

    def _peek_parity(self) -> int:
        return getattr(self, "_steps", 0) & 1

    def close(self):
        if self._handle is None:
            return
        torch.cuda.synchronize(self.dev)
        if self.world > 1 and dist.is_initialized():
            dist.barrier(group=self.group)
        self.lib.b200r_exchange_destroy(self._handle)
        self._handle = None
        for r, p in enumerate(self._arena):
            if r != self.rank and p:
                self.lib.b200r_peer_close(p)
        if self._own:
            self.lib.b200r_peer_free(self._own)
        self._arena, self._own, self._out = [], None, [None, None]
