"""Batch-sharded rasterization over the GPUs of one node (one process per GPU, torch.distributed).

The reference has no multi-GPU code on this path: every mesh of a batch renders to
its own image, forward and backward are independent per mesh, so the batch shards with NO data-path
collective.  The only communication is the one BASELINE.json's north_star names: an optional all-gather
of the rendered frames (NCCL over NVLink/NVSwitch on GPUs, gloo in the CPU tests).

    plan  = ShardPlan.build(first, num, world_size)                  # greedy LPT on a cost model
    local = plan.local_inputs(face_verts, rank)                      # this rank's packed slice
    frag  = raster_fn(local.face_verts, local.first, local.num, ...) # any rasterize_meshes op
    p2f   = plan.rebase(frag[0], rank)                               # pix_to_face -> global packed ids
    h     = FrameGather(plan, rank).start((p2f,) + frag[1:])         # whole batch on every rank, on a side stream
    full  = h.wait()                                                 # ... while the backward pass runs

`pix_to_face` of the gathered result is bit-identical to a single-GPU render of the whole batch.
Two transports: `FrameGather` (dense NCCL all-gather, pix_to_face narrowed to int32 on the wire) and
`peer.PackedFrameExchange` (the valid slots only, pushed into every peer's memory over NVLink by one kernel).
"""
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def lpt_partition(costs: Sequence[float], world_size: int) -> List[List[int]]:
    """Longest-processing-time-first assignment of items to `world_size` bins (ties: lower rank)."""
    order = sorted(range(len(costs)), key=lambda i: (-float(costs[i]), i))
    loads = [0.0] * world_size
    bins: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (loads[k], k))
        bins[r].append(i)
        loads[r] += float(costs[i])
    return [sorted(b) for b in bins]


@dataclass
class LocalInputs:
    face_verts: torch.Tensor  # (F_local, 3, 3)
    first: torch.Tensor  # (n_local,) first packed face of each local mesh, in the LOCAL packing
    num: torch.Tensor  # (n_local,)
    mesh_ids: List[int]  # global mesh index of each local mesh


class ShardPlan:
    """Which meshes each rank renders, and how local results map back to the global batch."""

    def __init__(self, assignment: List[List[int]], first: Sequence[int], num: Sequence[int]):
        self.assignment = assignment
        self.first = [int(v) for v in first]
        self.num = [int(v) for v in num]
        self.world_size = len(assignment)
        self.n_meshes = len(self.num)
        self.max_local = max((len(a) for a in assignment), default=0)

    @staticmethod
    def build(first: Sequence[int], num: Sequence[int], world_size: int, pixels_per_image: int = 0,
              alpha: float = 1.0, beta: float = 0.0) -> "ShardPlan":
        """cost(mesh) = alpha * faces + beta * pixels; defaults weigh faces only."""
        costs = [alpha * float(n) + beta * float(pixels_per_image) for n in num]
        return ShardPlan(lpt_partition(costs, world_size), first, num)

    def local_inputs(self, face_verts: torch.Tensor, rank: int) -> LocalInputs:
        ids = self.assignment[rank]
        parts = [face_verts[self.first[i]: self.first[i] + self.num[i]] for i in ids]
        fv = torch.cat(parts, 0) if parts else face_verts[:0]
        num = torch.tensor([self.num[i] for i in ids], dtype=torch.int64, device=face_verts.device)
        first = torch.zeros_like(num)
        if len(ids) > 1:
            first[1:] = torch.cumsum(num, 0)[:-1]
        return LocalInputs(fv.contiguous(), first, num, list(ids))

    def local_shifts(self, rank: int) -> List[int]:
        """Per local mesh: (first global packed face) - (first local packed face)."""
        out, local_first = [], 0
        for i in self.assignment[rank]:
            out.append(self.first[i] - local_first)
            local_first += self.num[i]
        return out

    def rebase(self, pix_to_face: torch.Tensor, rank: int) -> torch.Tensor:
        """Local packed face ids -> global packed face ids (padding -1 kept); one fused elementwise pass."""
        shifts = self.local_shifts(rank)
        if not shifts or not any(shifts):
            return pix_to_face
        s = torch.tensor(shifts, dtype=pix_to_face.dtype, device=pix_to_face.device).view(-1, 1, 1, 1)
        return torch.where(pix_to_face >= 0, pix_to_face + s, pix_to_face)

    def scatter_to_batch_order(self, gathered: torch.Tensor) -> torch.Tensor:
        """(world, max_local, ...) per-rank blocks -> (n_meshes, ...) in batch order."""
        if all(self.assignment[r] == list(range(r * self.max_local, (r + 1) * self.max_local))
               for r in range(self.world_size)):
            return gathered.reshape((self.n_meshes,) + tuple(gathered.shape[2:]))  # already in order: a view
        full = gathered.new_empty((self.n_meshes,) + tuple(gathered.shape[2:]))
        for r, ids in enumerate(self.assignment):
            if ids:
                full[torch.tensor(ids, device=gathered.device)] = gathered[r, : len(ids)]
        return full

    def all_gather(self, tensors: Sequence[torch.Tensor], rank: int, group=None) -> List[torch.Tensor]:
        """Synchronous all-gather of per-rank (n_local, H, W, ...) tensors into (n_meshes, H, W, ...) in batch
        order (see FrameGather for the overlapped form)."""
        return FrameGather(self, rank, group=group, overlap=False).start(tensors).wait()


class _GatherHandle:
    def __init__(self, outs, event, stream):
        self._outs, self._event, self._stream = outs, event, stream

    def wait(self) -> List[torch.Tensor]:
        """Make the current stream wait for the gather; returns the full-batch tensors."""
        if self._event is not None:
            torch.cuda.current_stream().wait_event(self._event)
            for t in self._outs:
                t.record_stream(torch.cuda.current_stream())
        return self._outs


class FrameGather:
    """All-gather of the rendered frames (the path's only collective), off the critical path.

    Ranks may own different numbers of meshes: each contribution is padded to `max_local` images so that one
    fixed-size collective per tensor suffices (all_gather_into_tensor on NCCL).  int64 tensors whose values fit
    32 bits (pix_to_face: packed face ids < 2^31) travel as int32 and are widened on arrival: 24 instead of 28
    bytes per slot on the wire.  With `overlap` (CUDA only) everything runs on a side stream: `start()` returns at
    once and the caller's stream only waits in `handle.wait()` -- e.g. after the backward pass of the same batch."""

    def __init__(self, plan: ShardPlan, rank: int, group=None, overlap: bool = True):
        self.plan, self.rank, self.group = plan, rank, group
        self.overlap = overlap
        self._stream = None

    def _gather_one(self, t: torch.Tensor) -> torch.Tensor:
        plan = self.plan
        narrow = t.dtype == torch.int64
        if narrow:
            t = t.to(torch.int32)
        pad = plan.max_local - t.shape[0]
        if pad > 0:
            t = torch.cat([t, t.new_zeros((pad,) + tuple(t.shape[1:]))], 0)
        t = t.contiguous()
        if plan.world_size == 1 or not dist.is_initialized():
            gathered = t.unsqueeze(0)
        else:
            buf = t.new_empty((plan.world_size,) + tuple(t.shape))
            if dist.get_backend(self.group) == "nccl":
                dist.all_gather_into_tensor(buf, t, group=self.group)
            else:
                chunks = list(buf.unbind(0))
                dist.all_gather(chunks, t, group=self.group)
                buf = torch.stack(chunks, 0)
            gathered = buf
        full = plan.scatter_to_batch_order(gathered)
        return full.to(torch.int64) if narrow else full

    def start(self, tensors: Sequence[torch.Tensor]) -> _GatherHandle:
        tensors = list(tensors)
        if not (self.overlap and tensors and tensors[0].is_cuda):
            return _GatherHandle([self._gather_one(t) for t in tensors], None, None)
        dev = tensors[0].device
        if self._stream is None:
            self._stream = torch.cuda.Stream(device=dev)
        side = self._stream
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for t in tensors:
                t.record_stream(side)
            outs = [self._gather_one(t) for t in tensors]
            ev = torch.cuda.Event()
            ev.record(side)
        return _GatherHandle(outs, ev, side)


def rasterize_meshes_sharded(face_verts: torch.Tensor, first: torch.Tensor, num: torch.Tensor, image_size,
                             blur_radius: float, faces_per_pixel: int, perspective_correct: bool = False,
                             clip_barycentric_coords: bool = False, cull_backfaces: bool = False,
                             rank: Optional[int] = None, world_size: Optional[int] = None, group=None,
                             gather: bool = True, raster_fn=None) -> Tuple[torch.Tensor, ...]:
    """Render this rank's share of the batch; optionally all-gather the frames.

    `first` / `num` describe the WHOLE batch (same on every rank); `face_verts` may be the whole packed
    tensor (only this rank's slices are read).  The local render goes through the autograd Function of
    `pytorch3d_b200.rasterize_meshes`, so zbuf / bary / dists of THIS rank's meshes carry gradients back to
    `face_verts` (no collective in the backward pass: every rank owns its meshes' vertices); frames of other
    ranks arrive as constants.  `raster_fn` (signature of pytorch3d_b200._C.rasterize_meshes) replaces the
    native op -- the CPU tests inject the oracle; it is not differentiable.

    Returns (pix_to_face, zbuf, bary, dists, plan): local frames if `gather` is False, else the whole batch."""
    if rank is None:
        rank = dist.get_rank(group) if dist.is_initialized() else 0
    if world_size is None:
        world_size = dist.get_world_size(group) if dist.is_initialized() else 1
    H, W = (image_size, image_size) if isinstance(image_size, int) else image_size
    plan = ShardPlan.build(first.tolist(), num.tolist(), world_size, pixels_per_image=H * W)
    loc = plan.local_inputs(face_verts, rank)
    nb = torch.full((loc.face_verts.shape[0],), -1, dtype=torch.int64, device=face_verts.device)
    nb._b200_all_minus_one = True
    if raster_fn is None:
        from .rasterize_meshes import _RasterizeFaceVerts
        p2f, zbuf, bary, dists = _RasterizeFaceVerts.apply(
            loc.face_verts, loc.first, loc.num, nb, (H, W), blur_radius, faces_per_pixel, 0, 0, perspective_correct,
            clip_barycentric_coords, cull_backfaces)
    else:
        p2f, zbuf, bary, dists = raster_fn(loc.face_verts, loc.first, loc.num, nb, (H, W), blur_radius,
                                           faces_per_pixel, 0, 0, perspective_correct, clip_barycentric_coords,
                                           cull_backfaces)
    p2f = plan.rebase(p2f, rank)
    if not gather:
        return p2f, zbuf, bary, dists, plan
    full = FrameGather(plan, rank, group=group).start(
        [p2f, zbuf.detach(), bary.detach(), dists.detach()]).wait()
    if zbuf.requires_grad and loc.mesh_ids:
        # this rank's own frames stay differentiable inside the gathered batch
        ids = torch.tensor(loc.mesh_ids, device=zbuf.device)
        full[1] = full[1].index_copy(0, ids, zbuf)
        full[2] = full[2].index_copy(0, ids, bary)
        full[3] = full[3].index_copy(0, ids, dists)
    return full[0], full[1], full[2], full[3], plan
