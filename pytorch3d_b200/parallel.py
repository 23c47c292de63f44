"""Batch-sharded rasterization over the GPUs of one node (one process per GPU, torch.distributed).

The reference has no multi-GPU code on this path: every mesh of a batch renders to
its own image, forward and backward are independent per mesh, so the batch shards with NO data-path
collective.  The only communication is the one BASELINE.json's north_star names: an optional all-gather
of the rendered frames (NCCL over NVLink/NVSwitch on GPUs, gloo in the CPU tests).

    plan  = ShardPlan.build(num_faces_per_mesh, world_size)          # greedy LPT on a cost model
    local = plan.local_inputs(face_verts, first, num, rank)          # this rank's packed slice
    frag  = raster_fn(local.face_verts, local.first, local.num, ...) # any rasterize_meshes op
    frag  = plan.rebase(frag, rank)                                  # pix_to_face -> global packed ids
    full  = plan.all_gather(frag, group)                             # optional: whole batch on every rank

`pix_to_face` of the gathered result is bit-identical to a single-GPU render of the whole batch.
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

    def rebase(self, pix_to_face: torch.Tensor, rank: int) -> torch.Tensor:
        """Local packed face ids -> global packed face ids (padding -1 kept)."""
        ids = self.assignment[rank]
        if not ids:
            return pix_to_face
        out = pix_to_face.clone()
        local_first = 0
        for j, i in enumerate(ids):
            shift = self.first[i] - local_first
            img = out[j]
            img[img >= 0] += shift
            local_first += self.num[i]
        return out

    def all_gather(self, tensors: Sequence[torch.Tensor], rank: int, group=None) -> List[torch.Tensor]:
        """All-gather per-rank (n_local, H, W, ...) tensors into (n_meshes, H, W, ...) in batch order.

        Ranks may own different numbers of meshes: each contribution is padded to `max_local` images so
        that one fixed-size collective per tensor suffices (all_gather_into_tensor on NCCL)."""
        outs = []
        for t in tensors:
            pad = self.max_local - t.shape[0]
            if pad > 0:
                t = torch.cat([t, t.new_zeros((pad,) + tuple(t.shape[1:]))], 0)
            t = t.contiguous()
            if self.world_size == 1 or not dist.is_initialized():
                gathered = t.unsqueeze(0)
            else:
                buf = t.new_empty((self.world_size,) + tuple(t.shape))
                if dist.get_backend(group) == "nccl":
                    dist.all_gather_into_tensor(buf, t, group=group)
                else:
                    chunks = list(buf.unbind(0))
                    dist.all_gather(chunks, t, group=group)
                    buf = torch.stack(chunks, 0)
                gathered = buf
            full = t.new_empty((self.n_meshes,) + tuple(t.shape[1:]))
            for r, ids in enumerate(self.assignment):
                for j, i in enumerate(ids):
                    full[i] = gathered[r, j]
            outs.append(full)
        return outs


def rasterize_meshes_sharded(face_verts: torch.Tensor, first: torch.Tensor, num: torch.Tensor, image_size,
                             blur_radius: float, faces_per_pixel: int, perspective_correct: bool = False,
                             clip_barycentric_coords: bool = False, cull_backfaces: bool = False,
                             rank: Optional[int] = None, world_size: Optional[int] = None, group=None,
                             gather: bool = True, raster_fn=None) -> Tuple[torch.Tensor, ...]:
    """Render this rank's share of the batch; optionally all-gather the frames.

    `first` / `num` describe the WHOLE batch (same on every rank); `face_verts` may be the whole packed
    tensor (only this rank's slices are read).  `raster_fn` has the signature of
    pytorch3d_b200._C.rasterize_meshes (default); the CPU tests inject the oracle."""
    if raster_fn is None:
        from . import _C
        raster_fn = _C.rasterize_meshes
    if rank is None:
        rank = dist.get_rank(group) if dist.is_initialized() else 0
    if world_size is None:
        world_size = dist.get_world_size(group) if dist.is_initialized() else 1
    H, W = (image_size, image_size) if isinstance(image_size, int) else image_size
    plan = ShardPlan.build(first.tolist(), num.tolist(), world_size, pixels_per_image=H * W)
    loc = plan.local_inputs(face_verts, rank)
    nb = torch.full((loc.face_verts.shape[0],), -1, dtype=torch.int64, device=face_verts.device)
    nb._b200_all_minus_one = True
    p2f, zbuf, bary, dists = raster_fn(loc.face_verts, loc.first, loc.num, nb, (H, W), blur_radius, faces_per_pixel,
                                       0, 0, perspective_correct, clip_barycentric_coords, cull_backfaces)
    p2f = plan.rebase(p2f, rank)
    if not gather:
        return p2f, zbuf, bary, dists, plan
    full = plan.all_gather([p2f, zbuf, bary, dists], rank, group)
    return full[0], full[1], full[2], full[3], plan
