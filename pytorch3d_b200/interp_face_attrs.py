"""`interpolate_face_attributes`, same API as pytorch3d/ops/interp_face_attrs.py:15-102."""
import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import _C


def interpolate_face_attributes(pix_to_face: torch.Tensor, barycentric_coords: torch.Tensor,
                                face_attributes: torch.Tensor) -> torch.Tensor:
    """pix_to_face (N,H,W,K) int64, barycentric_coords (N,H,W,K,3), face_attributes (F,3,D) -> (N,H,W,K,D);
    slots with pix_to_face < 0 give 0.  Same checks and messages as the reference."""
    F, FV, D = face_attributes.shape
    if FV != 3:
        raise ValueError("Faces can only have three vertices; got %r" % FV)
    N, H, W, K, _ = barycentric_coords.shape
    if pix_to_face.shape != (N, H, W, K):
        msg = "pix_to_face must have shape (batch_size, H, W, K); got %r"
        raise ValueError(msg % (pix_to_face.shape,))
    out = _InterpFaceAttrs.apply(pix_to_face.reshape(-1), barycentric_coords.reshape(N * H * W * K, 3),
                                 face_attributes)
    return out.view(N, H, W, K, -1)


class _InterpFaceAttrs(Function):
    @staticmethod
    def forward(ctx, pix_to_face, barycentric_coords, face_attrs):
        args = (pix_to_face, barycentric_coords, face_attrs)
        ctx.save_for_backward(*args)
        return _C.interp_face_attrs_forward(*args)

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_pix_attrs):
        grad_bary, grad_attrs = _C.interp_face_attrs_backward(*ctx.saved_tensors, grad_pix_attrs)
        return None, grad_bary, grad_attrs
