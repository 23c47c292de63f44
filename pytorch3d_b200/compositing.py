"""Compositing of point features, same API as the reference:
pytorch3d/renderer/compositing.py:19-250 (`alpha_composite`, `norm_weighted_sum`, `weighted_sum`) and
points/compositor.py:22-116 (`AlphaCompositor`, `NormWeightedCompositor`).
"""
import torch
import torch.nn as nn

from . import _C


class _CompositeAlphaPoints(torch.autograd.Function):
    """weighted_fs[b,c,i,j] = sum_k cum_alpha_k * features[c, pointsidx[b,k,i,j]],
    cum_alpha_k = alphas[b,k,i,j] * prod_{l<k} (1 - alphas[b,l,i,j])   (compositing.py:19-63 of the reference)."""

    @staticmethod
    def forward(ctx, features, alphas, points_idx):
        pt_cld = _C.accum_alphacomposite(features, alphas, points_idx)
        ctx.save_for_backward(features.clone(), alphas.clone(), points_idx.clone())
        return pt_cld

    @staticmethod
    def backward(ctx, grad_output):
        features, alphas, points_idx = ctx.saved_tensors
        grad_features, grad_alphas = _C.accum_alphacomposite_backward(grad_output, features, alphas, points_idx)
        return grad_features, grad_alphas, None


def alpha_composite(pointsidx, alphas, pt_clds) -> torch.Tensor:
    """pointsidx (N,K,H,W) int64, alphas (N,K,H,W) in [0,1], pt_clds (C,P) packed features -> (N,C,H,W)."""
    return _CompositeAlphaPoints.apply(pt_clds, alphas, pointsidx)


class _RenderPointsAlpha(torch.autograd.Function):
    """images = alpha_composite(idx, 1 - dists / radius**2, features) in one kernel per direction, on the rasterizer's
    (N,H,W,K) layout (see `render_points_alpha`)."""

    @staticmethod
    def forward(ctx, features, dists, idx, radius):
        images = _C.points_alpha_render(features, idx, dists, radius)
        ctx.save_for_backward(features, dists, idx)
        ctx.radius = radius
        ctx.mark_non_differentiable(idx)
        return images

    @staticmethod
    def backward(ctx, grad_images):
        features, dists, idx = ctx.saved_tensors
        grad_features, grad_dists = _C.points_alpha_render_backward(grad_images, features, idx, dists, ctx.radius)
        return grad_features, grad_dists, None, None


def render_points_alpha(fragments, features, radius: float) -> torch.Tensor:
    """Fused form of what `PointsRenderer.forward` does with an `AlphaCompositor`
    (pytorch3d/renderer/points/renderer.py:63-73 of the reference):

        weights = 1 - fragments.dists.permute(0, 3, 1, 2) / (radius * radius)
        images = alpha_composite(fragments.idx.long().permute(0, 3, 1, 2), weights, features)

    `fragments`: PointFragments (or (idx, zbuf, dists)); `features`: (C, P) packed features (the reference passes
    `features_packed().permute(1, 0)`); scalar `radius`.  Returns (N, C, H, W) like the compositor (the renderer then
    permutes to (N, H, W, C)).  Forward values are bit-identical to the unfused chain; gradients flow to `features` and
    to `fragments.dists`."""
    idx = fragments.idx if hasattr(fragments, "idx") else fragments[0]
    dists = fragments.dists if hasattr(fragments, "dists") else fragments[2]
    return _RenderPointsAlpha.apply(features, dists, idx, float(radius))


def _make_composite(forward_op, backward_op, doc):
    class _Composite(torch.autograd.Function):
        @staticmethod
        def forward(ctx, features, alphas, points_idx):
            pt_cld = forward_op(features, alphas, points_idx)
            ctx.save_for_backward(features.clone(), alphas.clone(), points_idx.clone())
            return pt_cld

        @staticmethod
        def backward(ctx, grad_output):
            features, alphas, points_idx = ctx.saved_tensors
            grad_features, grad_alphas = backward_op(grad_output, features, alphas, points_idx)
            return grad_features, grad_alphas, None

    _Composite.__doc__ = doc
    return _Composite


_CompositeNormWeightedSumPoints = _make_composite(
    _C.accum_weightedsumnorm, _C.accum_weightedsumnorm_backward,
    "weighted_fs[b,c,i,j] = sum_k alphas[b,k,i,j] * features[c, pointsidx[b,k,i,j]] / max(sum_k alphas[b,k,i,j], 1e-4)"
    "   (compositing.py:99-146 of the reference).")
_CompositeWeightedSumPoints = _make_composite(
    _C.accum_weightedsum, _C.accum_weightedsum_backward,
    "weighted_fs[b,c,i,j] = sum_k alphas[b,k,i,j] * features[c, pointsidx[b,k,i,j]]"
    "   (compositing.py:177-222 of the reference).")


def norm_weighted_sum(pointsidx, alphas, pt_clds) -> torch.Tensor:
    """Normalised weighted sum (compositing.py:149-174 of the reference): same arguments as `alpha_composite`."""
    return _CompositeNormWeightedSumPoints.apply(pt_clds, alphas, pointsidx)


def weighted_sum(pointsidx, alphas, pt_clds) -> torch.Tensor:
    """Weighted sum (compositing.py:225-250 of the reference): same arguments as `alpha_composite`."""
    return _CompositeWeightedSumPoints.apply(pt_clds, alphas, pointsidx)


class AlphaCompositor(nn.Module):
    """Accumulate points using alpha compositing (points/compositor.py:22-66 of the reference)."""

    def __init__(self, background_color=None) -> None:
        super().__init__()
        self.background_color = background_color

    def forward(self, fragments, alphas, ptclds, **kwargs) -> torch.Tensor:
        background_color = kwargs.get("background_color", self.background_color)
        images = alpha_composite(fragments, alphas, ptclds)
        if background_color is not None:
            images = _add_background_color_to_images(fragments, images, background_color)
        return images


class NormWeightedCompositor(nn.Module):
    """Accumulate points using a normalised weighted sum (points/compositor.py:69-116 of the reference)."""

    def __init__(self, background_color=None) -> None:
        super().__init__()
        self.background_color = background_color

    def forward(self, fragments, alphas, ptclds, **kwargs) -> torch.Tensor:
        background_color = kwargs.get("background_color", self.background_color)
        images = norm_weighted_sum(fragments, alphas, ptclds)
        if background_color is not None:
            images = _add_background_color_to_images(fragments, images, background_color)
        return images


def _add_background_color_to_images(pix_idxs, images, background_color):
    """Pixels that no point covers take the background colour (points/compositor.py:119-160 of the reference)."""
    background_mask = pix_idxs[:, 0] < 0  # (N, H, W)
    if not torch.is_tensor(background_color):
        background_color = images.new_tensor(background_color)
    background_color = background_color.to(images)
    if background_color.ndim == 0:
        background_color = background_color.expand(images.shape[1])
    if background_color.ndim > 1:
        raise ValueError("Wrong shape of background_color")
    if background_color.shape[0] + 1 == images.shape[1]:
        alpha = images.new_ones(1)
        background_color = torch.cat([background_color, alpha])
    elif background_color.shape[0] != images.shape[1]:
        raise ValueError("Background color has %s channels not %s" % (background_color.shape[0], images.shape[1]))
    num_background_pixels = background_mask.sum()
    masked_images = images.permute(0, 2, 3, 1).masked_scatter(
        background_mask[..., None], background_color[None, :].expand(num_background_pixels, -1))
    return masked_images.permute(0, 3, 1, 2)
