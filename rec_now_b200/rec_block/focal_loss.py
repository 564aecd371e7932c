"""Drop-in for rec_now/rec_block/focal_loss.py of the reference on torch CUDA tensors (cited as FL:n), and the fused
joint objective the production models use (a pointwise focal loss next to the in-batch pairwise loss, README.md:57).

``focal_crossentropy_loss`` is a pointwise elementwise loss: here it is a handful of torch ops with the reference's
signature.  What this repo adds is ``pairwise_loss_with_focal``: ONE C-ABI call computes
``pairwise_loss(...) + focal_weight * focal_crossentropy_loss(labels, outputs, ...)`` and its gradient, the focal part
riding in the pair kernel's first and last pass (rn_pairwise_args.focal_*).  There is no CPU path.
"""
from __future__ import annotations

import torch

from .. import ops
from .pairwise_loss_from_batch import FusedPairWeight, _as_cuda, _match_bpr, bpr_loss_func


def focal_crossentropy_loss(labels, logits, alpha=0.25, gamma=2.0, stop_weight_gradient=False, return_mean=True):
    """FL:12-66.  Same arguments, defaults and errors as the reference."""
    if alpha and (alpha <= 0.0 or alpha >= 1.0):
        raise ValueError("Value of alpha should be greater than zero and less than one.")          # FL:43-44
    if gamma and gamma < 0:
        raise ValueError("Value of gamma should be greater than or equal to zero.")                # FL:45-46
    labels, logits = _as_cuda(labels), _as_cuda(logits)
    # sigmoid_cross_entropy_with_logits: max(x, 0) - x z + log1p(exp(-|x|))                        FL:48
    focal_loss = torch.clamp(logits, min=0) - logits * labels + torch.log1p(torch.exp(-torch.abs(logits)))
    if alpha:
        focal_loss = (labels * alpha + (1 - labels) * (1 - alpha)) * focal_loss                    # FL:50-53
    if gamma:
        pred_prob = torch.sigmoid(logits)                                                          # FL:56
        pred_sim = labels * pred_prob + (1 - labels) * (1 - pred_prob)                             # FL:57
        modulating_factor = torch.pow(1.0 - pred_sim, gamma)                                       # FL:59
        if stop_weight_gradient:
            modulating_factor = modulating_factor.detach()                                         # FL:60-61
        focal_loss = modulating_factor * focal_loss                                                # FL:62
    if return_mean:
        focal_loss = torch.mean(focal_loss)                                                        # FL:64-65
    return focal_loss


class _FusedPairwiseFocal(torch.autograd.Function):
    @staticmethod
    def forward(ctx, outputs, labels, keys, row_ok, rw_pos, label_func, factor, reduce_mean, power, focal):
        out = ops.pairwise_fwd_bwd(outputs, labels, keys, row_ok=row_ok, rw_pos=rw_pos, label_func=label_func,
                                   factor=factor, power=power, reduce_mean=reduce_mean, focal=focal)
        ctx.save_for_backward(out["dlogits"])
        ctx.out_shape, ctx.out_dtype = outputs.shape, outputs.dtype
        n_pair = out["n_pair_f32"]
        ctx.mark_non_differentiable(n_pair)
        return out["loss"], n_pair

    @staticmethod
    def backward(ctx, g_loss, _g_n):
        (dlogits,) = ctx.saved_tensors
        return ((g_loss * dlogits).reshape(ctx.out_shape).to(ctx.out_dtype),) + (None,) * 9


def pairwise_loss_with_focal(outputs, labels, groups, focal_weight=1.0, alpha=0.25, gamma=2.0,
                             stop_weight_gradient=False, pairloss_func=bpr_loss_func, return_num_pair=False,
                             click_occurance_power=0.0, mask=None, label_pair_to_weight_func=None, **kwargs):
    """``pairwise_loss(outputs, labels, groups, ...) + focal_weight * focal_crossentropy_loss(labels, outputs, alpha,
    gamma, stop_weight_gradient)`` in one fused call (the pairwise arguments as pairwise_loss_from_batch.py:228-236; the
    fused menu only: bpr_loss_func or a partial of it, a FusedPairWeight with a positive-side weight or none)."""
    if alpha and (alpha <= 0.0 or alpha >= 1.0):
        raise ValueError("Value of alpha should be greater than zero and less than one.")
    if gamma and gamma < 0:
        raise ValueError("Value of gamma should be greater than or equal to zero.")
    bpr = _match_bpr(pairloss_func)
    if bpr is None or bpr[2] is not None:
        raise NotImplementedError("the fused joint loss needs bpr_loss_func (or a functools.partial of it)")
    outputs, labels = _as_cuda(outputs), _as_cuda(labels)
    gl = [_as_cuda(g) for g in groups] if isinstance(groups, list) else [_as_cuda(groups)]
    keys, row_ok = ops.canon_keys(gl, None if mask is None else _as_cuda(mask).reshape(-1).to(torch.bool))
    label_func, rw_pos = "step", None
    if label_pair_to_weight_func is not None:
        if not isinstance(label_pair_to_weight_func, FusedPairWeight) or label_pair_to_weight_func.neg_kw is not None:
            raise NotImplementedError("the fused joint loss supports FusedPairWeight with a positive-side weight only")
        label_func = label_pair_to_weight_func.label_func
        if label_pair_to_weight_func.pos_kw is not None:
            rw_pos = _as_cuda(kwargs[label_pair_to_weight_func.pos_kw])
    focal = (float(focal_weight), float(alpha or 0.0), float(gamma or 0.0), bool(stop_weight_gradient))
    loss, n = _FusedPairwiseFocal.apply(outputs, labels, keys, row_ok, rw_pos, label_func, bpr[0], bpr[1],
                                        float(click_occurance_power), focal)
    return (loss, n) if return_num_pair else loss
