"""Drop-in for rec_now/rec_block/listwise_loss_from_batch.py of the reference, on torch CUDA tensors.

Same public names, argument order, defaults and return arity as the reference
(/root/reference/rec_now/rec_block/listwise_loss_from_batch.py, cited as LW:n).

``to_listwise_sample`` returns three *lazy* dense views.  Passed on to
``listwise_loss_via_softmax_cross_entropy_with_logits`` (the composed use of the reference's tests,
tests/rec_block/test_listwise_loss_from_batch.py:26-31) they never materialise: the loss runs as one
segmented log-sum-exp kernel over the batch (rn_listwise_fwd_bwd).  Touched as tensors (``.shape``, torch
functions, indexing) they materialise the reference's (V,B) layout through rn_listwise_dense.
There is no CPU path.
"""
from __future__ import annotations

import torch

from .. import ops
from .pairwise_loss_from_batch import _as_cuda


def _f32(x):
    x = _as_cuda(x)
    return x if x.dtype == torch.float32 else x.to(torch.float32)


def row_not_all_zero(x):
    """LW:13-31."""
    return torch.sum((_f32(x) != 0.0).to(torch.int32), dim=-1) > 0


def row_has_value_greater_than(x, threshold):
    """LW:34-53."""
    return torch.sum((_f32(x) > threshold).to(torch.int32), dim=-1) > 0


def row_has_value_less_than(x, threshold):
    """LW:56-71."""
    return torch.sum((_f32(x) < threshold).to(torch.int32), dim=-1) > 0


def nan_to_zero(val):
    """LW:74-86 (scalar only)."""
    val = val.dense() if isinstance(val, LazyDense) else val
    if len(val.shape) != 0:
        raise ValueError('input muust be a scalar tf.Tensor')
    return torch.where(torch.isnan(val), torch.zeros_like(val), val)


class _ListwiseBatch:
    """One to_listwise_sample call: inputs, options and the cached kernel results."""

    def __init__(self, group_ids, labels, logits, do_mask_logits, value_of_masked_logit, pos_neg_th):
        self.logits_in = _as_cuda(logits)
        self.labels_in = _as_cuda(labels)
        g = _as_cuda(group_ids)
        self.keys, self.row_ok = ops.canon_keys(g.reshape(-1), inf_is_id=True)
        self.do_mask_logits, self.value_of_masked_logit = bool(do_mask_logits), float(value_of_masked_logit)
        self.th = float(pos_neg_th)
        self._fwd = {}
        self._dense = None

    def fwd(self, weights=None, do_reduce=True, sorted_form=False, temperature=1.0):
        key = (None if weights is None else weights.data_ptr(), bool(do_reduce), bool(sorted_form), float(temperature))
        if key not in self._fwd:
            self._fwd[key] = ops.listwise_fwd_bwd(self.keys[0], self.labels_in, self.logits_in, row_ok=self.row_ok,
                                                  list_w=weights, pos_neg_th=self.th, do_reduce=do_reduce,
                                                  sorted_form=sorted_form, temperature=temperature)
        return self._fwd[key]

    def n_valid(self) -> int:
        f = next(iter(self._fwd.values())) if self._fwd else self.fwd()
        return int(f["n_valid"].item())                   # synchronises (V is data dependent)

    def dense(self):
        if self._dense is None:
            f = self.fwd(sorted_form=True)                # (the dense layout is filled from the sorted form's arena)
            v = int(f["n_valid"].item())
            dm, dl, dz = ops.listwise_dense(f, v, self.do_mask_logits, self.value_of_masked_logit)
            dz = _DenseLogits.apply(self.logits_in, dz, dm)
            self._dense = (dm, dl, dz)
        return self._dense


class _DenseLogits(torch.autograd.Function):
    """Identity on the materialised dense logits that routes d/d dense_logits back to the batch logits
    (only member columns depend on them, LW:133, LW:139-140)."""

    @staticmethod
    def forward(ctx, logits, dense_logits, dense_mask):
        ctx.save_for_backward(dense_mask)
        ctx.in_shape, ctx.in_dtype = logits.shape, logits.dtype
        return dense_logits.view_as(dense_logits)

    @staticmethod
    def backward(ctx, g):
        (dm,) = ctx.saved_tensors
        return (g * dm).sum(0).reshape(ctx.in_shape).to(ctx.in_dtype), None, None


class LazyDense:
    """Tensor-like handle on one of to_listwise_sample's three outputs (which: 0 mask, 1 labels, 2 logits)."""

    def __init__(self, batch: _ListwiseBatch, which: int):
        self._batch, self._which = batch, which

    def dense(self) -> torch.Tensor:
        return self._batch.dense()[self._which]

    @property
    def shape(self):
        return torch.Size((self._batch.n_valid(), self._batch.logits_in.numel()))

    @property
    def dtype(self):
        return torch.bool if self._which == 0 else torch.float32

    def __getattr__(self, name):                  # everything else: behave like the dense tensor
        return getattr(self.dense(), name)

    def __getitem__(self, idx):
        return self.dense()[idx]

    def __len__(self):
        return self.shape[0]

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        unwrap = lambda a: a.dense() if isinstance(a, LazyDense) else a
        args = tuple(unwrap(a) for a in args)
        kwargs = {k: unwrap(v) for k, v in (kwargs or {}).items()}
        return func(*args, **kwargs)


def to_listwise_sample(group_ids, labels, logits, do_mask_logits=True, value_of_masked_logit=-1E9, pos_neg_th=0.5):
    """LW:89-148.  Returns (dense_mask, dense_labels, dense_logits) as lazy (V,B) views."""
    batch = _ListwiseBatch(group_ids, labels, logits, do_mask_logits, value_of_masked_logit, pos_neg_th)
    return LazyDense(batch, 0), LazyDense(batch, 1), LazyDense(batch, 2)


class _FusedListwiseLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, batch, weights, do_reduce, temperature=1.0):
        out = batch.fwd(weights, do_reduce, temperature=temperature)
        ctx.in_shape, ctx.in_dtype, ctx.do_reduce = logits.shape, logits.dtype, do_reduce
        if do_reduce:
            ctx.save_for_backward(out["dlogits"])
            return out["loss"]
        v = int(out["n_valid"].item())
        ctx.batch, ctx.weights = batch, weights
        ctx.save_for_backward(out["dlogits"])
        return out["list_loss"][:v].clone()

    @staticmethod
    def backward(ctx, g):
        (dlogits,) = ctx.saved_tensors
        if ctx.do_reduce:
            return (g * dlogits).reshape(ctx.in_shape).to(ctx.in_dtype), None, None, None, None
        # per-list upstream gradients: scale each member's gradient by its list's g (via the dense mask)
        dm = ctx.batch.dense()[0]
        per_row = (dm.to(g.dtype) * g.reshape(-1, 1)).sum(0)
        return (per_row * dlogits).reshape(ctx.in_shape).to(ctx.in_dtype), None, None, None, None


def listwise_loss_via_softmax_cross_entropy_with_logits(labels_for_softmax,
                                                        logits_for_softmax,
                                                        weights=None,
                                                        do_reduce=True):
    """LW:151-173."""
    lazy = (isinstance(labels_for_softmax, LazyDense) and isinstance(logits_for_softmax, LazyDense)
            and labels_for_softmax._batch is logits_for_softmax._batch
            and labels_for_softmax._which == 1 and logits_for_softmax._which == 2)
    # (the segmented kernels drop the non-member columns, which is what the reference computes when their logit is
    # value_of_masked_logit = -1e9: exp underflows to 0; a mild value such as -10 keeps (B - n) exp(value) in every
    # denominator, LW:139-140, 167 -- that case takes the dense formula below)
    if (lazy and logits_for_softmax._batch.do_mask_logits and logits_for_softmax._batch.th >= 0
            and logits_for_softmax._batch.value_of_masked_logit <= -1.0e4):
        batch = logits_for_softmax._batch
        w = None if weights is None else _f32(weights).reshape(-1).contiguous()
        return _FusedListwiseLoss.apply(batch.logits_in, batch, w, bool(do_reduce))
    # explicit dense tensors (or options the segmented form does not cover): the reference's dense formula
    labels_d = labels_for_softmax.dense() if isinstance(labels_for_softmax, LazyDense) else _as_cuda(labels_for_softmax)
    logits_d = logits_for_softmax.dense() if isinstance(logits_for_softmax, LazyDense) else _as_cuda(logits_for_softmax)
    labels_d = labels_d.detach()                                                      # LW:166
    listwise_loss = -(labels_d * torch.log_softmax(logits_d, dim=-1)).sum(-1)         # LW:167
    if weights is not None:
        listwise_loss = listwise_loss * _as_cuda(weights)                             # LW:168-169
    if do_reduce:
        listwise_loss = torch.mean(listwise_loss)                                     # LW:171 (empty -> NaN)
        listwise_loss = nan_to_zero(listwise_loss)                                    # LW:172
    return listwise_loss


def listwise_loss_from_batch(group_ids, labels, logits, weights=None, do_reduce=True, pos_neg_th=0.5, temperature=1.0):
    """``listwise_loss_via_softmax_cross_entropy_with_logits(*to_listwise_sample(group_ids, labels, logits)[1:], weights,
    do_reduce)`` (LW:89-173) as ONE call that never builds the (V, B) tensors, with an optional softmax temperature
    (the loss is taken on ``logits / temperature``; the gradient reaches the unscaled logits).  Returns
    ``(loss, n_valid_lists)`` -- n_valid_lists as a device tensor (no synchronisation)."""
    if not temperature > 0:
        raise ValueError("temperature must be positive")
    batch = _ListwiseBatch(group_ids, labels, logits, True, -1E9, pos_neg_th)
    w = None if weights is None else _f32(weights).reshape(-1).contiguous()
    loss = _FusedListwiseLoss.apply(batch.logits_in, batch, w, bool(do_reduce), float(temperature))
    return loss, batch.fwd(w, bool(do_reduce), temperature=float(temperature))["n_valid"]
