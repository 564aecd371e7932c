"""Drop-in for the slot-pooling part of rec_now/rec_block/embedding_util.py of the reference, on torch CUDA tensors.

Same names, argument order and defaults as the reference (/root/reference/rec_now/rec_block/embedding_util.py, cited
as EU:n): ``sparse_batch_segment_ids_of_targets`` (EU:127-198) and ``embedding_using_sparse_batch_segment_ids``
(EU:254-324).  The reference chains a hash-table lookup, boolean_mask, unique, gather, a weight multiply and
unsorted_segment_sum / _mean; here the gather, the multiply and the segment reduction are ONE kernel of
librecnow_b200.so (rn_segment_pool_fwd / _bwd) and nothing of size [kept, D] is materialised.

``embedding_func``:
  * a :class:`TableLookup` (a [V, D] float32 table, e.g. ``nn.Embedding.weight``): fully fused, the gradient reaches the
    table through rn_segment_pool_bwd;
  * any other callable: it is called ONCE on the unique target ids, exactly as in the reference (EU:305-309), and its
    output is pooled by the same kernel (autograd continues through the callable).
There is no CPU path: CPU tensors raise.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from .. import _lib
from ..ops import _need_cuda, _on_device, _stream


class TableLookup:
    """``embedding_func`` that is a plain table lookup: ids -> table[ids]  (table: float32 [V, D] on the GPU)."""

    def __init__(self, table: torch.Tensor):
        self.table = table

    def __call__(self, ids):
        return torch.nn.functional.embedding(ids.long(), self.table)


_target_cache: dict = {}


def _targets(target_slots, device) -> torch.Tensor:
    """The target slots as an int32 device tensor; cached per (slots, device) -- building it is a synchronous
    host-to-device copy, a quarter of the pooling kernel's own time."""
    if not isinstance(target_slots, list):
        target_slots = list(target_slots)
    key = (tuple(target_slots), str(device))
    t = _target_cache.get(key)
    if t is None:
        if len(set(target_slots)) != len(target_slots):
            raise ValueError("target_slots must be distinct (the reference's StaticHashTable rejects duplicate keys)")
        if len(_target_cache) > 256:
            _target_cache.clear()
        t = _target_cache[key] = torch.tensor(target_slots, dtype=torch.int32, device=device)
    return t


def sparse_batch_segment_ids_of_targets(slots, target_slots):
    """EU:127-198.  Returns (mask, sp_segment_ids, num_rows, num_ids, num_segments); compat helper -- the pooling below
    never builds these."""
    slots = torch.as_tensor(slots)
    _need_cuda(slots)
    tgt = _targets(target_slots, slots.device).to(slots.dtype)
    hit = slots.unsqueeze(-1) == tgt                                  # [B, C, T]
    mask = hit.any(-1)
    seg = hit.to(torch.int32).argmax(-1).to(torch.int32)
    num_rows, num_ids = slots.shape[0], tgt.numel()
    rows = torch.nonzero(mask)[:, 0].to(torch.int32)
    return mask, rows * num_ids + seg[mask], num_rows, num_ids, num_rows * num_ids


def _pool_args(slots, ids, weights, tgt, table, mean) -> _lib.PoolArgs:
    b, c = slots.shape
    return _lib.PoolArgs(B=b, C=c, T=tgt.numel(), D=table.shape[1], mean=1 if mean else 0, slots=slots.data_ptr(),
                         ids=ids.data_ptr(), weights=None if weights is None else weights.data_ptr(),
                         target_slots=tgt.data_ptr(), table=table.data_ptr(), V=table.shape[0])


_err_flags: dict = {}


def out_of_range_ids_seen(device=None) -> bool:
    """True if any pooling call on `device` met an id outside [0, V) (such ids contribute nothing; tf.gather would have
    raised).  Reads one word back (synchronises)."""
    idx = torch.cuda.current_device() if device is None else torch.device(device).index
    f = _err_flags.get(idx)
    return bool(f is not None and int(f.item()) != 0)


class _SegmentPool(torch.autograd.Function):
    @staticmethod
    def forward(ctx, table, weights, slots, ids, tgt, mean):
        b = slots.shape[0]
        out = torch.empty((b, tgt.numel(), table.shape[1]), dtype=torch.float32, device=table.device)
        err = _err_flags.get(table.device.index)       # (one flag word per device, set by any call that saw an id outside the table)
        if err is None:
            err = _err_flags[table.device.index] = torch.zeros(1, dtype=torch.int32, device=table.device)
        a = _pool_args(slots, ids, weights, tgt, table, mean)
        with _on_device(table.device):
            _lib.check(_lib.lib().rn_segment_pool_fwd(C.byref(a), out.data_ptr(), err.data_ptr(), _stream(table.device)),
                       "rn_segment_pool_fwd")
        ctx.save_for_backward(table, weights if weights is not None else torch.empty(0, device=table.device), slots, ids, tgt)
        ctx.mean, ctx.has_w = mean, weights is not None
        return out

    @staticmethod
    def backward(ctx, d_out):
        table, weights, slots, ids, tgt = ctx.saved_tensors
        weights = weights if ctx.has_w else None
        need_t, need_w = ctx.needs_input_grad[0], ctx.has_w and ctx.needs_input_grad[1]
        if not (need_t or need_w):
            return (None,) * 6
        d_out = d_out.to(torch.float32).contiguous()
        d_table = torch.zeros_like(table) if need_t else None
        d_w = torch.zeros_like(weights) if need_w else None
        a = _pool_args(slots, ids, weights, tgt, table, ctx.mean)
        with _on_device(table.device):
            _lib.check(_lib.lib().rn_segment_pool_bwd(C.byref(a), d_out.data_ptr(),
                                                      None if d_table is None else d_table.data_ptr(),
                                                      None if d_w is None else d_w.data_ptr(), _stream(table.device)),
                       "rn_segment_pool_bwd")
        return d_table, d_w, None, None, None, None


def segment_pool(table, slots, target_slots, ids, weights=None, method="sum"):
    """out[b, t, :] = sum (or mean) over the columns c of row b with slots[b, c] == target_slots[t] of
    weights[b, c] * table[ids[b, c]] -- one kernel (rn_segment_pool_fwd); differentiable in table and weights."""
    if method not in ("sum", "mean"):
        raise ValueError("method must be 'sum' or 'mean'")
    _need_cuda(table, slots, ids, weights)
    if table.dim() != 2:
        raise ValueError("table must be [V, D]")
    table = table if (table.dtype is torch.float32 and table.is_contiguous()) else table.to(torch.float32).contiguous()
    slots = torch.as_tensor(slots)
    if slots.dim() != 2 or ids.shape != slots.shape or (weights is not None and weights.shape != slots.shape):
        raise ValueError("slots, ids (and weights) must share one [B, C] shape")
    slots_i = slots.to(torch.int32).contiguous()
    ids_l = ids.to(torch.int64).contiguous()
    w = None if weights is None else weights.to(torch.float32).contiguous()
    tgt = _targets(target_slots, table.device)
    return _SegmentPool.apply(table, w, slots_i, ids_l, tgt, method == "mean")


def embedding_using_sparse_batch_segment_ids(embedding_func, slots, target_slots, ids, weights=None, method="sum",
                                             use_unique=True):
    """EU:254-324.  Returns pooled_embedding [B, T, D]."""
    slots = torch.as_tensor(slots)
    ids = torch.as_tensor(ids)
    _need_cuda(slots, ids, weights)
    if isinstance(embedding_func, TableLookup):
        return segment_pool(embedding_func.table, slots, target_slots, ids, weights, method)
    # an arbitrary callable: called once, on the (unique) ids of the target slots only, as in the reference (EU:303-312)
    tgt = _targets(target_slots, slots.device).to(slots.dtype)
    mask = (slots.unsqueeze(-1) == tgt).any(-1)
    sp_ids = ids[mask]
    if use_unique:
        uniq, inv = torch.unique(sp_ids, return_inverse=True)       # (sorted, not first-occurrence, order: the callable
        emb = embedding_func(uniq)                                  #  sees the same SET of ids; rows are gathered back)
    else:
        inv = torch.arange(sp_ids.numel(), device=ids.device)
        emb = embedding_func(sp_ids)
    index = torch.zeros(ids.shape, dtype=torch.int64, device=ids.device)
    index[mask] = inv
    if emb.shape[0] == 0:                                           # no target slot anywhere: all segments empty
        return torch.zeros((slots.shape[0], tgt.numel(), emb.shape[-1] if emb.dim() == 2 else 1), device=ids.device)
    return segment_pool(emb, slots, target_slots, index, weights, method)
