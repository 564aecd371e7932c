"""GAUC (group AUC) on the GPU: the evaluation side of the in-batch ranking losses.

The reference quotes the online GAUC uplift of its in-batch pairwise loss (README.md:5, 8) but ships no implementation
of the metric; the definition used here is written in include/recnow_b200.h (rn_gauc) and restated in float64 by
oracle/seg_ref.py::gauc: per group AUC over the label-ordered pairs (ties in the scores count one half), weighted by the
group's rows.  Same segmentation kernel as pairwise_loss, then one pair-counting kernel; exact integer counts.
There is no CPU path.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib, ops


def gauc(outputs, labels, groups, mask=None, return_details: bool = False):
    """GAUC of ``outputs`` (scores) against ``labels`` inside the groups given by ``groups`` (one id tensor or a list of
    them: composite key, as pairwise_loss).  Returns a 0-d float32 tensor, or a dict with ``gauc``, ``auc_mean``
    (unweighted mean of the per-group AUCs), ``n_valid_groups``, ``n_pair`` and ``concordant2`` (2 x concordant + ties)."""
    from .rec_block.pairwise_loss_from_batch import _as_cuda
    s = ops._f32(_as_cuda(outputs).detach())
    y = ops._f32(_as_cuda(labels).detach())
    gl = [_as_cuda(g) for g in groups] if isinstance(groups, (list, tuple)) else [_as_cuda(groups)]
    keys, row_ok = ops.canon_keys(gl, None if mask is None else _as_cuda(mask).reshape(-1).to(torch.bool))
    b = s.numel()
    if keys.dim() != 2 or not keys.is_contiguous():
        keys = keys.reshape(-1, b).contiguous()
    kk = keys.shape[0]
    dev = s.device
    out_f = torch.empty(2, dtype=torch.float32, device=dev)            # gauc, auc_mean
    out_i = torch.empty(6, dtype=torch.int32, device=dev)              # n_valid (i32) | pad | n_pair (i64) | conc2 (i64)
    lib = _lib.lib()
    nbytes = lib.rn_gauc_scratch_bytes(b, kk)
    st = torch.cuda.current_stream(dev).cuda_stream
    scratch = ops._scratch(nbytes, dev, st, ("gauc", kk))
    a = _lib.GaucArgs(B=b, K=kk, scratch_persistent=1, keys=keys.data_ptr(), scores=s.data_ptr(), labels=y.data_ptr(),
                      row_ok=ops._ptr(row_ok), gauc=out_f.data_ptr(), auc_mean=out_f.data_ptr() + 4,
                      n_valid_groups=out_i.data_ptr(), n_pair=out_i.data_ptr() + 8, concordant2=out_i.data_ptr() + 16)
    with ops._on_device(dev):
        _lib.check(lib.rn_gauc(C.byref(a), scratch.data_ptr(), nbytes, C.c_void_p(st)), "rn_gauc")
    if not return_details:
        return out_f[0]
    wide = out_i[2:6].view(torch.int64)
    return dict(gauc=out_f[0], auc_mean=out_f[1], n_valid_groups=out_i[0], n_pair=wide[0], concordant2=wide[1],
                _scratch=scratch, _keep=(s, y, keys, row_ok))
