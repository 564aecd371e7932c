"""Op-for-op torch-CPU float32 restatement of the reference's dense pairwise / listwise graphs, with autograd
for the backward pass.  BASELINE INFRASTRUCTURE (see oracle/__init__.py): this is the timed
"reference restatement (torch CPU), not TensorFlow" of BASELINE.md section 4 -- TensorFlow is absent from the
image, so the reference's own modules cannot be imported.  It performs the same sequence of dense (B,B) /
(G,B) tensor ops as /root/reference/rec_now/rec_block/pairwise_loss_from_batch.py (PW:n) and
listwise_loss_from_batch.py (LW:n), multi-threaded like TF's Eigen CPU kernels.  Never imported by the product.
"""
from __future__ import annotations

import torch


def _col(v):
    return v.reshape(-1, 1)


def generate_pair_mask(groups):
    keys = groups if isinstance(groups, list) else [groups]
    out = None
    for g in keys:
        g = _col(g)
        diff = g - g.t()                                                   # PW:33
        same = (diff == 0.0).to(torch.float32)                             # PW:35
        m = (same - torch.eye(g.numel())).to(torch.bool)                   # PW:36-37
        out = m if out is None else torch.logical_and(out, m)              # PW:73
    return out


def vec_to_matrix_pair(v):
    v = _col(v)
    mat = v.repeat(1, v.numel())                                           # PW:90-92 (tile materialises)
    return mat, mat.t()


def bpr_loss_func(pos, neg, weights=None, factor=1.0, reduce_mean=True):
    x = pos - neg
    if factor != 1.0:
        x = x * factor
    losses = torch.clamp(x, min=0) - x + torch.log1p(torch.exp(-torch.abs(x)))   # PW:120-121
    if weights is not None:
        losses = losses * weights
    loss = losses.sum()
    if reduce_mean:
        loss = loss / (float(losses.numel()) + 1.0e-10)
    return loss


def occurance_power_weight(group_id, power):
    _, idx, count = torch.unique(group_id, return_inverse=True, return_counts=True)
    w = count.to(torch.float32)
    if power != 1.0:
        w = torch.pow(w, power)
    return w[idx]


def pairwise_loss(outputs, labels, groups, factor=1.0, only_wrong=False, power=0.0, mask=None,
                  weight_func=None, **kwargs):
    """PW:228-279 with pairloss_func = bpr_loss_func(factor).  Returns (loss, n_pair)."""
    pair_mask = generate_pair_mask(groups)                                 # PW:254
    if mask is not None:                                                   # PW:255
        m, mt = vec_to_matrix_pair(mask)
        pair_mask = pair_mask & (m & mt)
    s, st = vec_to_matrix_pair(outputs)                                    # PW:256
    y, yt = vec_to_matrix_pair(labels)                                     # PW:187
    if weight_func is None:
        cond, wmat = y > yt, None                                          # PW:189
    else:
        wmat = weight_func(y, yt, **kwargs)                                # PW:192
        cond = wmat > 0
    pair_mask = pair_mask & cond                                           # PW:259
    if only_wrong:
        pair_mask = pair_mask & (s < st)                                   # PW:200-202
    flat = pair_mask.reshape(-1).detach()                                  # PW:263-264
    weights = None if wmat is None else wmat.reshape(-1)[flat]             # PW:266
    if power != 0.0:                                                       # PW:285-290
        g = groups[0] if isinstance(groups, list) else groups
        gm, _ = vec_to_matrix_pair(g)
        occ = occurance_power_weight(gm.reshape(-1)[flat], power)
        weights = occ if weights is None else weights * occ
    if weights is not None:
        weights = weights.detach()                                         # PW:270
    pos = s.reshape(-1)[flat]                                              # PW:272
    neg = st.reshape(-1)[flat]                                             # PW:273
    return bpr_loss_func(pos, neg, weights, factor), int(pos.numel())


def pairwise_fwd_bwd(s, y, g, **kw):
    """One forward + backward step; returns (loss float, n_pair, grad tensor)."""
    s = s.detach().clone().requires_grad_(True)
    loss, n = pairwise_loss(s, y, g, **kw)
    loss.backward()
    return float(loss), n, s.grad


def listwise_fwd_bwd(g, y, s, th=0.5, masked=-1e9):
    """to_listwise_sample + listwise_loss_via_softmax_cross_entropy_with_logits (LW:89-173), fwd + bwd."""
    s = s.detach().clone().requires_grad_(True)
    uniq, idx = torch.unique(g.reshape(-1), return_inverse=True)           # LW:109 (order irrelevant to the loss)
    n, ng = idx.numel(), uniq.numel()
    cols = torch.arange(n)

    def gen_dense(values):                                                 # LW:123-129
        d = torch.zeros((ng, n), dtype=values.dtype)
        return d.index_put((idx, cols), values.reshape(-1))

    dense_mask = gen_dense(torch.ones(n, dtype=torch.bool))
    dense_labels = gen_dense(y.reshape(-1))
    dense_logits = gen_dense(s.reshape(-1))
    has_pos = (dense_labels > th).to(torch.int32).sum(-1) > 0              # LW:135
    has_neg = (gen_dense(y.reshape(-1) - th) < 0).to(torch.int32).sum(-1) > 0
    row_mask = has_pos & has_neg
    dense_logits = dense_logits + (1.0 - dense_mask.to(torch.float32)) * masked   # LW:139-140
    dense_labels = dense_labels[row_mask]
    dense_labels = (dense_labels / dense_labels.sum(-1, keepdim=True)).detach()   # LW:144, LW:147
    dense_logits = dense_logits[row_mask]
    per_list = -(dense_labels * torch.log_softmax(dense_logits, -1)).sum(-1)      # LW:167
    loss = per_list.mean() if per_list.numel() else s.sum() * 0.0
    loss.backward()
    return float(loss), int(row_mask.sum()), s.grad
