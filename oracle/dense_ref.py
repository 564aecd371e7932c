"""Dense float32 NumPy restatement of the reference's in-batch ranking losses.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Op-for-op: every function below
performs the same sequence of dense (B,B) / (G,B) operations, in float32, as the
TensorFlow function it cites, so masks, pair counts and the row-major pair
order are bit-exact restatements.  Citations use the SURVEY.md shorthand

    PW:n = /root/reference/rec_now/rec_block/pairwise_loss_from_batch.py:n
    LW:n = /root/reference/rec_now/rec_block/listwise_loss_from_batch.py:n

TensorFlow itself is a third-party, unpinned dependency of the reference and is
absent from this image; the numerics of the tf.nn ops are restated from their
documented formulas:
  * sigmoid_cross_entropy_with_logits(z, x) = max(x,0) - x*z + log1p(exp(-|x|))
  * softmax_cross_entropy_with_logits(p, z) = sum_i p_i*(log(sum exp(z-max)) - (z_i-max)),
    backprop = softmax(z) * sum(p) - p  (== softmax - p for normalised p)
  * unique_with_counts returns values in first-occurrence order.
Memory is Theta(B^2): keep B <= ~8192.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32
SMALL_POSITIVE_FLOAT = 1.0e-10  # PW:13


# ----------------------------------------------------------------------------
# helpers restating stock TF ops
# ----------------------------------------------------------------------------
def _col(v):
    """reshape(v, [-1, 1]) keeping dtype."""
    return np.asarray(v).reshape(-1, 1)


def unique_with_counts(x):
    """tf.unique_with_counts: (values in first-occurrence order, idx, counts)."""
    x = np.asarray(x).reshape(-1)
    # tf.unique hashes with operator==: NaN != NaN, so every NaN is its own value
    kw = dict(equal_nan=False) if x.dtype.kind == "f" else {}
    vals, first, inv, cnt = np.unique(x, return_index=True, return_inverse=True, return_counts=True, **kw)
    order = np.argsort(first, kind="stable")           # sorted-unique -> first-occurrence rank
    rank = np.empty_like(order)
    rank[order] = np.arange(order.size)
    return vals[order], rank[inv.reshape(-1)].astype(np.int32), cnt[order].astype(np.int32)


def sigmoid_cross_entropy_with_logits(labels, logits):
    """TF's numerically stable form, float32."""
    x = np.asarray(logits, F32)
    z = np.asarray(labels, F32)
    return (np.maximum(x, F32(0)) - x * z + np.log1p(np.exp(-np.abs(x)))).astype(F32)


# ----------------------------------------------------------------------------
# pairwise_loss_from_batch.py
# ----------------------------------------------------------------------------
def _generate_pair_mask(group, only_upper_band=False):
    """PW:16-40.  M = bool((g - g^T == 0.0) - eye(B))."""
    g = _col(group)
    n = g.size
    with np.errstate(invalid="ignore"):
        diff = g - g.T                                  # PW:33  (dtype of g)
    same = (diff == 0.0).astype(F32)                    # PW:35
    m = (same - np.eye(n, dtype=F32)).astype(bool)      # PW:36-37 (-1 -> True for NaN/inf ids)
    if only_upper_band:                                 # PW:38-39  band_part(., 0, 1)
        r, c = np.indices((n, n))
        m = m & (c - r >= 0) & (c - r <= 1)
    return m


def generate_pair_mask(group_tensor_or_list, only_upper_band=False):
    """PW:43-74.  AND of the per-key masks."""
    keys = group_tensor_or_list if isinstance(group_tensor_or_list, list) else [group_tensor_or_list]
    out = None
    for g in keys:
        one = _generate_pair_mask(g, only_upper_band)
        out = one if out is None else np.logical_and(out, one)
    return out


def vec_to_matrix_pair(vec):
    """PW:77-93.  mat[i,j] = v_i ; mat_T[i,j] = v_j."""
    v = _col(vec)
    mat = np.tile(v, (1, v.size))
    return mat, mat.T


def bpr_loss_func(outputs_pos, outputs_neg, weights=None, factor=1.0, reduce_mean=True):
    """PW:96-127."""
    logits = (np.asarray(outputs_pos, F32) - np.asarray(outputs_neg, F32)).astype(F32)
    if factor != 1.0:
        logits = (logits * F32(factor)).astype(F32)
    losses = sigmoid_cross_entropy_with_logits(np.ones_like(logits), logits)
    if weights is not None:
        losses = (losses * np.asarray(weights, F32)).astype(F32)
    loss = np.sum(losses, dtype=F32)
    if reduce_mean:
        loss = F32(loss / (F32(losses.size) + F32(SMALL_POSITIVE_FLOAT)))
    return F32(loss)


def hinge_loss_func(outputs_pos, outputs_neg, weights=None, margin=1.0, factor=1.0, reduce_mean=True):
    """Margin ranking loss behind the reference's pairloss_func hook (PW:229, called as PW:274: (pos, neg, weights)):
    max(0, margin - (pos - neg) * factor), weighted and reduced exactly as bpr_loss_func (PW:117-126).  The reference
    ships only bpr_loss_func; this is SURVEY 8f N2's first "other pair loss"."""
    logits = (np.asarray(outputs_pos, F32) - np.asarray(outputs_neg, F32)).astype(F32)
    if factor != 1.0:
        logits = (logits * F32(factor)).astype(F32)
    losses = np.maximum(F32(margin) - logits, F32(0)).astype(F32)
    if weights is not None:
        losses = (losses * np.asarray(weights, F32)).astype(F32)
    loss = np.sum(losses, dtype=F32)
    if reduce_mean:
        loss = F32(loss / (F32(losses.size) + F32(SMALL_POSITIVE_FLOAT)))
    return F32(loss)


def occurance_power_weight(group_id, power=0.0):
    """PW:130-151."""
    _, idx, count = unique_with_counts(group_id)
    w = count.astype(F32)
    if power != 1.0:
        w = np.power(w, F32(power)).astype(F32)
    return w[idx]


def _apply_sample_mask(pair_mask, mask):
    """PW:154-172."""
    if mask is None:
        return pair_mask
    m, mt = vec_to_matrix_pair(np.asarray(mask, bool))
    return np.logical_and(pair_mask, np.logical_and(m, mt))


def _calc_label_cond_and_weights(labels, label_pair_to_weight_func, **kwargs):
    """PW:175-194."""
    y, yt = vec_to_matrix_pair(np.asarray(labels, F32))
    if label_pair_to_weight_func is None:
        return y > yt, None
    w = np.asarray(label_pair_to_weight_func(y, yt, **kwargs), F32)
    return w > 0, w


def _apply_pair_mask(mat, flat_mask):
    """PW:206-217: boolean_mask(reshape(mat,[-1]), mask) -> row-major pair order."""
    if mat is None:
        return None
    return np.asarray(mat).reshape(-1)[flat_mask]


def _merge_weights_by_mul(w1, w2):
    """PW:220-225."""
    if w1 is None:
        return w2
    if w2 is None:
        return w1
    return (w1 * w2).astype(F32)


def _final_pair_mask(outputs, labels, groups, only_use_wrong_order_pair, mask,
                     label_pair_to_weight_func, **kwargs):
    """PW:254-264 up to the flattened, final pair mask (+ the dense weight matrix)."""
    pair_mask = generate_pair_mask(groups)                                   # PW:254
    pair_mask = _apply_sample_mask(pair_mask, mask)                          # PW:255
    s, st = vec_to_matrix_pair(np.asarray(outputs, F32))                     # PW:256
    cond, wmat = _calc_label_cond_and_weights(labels, label_pair_to_weight_func, **kwargs)  # PW:257
    pair_mask = np.logical_and(pair_mask, cond)                              # PW:259
    if only_use_wrong_order_pair:                                            # PW:197-203
        pair_mask = np.logical_and(pair_mask, s < st)
    return pair_mask.reshape(-1), wmat, s, st


def pairwise_loss(outputs, labels, groups,
                  pairloss_func=bpr_loss_func,
                  only_use_wrong_order_pair=False,
                  return_num_pair=False,
                  click_occurance_power=0.0,
                  mask=None,
                  label_pair_to_weight_func=None,
                  **kwargs):
    """PW:228-279, same positional order / defaults / return arity."""
    flat, wmat, s, st = _final_pair_mask(outputs, labels, groups, only_use_wrong_order_pair,
                                         mask, label_pair_to_weight_func, **kwargs)
    weights = _apply_pair_mask(wmat, flat)                                   # PW:266
    weights = _apply_occurance_weights(groups, click_occurance_power, flat, weights)  # PW:267
    pos = _apply_pair_mask(s, flat)                                          # PW:272
    neg = _apply_pair_mask(st, flat)                                         # PW:273
    loss = pairloss_func(pos, neg, weights)                                  # PW:274
    if return_num_pair:
        return loss, F32(pos.size)                                           # PW:276
    return loss


def _apply_occurance_weights(groups, power, flat_mask, weights):
    """PW:282-291."""
    if power != 0.0:
        g = groups[0] if isinstance(groups, list) else groups
        gm, _ = vec_to_matrix_pair(g)
        gpos = _apply_pair_mask(gm, flat_mask)
        weights = _merge_weights_by_mul(weights, occurance_power_weight(gpos, power=power))
    return weights


def pairwise_full(outputs, labels, groups, factor=1.0, reduce_mean=True,
                  only_use_wrong_order_pair=False, click_occurance_power=0.0,
                  mask=None, label_pair_to_weight_func=None, **kwargs):
    """Everything the parity tests compare, for pairloss_func = bpr_loss_func(factor, reduce_mean).

    Returns dict(loss f32, n_pair int, pos_idx, neg_idx (row-major order, int64),
    weights f32[P] or None, grad f32[B] = d loss / d outputs  (what TF autodiff yields:
    only through outputs_pos / outputs_neg, PW:264, PW:270)).
    """
    flat, wmat, s, st = _final_pair_mask(outputs, labels, groups, only_use_wrong_order_pair,
                                         mask, label_pair_to_weight_func, **kwargs)
    b = s.shape[0]
    idx = np.flatnonzero(flat)
    pos_idx, neg_idx = idx // b, idx % b
    weights = _apply_pair_mask(wmat, flat)
    weights = _apply_occurance_weights(groups, click_occurance_power, flat, weights)
    pos, neg = _apply_pair_mask(s, flat), _apply_pair_mask(st, flat)
    loss = bpr_loss_func(pos, neg, weights, factor=factor, reduce_mean=reduce_mean)
    # analytic backward of bpr_loss_func (float32, as TF's kernels would run it)
    x = (pos - neg).astype(F32)
    if factor != 1.0:
        x = (x * F32(factor)).astype(F32)
    with np.errstate(over="ignore"):
        sig_neg = (F32(1) / (F32(1) + np.exp(x))).astype(F32)     # sigma(-x) = -d l/dx
    d = sig_neg if weights is None else (sig_neg * weights).astype(F32)
    denom = F32(F32(pos.size) + F32(SMALL_POSITIVE_FLOAT)) if reduce_mean else F32(1)
    d = (d * F32(factor) / denom).astype(F32)
    grad = np.zeros(b, np.float64)
    np.add.at(grad, pos_idx, -d.astype(np.float64))
    np.add.at(grad, neg_idx, d.astype(np.float64))
    return dict(loss=F32(loss), n_pair=int(pos.size), pos_idx=pos_idx.astype(np.int64),
                neg_idx=neg_idx.astype(np.int64), weights=weights, grad=grad.astype(F32))


# ----------------------------------------------------------------------------
# listwise_loss_from_batch.py
# ----------------------------------------------------------------------------
def row_not_all_zero(x):
    """LW:13-31."""
    x = np.asarray(x, F32)
    return np.sum((x != 0.0).astype(np.int32), axis=-1) > 0


def row_has_value_greater_than(x, threshold):
    """LW:34-53."""
    x = np.asarray(x, F32)
    return np.sum((x > F32(threshold)).astype(np.int32), axis=-1) > 0


def row_has_value_less_than(x, threshold):
    """LW:56-71."""
    x = np.asarray(x, F32)
    return np.sum((x < F32(threshold)).astype(np.int32), axis=-1) > 0


def nan_to_zero(val):
    """LW:74-86 (scalar only)."""
    if np.ndim(val) != 0:
        raise ValueError("input muust be a scalar tf.Tensor")
    return type(val)(0.0) if np.isnan(val) else val


def to_listwise_sample(group_ids, labels, logits, do_mask_logits=True,
                       value_of_masked_logit=-1e9, pos_neg_th=0.5):
    """LW:89-148.  Returns (dense_mask (V,B) bool, dense_labels (V,B) f32, dense_logits (V,B) f32)."""
    g = np.asarray(group_ids).reshape(-1)                                   # LW:107-108
    y, idx, _ = unique_with_counts(g)                                       # LW:109
    n = idx.size
    cols = np.arange(n)

    def gen_dense(values):                                                  # LW:123-129
        d = np.zeros((y.size, n), dtype=np.asarray(values).dtype)
        d[idx, cols] = np.asarray(values).reshape(-1)
        return d

    labels = np.asarray(labels, F32).reshape(-1)
    logits = np.asarray(logits, F32).reshape(-1)
    dense_mask = gen_dense((cols.astype(F32) + 1) > 0)                      # LW:122, LW:131
    dense_labels = gen_dense(labels)                                        # LW:132
    dense_logits = gen_dense(logits)                                        # LW:133
    has_pos = row_has_value_greater_than(dense_labels, pos_neg_th)          # LW:135
    has_neg = row_has_value_less_than(gen_dense((labels - F32(pos_neg_th)).astype(F32)), 0.0)  # LW:136
    row_mask = np.logical_and(has_pos, has_neg)                             # LW:137
    if do_mask_logits:                                                      # LW:139-140
        dense_logits = (dense_logits
                        + (F32(1.0) - dense_mask.astype(F32)) * F32(value_of_masked_logit)).astype(F32)
    dense_mask = dense_mask[row_mask]                                       # LW:142
    dense_labels = dense_labels[row_mask]                                   # LW:143
    with np.errstate(invalid="ignore", divide="ignore"):
        dense_labels = (dense_labels / np.sum(dense_labels, axis=-1, keepdims=True, dtype=F32)).astype(F32)  # LW:144
    dense_logits = dense_logits[row_mask]                                   # LW:145
    return dense_mask, dense_labels, dense_logits


def softmax_cross_entropy_with_logits(labels, logits):
    """tf.nn.softmax_cross_entropy_with_logits, float32, rows = lists."""
    z = np.asarray(logits, F32)
    p = np.asarray(labels, F32)
    if z.shape[0] == 0:
        return np.zeros((0,), F32)
    zs = (z - np.max(z, axis=-1, keepdims=True)).astype(F32)
    lse = np.log(np.sum(np.exp(zs), axis=-1, keepdims=True, dtype=F32)).astype(F32)
    return np.sum(p * (lse - zs), axis=-1, dtype=F32).astype(F32)


def listwise_loss_via_softmax_cross_entropy_with_logits(labels_for_softmax, logits_for_softmax,
                                                        weights=None, do_reduce=True):
    """LW:151-173."""
    loss = softmax_cross_entropy_with_logits(labels_for_softmax, logits_for_softmax)   # LW:167
    if weights is not None:
        loss = (loss * np.asarray(weights, F32)).astype(F32)                           # LW:168-169
    if do_reduce:
        with np.errstate(invalid="ignore"), np.testing.suppress_warnings() as sup:
            sup.filter(RuntimeWarning)
            loss = F32(np.mean(loss, dtype=F32)) if loss.size else F32(np.nan)          # LW:171 (mean of empty = NaN)
        loss = nan_to_zero(loss)                                                        # LW:172
    return loss


def listwise_full(group_ids, labels, logits, weights=None, do_mask_logits=True,
                  value_of_masked_logit=-1e9, pos_neg_th=0.5):
    """Composed use (tests/rec_block/test_listwise_loss_from_batch.py:26-31) + analytic backward.

    Returns dict(loss f32, n_valid int, list_loss f32[V] (first-occurrence order),
    valid_group_values, grad f32[B] = d loss / d logits).
    """
    g = np.asarray(group_ids).reshape(-1)
    dm, dl, dz = to_listwise_sample(g, labels, logits, do_mask_logits, value_of_masked_logit, pos_neg_th)
    per_list = listwise_loss_via_softmax_cross_entropy_with_logits(dl, dz, weights, do_reduce=False)
    loss = listwise_loss_via_softmax_cross_entropy_with_logits(dl, dz, weights, do_reduce=True)
    v = dz.shape[0]
    grad = np.zeros(g.size, F32)
    if v:
        zs = dz - dz.max(axis=-1, keepdims=True)
        e = np.exp(zs.astype(np.float64))
        sm = e / e.sum(axis=-1, keepdims=True)
        w = np.ones(v) if weights is None else np.asarray(weights, np.float64).reshape(-1)
        # TF xent backprop: softmax - labels (labels rows sum to 1); only member columns reach `logits`
        gd = (sm * dl.astype(np.float64).sum(axis=-1, keepdims=True) - dl) * (w / v)[:, None]
        grad = np.where(dm, gd, 0.0).sum(axis=0).astype(F32)
    return dict(loss=F32(loss), n_valid=int(v), list_loss=per_list, dense_mask=dm, grad=grad)


# ---- focal loss (rec_now/rec_block/focal_loss.py; fused by the product into the pairwise call as an extra term) ----------
def focal_crossentropy_loss(labels, logits, alpha=0.25, gamma=2.0, stop_weight_gradient=False, return_mean=True):
    """focal_loss.py:12-66, op for op in float32 (stop_weight_gradient only matters for the gradient)."""
    if alpha and (alpha <= 0.0 or alpha >= 1.0):
        raise ValueError("Value of alpha should be greater than zero and less than one.")           # focal_loss.py:43-44
    if gamma and gamma < 0:
        raise ValueError("Value of gamma should be greater than or equal to zero.")                 # focal_loss.py:45-46
    y = np.asarray(labels, F32)
    z = np.asarray(logits, F32)
    loss = sigmoid_cross_entropy_with_logits(y, z)                                                  # focal_loss.py:48
    if alpha:
        a = F32(alpha)
        loss = (y * a + (F32(1) - y) * (F32(1) - a)) * loss                                         # focal_loss.py:50-53
    if gamma:
        p = (F32(1) / (F32(1) + np.exp(-z))).astype(F32)                                            # focal_loss.py:56
        sim = y * p + (F32(1) - y) * (F32(1) - p)                                                   # focal_loss.py:57
        loss = np.power(F32(1) - sim, F32(gamma)).astype(F32) * loss                                # focal_loss.py:59-62
    return F32(np.mean(loss, dtype=F32)) if return_mean else loss                                   # focal_loss.py:64-66
