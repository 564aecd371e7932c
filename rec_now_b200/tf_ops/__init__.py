"""TensorFlow 2 front end: the reference's public functions on top of the custom ops of recnow_tf_ops.so.

STATUS: untested -- TensorFlow is not installed in the development image.  Importing this module needs
TensorFlow and a built recnow_tf_ops.so (python -m rec_now_b200.tf_ops.build); without them it raises.  The
functions mirror rec_now/rec_block/pairwise_loss_from_batch.py (PW:n) and listwise_loss_from_batch.py (LW:n):
same names, argument order, defaults and return arity, so `from rec_now_b200.tf_ops import pairwise_loss` is a
drop-in for `from rec_now.rec_block.pairwise_loss_from_batch import pairwise_loss` in eager mode, tf.function
and TF1 graphs (gradients are registered with tf.RegisterGradient, which all three honour).
"""
from __future__ import annotations

import functools
import os

import tensorflow as tf

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "recnow_tf_ops.so")
if not os.path.exists(_SO):
    raise ImportError(f"{_SO} is missing: build it with `python -m rec_now_b200.tf_ops.build` (there is no CPU fallback)")
_ops = tf.load_op_library(_SO)

SMALL_POSIVITE_FLOAT = 1.0E-10      # (sic) PW:13
RN_LABEL_STEP, RN_LABEL_DIFF, RN_LABEL_LUT = 0, 1, 3


@tf.RegisterGradient("RecNowPairwiseLoss")
def _pairwise_grad(op, g_loss, g_n, g_ni, g_dlogits):
    # d loss / d logits was computed by the forward kernel; labels, keys, masks and weights get no gradient
    # (PW:264, PW:270).  Second-order terms through dlogits are not provided.
    return [g_loss * op.outputs[3], None, None, None, None, None, None]


@tf.RegisterGradient("RecNowListwiseLoss")
def _listwise_grad(op, g_loss, g_list_loss, g_nv, g_ng, g_dlogits):
    return [None, None, None, g_loss * op.outputs[4], None]           # LW:147, LW:166


tf.no_gradient("RecNowCanonKeys")
tf.no_gradient("RecNowPairIndices")


def _empty(dtype):
    return tf.zeros([0], dtype=dtype)


def _canon(groups, mask):
    cols = groups if isinstance(groups, list) else [groups]
    ok = _empty(tf.uint8) if mask is None else tf.cast(tf.reshape(mask, [-1]), tf.uint8)
    keys = []
    for g in cols:
        g = tf.reshape(g, [-1])
        if g.dtype in (tf.float32, tf.float64):
            k, ok = _ops.rec_now_canon_keys(ids=g, row_ok_in=ok)
        elif g.dtype in (tf.float16, tf.bfloat16):
            k, ok = _ops.rec_now_canon_keys(ids=tf.cast(g, tf.float32), row_ok_in=ok)
        else:
            k = tf.cast(g, tf.int64)
        keys.append(k)
    return tf.stack(keys), ok


def bpr_loss_func(outputs_pos, outputs_neg, weights=None, factor=1.0, reduce_mean=True):
    """PW:96-127 on explicit pair vectors."""
    logits = outputs_pos - outputs_neg
    if factor != 1.0:
        logits = logits * factor
    losses = tf.nn.sigmoid_cross_entropy_with_logits(labels=tf.ones_like(logits), logits=logits)
    if weights is not None:
        losses = losses * weights
    loss = tf.reduce_sum(losses)
    if reduce_mean:
        loss = loss / (tf.cast(tf.size(losses), tf.float32) + SMALL_POSIVITE_FLOAT)
    return loss


class FusedPairWeight:
    """label_pair_to_weight_func the fused kernel understands: phi(y_i, y_j) * kwargs[pos_kw][i] * kwargs[neg_kw][j],
    phi = [y_i > y_j] ("step"), (y_i - y_j)[y_i > y_j] ("diff") or table[y_i + 1][y_j + 1][y_i > y_j] ("lut": any
    label-only weight function as an 8 x 8 table over the label levels -1 .. 6, see from_callable).  Also a plain callable
    with the reference's contract, so the same object works with the reference implementation."""

    def __init__(self, label_func="step", pos_kw=None, neg_kw=None, table=None):
        self.label_func, self.pos_kw, self.neg_kw = label_func, pos_kw, neg_kw
        self.table = None if table is None else tf.reshape(tf.cast(table, tf.float32), [8, 8])

    @classmethod
    def from_callable(cls, f, pos_kw=None, **kwargs):
        lev = tf.range(-1.0, 7.0, dtype=tf.float32)
        return cls("lut", pos_kw=pos_kw, table=f(tf.tile(lev[:, None], [1, 8]), tf.tile(lev[None, :], [8, 1]), **kwargs))

    def __call__(self, label_matrix, label_matrix_transpose, **kwargs):
        gt = tf.cast(label_matrix > label_matrix_transpose, tf.float32)
        if self.label_func == "lut":
            li = tf.clip_by_value(tf.cast(label_matrix + 1.0, tf.int32), 0, 7)
            lj = tf.clip_by_value(tf.cast(label_matrix_transpose + 1.0, tf.int32), 0, 7)
            w = tf.gather_nd(self.table, tf.stack([li, lj], axis=-1)) * gt
        else:
            w = (label_matrix - label_matrix_transpose) * gt if self.label_func == "diff" else gt
        if self.pos_kw is not None:
            w = w * tf.reshape(kwargs[self.pos_kw], [-1, 1])
        if self.neg_kw is not None:
            w = w * tf.reshape(kwargs[self.neg_kw], [1, -1])
        return w


label_gain_times_sample_weight = FusedPairWeight("diff", pos_kw="sample_weight")


def _match_bpr(f):
    if f is bpr_loss_func:
        return 1.0, True
    if isinstance(f, functools.partial) and f.func is bpr_loss_func and not f.args and set(f.keywords) <= {"factor", "reduce_mean"}:
        return float(f.keywords.get("factor", 1.0)), bool(f.keywords.get("reduce_mean", True))
    return None


def pairwise_loss(outputs, labels, groups, pairloss_func=bpr_loss_func, only_use_wrong_order_pair=False,
                  return_num_pair=False, click_occurance_power=0.0, mask=None, label_pair_to_weight_func=None, **kwargs):
    """PW:228-279."""
    s = tf.reshape(tf.cast(outputs, tf.float32), [-1])
    y = tf.reshape(tf.cast(labels, tf.float32), [-1])
    keys, ok = _canon(groups, mask)
    fw = label_pair_to_weight_func if isinstance(label_pair_to_weight_func, FusedPairWeight) else None
    wp = _empty(tf.float32) if fw is None or fw.pos_kw is None else tf.reshape(tf.cast(kwargs[fw.pos_kw], tf.float32), [-1])
    wn = _empty(tf.float32) if fw is None or fw.neg_kw is None else tf.reshape(tf.cast(kwargs[fw.neg_kw], tf.float32), [-1])
    lf = RN_LABEL_DIFF if (fw is not None and fw.label_func == "diff") else RN_LABEL_STEP
    lut = _empty(tf.float32)
    if fw is not None and fw.label_func == "lut" and not only_use_wrong_order_pair:
        lf, lut = RN_LABEL_LUT, tf.reshape(fw.table, [-1])       # (labels off the menu -1 .. 6 fail the op: loss = NaN)
    elif fw is not None and fw.label_func == "lut":
        label_pair_to_weight_func, fw = fw, None                 # general path: the table object as a plain callable
        wp = _empty(tf.float32)
    bpr = _match_bpr(pairloss_func)
    if bpr is not None and (label_pair_to_weight_func is None or fw is not None):
        loss, n, _, _ = _ops.rec_now_pairwise_loss(logits=s, labels=y, group_keys=keys, row_ok=ok, rw_pos=wp, rw_neg=wn,
                                                   weight_lut=lut, label_func=lf, factor=bpr[0], power=float(click_occurance_power),
                                                   only_wrong=bool(only_use_wrong_order_pair), reduce_mean=bpr[1])
        return (loss, n) if return_num_pair else loss
    # general path: materialised pairs (row-major order, PW:217) + the caller's callables
    custom = label_pair_to_weight_func is not None and fw is None
    pos, neg, w = _ops.rec_now_pair_indices(logits=s, labels=y, group_keys=keys, row_ok=ok, rw_pos=wp, rw_neg=wn, label_func=lf,
                                            only_wrong=bool(only_use_wrong_order_pair) and not custom, label_cond=not custom)
    weights = w if fw is not None else None
    if custom:
        kw = {k: (tf.gather(tf.reshape(v, [-1]), neg if (len(v.shape) == 2 and v.shape[0] == 1) else pos)
                  if isinstance(v, tf.Tensor) and v.shape.num_elements() == s.shape.num_elements() else v)
              for k, v in kwargs.items()}
        wmat = label_pair_to_weight_func(tf.gather(y, pos), tf.gather(y, neg), **kw)          # PW:192
        keep = wmat > 0                                                                       # PW:193
        if only_use_wrong_order_pair:
            keep = tf.logical_and(keep, tf.gather(s, pos) < tf.gather(s, neg))                # PW:200-202
        pos, neg, weights = tf.boolean_mask(pos, keep), tf.boolean_mask(neg, keep), tf.boolean_mask(wmat, keep)
    if click_occurance_power != 0.0:                                                          # PW:285-290
        _, idx, cnt = tf.unique_with_counts(tf.gather(keys[0], pos))
        occ = tf.gather(tf.cast(cnt, tf.float32) if click_occurance_power == 1.0
                        else tf.pow(tf.cast(cnt, tf.float32), click_occurance_power), idx)
        weights = occ if weights is None else weights * occ
    if weights is not None:
        weights = tf.stop_gradient(weights)                                                   # PW:270
    loss = pairloss_func(tf.gather(s, pos), tf.gather(s, neg), weights)                       # PW:272-274
    return (loss, tf.cast(tf.size(pos), tf.float32)) if return_num_pair else loss


def to_listwise_sample(group_ids, labels, logits, do_mask_logits=True, value_of_masked_logit=-1E9, pos_neg_th=0.5):
    """LW:89-148: the dense (V,B) outputs (one stream synchronisation: V is data dependent)."""
    keys, ok = _canon(group_ids, None)
    y = tf.reshape(tf.cast(labels, tf.float32), [-1])
    s = tf.reshape(tf.cast(logits, tf.float32), [-1])
    dm, dl, dz = _ops.rec_now_listwise_dense(group_keys=keys[0], row_ok=ok, labels=y, logits=s, pos_neg_th=float(pos_neg_th),
                                             do_mask_logits=bool(do_mask_logits), value_of_masked_logit=float(value_of_masked_logit))
    # dense_logits must carry the gradient back to `logits` (LW:133, 139-140): member columns are the logits
    dz = tf.where(dm, tf.broadcast_to(tf.reshape(s, [1, -1]), tf.shape(dz)), dz)
    return dm, tf.stop_gradient(dl), dz


def listwise_loss_from_batch(group_ids, labels, logits, weights=None, do_reduce=True, pos_neg_th=0.5):
    """Fused to_listwise_sample + listwise_loss_via_softmax_cross_entropy_with_logits (LW:89-173) without the
    (V,B) tensors: returns (loss or per-list losses[:V], n_valid_list)."""
    if not do_reduce:
        # the registered gradient scales the saved d mean-loss / d logits by the upstream of output 0 only; per-list
        # upstream gradients (output 1) would need the rows' list ranks from the op -- refused until that exists and a
        # machine with TensorFlow has tested it (the torch drop-in handles this case through the dense mask)
        raise NotImplementedError("do_reduce=False is not differentiable through the fused TF op; use to_listwise_sample + "
                                  "listwise_loss_via_softmax_cross_entropy_with_logits(do_reduce=False)")
    keys, ok = _canon(group_ids, None)
    y = tf.reshape(tf.cast(labels, tf.float32), [-1])
    s = tf.reshape(tf.cast(logits, tf.float32), [-1])
    lw = _empty(tf.float32) if weights is None else tf.reshape(tf.cast(weights, tf.float32), [-1])
    loss, list_loss, nv, _, _ = _ops.rec_now_listwise_loss(group_keys=keys[0], row_ok=ok, labels=y, logits=s, list_w=lw,
                                                           pos_neg_th=float(pos_neg_th), do_reduce=bool(do_reduce))
    return (loss if do_reduce else list_loss[:nv]), nv


def nan_to_zero(val):
    """LW:74-86."""
    if len(val.shape) != 0:
        raise ValueError('input muust be a scalar tf.Tensor')
    return tf.cond(tf.math.is_nan(val), lambda: tf.zeros_like(val), lambda: val)


def listwise_loss_via_softmax_cross_entropy_with_logits(labels_for_softmax, logits_for_softmax, weights=None, do_reduce=True):
    """LW:151-173 on explicit dense tensors (use listwise_loss_from_batch for the fused segmented path)."""
    labels_for_softmax = tf.stop_gradient(labels_for_softmax)
    listwise_loss = tf.nn.softmax_cross_entropy_with_logits(labels=labels_for_softmax, logits=logits_for_softmax)
    if weights is not None:
        listwise_loss = listwise_loss * weights
    if do_reduce:
        listwise_loss = nan_to_zero(tf.reduce_mean(listwise_loss))
    return listwise_loss
