"""Segmented float64 restatement of the reference's in-batch ranking losses (NumPy).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Same semantics as dense_ref.py /
the reference (closed forms in SURVEY.md section 8a), but evaluated group by
group in float64, so it is the truth for loss and gradient at any batch size
and never builds a (B,B) tensor.  Integer outputs (pair counts per row, per
primary group and in total; row-major pair lists) are exact.

Citations: PW:n = /root/reference/rec_now/rec_block/pairwise_loss_from_batch.py:n,
LW:n = /root/reference/rec_now/rec_block/listwise_loss_from_batch.py:n.

The fused weight menu mirrored here (and in include/recnow_b200.h):
    label_func "step":  C = y_i > y_j,              W = None      (PW:188-190, the default)
    label_func "diff":  W = (y_i - y_j) * [y_i > y_j] [* rw_pos_i] [* rw_neg_j],  C = W > 0   (PW:192-193)
    label_func "step" with row weights:  W = [y_i > y_j] [* rw_pos_i] [* rw_neg_j], C = W > 0
    label_func "gain2": W = (2^y_i - 2^y_j) * [y_i > y_j] [* rw ...]   (exponential gains; float32 exp2 as the product)
i.e. what a reference user writes as label_pair_to_weight_func(Y, Yt, sample_weight=w).
    label_func "callable": W = weight_func(Y_i, Y_j) [* rw ...], C = W > 0 -- ANY label_pair_to_weight_func of the two label
                        matrices (PW:192-193), evaluated block by block in float32; the truth for the level-table path
                        (RN_LABEL_LUT) of the product.
    label_func "lambda": LambdaRank, W = (2^y_i - 2^y_j) * |D(r_i) - D(r_j)| / IDCG_g [* rw_pos_i] on the pair set y_i > y_j
                        (defined in include/recnow_b200.h -- the reference ships only the hook): D(r) =
                        1 / log2(1 + r) rounded to float32, r = 1-based rank by score inside the group among the rows that
                        can pair (descending, ties by row index), IDCG_g = DCG of the group's labels in descending
                        order with gains 2^y - 1 (float64); IDCG_g <= 0 gives weight 0.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np

F32 = np.float32


@dataclass
class PairSpec:
    factor: float = 1.0                 # bpr_loss_func factor            PW:118-119
    reduce_mean: bool = True            # bpr_loss_func reduce_mean       PW:125-126
    only_wrong: bool = False            # only_use_wrong_order_pair       PW:197-203
    power: float = 0.0                  # click_occurance_power           PW:282-291
    label_func: str = "step"            # "step" | "diff" | "gain2" | "callable" | "lambda"
    weight_func: Optional[object] = None  # label_func "callable": f(label_matrix, label_matrix_transpose) -> weights   PW:192
    rw_pos: Optional[np.ndarray] = None  # per-sample weight applied on the positive (row) side
    rw_neg: Optional[np.ndarray] = None  # per-sample weight applied on the negative (column) side
    # pairloss_func menu beyond bpr_loss_func (the reference only ships the hook, PW:229, 274):
    pair_loss: str = "logistic"         # "logistic" = bpr_loss_func | "hinge" = max(0, margin - x), x as PW:117-119
    margin: float = 1.0


def canonical_keys(groups, inf_is_id: bool = False) -> tuple[np.ndarray, np.ndarray]:
    """Composite-key canonicalisation: value equality per key (PW:33-35), -0.0 == +0.0,
    NaN/+-inf match nothing (inf_is_id: +-inf equal themselves -- tf.unique of the listwise path, LW:109).
    Returns (int64 keys [K,B], row_ok bool[B])."""
    cols = groups if isinstance(groups, (list, tuple)) else [groups]
    out, ok = [], None
    for g in cols:
        g = np.asarray(g).reshape(-1)
        fin = np.ones(g.size, bool)
        if g.dtype.kind == "f":
            fin = ~np.isnan(g) if inf_is_id else np.isfinite(g)
            gz = np.where(fin, g, 0) + g.dtype.type(0)          # -0.0 -> +0.0
            k = gz.astype(np.float64).view(np.int64) if g.dtype != np.float64 else gz.view(np.int64)
        else:
            k = g.astype(np.int64)
        out.append(np.ascontiguousarray(k))
        ok = fin if ok is None else (ok & fin)
    return np.stack(out), ok


def _group_members(keys: np.ndarray, row_ok: np.ndarray):
    """Lists of member rows (ascending) per composite group, groups in first-occurrence order."""
    rows = np.flatnonzero(row_ok)
    if rows.size == 0:
        return []
    sub = keys[:, rows]
    order = np.lexsort(tuple(sub[k] for k in range(sub.shape[0] - 1, -1, -1)) )  # stable -> rows ascending
    ks = sub[:, order]
    brk = np.flatnonzero(np.any(ks[:, 1:] != ks[:, :-1], axis=0)) + 1
    segs = np.split(rows[order], brk)
    segs.sort(key=lambda m: m[0])
    return segs


def _softplus_neg(x):
    """sigmoid-CE with label 1 = softplus(-x), TF stable form (PW:120-121), float64."""
    return np.maximum(x, 0.0) - x + np.log1p(np.exp(-np.abs(x)))


def _sigma_neg(x):
    """sigma(-x) = -d softplus(-x)/dx, overflow-safe."""
    e = np.exp(-np.abs(x))
    return np.where(x >= 0, e / (1.0 + e), 1.0 / (1.0 + e))


def pairwise(outputs, labels, groups, spec: PairSpec = PairSpec(), mask=None,
             want_pairs: bool = False, chunk_elems: int = 1 << 22):
    """float64 truth for pairwise_loss (PW:228-279) with pairloss_func = bpr_loss_func.

    Returns dict: loss, grad[B], grad_abs[B] (A_i of SURVEY 8d), n_pair, row_pairs[B] (pairs with the
    row on the positive side), prim_count{primary key -> c_h}, and if want_pairs: pos_idx, neg_idx, w
    (row-major order, as PW:217 defines it).
    """
    s32 = np.asarray(outputs, F32).reshape(-1)
    y32 = np.asarray(labels, F32).reshape(-1)
    b = s32.size
    keys, ok = canonical_keys(groups)
    if mask is not None:
        ok = ok & np.asarray(mask, bool).reshape(-1)                         # PW:154-172
    rwp = None if spec.rw_pos is None else np.asarray(spec.rw_pos, F32).reshape(-1)
    rwn = None if spec.rw_neg is None else np.asarray(spec.rw_neg, F32).reshape(-1)
    has_w = spec.label_func in ("diff", "gain2", "callable", "lambda") or rwp is not None or rwn is not None
    g32 = np.exp2(y32).astype(F32) if spec.label_func in ("gain2", "lambda") else y32
    segs = _group_members(keys, ok)
    disc32 = np.zeros(b, F32)            # "lambda": rank discount of every row, 1 / IDCG of its group
    inv_idcg32 = np.zeros(b, F32)
    if spec.label_func == "lambda":
        for m in segs:
            mm = m[~np.isnan(y32[m])]                                        # (a NaN label pairs with nothing: not ranked)
            if mm.size == 0:
                continue
            order = np.lexsort((mm, -s32[mm].astype(np.float64)))            # score descending, ties by row index
            rank = np.empty(mm.size, np.float64)
            rank[order] = np.arange(1, mm.size + 1)
            disc32[mm] = (1.0 / np.log2(1.0 + rank)).astype(F32)
            gains = np.sort(g32[mm].astype(np.float64) - 1.0)[::-1]
            idcg = float((gains / np.log2(1.0 + np.arange(1, mm.size + 1))).sum())
            inv_idcg32[mm] = F32(1.0 / idcg) if idcg > 0.0 else F32(0.0)

    row_pairs = np.zeros(b, np.int64)
    # pass 1: exact counts (needed for c_h before any weight can be formed, PW:288-289)
    per_seg = []
    for m in segs:
        if m.size < 2:
            continue
        per_seg.append(m)
    prim_count: dict[int, int] = {}

    def pair_block(mi, m):
        """cond (bool) and float32 weight matrix (or None) for positive rows mi x negative rows m."""
        yi, yj = y32[mi][:, None], y32[m][None, :]
        gt = yi > yj
        if not has_w:
            cond, w = gt, None                                               # PW:188-190
        else:
            if spec.label_func in ("diff", "gain2", "lambda"):
                w = ((g32[mi][:, None] - g32[m][None, :]).astype(F32) * gt.astype(F32)).astype(F32)
            elif spec.label_func == "callable":
                yim, yjm = np.broadcast_arrays(yi, yj)
                w = np.asarray(spec.weight_func(yim, yjm), F32)               # PW:192 (no label condition of its own)
            else:
                w = gt.astype(F32)
            if rwp is not None:
                w = (w * rwp[mi][:, None]).astype(F32)
            if rwn is not None:
                w = (w * rwn[m][None, :]).astype(F32)
            cond = w > 0                                                     # PW:193
            if spec.label_func == "lambda":
                # (the pair set is y_i > y_j -- the label gain times the row weight, tested above, decides it; the rank
                # part of the weight only scales the pairs)
                w = (w * inv_idcg32[mi][:, None]).astype(F32)
                w = (w * np.abs(disc32[mi][:, None] - disc32[m][None, :]).astype(F32)).astype(F32)
        cond = cond & (mi[:, None] != m[None, :])                            # PW:36 (identity removed)
        if spec.only_wrong:
            cond = cond & (s32[mi][:, None] < s32[m][None, :])               # PW:200-202
        return cond, w

    def chunks(m):
        step = max(1, chunk_elems // m.size)
        for a in range(0, m.size, step):
            yield m[a:a + step]

    for m in per_seg:
        for mi in chunks(m):
            cond, _ = pair_block(mi, m)
            row_pairs[mi] += cond.sum(axis=1)
    n = int(row_pairs.sum())
    prim = keys[0]
    if n:
        nz = np.flatnonzero(row_pairs)
        for k, c in zip(prim[nz].tolist(), row_pairs[nz].tolist()):
            prim_count[k] = prim_count.get(k, 0) + c

    loss = 0.0
    grad = np.zeros(b, np.float64)
    gabs = np.zeros(b, np.float64)
    out_pos, out_neg, out_w = [], [], []
    f = float(spec.factor)
    for m in per_seg:
        for mi in chunks(m):
            cond, w = pair_block(mi, m)
            if not cond.any():
                continue
            x = (s32[mi][:, None] - s32[m][None, :]).astype(F32)             # PW:117 (float32 subtract)
            if spec.factor != 1.0:
                x = (x * F32(spec.factor)).astype(F32)                        # PW:118-119
            x = x.astype(np.float64)
            wt = np.ones_like(x) if w is None else w.astype(np.float64)
            if spec.power != 0.0:                                            # PW:285-290
                c = np.array([prim_count.get(k, 1) for k in prim[mi].tolist()], np.float64)
                wt = wt * np.power(c, float(spec.power))[:, None]
            wt = np.where(cond, wt, 0.0)
            if spec.pair_loss == "hinge":
                v = float(spec.margin) - x
                loss += float((wt * np.maximum(v, 0.0)).sum())
                d = wt * (v > 0) * f
            else:
                loss += float((wt * _softplus_neg(x)).sum())
                d = wt * _sigma_neg(x) * f
            grad[mi] -= d.sum(axis=1)
            np.add.at(grad, m, d.sum(axis=0))
            gabs[mi] += np.abs(d).sum(axis=1)
            np.add.at(gabs, m, np.abs(d).sum(axis=0))
            if want_pairs:
                ii, jj = np.nonzero(cond)
                out_pos.append(mi[ii]); out_neg.append(m[jj]); out_w.append(wt[ii, jj])
    denom = float(F32(F32(n) + F32(1e-10))) if spec.reduce_mean else 1.0     # PW:125-126, PW:13
    res = dict(loss=loss / denom, grad=grad / denom, grad_abs=gabs / denom, n_pair=n,
               row_pairs=row_pairs, prim_count=prim_count)
    if want_pairs:
        if out_pos:
            p, q, w = np.concatenate(out_pos), np.concatenate(out_neg), np.concatenate(out_w)
            o = np.lexsort((q, p))                                           # row-major: i asc, then j asc
            res.update(pos_idx=p[o].astype(np.int64), neg_idx=q[o].astype(np.int64), w=w[o])
        else:
            res.update(pos_idx=np.zeros(0, np.int64), neg_idx=np.zeros(0, np.int64), w=np.zeros(0))
    return res


def listwise(group_ids, labels, logits, weights=None, pos_neg_th=0.5):
    """float64 truth for to_listwise_sample + listwise_loss_via_softmax_cross_entropy_with_logits with the
    default masking (do_mask_logits=True, th >= 0): closed form of SURVEY 8a (LW:109-148, LW:166-172).

    `weights`: optional per-VALID-list weights in first-occurrence order (LW:168-169).
    Returns dict(loss, grad[B], n_valid, n_group, list_loss[V], list_first_row[V], list_size[V]).
    """
    s32 = np.asarray(logits, F32).reshape(-1)
    y32 = np.asarray(labels, F32).reshape(-1)
    keys, ok = canonical_keys(group_ids, inf_is_id=True)
    # tf.unique puts every row in some list (values compared with ==: +-inf ids equal themselves); a NaN id equals
    # nothing, i.e. is a singleton list, and a singleton is never valid (needs a positive AND a negative) -> dropping
    # those rows gives the same result
    segs = _group_members(keys, ok)
    th = F32(pos_neg_th)
    losses, firsts, sizes, members = [], [], [], []
    for m in segs:
        ym = y32[m]
        if not (np.any(ym > th) and np.any((ym - th).astype(F32) < 0)):      # LW:135-137
            continue
        z = s32[m].astype(np.float64)
        p = (ym / np.sum(ym, dtype=F32)).astype(F32).astype(np.float64)      # LW:144 (float32 normalise)
        zs = z - z.max()
        lse = np.log(np.exp(zs).sum())
        losses.append(float((p * (lse - zs)).sum()))
        firsts.append(int(m[0])); sizes.append(int(m.size)); members.append((m, p, zs, lse))
    v = len(losses)
    grad = np.zeros(s32.size, np.float64)
    ll = np.asarray(losses, np.float64)
    w = np.ones(v) if weights is None else np.asarray(weights, np.float64).reshape(-1)
    if v:
        lw = ll * w
        loss = float(lw.mean())
        for r, (m, p, zs, lse) in enumerate(members):
            sm = np.exp(zs - lse)
            grad[m] = (w[r] / v) * (sm * p.sum() - p)
    else:
        lw, loss = ll, 0.0                                                   # LW:172 nan_to_zero
    return dict(loss=loss, grad=grad, n_valid=v, n_group=len(segs) + int((~ok).sum()), list_loss=lw,
                list_first_row=np.asarray(firsts, np.int64), list_size=np.asarray(sizes, np.int64))


def gauc(scores, labels, groups, mask=None):
    """float64 truth for the GAUC metric of rec_now_b200.metrics.gauc / rn_gauc (the reference quotes the metric,
    README.md:5, 8, without implementing it; definition: include/recnow_b200.h).  Per group of rows that can pair (finite
    key, mask, non-NaN label): pairs = {(i, j): y_i > y_j} (PW:189), AUC = (#{s_i > s_j} + #{s_i == s_j} / 2) / #pairs;
    GAUC = sum |g| AUC_g / sum |g| over the groups with pairs.  Counts are exact integers.
    Returns dict(gauc, auc_mean, n_valid_groups, n_pair, concordant2)."""
    s = np.asarray(scores, F32).reshape(-1)
    y = np.asarray(labels, F32).reshape(-1)
    keys, ok = canonical_keys(groups)
    ok = ok & ~np.isnan(y)
    if mask is not None:
        ok = ok & np.asarray(mask, bool).reshape(-1)
    num = den = asum = 0.0
    n_pair = conc2 = nv = 0
    for m in _group_members(keys, ok):
        ym, sm = y[m], s[m]
        pm = ym[:, None] > ym[None, :]
        n = int(pm.sum())
        if not n:
            continue
        c2 = int((2 * (sm[:, None] > sm[None, :])[pm]).sum() + (sm[:, None] == sm[None, :])[pm].sum())
        auc = c2 / (2.0 * n)
        num += m.size * auc; den += m.size; asum += auc
        n_pair += n; conc2 += c2; nv += 1
    return dict(gauc=num / den if den else 0.0, auc_mean=asum / nv if nv else 0.0, n_valid_groups=nv, n_pair=n_pair,
                concordant2=conc2)


def focal(labels, logits, alpha=0.25, gamma=2.0, stop_weight_gradient=False):
    """float64 truth for focal_crossentropy_loss(return_mean=True) (focal_loss.py:12-66) and its gradient with respect
    to the logits.  Returns dict(loss, grad[B])."""
    y = np.asarray(labels, F32).reshape(-1).astype(np.float64)
    z = np.asarray(logits, F32).reshape(-1).astype(np.float64)
    ce = np.maximum(z, 0) - z * y + np.log1p(np.exp(-np.abs(z)))
    p = 1.0 / (1.0 + np.exp(-z))
    af = (y * alpha + (1 - y) * (1 - alpha)) if alpha else np.ones_like(z)
    mod, dmod = np.ones_like(z), np.zeros_like(z)
    if gamma:
        om = 1.0 - (y * p + (1 - y) * (1 - p))
        mod = om ** gamma
        if not stop_weight_gradient:
            with np.errstate(divide="ignore", invalid="ignore"):
                dmod = -gamma * om ** (gamma - 1.0) * (2 * y - 1) * p * (1 - p)
    fl = af * mod * ce
    grad = af * (mod * (p - y) + ce * dmod) / z.size
    return dict(loss=float(fl.mean()), grad=grad)
