"""Drop-in for rec_now/rec_block/pairwise_loss_from_batch.py of the reference, on torch CUDA tensors.

Same public names, positional order, defaults and return arity as the reference
(/root/reference/rec_now/rec_block/pairwise_loss_from_batch.py, cited below as PW:n); the dense (B,B)
TensorFlow graph is replaced by librecnow_b200.so (segmentation + fused pair kernels, sm_100a).

Dispatch of ``pairwise_loss`` (PW:228-279):
  * fused path  -- ``pairloss_func`` is ``bpr_loss_func`` or a ``functools.partial`` of it, and
    ``label_pair_to_weight_func`` is None or a :class:`FusedPairWeight`: one C-ABI call computes loss, n_pair
    and d loss / d outputs; autograd just scales the saved gradient.
  * general path -- any other callable (e.g. the wrapper the reference's own test uses, tests/rec_block/
    test_pairwise_loss_from_batch.py:38-39): pairs are materialised on the GPU in the reference's row-major
    order (rn_pair_indices_*), the caller's functions run on the gathered pair vectors, torch autograd does the
    rest.  An arbitrary ``label_pair_to_weight_func`` must be elementwise in its two label arguments (the
    documented contract, PW:180-182); tensor kwargs shaped (B,1)/(B,) follow the positive side, (1,B) the
    negative side.
There is no CPU path: CPU tensors raise.
"""
from __future__ import annotations

import functools
from typing import Callable, Optional

import torch

from .. import ops

SMALL_POSIVITE_FLOAT = 1.0E-10      # (sic) PW:13


# --------------------------------------------------------------------------------------------------------
# small helpers
# --------------------------------------------------------------------------------------------------------
def _as_cuda(x, dtype=None) -> torch.Tensor:
    if isinstance(x, torch.Tensor):
        if not x.is_cuda:
            raise RuntimeError("rec_now_b200 needs CUDA tensors (there is no CPU fallback)")
        return x if dtype is None else x.to(dtype)
    return torch.as_tensor(x, dtype=dtype, device="cuda")


def _generate_pair_mask(sample_group_idx_var, only_upper_band=False):
    """PW:16-40 (dense compat helper)."""
    g = _as_cuda(sample_group_idx_var).reshape(-1, 1)
    n = g.numel()
    same = ((g - g.t()) == 0.0).to(torch.float32)
    m = (same - torch.eye(n, device=g.device)).to(torch.bool)
    if only_upper_band:
        m = torch.triu(m, 0) & torch.tril(m, 1)          # band_part(., 0, 1)
    return m


def generate_pair_mask(group_tensor_or_list, only_upper_band=False):
    """PW:43-74.  Dense (B,B) bool mask -- compat helper; pairwise_loss itself never builds it."""
    if not isinstance(group_tensor_or_list, list):
        group_tensor_or_list = [group_tensor_or_list]
    pair_mask = None
    for group in group_tensor_or_list:
        one = _generate_pair_mask(group, only_upper_band)
        pair_mask = one if pair_mask is None else torch.logical_and(pair_mask, one)
    return pair_mask


def vec_to_matrix_pair(vec):
    """PW:77-93 (dense compat helper): mat[i,j] = v_i, mat_T[i,j] = v_j."""
    v = _as_cuda(vec).reshape(-1, 1)
    mat = v.expand(-1, v.numel())
    return mat, mat.t()


def bpr_loss_func(outputs_pos, outputs_neg, weights=None, factor=1.0, reduce_mean=True):
    """PW:96-127 on explicit pair vectors (used by the general path and by user wrappers)."""
    logits = outputs_pos - outputs_neg
    if factor != 1.0:
        logits = logits * factor
    # sigmoid_cross_entropy_with_logits(labels=1): max(x,0) - x + log1p(exp(-|x|))
    losses = torch.clamp(logits, min=0) - logits + torch.log1p(torch.exp(-torch.abs(logits)))
    if weights is not None:
        losses = losses * weights
    loss = torch.sum(losses)
    if reduce_mean:
        loss = loss / (float(losses.numel()) + SMALL_POSIVITE_FLOAT)
    return loss


def hinge_loss_func(outputs_pos, outputs_neg, weights=None, margin=1.0, factor=1.0, reduce_mean=True):
    """Margin ranking loss for the ``pairloss_func`` hook (PW:229, called as PW:274): max(0, margin - (pos - neg) *
    factor), weighted and reduced exactly as bpr_loss_func (PW:122-126).  Not in the reference (which ships only
    bpr_loss_func); ``pairwise_loss(..., pairloss_func=hinge_loss_func)`` or a keyword partial of it takes the fused
    kernel."""
    logits = outputs_pos - outputs_neg
    if factor != 1.0:
        logits = logits * factor
    losses = torch.clamp(margin - logits, min=0)
    if weights is not None:
        losses = losses * weights
    loss = torch.sum(losses)
    if reduce_mean:
        loss = loss / (float(losses.numel()) + SMALL_POSIVITE_FLOAT)
    return loss


def occurance_power_weight(group_id, power=0.0):
    """PW:130-151: count(group_id == group_id[i]) ** power, float32."""
    g = _as_cuda(group_id)
    keys, ok = ops.canon_keys(g)
    k0 = keys[0]
    if ok is not None:
        # NaN/inf ids equal nothing (each is its own group): give them distinct keys in the NaN-payload range,
        # which no canonical finite key can take
        uniq = torch.arange(k0.numel(), device=k0.device, dtype=torch.int64) + 0x7FF8000000000000
        k0 = torch.where(ok.bool(), k0, uniq)
    return ops.occurrence_power_weight(k0, float(power)).reshape(g.shape)


class FusedPairWeight:
    """A ``label_pair_to_weight_func`` the fused kernel understands.

    ``W[i,j] = phi(y_i, y_j) * kwargs[pos_kw][i] * kwargs[neg_kw][j]`` with ``phi`` = ``[y_i > y_j]`` ("step"),
    ``(y_i - y_j) * [y_i > y_j]`` ("diff") or ``(2^y_i - 2^y_j) * [y_i > y_j]`` ("gain2": NDCG-style exponential gains);
    "lut": a table over the label levels (any label-only function, see ``from_callable``); "lambda": LambdaRank,
    ``phi = (2^y_i - 2^y_j) * |D(r_i) - D(r_j)| / IDCG_group * [y_i > y_j]`` with the rows' score ranks inside their group
    (include/recnow_b200.h RN_LABEL_LAMBDA; fused path only -- the reference's callable contract has no access to the scores).  It is also a plain callable with the reference's contract
    (label_matrix, label_matrix_transpose, **kwargs) -> weights, so the same object works with the reference.
    """

    #: label levels of the table form: integer labels -1 .. 6 (binary clicks, graded relevance), level = label + 1
    LEVEL_LABELS = tuple(float(v) for v in range(-1, 7))

    def __init__(self, label_func: str = "step", pos_kw: Optional[str] = None, neg_kw: Optional[str] = None, table=None):
        if label_func not in ("step", "diff", "gain2", "lut", "lambda"):
            raise ValueError("label_func must be 'step', 'diff', 'gain2', 'lut' or 'lambda'")
        if label_func == "lambda" and neg_kw is not None:
            raise ValueError("LambdaRank weights take no negative-side weights")
        self.label_func, self.pos_kw, self.neg_kw = label_func, pos_kw, neg_kw
        self.table = None
        self._table_dev: dict = {}
        if label_func == "lut":
            # W[i,j] = table[level(y_i)][level(y_j)] * [y_i > y_j]: ANY label-only weight function, evaluated once on the
            # 8 x 8 grid of label levels (rn_pairwise_args.weight_lut).  Entries with level_i > level_j must be finite
            # and > 0 (so that C = W > 0 is exactly y_i > y_j, PW:193); the others are not used.
            t = torch.as_tensor(table, dtype=torch.float32).detach().cpu().reshape(8, 8).clone()
            lower = torch.tril(torch.ones(8, 8, dtype=torch.bool), -1)
            if not bool((torch.isfinite(t[lower]) & (t[lower] > 0)).all()):
                raise ValueError("table entries with level_i > level_j must be finite and > 0")
            self.table = torch.where(lower, t, torch.zeros(()))
            if neg_kw is not None:
                raise ValueError("the table form has no negative-side weights")

    @classmethod
    def from_callable(cls, f, pos_kw: Optional[str] = None, **kwargs):
        """The table form of a label-only ``label_pair_to_weight_func`` ``f`` (evaluated on the label levels -1 .. 6)."""
        y = torch.tensor(cls.LEVEL_LABELS, dtype=torch.float32)
        return cls("lut", pos_kw=pos_kw, table=f(y.reshape(-1, 1).expand(8, 8), y.reshape(1, -1).expand(8, 8), **kwargs))

    def table_on(self, device) -> torch.Tensor:
        t = self._table_dev.get(device)
        if t is None:
            t = self._table_dev[device] = self.table.to(device).contiguous()
        return t

    def __call__(self, label_matrix, label_matrix_transpose, **kwargs):
        if self.label_func == "lambda":
            # |delta NDCG| needs the scores' ranks inside the groups: not a function of the label matrices
            raise NotImplementedError("LambdaRank weights exist on the fused path of pairwise_loss only")
        gt = (label_matrix > label_matrix_transpose).to(torch.float32)
        if self.label_func == "lut":
            # (a label off the level menu has no table entry: weight 0, the pair is dropped)
            def level(y):
                on = (y == torch.round(y)) & (y >= -1) & (y <= 6)
                return torch.where(on, y + 1, torch.zeros_like(y)).long(), on
            (li, oi), (lj, oj) = level(label_matrix), level(label_matrix_transpose)
            w = self.table.to(label_matrix.device)[li, lj] * gt * (oi & oj).to(torch.float32)
        elif self.label_func == "diff":
            w = (label_matrix - label_matrix_transpose) * gt
        elif self.label_func == "gain2":
            w = (torch.exp2(label_matrix) - torch.exp2(label_matrix_transpose)) * gt
        else:
            w = gt
        # ((B, B) label matrices as the reference passes them; pair vectors on this module's general path)
        if self.pos_kw is not None:
            w = w * (kwargs[self.pos_kw].reshape(-1, 1) if w.dim() == 2 else kwargs[self.pos_kw].reshape(w.shape))
        if self.neg_kw is not None:
            w = w * (kwargs[self.neg_kw].reshape(1, -1) if w.dim() == 2 else kwargs[self.neg_kw].reshape(w.shape))
        return w


#: cfg3 of BASELINE.json: W = (y_i - y_j) * [y_i > y_j] * sample_weight_i
label_gain_times_sample_weight = FusedPairWeight("diff", pos_kw="sample_weight")


class _FusedPairwiseLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, outputs, labels, keys, row_ok, rw_pos, rw_neg, label_func, factor, reduce_mean,
                only_wrong, power, hinge_margin=None, weight_lut=None):
        out = ops.pairwise_fwd_bwd(outputs, labels, keys, row_ok=row_ok, rw_pos=rw_pos, rw_neg=rw_neg,
                                   label_func=label_func, factor=factor, power=power, only_wrong=only_wrong,
                                   reduce_mean=reduce_mean, pair_loss="logistic" if hinge_margin is None else "hinge",
                                   margin=1.0 if hinge_margin is None else hinge_margin, weight_lut=weight_lut)
        ctx.save_for_backward(out["dlogits"])
        ctx.out_shape, ctx.out_dtype = outputs.shape, outputs.dtype
        n_pair = out["n_pair_f32"]
        ctx.mark_non_differentiable(n_pair)
        return out["loss"], n_pair

    @staticmethod
    def backward(ctx, g_loss, _g_n):
        (dlogits,) = ctx.saved_tensors
        g = g_loss * dlogits
        if g.shape != ctx.out_shape:
            g = g.reshape(ctx.out_shape)
        if g.dtype != ctx.out_dtype:
            g = g.to(ctx.out_dtype)
        return (g,) + (None,) * 12


def _match_bpr(pairloss_func) -> Optional[tuple]:
    """(factor, reduce_mean, hinge margin or None) if pairloss_func is bpr_loss_func / hinge_loss_func or a keyword-only
    partial of one of them."""
    if pairloss_func is bpr_loss_func:
        return 1.0, True, None
    if pairloss_func is hinge_loss_func:
        return 1.0, True, 1.0
    if isinstance(pairloss_func, functools.partial) and not pairloss_func.args:
        kw = dict(pairloss_func.keywords)
        if pairloss_func.func is bpr_loss_func and set(kw) <= {"factor", "reduce_mean"}:
            return float(kw.get("factor", 1.0)), bool(kw.get("reduce_mean", True)), None
        if pairloss_func.func is hinge_loss_func and set(kw) <= {"factor", "reduce_mean", "margin"} and \
                float(kw.get("margin", 1.0)) >= 0.0:
            return float(kw.get("factor", 1.0)), bool(kw.get("reduce_mean", True)), float(kw.get("margin", 1.0))
    return None


# --------------------------------------------------------------------------------------------------------
# arbitrary callables that ARE one of the fused forms
# --------------------------------------------------------------------------------------------------------
# The reference's API takes Python callables (PW:229, 234-235) and its own test passes a wrapper around bpr_loss_func and
# a lambda `float(y_i > y_j)` (TPW:38-39, 51-53).  A callable that depends only on its tensor arguments can be recognised
# by what it computes: it is evaluated ONCE on a fixed grid of probe values (a 16 x 16 grid of label pairs: negative,
# fractional, repeated and large values; random pair vectors for a loss function) and compared with the fused forms.  The
# verdict is cached per callable object (one small device-to-host read the first time, nothing afterwards).  A callable
# that matches on every probe point and still differs elsewhere would have to be built for the purpose;
# RN_PROBE_CALLABLES=0 switches the recognition off (everything unknown then takes the materialised-pair path).
_probe_cache: dict = {}
_PROBE_LABELS = [-3.0, -1.0, -0.5, 0.0, 0.0, 0.25, 0.5, 1.0, 1.0, 1.1, 2.0, 3.0, 4.0, 7.25, 100.0, 1.0e4]


def _probe_enabled() -> bool:
    import os
    return os.environ.get("RN_PROBE_CALLABLES", "1") != "0"


def _classify_weight_func(f, kwargs: dict, device) -> Optional["FusedPairWeight"]:
    """FusedPairWeight equivalent of a label-only ``label_pair_to_weight_func`` (no tensor kwargs), or None."""
    if not _probe_enabled() or any(isinstance(v, torch.Tensor) for v in kwargs.values()):
        return None
    try:
        key = ("w", id(f), tuple(sorted((k, repr(v)) for k, v in kwargs.items())))
    except Exception:
        return None
    hit = _probe_cache.get(key)
    if hit is not None and hit[0] is f:
        return hit[1]
    res = None
    try:
        t = torch.tensor(_PROBE_LABELS, dtype=torch.float32, device=device)
        y, yt = t.reshape(-1, 1).expand(-1, t.numel()).contiguous(), t.reshape(1, -1).expand(t.numel(), -1).contiguous()
        w = f(y, yt, **kwargs)
        if isinstance(w, torch.Tensor) and w.shape == y.shape:
            w = w.to(torch.float32)
            step = (y > yt).to(torch.float32)
            cands = (("step", step), ("diff", (y - yt) * step), ("gain2", (torch.exp2(y) - torch.exp2(yt)) * step))
            # (equal_nan: 2^1e4 overflows, inf - inf on the grid's diagonal is NaN in the callable and in the candidate alike)
            verdict = [bool(torch.equal(w, c)) or bool(torch.allclose(w, c, rtol=1e-6, atol=0.0, equal_nan=True)) for _, c in cands]
            for (name, _), ok in zip(cands, verdict):
                if ok:
                    res = FusedPairWeight(name)
                    break
            if res is None:
                res = _table_of_weight_func(f, kwargs, device)
    except Exception:
        res = None
    if len(_probe_cache) > 256:
        _probe_cache.clear()
    _probe_cache[key] = (f, res)
    return res


def _table_of_weight_func(f, kwargs: dict, device) -> Optional["FusedPairWeight"]:
    """Level-table form (FusedPairWeight 'lut') of a label-only callable that is none of the closed forms: ``f`` evaluated on
    the 8 x 8 grid of the label levels -1 .. 6.  Taken only if the pair set it implies is exactly ``y_i > y_j`` (entries
    with level_i > level_j finite and > 0, all others not > 0: PW:193) and if ``f`` is elementwise in its two arguments --
    checked on a shuffled (N, N) arrangement and on 1-D pair vectors against the table."""
    lev = torch.tensor(FusedPairWeight.LEVEL_LABELS, dtype=torch.float32, device=device)
    t = f(lev.reshape(-1, 1).expand(8, 8).contiguous(), lev.reshape(1, -1).expand(8, 8).contiguous(), **kwargs)
    if not isinstance(t, torch.Tensor) or t.shape != (8, 8):
        return None
    t = t.to(torch.float32)
    lower = torch.tril(torch.ones(8, 8, dtype=torch.bool, device=device), -1)
    if not bool((torch.isfinite(t[lower]) & (t[lower] > 0)).all()) or bool((t[~lower] > 0).any()):
        return None
    g = torch.Generator(device="cpu").manual_seed(20240607)
    idx = torch.randint(0, 8, (48,), generator=g).to(device)
    v = lev[idx]
    n = v.numel()
    w2 = f(v.reshape(-1, 1).expand(n, n).contiguous(), v.reshape(1, -1).expand(n, n).contiguous(), **kwargs)
    ia, ib = torch.randint(0, 8, (257,), generator=g).to(device), torch.randint(0, 8, (257,), generator=g).to(device)
    w1 = f(lev[ia], lev[ib], **kwargs)
    for w, want in ((w2, t[idx.reshape(-1, 1), idx.reshape(1, -1)]), (w1, t[ia, ib])):
        if not isinstance(w, torch.Tensor) or w.shape != want.shape:
            return None
        w = w.to(torch.float32)
        keep = want > 0
        if not (torch.equal(w > 0, keep) and torch.equal(w[keep], want[keep])):
            return None
    return FusedPairWeight("lut", table=t)


def _labels_on_level_menu(labels: torch.Tensor, row_ok: Optional[torch.Tensor]) -> bool:
    """True if every label that can take part in a pair is an integer in -1 .. 6 (one device reduction + one read-back)."""
    y = labels.reshape(-1).to(torch.float32)
    fine = ((y == torch.round(y)) & (y >= -1.0) & (y <= 6.0)) | torch.isnan(y)
    if row_ok is not None:
        fine = fine | ~row_ok.reshape(-1).bool()
    return bool(fine.all())


def _classify_pairloss_func(f, device) -> Optional[tuple]:
    """(factor, reduce_mean, hinge margin or None) if ``pairloss_func`` computes what bpr_loss_func(factor=1) or
    hinge_loss_func(margin=1, factor=1) computes (mean or sum), else None."""
    if not _probe_enabled():
        return None
    key = ("l", id(f))
    hit = _probe_cache.get(key)
    if hit is not None and hit[0] is f:
        return hit[1]
    res = None
    try:
        g = torch.Generator(device="cpu").manual_seed(1234)
        verdicts = {(1.0, True, None): True, (1.0, False, None): True, (1.0, True, 1.0): True, (1.0, False, 1.0): True}
        for n in (1, 7, 33):
            pos = (torch.randn(n, generator=g) * 3).to(device)
            neg = (torch.randn(n, generator=g) * 3).to(device)
            wts = torch.rand(n, generator=g).to(device) + 0.1
            for wv in (None, wts):
                out = f(pos, neg, wv)
                if not isinstance(out, torch.Tensor) or out.numel() != 1:
                    verdicts = {}
                    break
                for k in list(verdicts):
                    ref = bpr_loss_func(pos, neg, wv, factor=k[0], reduce_mean=k[1]) if k[2] is None else \
                        hinge_loss_func(pos, neg, wv, margin=k[2], factor=k[0], reduce_mean=k[1])
                    if not torch.allclose(out.reshape(()).to(torch.float32), ref, rtol=1e-6, atol=1e-7):
                        verdicts[k] = False
        for k in ((1.0, True, None), (1.0, False, None), (1.0, True, 1.0), (1.0, False, 1.0)):
            if verdicts.get(k):
                res = k
                break
    except Exception:
        res = None
    if len(_probe_cache) > 256:
        _probe_cache.clear()
    _probe_cache[key] = (f, res)
    return res


def _gather_kwargs(kwargs: dict, b: int, pos: torch.Tensor, neg: torch.Tensor) -> dict:
    out = {}
    for k, v in kwargs.items():
        if isinstance(v, torch.Tensor) and v.numel() == b and b > 1:
            out[k] = v.reshape(-1)[neg] if (v.dim() == 2 and v.shape[0] == 1) else v.reshape(-1)[pos]
        else:
            out[k] = v
    return out


def pairwise_loss(outputs, labels, groups,
                  pairloss_func=bpr_loss_func,
                  only_use_wrong_order_pair=False,
                  return_num_pair=False,
                  click_occurance_power=0.0,
                  mask=None,
                  label_pair_to_weight_func=None,
                  **kwargs
                  ):
    """PW:228-279.  Same arguments and return values as the reference."""
    outputs = _as_cuda(outputs)
    labels = _as_cuda(labels)
    group_list = [_as_cuda(g) for g in groups] if isinstance(groups, list) else [_as_cuda(groups)]
    b = outputs.numel()
    mask_t = None if mask is None else _as_cuda(mask).reshape(-1).to(torch.bool)
    keys, row_ok = ops.canon_keys(group_list, mask_t)
    users_weight_func = label_pair_to_weight_func
    if label_pair_to_weight_func is not None and not isinstance(label_pair_to_weight_func, FusedPairWeight):
        # a label-only callable that computes one of the fused forms (e.g. the reference's test lambda, TPW:51-53)
        recognised = _classify_weight_func(label_pair_to_weight_func, kwargs, outputs.device)
        if recognised is not None:
            label_pair_to_weight_func = recognised
    fused_w = label_pair_to_weight_func if isinstance(label_pair_to_weight_func, FusedPairWeight) else None
    rw_pos = rw_neg = None
    label_func = "step"
    if fused_w is not None:
        label_func = fused_w.label_func
        rw_pos = None if fused_w.pos_kw is None else _as_cuda(kwargs[fused_w.pos_kw])
        rw_neg = None if fused_w.neg_kw is None else _as_cuda(kwargs[fused_w.neg_kw])
        if fused_w.label_func == "step" and rw_pos is None and rw_neg is None:
            # W = [y_i > y_j], C = W > 0: identical to the default; keep the weight-free kernel
            fused_w = None
    bpr = _match_bpr(pairloss_func)
    if bpr is None and callable(pairloss_func):
        bpr = _classify_pairloss_func(pairloss_func, outputs.device)      # e.g. a wrapper around bpr_loss_func (TPW:38-39)
    power = float(click_occurance_power)
    weight_lut = None
    menu_w = label_pair_to_weight_func is None or isinstance(label_pair_to_weight_func, FusedPairWeight)
    if fused_w is not None and fused_w.label_func == "lut":
        # the level table: fused when the labels are on its menu (checked on the device: one read-back) and the pair set
        # does not depend on the scores; otherwise the table object is just another callable of the general path
        if bpr is not None and not only_use_wrong_order_pair and _labels_on_level_menu(labels, row_ok):
            weight_lut = fused_w.table_on(outputs.device)
        else:
            menu_w = False
            label_pair_to_weight_func = users_weight_func          # (the caller's own function, where the table stood in for one)
            rw_pos = None

    if fused_w is not None and fused_w.label_func == "lambda" and (
            bpr is None or bpr[2] is not None or only_use_wrong_order_pair):
        raise ValueError("LambdaRank weights need the logistic pair loss (bpr_loss_func) and the label-ordered pair set")

    if bpr is not None and menu_w:
        factor, reduce_mean, hinge_margin = bpr
        loss, n_pair = _FusedPairwiseLoss.apply(outputs, labels, keys, row_ok, rw_pos, rw_neg, label_func, factor,
                                                reduce_mean, bool(only_use_wrong_order_pair), power, hinge_margin,
                                                weight_lut)
        return (loss, n_pair) if return_num_pair else loss

    # ---- general path: materialised pairs + the caller's callables ------------------------------------
    flat_out = outputs.reshape(-1)
    if menu_w:
        want_w = isinstance(label_pair_to_weight_func, FusedPairWeight)
        pos, neg, w = ops.pair_indices(outputs, labels, keys, row_ok=row_ok, rw_pos=rw_pos, rw_neg=rw_neg,
                                       label_func=label_func, only_wrong=only_use_wrong_order_pair,
                                       label_cond=True, want_weights=want_w)
        pos, neg = pos.long(), neg.long()
        weights = w
    else:
        pos, neg, _ = ops.pair_indices(outputs, labels, keys, row_ok=row_ok, label_cond=False)
        pos, neg = pos.long(), neg.long()
        flat_y = labels.reshape(-1).to(torch.float32)
        wmat = label_pair_to_weight_func(flat_y[pos], flat_y[neg], **_gather_kwargs(kwargs, b, pos, neg))  # PW:192
        keep = wmat > 0                                                                                   # PW:193
        if only_use_wrong_order_pair:
            keep = keep & (flat_out.detach()[pos] < flat_out.detach()[neg])                               # PW:200-202
        pos, neg, weights = pos[keep], neg[keep], wmat[keep].to(torch.float32)
    if power != 0.0:                                                                                      # PW:285-290
        occ = ops.occurrence_power_weight(keys[0][pos], power) if pos.numel() else torch.empty(0, device=outputs.device)
        weights = occ if weights is None else weights * occ
    if weights is not None:
        weights = weights.detach()                                                                        # PW:270
    outputs_pos, outputs_neg = flat_out[pos], flat_out[neg]                                               # PW:272-273
    loss = pairloss_func(outputs_pos, outputs_neg, weights)                                               # PW:274
    if return_num_pair:
        return loss, torch.tensor(float(pos.numel()), dtype=torch.float32, device=outputs.device)          # PW:276
    return loss
