"""GPU parity of the pair losses / label gains beyond the reference's defaults: the hinge (margin) loss
behind the ``pairloss_func`` hook (pairwise_loss_from_batch.py:229, 274) and exponential label gains behind
``label_pair_to_weight_func`` (:175-194), through the C ABI against the float64 oracle -- which tests/
test_oracle_consistency.py ties to the dense op-for-op pipeline with those callables plugged in."""
import functools

import numpy as np
import pytest
import torch

from oracle import dense_ref as D
from oracle import generators as G
from oracle import seg_ref as S
from tests.util import check_pairwise, dev, run_pairwise

pytestmark = pytest.mark.gpu

HINGE = {
    "hinge": dict(),
    "hinge_m0": dict(margin=0.0),
    "hinge_factor_sum": dict(factor=2.5, reduce_mean=False, margin=0.3),
    "hinge_power": dict(power=-0.5),
    "hinge_wrong": dict(only_wrong=True, power=1.0),
    "hinge_diff_rwp": dict(label_func="diff", rw_pos="w"),
    "hinge_rwn": dict(rw_neg="w", margin=2.0),
    "hinge_gain2": dict(label_func="gain2", power=-0.5),
    "gain2": dict(label_func="gain2", pair_loss="logistic"),
    "gain2_rwp_power": dict(label_func="gain2", pair_loss="logistic", rw_pos="w", power=-0.5),
}


@pytest.mark.parametrize("name", list(HINGE))
@pytest.mark.parametrize("b", [700, 3000, 40000])
def test_hinge_and_gains(name, b):
    rng = np.random.default_rng(len(name) * 31 + b)
    gidx = G.zipf_groups(rng, b, max(8, b // 30))
    s = rng.standard_normal(b).astype(np.float32) * 1.5
    s[rng.integers(0, b, b // 10)] = 0.25                     # tied scores: the hinge's kink at margin = 0
    y = rng.integers(0, 5, b).astype(np.float32)
    w = rng.uniform(0.5, 1.5, b).astype(np.float32)
    kw = dict(pair_loss="hinge")
    kw.update({k: (w if v == "w" else v) for k, v in HINGE[name].items()})
    spec = S.PairSpec(**kw)
    ids = gidx.astype(np.int64) * 104729 + 7
    out = run_pairwise(s, y, ids, spec)
    check_pairwise(out, S.pairwise(s, y, ids, spec), ctx=f"{name} B={b}")


def test_hinge_non_integer_labels_and_masks():
    """Labels outside the counting path's level menu (radix segmentation), a sample mask, float ids."""
    rng = np.random.default_rng(3)
    b = 5000
    g = rng.integers(0, 60, b).astype(np.float32)
    s = rng.standard_normal(b).astype(np.float32)
    y = rng.uniform(0, 1, b).astype(np.float32).round(1)
    mask = rng.random(b) < 0.85
    spec = S.PairSpec(pair_loss="hinge", margin=0.7, power=-0.5)
    out = run_pairwise(s, y, g, spec, mask=mask)
    check_pairwise(out, S.pairwise(s, y, g, spec, mask=mask), ctx="hinge radix")


def test_dropin_hinge_and_gain2():
    """The torch drop-in: hinge_loss_func and partials of it, a user wrapper recognised by what it computes, and the
    gain2 weight function -- all on the fused path -- against the dense op-for-op oracle with the same callables."""
    from rec_now_b200 import ops
    from rec_now_b200.rec_block import pairwise_loss_from_batch as PW
    rng = np.random.default_rng(11)
    b = 1500
    g = rng.integers(0, 25, b).astype(np.float32)
    s = rng.standard_normal(b).astype(np.float32)
    y = rng.integers(0, 4, b).astype(np.float32)
    tg, ts, ty = (torch.tensor(v, device="cuda") for v in (g, s, y))
    real, calls = ops.pair_indices, {"n": 0}

    def counting(*a, **k):
        calls["n"] += 1
        return real(*a, **k)
    ops.pair_indices = counting
    try:
        ts.requires_grad_(True)
        loss, n = PW.pairwise_loss(ts, ty, tg, functools.partial(PW.hinge_loss_func, margin=0.5, factor=2.0),
                                   return_num_pair=True, click_occurance_power=-0.5)
        ref, nref = D.pairwise_loss(s, y, g, functools.partial(D.hinge_loss_func, margin=0.5, factor=2.0),
                                    return_num_pair=True, click_occurance_power=-0.5)
        assert float(n) == float(nref) and abs(loss.item() - float(ref)) <= 1e-5 * abs(float(ref))
        loss.backward()
        r = S.pairwise(s, y, g, S.PairSpec(pair_loss="hinge", margin=0.5, factor=2.0, power=-0.5))
        assert np.abs(ts.grad.cpu().numpy() - r["grad"]).max() <= 1e-5 * r["grad_abs"].max()

        def my_hinge(pos, neg, weights):                      # (a wrapper, as the reference's test wraps bpr_loss_func)
            return PW.hinge_loss_func(pos, neg, weights)
        a = PW.pairwise_loss(ts.detach(), ty, tg, my_hinge)
        assert abs(a.item() - float(D.pairwise_loss(s, y, g, D.hinge_loss_func))) <= 1e-5 * abs(a.item())

        def exp_gain(lm, lmt):
            return (torch.exp2(lm) - torch.exp2(lmt)) * (lm > lmt).to(torch.float32)
        c = PW.pairwise_loss(ts.detach(), ty, tg, label_pair_to_weight_func=exp_gain)
        d = PW.pairwise_loss(ts.detach(), ty, tg, label_pair_to_weight_func=PW.FusedPairWeight("gain2"))
        refc = D.pairwise_loss(s, y, g, label_pair_to_weight_func=lambda lm, lmt: ((np.exp2(lm) - np.exp2(lmt)) * (lm > lmt)).astype(np.float32))
        assert abs(c.item() - float(refc)) <= 1e-5 * abs(float(refc)) and abs(d.item() - c.item()) <= 2e-6 * abs(c.item())
        assert calls["n"] == 0                                # nothing was materialised
    finally:
        ops.pair_indices = real


# ---- LambdaRank |delta NDCG| weights (RN_LABEL_LAMBDA) ----------------------------------------------------
@pytest.mark.parametrize("b,ng,rw", [(400, 6, False), (7000, 50, True), (1000, 1, False)])
def test_lambdarank_weights_match_oracle(b, ng, rw):
    """Loss, gradient and exact counts of the LambdaRank-weighted logistic pair loss against the float64 oracle (ranks by
    score inside the group, ties by row; discounts rounded to float32 as the product rounds them)."""
    rng = np.random.default_rng(b + ng)
    g = rng.integers(0, ng, b).astype(np.float32)
    s = rng.standard_normal(b).astype(np.float32)
    s[rng.integers(0, b, b // 10)] = 0.25                              # tied scores: ranks by row index
    y = rng.integers(0, 5, b).astype(np.float32)
    spec = S.PairSpec(label_func="lambda", rw_pos=rng.uniform(0.5, 1.5, b).astype(np.float32) if rw else None,
                      power=-0.5 if rw else 0.0)
    out = run_pairwise(s, y, g, spec)
    check_pairwise(out, S.pairwise(s, y, g, spec), ctx=f"lambda B={b}")


def test_lambdarank_cfg3_full_size_and_mask():
    """cfg3's batch (B = 65 536, groups of up to ~7 500 rows: neighbouring ranks of a long group differ in the sixth digit
    of their discounts) on the counting segmentation; then a masked batch with NaN labels on two key columns (radix)."""
    from oracle import generators as G
    c = G.cfg3()
    spec = S.PairSpec(label_func="lambda", rw_pos=c["w"], power=-0.5)
    out = run_pairwise(c["s"], c["y"], c["g_f32"], spec)
    check_pairwise(out, S.pairwise(c["s"], c["y"], c["g_f32"], spec), ctx="lambda cfg3")
    rng = np.random.default_rng(11)
    b = 9000
    g1, g2 = rng.integers(0, 25, b).astype(np.float32), rng.integers(0, 3, b).astype(np.float32)
    s = rng.standard_normal(b).astype(np.float32)
    y = rng.integers(0, 4, b).astype(np.float32)
    y[rng.random(b) < 0.02] = np.nan
    mask = rng.random(b) < 0.85
    spec = S.PairSpec(label_func="lambda", factor=2.0, reduce_mean=False)
    out = run_pairwise(s, y, [g1, g2], spec, mask=mask)
    check_pairwise(out, S.pairwise(s, y, [g1, g2], spec, mask=mask), ctx="lambda K=2 masked")


def test_lambdarank_dropin_and_menu():
    from rec_now_b200 import ops
    from rec_now_b200.rec_block import pairwise_loss_from_batch as PW
    rng = np.random.default_rng(12)
    b = 3000
    g = rng.integers(0, 30, b).astype(np.float32)
    s = rng.standard_normal(b).astype(np.float32)
    y = rng.integers(0, 5, b).astype(np.float32)
    ts = torch.tensor(s, device="cuda", requires_grad=True)
    loss = PW.pairwise_loss(ts, dev(y), dev(g), label_pair_to_weight_func=PW.FusedPairWeight("lambda"))
    loss.backward()
    ref = S.pairwise(s, y, g, S.PairSpec(label_func="lambda"))
    assert abs(float(loss.detach()) - ref["loss"]) <= 1e-5 * abs(ref["loss"])
    err = np.abs(ts.grad.cpu().numpy().astype(np.float64) - ref["grad"])
    assert (err <= 1e-5 * ref["grad_abs"] + 1e-12).all()
    with pytest.raises(ValueError):
        PW.pairwise_loss(dev(s), dev(y), dev(g), PW.hinge_loss_func, label_pair_to_weight_func=PW.FusedPairWeight("lambda"))
    with pytest.raises(ValueError):
        PW.pairwise_loss(dev(s), dev(y), dev(g), only_use_wrong_order_pair=True,
                         label_pair_to_weight_func=PW.FusedPairWeight("lambda"))
    keys, _ = ops.canon_keys([dev(g)])
    with pytest.raises(Exception):
        ops.pairwise_fwd_bwd(dev(s), dev(y), keys, label_func="lambda", pair_loss="hinge")
