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
from tests.util import check_pairwise, run_pairwise

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
