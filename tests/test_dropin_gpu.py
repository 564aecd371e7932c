"""The reference's own hot-path unit tests, run against the drop-in modules on the GPU.

Bodies follow /root/reference/tests/rec_block/test_pairwise_loss_from_batch.py:19-74 and
test_listwise_loss_from_batch.py:18-51 with tf.constant -> torch.tensor(device="cuda") (and assertEquals ->
assertEqual, removed in Python 3.12); golden numbers and tolerances are the reference's.
"""
import functools
import unittest

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def colt(v, dtype=torch.float32):
    return torch.tensor([v], dtype=dtype, device="cuda").t()


class TestPairwiseLossFromBatch(unittest.TestCase):
    def test_occurance_power_weight(self):
        from rec_now_b200.rec_block.pairwise_loss_from_batch import occurance_power_weight
        group_id = [1, 1, 2, 4, 4, 4]
        weights1 = occurance_power_weight(group_id, power=-1)
        result1 = [0.5, 0.5, 1., 0.33333334, 0.33333334, 0.33333334]
        weights2 = occurance_power_weight(group_id, power=2)
        result2 = [4., 4., 1., 9., 9., 9.]
        for e, r in zip(result1, weights1.cpu().numpy()):
            self.assertAlmostEqual(e, r, delta=0.0001)
        for e, r in zip(result2, weights2.cpu().numpy()):
            self.assertAlmostEqual(e, r, delta=0.0001)

    def test_pairwise_loss(self):
        from rec_now_b200.rec_block.pairwise_loss_from_batch import bpr_loss_func, pairwise_loss
        sample_group_idx_var = colt([1, 1, 2, 2, 2])
        logits = colt([0, 1, 2, 3, 4])
        label = colt([1.1, 0, 0, 1, 1])

        def pairwise_loss_func(outputs_pos, outputs_neg, weights):
            return bpr_loss_func(outputs_pos, outputs_neg, weights, 1.0)

        pairloss = pairwise_loss(logits, label, sample_group_idx_var, pairwise_loss_func,
                                 only_use_wrong_order_pair=False, click_occurance_power=-0.5)
        self.assertAlmostEqual(pairloss.item(), 0.5415076, delta=1e-4)

        def _label_pair_to_weight_func(label_matrix, label_matrix_transpose, **kwargs):
            return (label_matrix > label_matrix_transpose).to(torch.float32)

        pairloss_with_weight = pairwise_loss(logits, label, sample_group_idx_var, pairwise_loss_func,
                                             only_use_wrong_order_pair=False, click_occurance_power=-0.5,
                                             label_pair_to_weight_func=_label_pair_to_weight_func)
        self.assertAlmostEqual(pairloss_with_weight.item(), 0.5415076, delta=1e-4)

        mask = colt([True, True, False, False, False], torch.bool)
        pairloss_with_sample_mask = pairwise_loss(logits, label, sample_group_idx_var, pairwise_loss_func,
                                                  only_use_wrong_order_pair=False, click_occurance_power=-0.5,
                                                  mask=mask)
        self.assertAlmostEqual(pairloss_with_sample_mask.item(), 1.3132617, delta=1e-4)

    def test_pairwise_loss_fused_defaults(self):
        """Same three cases through the fused path (pairloss_func left at its default)."""
        from rec_now_b200.rec_block.pairwise_loss_from_batch import FusedPairWeight, bpr_loss_func, pairwise_loss
        g, logits, label = colt([1, 1, 2, 2, 2]), colt([0, 1, 2, 3, 4]), colt([1.1, 0, 0, 1, 1])
        self.assertAlmostEqual(pairwise_loss(logits, label, g, click_occurance_power=-0.5).item(), 0.5415076, delta=1e-5)
        loss, n = pairwise_loss(logits, label, g, functools.partial(bpr_loss_func, factor=1.0),
                                return_num_pair=True, click_occurance_power=-0.5,
                                label_pair_to_weight_func=FusedPairWeight("step"))
        self.assertAlmostEqual(loss.item(), 0.5415076, delta=1e-5)
        self.assertEqual(n.item(), 3.0)
        mask = colt([True, True, False, False, False], torch.bool)
        self.assertAlmostEqual(pairwise_loss(logits, label, g, click_occurance_power=-0.5, mask=mask).item(),
                               1.3132617, delta=1e-5)


class TestListwiseLoss(unittest.TestCase):
    def test_listwise_loss(self):
        from rec_now_b200.rec_block.listwise_loss_from_batch import (
            listwise_loss_via_softmax_cross_entropy_with_logits, nan_to_zero, to_listwise_sample)
        sample_group_idx_var = colt([1, 1, 2, 1, 2, 2, 3, 4])
        labels = colt([1, 1, 1, 0, 0, 0, 1, 0])
        logits = colt([0.1, 0.01, 0.2, 0.001, 0.02, 0.002, 0.3, 0.4])
        sample_mask, labels_for_softmax, logits_for_softmax = to_listwise_sample(sample_group_idx_var, labels, logits)
        n_valid_list = labels_for_softmax.shape[0]
        n_sample_per_valid_group = torch.mean(torch.sum(sample_mask.to(torch.float32), dim=-1))
        n_sample_per_valid_group = nan_to_zero(n_sample_per_valid_group)
        listwise_loss = listwise_loss_via_softmax_cross_entropy_with_logits(
            labels_for_softmax=labels_for_softmax, logits_for_softmax=logits_for_softmax)
        self.assertEqual(n_valid_list, 2)
        self.assertAlmostEqual(n_sample_per_valid_group.item(), 3.0, delta=1e-6)
        self.assertAlmostEqual(listwise_loss.item(), 1.0291535, delta=1e-4)

    def test_listwise_loss_case2(self):
        from rec_now_b200.rec_block.listwise_loss_from_batch import (
            listwise_loss_via_softmax_cross_entropy_with_logits, nan_to_zero, to_listwise_sample)
        sample_group_idx_var = colt([3, 4])
        labels = colt([1, 0])
        logits = colt([0.3, 0.4])
        sample_mask, labels_for_softmax, logits_for_softmax = to_listwise_sample(sample_group_idx_var, labels, logits)
        n_valid_list = labels_for_softmax.shape[0]
        n_sample_per_valid_group = torch.mean(torch.sum(sample_mask.to(torch.float32), dim=-1))
        n_sample_per_valid_group = nan_to_zero(n_sample_per_valid_group)
        listwise_loss = listwise_loss_via_softmax_cross_entropy_with_logits(
            labels_for_softmax=labels_for_softmax, logits_for_softmax=logits_for_softmax)
        self.assertEqual(n_valid_list, 0)
        self.assertAlmostEqual(listwise_loss.item(), 0.0, delta=1e-4)
        self.assertEqual(n_sample_per_valid_group.item(), 0.0)


def test_autograd_matches_oracle():
    """Gradient through the drop-in API (fused and general path) vs the float64 oracle."""
    from oracle import generators as G, seg_ref as S
    from rec_now_b200.rec_block.pairwise_loss_from_batch import (
        bpr_loss_func, label_gain_times_sample_weight, pairwise_loss)
    d = G.cfg1(3)
    rng = np.random.default_rng(0)
    y = rng.integers(0, 4, d["s"].size).astype(np.float32)
    w = rng.uniform(0.5, 1.5, y.size).astype(np.float32)
    ref = S.pairwise(d["s"], y, d["g_f32"], S.PairSpec(power=-0.5, label_func="diff", rw_pos=w, factor=2.0))
    g = torch.tensor(d["g_f32"]).cuda().reshape(-1, 1)
    yt, wt = torch.tensor(y).cuda().reshape(-1, 1), torch.tensor(w).cuda().reshape(-1, 1)
    for general in (False, True):
        s = torch.tensor(d["s"]).cuda().reshape(-1, 1).requires_grad_(True)
        if general:
            fn = lambda p, n, wts: bpr_loss_func(p, n, wts, 2.0)                       # unrecognisable wrapper
            lw = lambda a, b, sample_weight: (a - b) * (a > b).float() * sample_weight  # arbitrary callable
        else:
            fn = functools.partial(bpr_loss_func, factor=2.0)
            lw = label_gain_times_sample_weight
        loss, n = pairwise_loss(s, yt, g, fn, return_num_pair=True, click_occurance_power=-0.5,
                                label_pair_to_weight_func=lw, sample_weight=wt)
        (3.0 * loss).backward()
        assert n.item() == float(ref["n_pair"])
        assert abs(loss.item() - ref["loss"]) <= 2e-5 * abs(ref["loss"])
        err = np.abs(s.grad.cpu().numpy().reshape(-1) / 3.0 - ref["grad"])
        assert (err <= 2e-5 * ref["grad_abs"] + 1e-12).all(), (general, err.max())


def test_cpu_tensors_raise():
    from rec_now_b200.rec_block.pairwise_loss_from_batch import pairwise_loss
    t = torch.zeros(4, 1)
    with pytest.raises(RuntimeError):
        pairwise_loss(t, t, t)


def test_reference_test_callables_take_the_fused_path(monkeypatch):
    """the reference's own test passes a wrapper around bpr_loss_func and a `float(y_i > y_j)` weight
    function (TPW:38-39, 51-53).  Both are recognised by what they compute (probed once, cached) and must NOT go through
    pair materialisation; a callable that is none of the fused forms still does."""
    from rec_now_b200 import ops
    from rec_now_b200.rec_block import pairwise_loss_from_batch as PW
    calls = {"n": 0}
    real = ops.pair_indices

    def counting_pair_indices(*a, **k):
        calls["n"] += 1
        return real(*a, **k)

    monkeypatch.setattr(ops, "pair_indices", counting_pair_indices)
    g, logits, label = colt([1, 1, 2, 2, 2]), colt([0, 1, 2, 3, 4]), colt([1.1, 0, 0, 1, 1])

    def pairwise_loss_func(outputs_pos, outputs_neg, weights):
        return PW.bpr_loss_func(outputs_pos, outputs_neg, weights, 1.0)

    def step_weights(label_matrix, label_matrix_transpose, **kwargs):
        return (label_matrix > label_matrix_transpose).to(torch.float32)

    def gain_weights(label_matrix, label_matrix_transpose):
        return (label_matrix - label_matrix_transpose) * (label_matrix > label_matrix_transpose).to(torch.float32)

    loss = PW.pairwise_loss(logits, label, g, pairwise_loss_func, click_occurance_power=-0.5,
                            label_pair_to_weight_func=step_weights)
    assert abs(loss.item() - 0.5415076) < 1e-5 and calls["n"] == 0
    # label-gain weights, a larger batch: same value as the explicit FusedPairWeight
    rng = np.random.default_rng(0)
    b = 3000
    gg = torch.tensor(rng.integers(0, 40, b).astype(np.float32), device="cuda")
    s = torch.tensor(rng.standard_normal(b).astype(np.float32), device="cuda")
    y = torch.tensor(rng.integers(0, 5, b).astype(np.float32), device="cuda")
    a = PW.pairwise_loss(s, y, gg, pairwise_loss_func, label_pair_to_weight_func=gain_weights)
    bb = PW.pairwise_loss(s, y, gg, label_pair_to_weight_func=PW.FusedPairWeight("diff"))
    assert calls["n"] == 0 and abs(a.item() - bb.item()) <= 2e-6 * abs(bb.item())
    # not one of the closed forms: squared label gain.  Labels on the level menu -> the level table, still fused
    # (tests/test_weight_lut_gpu.py); labels off the menu -> materialised pairs; the right value both ways
    sq = lambda lm, lmt: ((lm - lmt) ** 2) * (lm > lmt).to(torch.float32)
    sq_np = lambda lm, lmt: ((lm - lmt) ** 2) * (lm > lmt).astype(np.float32)
    from oracle import dense_ref as D
    c = PW.pairwise_loss(s, y, gg, label_pair_to_weight_func=sq)
    assert calls["n"] == 0
    ref = D.pairwise_loss(s.cpu().numpy(), y.cpu().numpy(), gg.cpu().numpy(), label_pair_to_weight_func=sq_np)
    assert abs(c.item() - float(ref)) <= 1e-4 * abs(float(ref))
    c = PW.pairwise_loss(s, y * 0.5, gg, label_pair_to_weight_func=sq)
    assert calls["n"] == 1
    ref = D.pairwise_loss(s.cpu().numpy(), (y * 0.5).cpu().numpy(), gg.cpu().numpy(), label_pair_to_weight_func=sq_np)
    assert abs(c.item() - float(ref)) <= 1e-4 * abs(float(ref))
    # the recognition can be switched off
    monkeypatch.setenv("RN_PROBE_CALLABLES", "0")
    PW.pairwise_loss(logits, label, g, pairwise_loss_func, click_occurance_power=-0.5)
    assert calls["n"] == 2
