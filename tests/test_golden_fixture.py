"""The reference's own known-answer tests (tests/golden/reference_known_answers.json, generated from the literals of
/root/reference/tests/rec_block/test_{pairwise,listwise}_loss_from_batch.py by scripts/make_golden.py):
 - CPU: both oracle restatements reproduce every golden number (this is what pins the oracle);
 - GPU: the product path reproduces them through the drop-in modules (C ABI underneath)."""
import json
import os

import numpy as np
import pytest

from oracle import dense_ref as D
from oracle import seg_ref as S

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "reference_known_answers.json")) as f:
    GOLD = json.load(f)
CASES = {c["name"]: c for c in GOLD["cases"]}
TOL = GOLD["tolerance"]


def col(v, dt=np.float32):
    return np.asarray([v], dtype=dt).T


def test_oracle_occurance_power_weight():
    c = CASES["occurance_power_weight"]
    assert np.allclose(D.occurance_power_weight(np.asarray(c["group_id"]), -1.0), c["power_-1"], atol=TOL)
    assert np.allclose(D.occurance_power_weight(np.asarray(c["group_id"]), 2.0), c["power_2"], atol=TOL)


def test_oracle_pairwise_loss():
    c = CASES["pairwise_loss"]
    g, s, y = col(c["groups"]), col(c["logits"]), col(c["labels"])
    p = c["click_occurance_power"]
    assert abs(float(D.pairwise_loss(s, y, g, click_occurance_power=p)) - c["expected_plain"]) < TOL
    wf = lambda a, b, **kw: (a > b).astype(np.float32)
    assert abs(float(D.pairwise_loss(s, y, g, click_occurance_power=p, label_pair_to_weight_func=wf))
               - c["expected_with_weight_func"]) < TOL
    assert abs(float(D.pairwise_loss(s, y, g, click_occurance_power=p, mask=col(c["mask"], bool)))
               - c["expected_with_mask"]) < TOL
    for mask, key in ((None, "expected_plain"), (np.asarray(c["mask"], bool), "expected_with_mask")):
        r = S.pairwise(s.ravel(), y.ravel(), g.ravel(), S.PairSpec(power=p), mask=mask)
        assert abs(r["loss"] - c[key]) < TOL


@pytest.mark.parametrize("name", ["listwise_loss", "listwise_loss_case2"])
def test_oracle_listwise(name):
    c = CASES[name]
    g, y, s = (np.asarray(c[k], np.float32) for k in ("groups", "labels", "logits"))
    r = S.listwise(g, y, s)
    assert r["n_valid"] == c["expected_n_valid_list"]
    assert abs(r["loss"] - c["expected_loss"]) < TOL
    dm, dl, dz = D.to_listwise_sample(col(g), col(y), col(s))
    assert dl.shape[0] == c["expected_n_valid_list"]
    assert abs(float(D.listwise_loss_via_softmax_cross_entropy_with_logits(dl, dz)) - c["expected_loss"]) < TOL


@pytest.mark.gpu
def test_product_matches_reference_known_answers():
    import torch
    from rec_now_b200.rec_block import listwise_loss_from_batch as LW
    from rec_now_b200.rec_block import pairwise_loss_from_batch as PW
    dev = lambda v, dt=torch.float32: torch.tensor([v], dtype=dt, device="cuda").t()
    c = CASES["occurance_power_weight"]
    assert np.allclose(PW.occurance_power_weight(c["group_id"], power=-1).cpu().numpy(), c["power_-1"], atol=TOL)
    assert np.allclose(PW.occurance_power_weight(c["group_id"], power=2).cpu().numpy(), c["power_2"], atol=TOL)
    c = CASES["pairwise_loss"]
    g, s, y = dev(c["groups"]), dev(c["logits"]), dev(c["labels"])

    def wrapper(outputs_pos, outputs_neg, weights):          # the reference test's wrapper (TPW:38-39)
        return PW.bpr_loss_func(outputs_pos, outputs_neg, weights, 1.0)

    for func in (wrapper, PW.bpr_loss_func):                 # general (materialised pairs) and fused paths
        kw = dict(only_use_wrong_order_pair=False, click_occurance_power=c["click_occurance_power"])
        assert abs(PW.pairwise_loss(s, y, g, func, **kw).item() - c["expected_plain"]) < TOL
        wf = lambda a, b, **k: (a > b).to(torch.float32)
        assert abs(PW.pairwise_loss(s, y, g, func, label_pair_to_weight_func=wf, **kw).item()
                   - c["expected_with_weight_func"]) < TOL
        assert abs(PW.pairwise_loss(s, y, g, func, mask=dev(c["mask"], torch.bool), **kw).item()
                   - c["expected_with_mask"]) < TOL
    for name in ("listwise_loss", "listwise_loss_case2"):
        c = CASES[name]
        m, lab, lgt = LW.to_listwise_sample(dev(c["groups"]), dev(c["labels"]), dev(c["logits"]))
        assert lab.shape[0] == c["expected_n_valid_list"]
        loss = LW.listwise_loss_via_softmax_cross_entropy_with_logits(labels_for_softmax=lab, logits_for_softmax=lgt)
        assert abs(loss.item() - c["expected_loss"]) < TOL
