"""GPU parity of the focal loss drop-in (rec_block/focal_loss.py of the reference: three golden values) and of the fused
joint objective pairwise_loss + focal_weight * focal_crossentropy_loss (rn_pairwise_args.focal_*) against the float64
oracles."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import generators as G
from oracle import seg_ref as S
from tests.util import dev

pytestmark = pytest.mark.gpu


def test_reference_known_answers():
    """tests/rec_block/test_focal_loss.py:17-32 of the reference, (4,1) float32 inputs as there."""
    from rec_now_b200.rec_block.focal_loss import focal_crossentropy_loss
    g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_known_answers.json")))
    c = [x for x in g["cases"] if x["name"] == "focal_crossentropy_loss"][0]
    y = dev(np.asarray(c["labels"], np.float32)).reshape(-1, 1)
    z = dev(np.asarray(c["logits"], np.float32)).reshape(-1, 1)
    assert abs(float(focal_crossentropy_loss(y, z, alpha=None, gamma=None)) - c["expected_alpha_none_gamma_none"]) < 1e-5
    assert abs(float(focal_crossentropy_loss(y, z, alpha=0.25, gamma=None)) - c["expected_alpha_0.25_gamma_none"]) < 1e-5
    assert abs(float(focal_crossentropy_loss(y, z, alpha=None, gamma=1)) - c["expected_alpha_none_gamma_1"]) < 1e-5
    with pytest.raises(ValueError):
        focal_crossentropy_loss(y, z, alpha=1.5)


@pytest.mark.parametrize("kw", [dict(alpha=0.25, gamma=2.0), dict(alpha=None, gamma=1.0), dict(alpha=0.6, gamma=None),
                                dict(alpha=0.25, gamma=2.0, stop_weight_gradient=True)])
def test_fused_joint_loss(kw):
    from rec_now_b200.rec_block.focal_loss import focal_crossentropy_loss, pairwise_loss_with_focal
    from rec_now_b200.rec_block.pairwise_loss_from_batch import pairwise_loss
    d = G.cfg2(1, b=6000, n_groups=300)
    mask = np.random.default_rng(0).random(6000) < 0.9
    fw = 0.7
    logits = dev(d["s"]).requires_grad_(True)
    loss, n = pairwise_loss_with_focal(logits, dev(d["y"]), dev(d["g"]), focal_weight=fw, click_occurance_power=-0.5,
                                       mask=dev(mask), return_num_pair=True, **kw)
    loss.backward()
    rp = S.pairwise(d["s"], d["y"], d["g"], S.PairSpec(power=-0.5), mask=mask)
    rf = S.focal(d["y"], d["s"], alpha=kw.get("alpha") or 0, gamma=kw.get("gamma") or 0,
                 stop_weight_gradient=kw.get("stop_weight_gradient", False))
    assert int(n.item()) == rp["n_pair"]
    ref_loss = rp["loss"] + fw * rf["loss"]
    assert abs(float(loss.item()) - ref_loss) <= 1e-5 * abs(ref_loss)
    ref_grad = rp["grad"] + fw * rf["grad"]
    err = np.abs(logits.grad.cpu().numpy().astype(np.float64) - ref_grad)
    assert (err <= 1e-5 * (rp["grad_abs"] + fw * np.abs(rf["grad"])) + 1e-12).all(), err.max()
    # the same objective from the two unfused drop-ins (autograd through the torch focal loss)
    l2 = dev(d["s"]).requires_grad_(True)
    ref2 = pairwise_loss(l2, dev(d["y"]), dev(d["g"]), click_occurance_power=-0.5, mask=dev(mask)) + \
        fw * focal_crossentropy_loss(dev(d["y"]), l2, **{"alpha": 0.25, "gamma": 2.0, **kw})
    ref2.backward()
    assert abs(float(ref2.item()) - float(loss.item())) <= 2e-6 * abs(float(loss.item()))
    assert np.abs((l2.grad - logits.grad).cpu().numpy()).max() <= 2e-6 * np.abs(ref_grad).max() + 1e-9


def test_fused_joint_loss_graded_weights_and_deterministic():
    from rec_now_b200 import ops
    d = G.cfg3(2, b=20000, n_groups=900)
    keys = dev(d["g"]).reshape(1, -1)
    focal = (0.3, 0.25, 2.0, False)
    outs = []
    for det in (False, True, True):
        out = ops.pairwise_fwd_bwd(dev(d["s"]), dev((d["y"] > 2).astype(np.float32)), keys, rw_pos=dev(d["w"]),
                                   power=-0.5, focal=focal, deterministic=det)
        outs.append((float(out["loss"].item()), out["dlogits"].cpu().numpy()))
    yb = (d["y"] > 2).astype(np.float32)
    rp = S.pairwise(d["s"], yb, d["g"], S.PairSpec(power=-0.5, rw_pos=d["w"]))
    rf = S.focal(yb, d["s"], alpha=0.25, gamma=2.0)
    for loss, g in outs:
        assert abs(loss - (rp["loss"] + 0.3 * rf["loss"])) <= 1e-5 * abs(rp["loss"] + 0.3 * rf["loss"])
        err = np.abs(g.astype(np.float64) - (rp["grad"] + 0.3 * rf["grad"]))
        assert (err <= 1e-5 * (rp["grad_abs"] + 0.3 * np.abs(rf["grad"])) + 1e-12).all()
    assert outs[1][0] == outs[2][0] and np.array_equal(outs[1][1], outs[2][1])      # deterministic mode: same bits
