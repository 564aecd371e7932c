"""GPU parity of the sort-free listwise kernel (k_lw_count, csrc/listwise.cu) and of the listwise drop-in's corner
cases (listwise_loss_from_batch.py:89-173 of the reference): NaN -> 0 on the reduced loss, mild masked-logit values."""
import numpy as np
import pytest
import torch

from oracle import dense_ref as D
from oracle import generators as G
from oracle import seg_ref as S
from tests.util import dev

pytestmark = pytest.mark.gpu


def run_counting(g, y, s, th=0.5):
    from rec_now_b200 import ops
    keys, ok = ops.canon_keys(dev(g), inf_is_id=True)
    out = ops.listwise_fwd_bwd(keys[0], dev(y), dev(s), row_ok=ok, pos_neg_th=th)
    assert not out["_sorted"]
    assert ops.last_segmentation_path(out["_scratch"]) == 1
    assert ops.device_error(out["_scratch"]) == 0
    return out


def check(out, ref, tol=1e-5):
    v = int(out["n_valid"].item())
    assert v == ref["n_valid"]
    assert int(out["n_group"].item()) == ref["n_group"]
    loss = float(out["loss"].item())
    assert abs(loss - ref["loss"]) <= tol * abs(ref["loss"]) + 1e-12, (loss, ref["loss"])
    g = out["dlogits"].cpu().numpy().astype(np.float64)
    scale = np.abs(ref["grad"]).max() if v else 1.0
    assert np.abs(g - ref["grad"]).max() <= tol * scale + 1e-12, np.abs(g - ref["grad"]).max()


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_small_vs_both_oracles(seed):
    rng = np.random.default_rng(seed)
    b = 900
    g = rng.integers(0, 70, b).astype(np.float32)
    g[rng.integers(0, b, 6)] = np.nan
    g[rng.integers(0, b, 3)] = np.inf
    y = (rng.random(b) < 0.3).astype(np.float32) * rng.integers(1, 3, b)
    s = rng.standard_normal(b).astype(np.float32) * 3
    for th in (0.5, 1.5):
        out = run_counting(g, y, s, th)
        check(out, S.listwise(g, y, s, pos_neg_th=th))
    dref = D.listwise_full(g, y, s)
    out = run_counting(g, y, s)
    assert dref["n_valid"] == int(out["n_valid"].item())
    assert abs(float(dref["loss"]) - float(out["loss"].item())) < 2e-6 * max(1, abs(float(dref["loss"])))
    assert np.abs(dref["grad"] - out["dlogits"].cpu().numpy()).max() < 1e-6


def test_cfg4_and_repeated_calls_on_one_arena():
    d = G.cfg4(0)
    ref = S.listwise(d["g"], d["y"], d["s"])
    ptrs = set()
    for _ in range(3):
        out = run_counting(d["g"], d["y"], d["s"])
        check(out, ref)
        ptrs.add(out["_scratch"].data_ptr())
    assert len(ptrs) == 1
    d1 = G.cfg4(1)                                   # other data, same arena
    out = run_counting(d1["g"], d1["y"], d1["s"])
    check(out, S.listwise(d1["g"], d1["y"], d1["s"]))
    assert out["_scratch"].data_ptr() in ptrs


def test_degenerate_and_large():
    out = run_counting(np.arange(5, dtype=np.float32), np.ones(5, np.float32), np.ones(5, np.float32))
    assert int(out["n_valid"].item()) == 0 and float(out["loss"].item()) == 0.0
    assert not out["dlogits"].cpu().numpy().any()
    rng = np.random.default_rng(0)
    b = 5000
    y = (rng.random(b) < 0.5).astype(np.float32)
    s = (rng.standard_normal(b) * 10).astype(np.float32)
    g = np.full(b, 7.0, np.float32)
    check(run_counting(g, y, s), S.listwise(g, y, s))
    # several tiles per CTA
    b = 200_000
    gi = rng.integers(0, 9000, b).astype(np.int64) * 977 + 3
    y = (rng.random(b) < 0.2).astype(np.float32)
    s = rng.standard_normal(b).astype(np.float32)
    check(run_counting(gi, y, s), S.listwise(gi, y, s))


def test_nan_to_zero_semantics():
    """LW:170-172: the reduced loss goes through nan_to_zero, so a NaN anywhere in it gives 0 -- and, tf.cond taking
    the constant branch, a zero gradient.  Two ways to get there: a NaN logit in a valid list, and a valid list whose
    labels sum to zero (p = y / 0)."""
    from rec_now_b200.rec_block import listwise_loss_from_batch as LW
    g = np.array([1, 1, 1, 2, 2, 2], np.float32)
    y = np.array([1, 0, 0, 1, 0, 0], np.float32)
    for form in ("counting", "sorted"):
        s = np.array([0.1, np.nan, 0.3, 0.2, 0.1, 0.0], np.float32)
        lg = dev(s).requires_grad_(True)
        _, lab, lgt = LW.to_listwise_sample(dev(g), dev(y), lg)
        w = None if form == "counting" else torch.ones(2, device="cuda")
        loss = LW.listwise_loss_via_softmax_cross_entropy_with_logits(lab, lgt, weights=w)
        loss.backward()
        assert float(loss.item()) == 0.0, form
        assert not lg.grad.cpu().numpy().any(), form
    y0 = np.array([1, -1, 0, 1, 0, 0], np.float32)      # list 1: labels sum to 0 but it has y > th and y < th
    s = np.array([0.1, 0.2, 0.3, 0.2, 0.1, 0.0], np.float32)
    lg = dev(s).requires_grad_(True)
    _, lab, lgt = LW.to_listwise_sample(dev(g), dev(y0), lg)
    loss = LW.listwise_loss_via_softmax_cross_entropy_with_logits(lab, lgt)
    loss.backward()
    assert float(loss.item()) == 0.0 and not lg.grad.cpu().numpy().any()
    dref = D.listwise_full(g, y0, s)
    assert float(dref["loss"]) == 0.0


def test_mild_masked_logit_takes_the_dense_formula():
    """value_of_masked_logit = -10 leaves (B - n) exp(-10) in every softmax denominator (LW:139-140, 167): the segmented
    kernels would drop it, so the drop-in must take the dense formula -- and agree with the dense oracle."""
    from rec_now_b200.rec_block import listwise_loss_from_batch as LW
    rng = np.random.default_rng(2)
    b = 64
    g = rng.integers(0, 6, b).astype(np.float32)
    y = (rng.random(b) < 0.4).astype(np.float32)
    s = rng.standard_normal(b).astype(np.float32)
    lg = dev(s).requires_grad_(True)
    _, lab, lgt = LW.to_listwise_sample(dev(g), dev(y), lg, value_of_masked_logit=-10.0)
    loss = LW.listwise_loss_via_softmax_cross_entropy_with_logits(lab, lgt)
    loss.backward()
    dref = D.listwise_full(g, y, s, value_of_masked_logit=-10.0)
    assert abs(float(loss.item()) - float(dref["loss"])) < 1e-5 * abs(float(dref["loss"]))
    assert np.abs(lg.grad.cpu().numpy() - dref["grad"]).max() < 1e-6
    dref9 = D.listwise_full(g, y, s)
    assert abs(float(dref9["loss"]) - float(dref["loss"])) > 1e-4       # (the two really differ)


@pytest.mark.parametrize("temperature", [0.5, 2.0])
@pytest.mark.parametrize("weighted", [False, True])
def test_listwise_temperature(temperature, weighted):
    """SURVEY 8f N2 (listwise variants): a softmax temperature inside the fused call = the oracle on logits / T, with the
    gradient scaled back to the unscaled logits; both kernels (the sort-free one and, with per-list weights, the sorted
    form)."""
    import torch
    from oracle import generators as G
    from oracle import seg_ref as S
    from rec_now_b200.rec_block import listwise_loss_from_batch as LW
    d = G.cfg4(3, b=20000, cap=128)
    ref = S.listwise(d["g"], d["y"], (d["s"] / np.float32(temperature)).astype(np.float32))
    w = None
    if weighted:
        rng = np.random.default_rng(1)
        wv = rng.uniform(0.5, 1.5, ref["n_valid"]).astype(np.float32)
        ref = S.listwise(d["g"], d["y"], (d["s"] / np.float32(temperature)).astype(np.float32), weights=wv)
        w = torch.tensor(wv, device="cuda")
    lg = torch.tensor(d["s"], device="cuda", requires_grad=True)
    loss, nv = LW.listwise_loss_from_batch(torch.tensor(d["g"], device="cuda"), torch.tensor(d["y"], device="cuda"), lg,
                                           weights=w, temperature=temperature)
    loss.backward()
    assert int(nv.item()) == ref["n_valid"]
    assert abs(loss.item() - ref["loss"]) <= 2e-5 * abs(ref["loss"])
    want = ref["grad"] / temperature
    assert np.abs(lg.grad.cpu().numpy() - want).max() <= 2e-5 * np.abs(want).max()
