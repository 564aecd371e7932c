"""GPU parity: listwise kernels, pair materialisation, occurrence weights vs the oracles."""
import numpy as np
import pytest
import torch

from oracle import dense_ref as D
from oracle import generators as G
from oracle import seg_ref as S
from tests.util import dev

pytestmark = pytest.mark.gpu


def run_listwise(g, y, s, list_w=None, th=0.5, do_reduce=True):
    from rec_now_b200 import ops
    keys, ok = ops.canon_keys(dev(g), inf_is_id=True)
    return ops.listwise_fwd_bwd(keys[0], dev(y), dev(s), row_ok=ok, list_w=None if list_w is None else dev(list_w),
                                pos_neg_th=th, do_reduce=do_reduce, want_list_loss=True)


def check_listwise(out, ref, tol=1e-5):
    v = int(out["n_valid"].item())
    assert v == ref["n_valid"]
    assert int(out["n_group"].item()) == ref["n_group"]
    ll = out["list_loss"].cpu().numpy()[:v].astype(np.float64)
    assert np.allclose(ll, ref["list_loss"], rtol=tol, atol=1e-7), np.abs(ll - ref["list_loss"]).max()
    loss = float(out["loss"].item())
    assert abs(loss - ref["loss"]) <= tol * abs(ref["loss"]) + 1e-12, (loss, ref["loss"])
    g = out["dlogits"].cpu().numpy().astype(np.float64)
    scale = np.abs(ref["grad"]).max() if v else 1.0
    assert np.abs(g - ref["grad"]).max() <= tol * scale + 1e-12, np.abs(g - ref["grad"]).max()


@pytest.mark.parametrize("seed", [0, 1])
def test_listwise_small_vs_both_oracles(seed):
    rng = np.random.default_rng(seed)
    b = 700
    g = rng.integers(0, 60, b).astype(np.float32)
    g[rng.integers(0, b, 5)] = np.nan
    y = (rng.random(b) < 0.3).astype(np.float32) * rng.integers(1, 3, b)
    s = rng.standard_normal(b).astype(np.float32) * 3
    out = run_listwise(g, y, s)
    ref = S.listwise(g, y, s)
    check_listwise(out, ref)
    dref = D.listwise_full(g, y, s)                    # op-for-op dense restatement
    assert dref["n_valid"] == int(out["n_valid"].item())
    assert abs(float(dref["loss"]) - float(out["loss"].item())) < 2e-6 * max(1, abs(float(dref["loss"])))
    assert np.abs(dref["grad"] - out["dlogits"].cpu().numpy()).max() < 1e-6
    # per-list weights, un-reduced output
    w = rng.uniform(0.5, 2.0, ref["n_valid"]).astype(np.float32)
    out = run_listwise(g, y, s, list_w=w)
    check_listwise(out, S.listwise(g, y, s, weights=w))
    out = run_listwise(g, y, s, list_w=w, do_reduce=False, th=1.5)
    r2 = S.listwise(g, y, s, weights=w[:S.listwise(g, y, s, pos_neg_th=1.5)["n_valid"]], pos_neg_th=1.5)
    v = int(out["n_valid"].item())
    assert v == r2["n_valid"]
    assert np.allclose(out["list_loss"].cpu().numpy()[:v], r2["list_loss"], rtol=1e-5, atol=1e-7)


def test_listwise_dense_layout():
    from rec_now_b200 import ops
    rng = np.random.default_rng(3)
    b = 300
    g = rng.integers(0, 40, b).astype(np.float32)
    y = (rng.random(b) < 0.3).astype(np.float32)
    s = rng.standard_normal(b).astype(np.float32)
    out = run_listwise(g, y, s)
    v = int(out["n_valid"].item())
    for do_mask in (True, False):
        dm, dl, dz = ops.listwise_dense(out, v, do_mask, -1e9)
        rm, rl, rz = D.to_listwise_sample(g, y, s, do_mask_logits=do_mask)
        assert np.array_equal(dm.cpu().numpy(), rm)
        assert np.array_equal(dl.cpu().numpy(), rl)       # bit-exact: same float32 divide
        assert np.array_equal(dz.cpu().numpy(), rz)


def test_listwise_cfg4():
    d = G.cfg4(0)
    out = run_listwise(d["g"], d["y"], d["s"])
    check_listwise(out, S.listwise(d["g"], d["y"], d["s"]))
    print("cfg4 lists", d["n_lists"], "valid", int(out["n_valid"].item()))


def test_listwise_degenerate():
    out = run_listwise(np.arange(5, dtype=np.float32), np.ones(5, np.float32), np.ones(5, np.float32))
    assert int(out["n_valid"].item()) == 0 and float(out["loss"].item()) == 0.0
    assert not out["dlogits"].cpu().numpy().any()
    # one list holding the whole batch
    rng = np.random.default_rng(0)
    b = 5000
    y = (rng.random(b) < 0.5).astype(np.float32)
    s = (rng.standard_normal(b) * 10).astype(np.float32)
    g = np.full(b, 7.0, np.float32)
    check_listwise(run_listwise(g, y, s), S.listwise(g, y, s))


@pytest.mark.parametrize("label_cond", [True, False])
def test_pair_indices_row_major(label_cond):
    from rec_now_b200 import ops
    rng = np.random.default_rng(11)
    b = 1500
    g = G.zipf_groups(rng, b, 40).astype(np.float32)
    y = rng.integers(0, 3, b).astype(np.float32)
    s = rng.standard_normal(b).astype(np.float32)
    w = rng.uniform(0.5, 1.5, b).astype(np.float32)
    mask = rng.random(b) < 0.9
    keys, ok = ops.canon_keys(dev(g), dev(mask))
    if label_cond:
        pos, neg, wt = ops.pair_indices(dev(s), dev(y), keys, row_ok=ok, rw_pos=dev(w), label_func="diff",
                                        only_wrong=True, label_cond=True, want_weights=True)
        ref = D.pairwise_full(s, y, g, only_use_wrong_order_pair=True, mask=mask,
                              label_pair_to_weight_func=lambda a, c: ((a - c) * (a > c) * w[:, None]).astype(np.float32))
        assert np.array_equal(wt.cpu().numpy(), ref["weights"])
    else:
        pos, neg, _ = ops.pair_indices(dev(s), dev(y), keys, row_ok=ok, label_cond=False)
        ref = D.pairwise_full(s, y, g, mask=mask, label_pair_to_weight_func=lambda a, c: np.ones_like(a))
    assert np.array_equal(pos.cpu().numpy(), ref["pos_idx"])      # bit-exact, row-major (PW:217)
    assert np.array_equal(neg.cpu().numpy(), ref["neg_idx"])


def test_occurrence_power_weight():
    from rec_now_b200.rec_block.pairwise_loss_from_batch import occurance_power_weight
    rng = np.random.default_rng(2)
    ids = rng.integers(0, 500, 20000)
    for p in (-1.0, -0.5, 0.0, 1.0, 2.0):
        out = occurance_power_weight(torch.tensor(ids).cuda(), power=p).cpu().numpy()
        assert np.allclose(out, D.occurance_power_weight(ids, p), rtol=2e-6)
