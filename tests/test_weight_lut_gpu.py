"""Level weight table (RN_LABEL_LUT, SURVEY 8b `weight_lut` / hard part 5): ANY label-only
label_pair_to_weight_func (PW:175-194) stays on the fused path as an 8 x 8 table over the label levels -1 .. 6.

Parity: the C-ABI call against the float64 segmented oracle with the callable itself plugged in (oracle label_func
"callable": W = f(Y, Yt), C = W > 0, PW:192-193); the drop-in against the same oracle, on the fused path for labels on the
menu and on the materialised-pair path otherwise.  Tolerances as SURVEY 8d: exact counts, loss 1e-5 relative, gradient
1e-5 of the per-row scale A_i."""
import numpy as np
import pytest
import torch

from oracle import generators as G
from oracle import seg_ref as S
from tests.util import check_pairwise, dev

pytestmark = pytest.mark.gpu


def np_weight(a, b):
    """A weight function that is none of the closed forms: squared label gain plus a level-dependent offset."""
    return (((a - b) ** 2 + 0.5 * a + 1.0) * (a > b)).astype(np.float32)


def torch_weight(a, b):
    return ((a - b) ** 2 + 0.5 * a + 1.0) * (a > b).to(torch.float32)


def table_of(f=np_weight):
    lev = np.arange(-1, 7, dtype=np.float32)
    return f(np.broadcast_to(lev[:, None], (8, 8)), np.broadcast_to(lev[None, :], (8, 8)))


def run_lut(s, y, groups, table, spec=S.PairSpec(), mask=None):
    from rec_now_b200 import ops
    cols = groups if isinstance(groups, list) else [groups]
    keys, ok = ops.canon_keys([dev(c) for c in cols], None if mask is None else dev(np.asarray(mask, bool)))
    return ops.pairwise_fwd_bwd(dev(s), dev(y), keys, row_ok=ok, rw_pos=None if spec.rw_pos is None else dev(spec.rw_pos),
                                label_func="lut", weight_lut=dev(np.asarray(table, np.float32)), factor=spec.factor,
                                power=spec.power, reduce_mean=spec.reduce_mean, want_row_pairs=True,
                                pair_loss=spec.pair_loss, margin=spec.margin)


def oracle_spec(**kw):
    return S.PairSpec(label_func="callable", weight_func=np_weight, **kw)


@pytest.mark.parametrize("b,ng", [(300, 7), (5000, 40), (16384, 1024)])
def test_table_matches_oracle(b, ng):
    rng = np.random.default_rng(b)
    g = rng.integers(0, ng, b).astype(np.float32)
    s = rng.standard_normal(b).astype(np.float32)
    y = rng.integers(-1, 7, b).astype(np.float32)                  # all eight levels
    out = run_lut(s, y, g, table_of())
    check_pairwise(out, S.pairwise(s, y, g, oracle_spec()), ctx=f"lut B={b}")
    from rec_now_b200 import ops
    assert ops.last_segmentation_path(out["_scratch"]) == (3 if b <= 1024 else 1)      # (the one-CTA kernel takes the table too)


def test_table_small_batch_kernel_options_and_errors():
    """The one-CTA kernel of batches up to 1024 rows with the level table: row weights, occurrence power, hinge, mask; a label
    off the menu or an unusable table entry fails the call there as well."""
    from rec_now_b200 import ops
    rng = np.random.default_rng(31)
    b = 900
    g = rng.integers(0, 12, b).astype(np.float32)
    s = rng.standard_normal(b).astype(np.float32)
    y = rng.integers(0, 5, b).astype(np.float32)
    w = rng.uniform(0.5, 1.5, b).astype(np.float32)
    mask = rng.random(b) < 0.9
    for kw in (dict(rw_pos=w, power=-0.5), dict(pair_loss="hinge", margin=0.5, factor=2.0, reduce_mean=False)):
        spec = oracle_spec(**kw)
        out = run_lut(s, y, g, table_of(), spec, mask=mask)
        check_pairwise(out, S.pairwise(s, y, g, spec, mask=mask), ctx=f"small lut {sorted(kw)}")
        assert ops.last_segmentation_path(out["_scratch"]) == 3
    y_off = y.copy(); y_off[5] = 2.5
    out = run_lut(s, y_off, g, table_of())
    assert np.isnan(float(out["loss"])) and ops.device_error(out["_scratch"]) & 8
    bad = table_of().copy(); bad[4, 2] = -1.0
    out = run_lut(s, y, g, bad)
    assert np.isnan(float(out["loss"])) and ops.device_error(out["_scratch"]) & 8
    out = run_lut(s, y, g, table_of())
    check_pairwise(out, S.pairwise(s, y, g, oracle_spec()), ctx="small lut after failed calls")


def test_table_cfg3_full_size_with_row_weights_and_power():
    """BASELINE cfg3's batch (B = 65 536, graded labels, per-sample weights, power -0.5) with the table in place of the
    label gain; counting segmentation."""
    from rec_now_b200 import ops
    c = G.cfg3()
    spec = oracle_spec(rw_pos=c["w"], power=-0.5)
    out = run_lut(c["s"], c["y"], c["g_f32"], table_of(), spec)
    check_pairwise(out, S.pairwise(c["s"], c["y"], c["g_f32"], spec), ctx="lut cfg3")
    assert ops.last_segmentation_path(out["_scratch"]) == 1


def test_table_equal_to_label_gain_reproduces_diff():
    """A table holding y_i - y_j gives what RN_LABEL_DIFF gives (same tiles, the weight looked up instead of subtracted)."""
    from tests.util import run_pairwise
    c = G.cfg3(seed=1, b=20000, n_groups=500)
    diff = lambda a, b: ((a - b) * (a > b)).astype(np.float32)
    a = run_lut(c["s"], c["y"], c["g_f32"], table_of(diff))
    d = run_pairwise(c["s"], c["y"], c["g_f32"], S.PairSpec(label_func="diff"))
    assert int(a["n_pair"]) == int(d["n_pair"])
    assert abs(float(a["loss"]) - float(d["loss"])) <= 2e-6 * abs(float(d["loss"]))
    ga, gd = a["dlogits"].cpu().numpy(), d["dlogits"].cpu().numpy()
    assert np.abs(ga - gd).max() <= 1e-5 * np.abs(gd).max()


def test_table_hinge_factor_sum_and_mask():
    rng = np.random.default_rng(5)
    b = 9000
    g = rng.integers(0, 60, b).astype(np.float32)
    s = rng.standard_normal(b).astype(np.float32)
    y = rng.integers(0, 5, b).astype(np.float32)
    mask = rng.random(b) < 0.8
    spec = oracle_spec(pair_loss="hinge", margin=0.7, factor=1.5, reduce_mean=False)
    out = run_lut(s, y, g, table_of(), spec, mask=mask)
    check_pairwise(out, S.pairwise(s, y, g, spec, mask=mask), ctx="lut hinge")
    spec = oracle_spec(factor=0.5)
    out = run_lut(s, y, g, table_of(), spec, mask=mask)
    check_pairwise(out, S.pairwise(s, y, g, spec, mask=mask), ctx="lut factor")


def test_table_two_key_columns_radix_path():
    """Several key columns take the radix segmentation: the label levels are worked out in its tail."""
    from rec_now_b200 import ops
    rng = np.random.default_rng(6)
    b = 12000
    g1, g2 = rng.integers(0, 30, b).astype(np.float32), rng.integers(0, 4, b).astype(np.float32)
    s = rng.standard_normal(b).astype(np.float32)
    y = rng.integers(0, 5, b).astype(np.float32)
    y[rng.random(b) < 0.01] = np.nan                                 # NaN labels pair with nothing
    out = run_lut(s, y, [g1, g2], table_of())
    check_pairwise(out, S.pairwise(s, y, [g1, g2], oracle_spec()), ctx="lut K=2")
    assert ops.last_segmentation_path(out["_scratch"]) == 2


def test_label_off_the_menu_or_bad_table_fails_the_call():
    """A label without a level, or a table entry that would change the pair set, fails the call on the device: loss = NaN
    and rn_last_device_error bit 8 -- never a plausible number."""
    from rec_now_b200 import ops
    rng = np.random.default_rng(7)
    b = 4000
    g = rng.integers(0, 20, b).astype(np.float32)
    s = rng.standard_normal(b).astype(np.float32)
    y = rng.integers(0, 5, b).astype(np.float32)
    y_off = y.copy(); y_off[17] = 0.5
    out = run_lut(s, y_off, g, table_of())
    assert np.isnan(float(out["loss"]))
    assert ops.device_error(out["_scratch"]) & 8
    bad = table_of().copy(); bad[3, 1] = 0.0
    out = run_lut(s, y, g, bad)
    assert np.isnan(float(out["loss"]))
    assert ops.device_error(out["_scratch"]) & 8
    # ... and the arena is still good for the next call
    out = run_lut(s, y, g, table_of())
    check_pairwise(out, S.pairwise(s, y, g, oracle_spec()), ctx="lut after a failed call")
    # host-side validation: the table and the label function go together; no table variant of the score-dependent pair set
    keys, _ = ops.canon_keys([dev(g)])
    with pytest.raises(Exception):
        ops.pairwise_fwd_bwd(dev(s), dev(y), keys, label_func="lut")
    with pytest.raises(Exception):
        ops.pairwise_fwd_bwd(dev(s), dev(y), keys, label_func="lut", weight_lut=dev(table_of()), only_wrong=True)


def test_dropin_arbitrary_label_callable_takes_the_fused_path(monkeypatch):
    """pairwise_loss(label_pair_to_weight_func=<any label-only function>) no longer materialises pairs when the labels are
    on the level menu; off the menu (or with the wrong-order filter) it still does, with the caller's own function."""
    from rec_now_b200 import ops
    from rec_now_b200.rec_block import pairwise_loss_from_batch as PW
    calls = {"n": 0}
    real = ops.pair_indices

    def counting_pair_indices(*a, **k):
        calls["n"] += 1
        return real(*a, **k)

    monkeypatch.setattr(ops, "pair_indices", counting_pair_indices)
    rng = np.random.default_rng(8)
    b = 6000
    g = rng.integers(0, 50, b).astype(np.float32)
    s = rng.standard_normal(b).astype(np.float32)
    y = rng.integers(0, 5, b).astype(np.float32)
    ts = torch.tensor(s, device="cuda", requires_grad=True)
    loss, n = PW.pairwise_loss(ts, dev(y), dev(g), label_pair_to_weight_func=torch_weight, return_num_pair=True,
                               click_occurance_power=-0.5)
    loss.backward()
    assert calls["n"] == 0
    ref = S.pairwise(s, y, g, oracle_spec(power=-0.5))
    assert float(n) == float(np.float32(ref["n_pair"]))
    assert abs(float(loss) - ref["loss"]) <= 1e-5 * abs(ref["loss"])
    err = np.abs(ts.grad.cpu().numpy().astype(np.float64) - ref["grad"])
    assert (err <= 1e-5 * ref["grad_abs"] + 1e-12).all()
    # labels off the menu: the general path with the caller's function, same oracle
    y2 = y + 0.5
    l2 = PW.pairwise_loss(dev(s), dev(y2), dev(g), label_pair_to_weight_func=torch_weight)
    assert calls["n"] == 1
    r2 = S.pairwise(s, y2, g, oracle_spec())
    assert abs(float(l2) - r2["loss"]) <= 1e-4 * abs(r2["loss"])
    # the wrong-order filter makes the pair set score dependent: general path as well
    l3 = PW.pairwise_loss(dev(s), dev(y), dev(g), label_pair_to_weight_func=torch_weight, only_use_wrong_order_pair=True)
    assert calls["n"] == 2
    r3 = S.pairwise(s, y, g, oracle_spec(only_wrong=True))
    assert abs(float(l3) - r3["loss"]) <= 1e-4 * abs(r3["loss"])
    # the explicit table object with per-sample weights on the positive side
    w = rng.uniform(0.5, 1.5, b).astype(np.float32)
    fw = PW.FusedPairWeight.from_callable(torch_weight, pos_kw="sample_weight")
    l4 = PW.pairwise_loss(dev(s), dev(y), dev(g), label_pair_to_weight_func=fw, sample_weight=dev(w))
    assert calls["n"] == 2
    r4 = S.pairwise(s, y, g, oracle_spec(rw_pos=w))
    assert abs(float(l4) - r4["loss"]) <= 1e-5 * abs(r4["loss"])
