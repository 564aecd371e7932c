"""The two oracles against each other (no GPU): the op-for-op dense float32 restatement and the segmented
float64 one must agree on masks / counts / pair order exactly and on loss / gradient to float32 accuracy."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from oracle import dense_ref as D
from oracle import generators as G
from oracle import seg_ref as S


def _dense_weight_func(spec, w_pos, w_neg):
    if spec.label_func == "step" and w_pos is None and w_neg is None:
        return None

    def f(a, b):
        w = ((a - b) * (a > b)).astype(np.float32) if spec.label_func == "diff" else (a > b).astype(np.float32)
        if w_pos is not None:
            w = (w * w_pos[:, None]).astype(np.float32)
        if w_neg is not None:
            w = (w * w_neg[None, :]).astype(np.float32)
        return w
    return f


@settings(max_examples=40, deadline=None)
@given(st.integers(0, 10 ** 6), st.integers(1, 120), st.integers(1, 12), st.booleans(), st.booleans(),
       st.sampled_from([0.0, -0.5, 1.0, 2.0]), st.sampled_from(["step", "diff"]), st.integers(0, 3))
def test_dense_vs_segmented(seed, b, ng, wrong, use_mask, power, label_func, wmode):
    rng = np.random.default_rng(seed)
    g = rng.integers(0, ng, b).astype(np.float32)
    s = (rng.standard_normal(b) * 3).astype(np.float32)
    y = rng.integers(0, 4, b).astype(np.float32)
    wp = rng.uniform(-0.2, 1.5, b).astype(np.float32) if wmode & 1 else None
    wn = rng.uniform(-0.2, 1.5, b).astype(np.float32) if wmode & 2 else None
    mask = (rng.random(b) < 0.8) if use_mask else None
    spec = S.PairSpec(factor=1.7, only_wrong=wrong, power=power, label_func=label_func, rw_pos=wp, rw_neg=wn)
    r = S.pairwise(s, y, g, spec, mask=mask, want_pairs=True)
    d = D.pairwise_full(s, y, g, factor=1.7, only_use_wrong_order_pair=wrong, click_occurance_power=power,
                        mask=mask, label_pair_to_weight_func=_dense_weight_func(spec, wp, wn))
    assert r["n_pair"] == d["n_pair"]
    assert np.array_equal(r["pos_idx"], d["pos_idx"]) and np.array_equal(r["neg_idx"], d["neg_idx"])
    assert np.array_equal(np.bincount(d["pos_idx"], minlength=b), r["row_pairs"])
    assert abs(r["loss"] - float(d["loss"])) <= 3e-6 * max(1.0, abs(r["loss"]))
    assert np.abs(r["grad"] - d["grad"]).max() <= 3e-6 * max(1.0, np.abs(r["grad"]).max())


def test_multi_key_and_nonfinite_ids():
    rng = np.random.default_rng(4)
    b = 200
    k0 = rng.integers(0, 6, b).astype(np.float32)
    k1 = rng.integers(0, 3, b).astype(np.float32)
    k0[[3, 50]] = np.nan
    k1[[7]] = np.inf
    s = rng.standard_normal(b).astype(np.float32)
    y = rng.integers(0, 3, b).astype(np.float32)
    r = S.pairwise(s, y, [k0, k1], S.PairSpec(power=-1.0), want_pairs=True)
    d = D.pairwise_full(s, y, [k0, k1], click_occurance_power=-1.0)
    assert r["n_pair"] == d["n_pair"] and np.array_equal(r["pos_idx"], d["pos_idx"])
    assert abs(r["loss"] - float(d["loss"])) < 1e-6


@settings(max_examples=25, deadline=None)
@given(st.integers(0, 10 ** 6), st.integers(1, 150), st.integers(1, 20))
def test_listwise_dense_vs_segmented(seed, b, ng):
    rng = np.random.default_rng(seed)
    g = rng.integers(0, ng, b).astype(np.float32)
    y = ((rng.random(b) < 0.4) * rng.integers(1, 3, b)).astype(np.float32)
    s = (rng.standard_normal(b) * 2).astype(np.float32)
    d = D.listwise_full(g, y, s)
    r = S.listwise(g, y, s)
    assert d["n_valid"] == r["n_valid"]
    assert abs(float(d["loss"]) - r["loss"]) <= 3e-6 * max(1.0, abs(r["loss"]))
    assert np.abs(d["grad"] - r["grad"]).max() <= 3e-6
    if r["n_valid"]:
        assert np.allclose(d["list_loss"], r["list_loss"], rtol=1e-5, atol=1e-6)


def test_generators_shapes_and_estimates():
    d = G.cfg2(0)
    assert d["g"].shape == (16384,) and d["y"].max() == 1.0
    r = S.pairwise(d["s"], d["y"], d["g"])
    assert 1.2e6 < r["n_pair"] < 1.8e6              # SURVEY 8d estimate: n ~ 1.48 M
    d4 = G.cfg4(0)
    assert d4["g"].shape == (65536,) and 1500 < d4["n_lists"] < 2600
    d5 = G.cfg5(2, rows_per_rank=1024, groups_per_rank=64)
    assert d5["g"].shape == (2048,)


def test_listwise_inf_ids_form_one_list():
    """tf.unique (LW:109) compares ids with ==: +inf ids equal each other (one list), NaN ids equal nothing.  The dense
    op-for-op restatement and the segmented one must agree on that."""
    from oracle import dense_ref as D
    from oracle import seg_ref as S
    g = np.array([1, np.inf, 1, np.inf, np.nan, np.nan, -np.inf, np.inf], np.float32)
    y = np.array([1, 1, 0, 0, 1, 0, 1, 0], np.float32)
    s = np.linspace(-1, 1, 8).astype(np.float32)
    d, r = D.listwise_full(g, y, s), S.listwise(g, y, s)
    assert d["n_valid"] == r["n_valid"] == 2                 # {0, 2} and the three +inf rows
    assert abs(float(d["loss"]) - r["loss"]) < 1e-6
    assert np.abs(d["grad"] - r["grad"]).max() < 1e-6


@settings(max_examples=25, deadline=None)
@given(st.integers(0, 10 ** 6), st.integers(2, 100), st.integers(1, 8), st.sampled_from([0.0, 0.5, 1.0]),
       st.sampled_from(["step", "gain2"]), st.sampled_from([0.0, -0.5]))
def test_hinge_and_gain2_dense_vs_segmented(seed, b, ng, margin, label_func, power):
    """the hinge pair loss behind the pairloss_func hook (PW:229, 274) and exponential label gains behind
    label_pair_to_weight_func (PW:175-194) -- the dense op-for-op pipeline with those callables plugged in against the
    segmented float64 form the GPU tests use."""
    import functools
    rng = np.random.default_rng(seed)
    g = rng.integers(0, ng, b).astype(np.float32)
    s = (rng.standard_normal(b) * 2).astype(np.float32)
    y = rng.integers(0, 4, b).astype(np.float32)
    spec = S.PairSpec(factor=1.3, power=power, label_func=label_func, pair_loss="hinge", margin=margin)
    r = S.pairwise(s, y, g, spec)
    wf = None if label_func == "step" else (lambda a, c: ((np.exp2(a) - np.exp2(c)) * (a > c)).astype(np.float32))
    loss, n = D.pairwise_loss(s, y, g, pairloss_func=functools.partial(D.hinge_loss_func, margin=margin, factor=1.3),
                              return_num_pair=True, click_occurance_power=power, label_pair_to_weight_func=wf)
    assert int(n) == r["n_pair"]
    assert abs(float(loss) - r["loss"]) <= 3e-6 * max(1.0, abs(r["loss"]))
    # gradient of the float64 form by central differences on the loss (away from the kinks)
    if r["n_pair"] and b <= 40:
        eps = 1e-3
        for i in rng.choice(b, size=min(b, 4), replace=False):
            sp, sm = s.astype(np.float64).copy(), s.astype(np.float64).copy()
            sp[i] += eps; sm[i] -= eps
            # (float32 inputs are part of the contract: skip rows whose +-eps move crosses a hinge kink)
            lp = S.pairwise(sp.astype(np.float32), y, g, spec)["loss"]
            lm = S.pairwise(sm.astype(np.float32), y, g, spec)["loss"]
            fd = (lp - lm) / (float(np.float32(sp[i]) - np.float32(sm[i])))
            if abs(fd - r["grad"][i]) > 0.05 * max(1.0, abs(r["grad"][i])):
                # a kink inside [s - eps, s + eps]: the one-sided slopes must bracket the analytic value
                l0 = r["loss"]
                f1 = (lp - l0) / float(np.float32(sp[i]) - s[i]); f2 = (l0 - lm) / float(s[i] - np.float32(sm[i]))
                assert min(f1, f2) - 0.05 <= r["grad"][i] <= max(f1, f2) + 0.05


@settings(max_examples=20, deadline=None)
@given(seed=st.integers(0, 10_000), b=st.integers(2, 90), ng=st.integers(1, 9), wrong=st.booleans(),
       power=st.sampled_from([0.0, -0.5, 1.0]))
def test_callable_weight_func_dense_vs_segmented(seed, b, ng, wrong, power):
    """The segmented oracle with an arbitrary label_pair_to_weight_func plugged in (label_func "callable": the truth for
    the level-table path RN_LABEL_LUT) against the op-for-op dense pipeline running the same callable (PW:192-193)."""
    rng = np.random.default_rng(seed)
    g = rng.integers(0, ng, b).astype(np.float32)
    s = rng.standard_normal(b).astype(np.float32)
    y = rng.integers(-1, 7, b).astype(np.float32)
    f = lambda a, c: (((a - c) ** 2 + 0.5 * a + 1.0) * (a > c)).astype(np.float32)
    dense, n = D.pairwise_loss(s, y, g, only_use_wrong_order_pair=wrong, return_num_pair=True,
                               click_occurance_power=power, label_pair_to_weight_func=f)
    seg = S.pairwise(s, y, g, S.PairSpec(label_func="callable", weight_func=f, only_wrong=wrong, power=power))
    assert seg["n_pair"] == int(n)
    assert abs(seg["loss"] - float(dense)) <= 2e-5 * max(abs(float(dense)), 1e-6)


@settings(max_examples=15, deadline=None)
@given(seed=st.integers(0, 10_000), b=st.integers(2, 60), ng=st.integers(1, 5))
def test_lambdarank_weights_are_the_swap_delta_ndcg(seed, b, ng):
    """The oracle's LambdaRank weight of a pair is |NDCG(ranking) - NDCG(ranking with rows i and j swapped)| of its group
    (gains 2^y - 1, discounts 1 / log2(1 + rank), ranks by score descending with ties by row index) -- the definition
    include/recnow_b200.h gives for RN_LABEL_LAMBDA, restated here by brute force."""
    rng = np.random.default_rng(seed)
    g = rng.integers(0, ng, b).astype(np.float32)
    s = np.round(rng.standard_normal(b), 1).astype(np.float32)              # (rounded: tied scores occur)
    y = rng.integers(0, 5, b).astype(np.float32)
    res = S.pairwise(s, y, g, S.PairSpec(label_func="lambda"), want_pairs=True)
    for i, j, w in zip(res["pos_idx"], res["neg_idx"], res["w"]):
        m = np.flatnonzero(g == g[i])
        order = np.lexsort((m, -s[m].astype(np.float64)))
        rank = np.empty(m.size)
        rank[order] = np.arange(1, m.size + 1)
        gain = 2.0 ** y[m].astype(np.float64) - 1.0
        idcg = (np.sort(gain)[::-1] / np.log2(1.0 + np.arange(1, m.size + 1))).sum()
        ii, jj = int(np.flatnonzero(m == i)[0]), int(np.flatnonzero(m == j)[0])
        swapped = rank.copy()
        swapped[ii], swapped[jj] = rank[jj], rank[ii]
        want = abs((gain / np.log2(1.0 + swapped)).sum() - (gain / np.log2(1.0 + rank)).sum()) / idcg
        assert abs(w - want) <= 2e-5 * want + 1e-9, (w, want)
