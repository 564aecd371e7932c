"""Host-side logic of the drop-in modules that needs no GPU: which callables are recognised as fused forms, the torch
restatements of the pair-loss callables against the NumPy oracle, argument validation, and the no-CPU-fallback rule."""
import functools

import numpy as np
import pytest
import torch

from oracle import dense_ref as D
from rec_now_b200.rec_block import pairwise_loss_from_batch as PW


def test_match_bpr_and_hinge_partials():
    assert PW._match_bpr(PW.bpr_loss_func) == (1.0, True, None)
    assert PW._match_bpr(functools.partial(PW.bpr_loss_func, factor=2.0, reduce_mean=False)) == (2.0, False, None)
    assert PW._match_bpr(PW.hinge_loss_func) == (1.0, True, 1.0)
    assert PW._match_bpr(functools.partial(PW.hinge_loss_func, margin=0.25, factor=3.0)) == (3.0, True, 0.25)
    assert PW._match_bpr(functools.partial(PW.hinge_loss_func, margin=-1.0)) is None            # (margin must be >= 0)
    assert PW._match_bpr(functools.partial(PW.bpr_loss_func, 1.0)) is None                      # (positional partial)
    assert PW._match_bpr(lambda p, n, w: PW.bpr_loss_func(p, n, w)) is None                     # (wrappers: by probing, on the GPU)


def test_pair_loss_callables_match_the_numpy_oracle():
    rng = np.random.default_rng(0)
    pos, neg = rng.standard_normal(257).astype(np.float32) * 2, rng.standard_normal(257).astype(np.float32) * 2
    w = rng.uniform(0.1, 2.0, 257).astype(np.float32)
    tp, tn, tw = torch.tensor(pos), torch.tensor(neg), torch.tensor(w)
    for kw in (dict(), dict(factor=2.5), dict(reduce_mean=False)):
        for wt, wn in ((None, None), (tw, w)):
            a = float(PW.bpr_loss_func(tp, tn, wt, **kw))
            b = float(D.bpr_loss_func(pos, neg, wn, **kw))
            assert abs(a - b) <= 2e-6 * max(1.0, abs(b))
            for margin in (0.0, 0.5, 1.0):
                a = float(PW.hinge_loss_func(tp, tn, wt, margin=margin, **kw))
                b = float(D.hinge_loss_func(pos, neg, wn, margin=margin, **kw))
                assert abs(a - b) <= 2e-6 * max(1.0, abs(b))


def test_fused_pair_weight_is_a_reference_style_callable():
    """FusedPairWeight works as label_pair_to_weight_func(label_matrix, label_matrix_transpose, **kwargs) (PW:175-194)."""
    y = torch.tensor([0.0, 1.0, 3.0, 1.0])
    ym, ymt = y.reshape(-1, 1).expand(-1, 4), y.reshape(1, -1).expand(4, -1)
    w = torch.tensor([1.0, 2.0, 3.0, 4.0])
    step = PW.FusedPairWeight("step")(ym, ymt)
    assert step.tolist() == (ym > ymt).float().tolist()
    diff = PW.FusedPairWeight("diff", pos_kw="sw")(ym, ymt, sw=w)
    assert torch.equal(diff, (ym - ymt) * (ym > ymt).float() * w.reshape(-1, 1))
    gain = PW.FusedPairWeight("gain2", neg_kw="sw")(ym, ymt, sw=w)
    assert torch.equal(gain, (torch.exp2(ym) - torch.exp2(ymt)) * (ym > ymt).float() * w.reshape(1, -1))
    with pytest.raises(ValueError):
        PW.FusedPairWeight("cube")


def test_level_table_form_of_label_only_weight_functions():
    """VERDICT r1 item 7 / SURVEY 8b weight_lut: ANY label-only label_pair_to_weight_func whose pair set is y_i > y_j has a
    level-table form (evaluated once on the label levels -1 .. 6); functions with another pair set, functions that are not
    elementwise and bad tables have none."""
    sq = lambda a, b: ((a - b) ** 2 + 0.5 * a + 1.0) * (a > b).float()
    f = PW._table_of_weight_func(sq, {}, "cpu")
    assert f is not None and f.label_func == "lut"
    lev = torch.tensor(PW.FusedPairWeight.LEVEL_LABELS)
    want = sq(lev.reshape(-1, 1).expand(8, 8), lev.reshape(1, -1).expand(8, 8))
    assert torch.equal(f.table, want)
    # the table object is a reference-style callable: (B, B) label matrices and pair vectors, labels off the menu -> 0
    y = torch.tensor([0.0, 1.0, 2.0, 5.0, 7.0, 0.5, -1.0])
    ym, ymt = y.reshape(-1, 1).expand(-1, 7), y.reshape(1, -1).expand(7, -1)
    on = ((y == y.round()) & (y <= 6)).float()
    assert torch.equal(f(ym, ymt), sq(ym, ymt) * on.reshape(-1, 1) * on.reshape(1, -1))
    assert torch.equal(f(y[[3, 2, 1]], y[[0, 6, 2]]), sq(y[[3, 2, 1]], y[[0, 6, 2]]))
    w = torch.arange(1.0, 8.0)
    g = PW.FusedPairWeight.from_callable(sq, pos_kw="sw")
    assert torch.equal(g(ym, ymt, sw=w), f(ym, ymt) * w.reshape(-1, 1))
    # pair set is not y_i > y_j: symmetric weights, weights with holes, negative gains
    assert PW._table_of_weight_func(lambda a, b: (a - b).abs(), {}, "cpu") is None
    assert PW._table_of_weight_func(lambda a, b: (a - b - 1.0) * (a > b).float(), {}, "cpu") is None
    assert PW._table_of_weight_func(lambda a, b: (b - a) * (a > b).float(), {}, "cpu") is None
    # not elementwise in the labels (depends on the position)
    pos_dep = lambda a, b: (a > b).float() * (1.0 + torch.arange(a.numel(), dtype=torch.float32).reshape(a.shape))
    assert PW._table_of_weight_func(pos_dep, {}, "cpu") is None
    with pytest.raises(ValueError):
        PW.FusedPairWeight("lut", table=torch.zeros(8, 8))
    with pytest.raises(ValueError):
        PW.FusedPairWeight("lut", table=torch.ones(8, 8), neg_kw="w")


def test_lambdarank_weight_object():
    """FusedPairWeight("lambda") selects LambdaRank weights on the fused path; it is not a function of the label matrices
    (|delta NDCG| needs the scores' ranks), so calling it the reference's way says so instead of returning something."""
    f = PW.FusedPairWeight("lambda", pos_kw="sample_weight")
    assert f.label_func == "lambda" and f.pos_kw == "sample_weight"
    y = torch.tensor([0.0, 1.0, 2.0])
    with pytest.raises(NotImplementedError):
        f(y.reshape(-1, 1).expand(3, 3), y.reshape(1, -1).expand(3, 3))
    with pytest.raises(ValueError):
        PW.FusedPairWeight("lambda", neg_kw="w")


def test_no_cpu_fallback_anywhere():
    from rec_now_b200.rec_block import embedding_util as EU
    from rec_now_b200.rec_block import listwise_loss_from_batch as LW
    s, y, g = torch.zeros(4), torch.tensor([1.0, 0, 1, 0]), torch.tensor([1.0, 1, 2, 2])
    with pytest.raises(RuntimeError, match="CUDA"):
        PW.pairwise_loss(s, y, g)
    with pytest.raises(RuntimeError, match="CUDA"):
        LW.to_listwise_sample(g, y, s)
    with pytest.raises(RuntimeError, match="CUDA"):
        EU.segment_pool(torch.zeros(4, 2), torch.zeros(2, 2, dtype=torch.int32), [0], torch.zeros(2, 2, dtype=torch.int64))
    with pytest.raises(RuntimeError, match="CUDA"):
        EU.embedding_using_sparse_batch_segment_ids(lambda i: i, torch.zeros(2, 2, dtype=torch.int32), [0],
                                                    torch.zeros(2, 2, dtype=torch.int64))


def test_pool_abi_argument_validation():
    import ctypes as C
    from rec_now_b200 import _lib
    lib = _lib.lib()
    a = _lib.PoolArgs()
    assert lib.rn_segment_pool_fwd(C.byref(a), None, None, None) == 1            # RN_ERR_ARG: empty struct
    assert lib.rn_segment_pool_bwd(C.byref(a), None, None, None, None) == 1
    a = _lib.PoolArgs(B=4, C=3, T=2, D=8, slots=16, ids=16, target_slots=16, table=16, V=10)
    assert lib.rn_segment_pool_fwd(C.byref(a), None, None, None) == 1            # no output buffer
    assert lib.rn_segment_pool_bwd(C.byref(a), 16, None, None, None) == 1        # neither gradient requested
