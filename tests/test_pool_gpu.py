"""GPU parity of the slot-embedding segment pooling (rn_segment_pool_fwd / _bwd) through the C ABI and the
drop-in embedding_using_sparse_batch_segment_ids against oracle/pool_ref.py (op-for-op restatement of
rec_block/embedding_util.py:127-324, pinned to the reference's test literals): forward bit-exact (same accumulation
order as TF's CPU kernels), gradients against float64 sums."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import pool_ref as P

pytestmark = pytest.mark.gpu


def _golden(name):
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_known_answers.json")))
    return [c for c in g["cases"] if c["name"] == name][0]


def test_reference_known_answers():
    """tests/rec_block/test_embedding_util.py:55-109 of the reference on the drop-in."""
    from rec_now_b200.rec_block import embedding_util as EU
    c = _golden("sparse_batch_segment_ids_of_targets")
    mask, sp, nr, ni, ns = EU.sparse_batch_segment_ids_of_targets(torch.tensor(c["slots"], device="cuda"), c["target_slots"])
    assert mask.cpu().tolist() == c["expected_mask"] and sp.cpu().tolist() == c["expected_sp_segment_ids"]
    assert (nr, ni, ns) == (c["num_rows"], c["num_ids"], c["num_segments"])
    c = _golden("embedding_using_sparse_batch_segment_ids")
    params = torch.tensor([[i, -i] for i in range(40)], dtype=torch.float32, device="cuda")
    ids = torch.tensor(c["ids"], device="cuda")
    slots = ((ids.to(torch.float64) + 0.5) / 10.0).to(torch.int32)
    weights = ids.to(torch.float32) * 10.0

    def embedding_func(i):                                     # (an arbitrary callable, as TEU:79)
        return torch.nn.functional.embedding(i, params)
    for f in (embedding_func, EU.TableLookup(params)):
        for uu in (True, False):
            out = EU.embedding_using_sparse_batch_segment_ids(f, slots, c["target_slots"], ids, weights=weights, use_unique=uu)
            assert out.cpu().tolist() == c["expected_with_weights"]
            out = EU.embedding_using_sparse_batch_segment_ids(f, slots, c["target_slots"], ids, use_unique=uu)
            assert out.cpu().tolist() == c["expected_without_weights"]


@pytest.mark.parametrize("b,c,t,d,v", [(64, 12, 3, 2, 50), (500, 40, 8, 16, 1000), (4096, 64, 16, 64, 20000), (300, 7, 5, 33, 97),
                                       (700, 70, 6, 32, 900), (600, 33, 4, 128, 500), (257, 20, 3, 256, 300),
                                       (200, 90, 40, 64, 500)])       # (T > 32: the per-target warp kernel)
@pytest.mark.parametrize("method", ["sum", "mean"])
@pytest.mark.parametrize("weighted", [False, True])
def test_pool_matches_oracle(b, c, t, d, v, method, weighted):
    from rec_now_b200.rec_block import embedding_util as EU
    rng = np.random.default_rng(b * 31 + c + d)
    n_slots = t + 5
    slots = rng.integers(0, n_slots, (b, c)).astype(np.int32)
    target = rng.permutation(n_slots)[:t].tolist()
    ids = rng.integers(0, v, (b, c)).astype(np.int64)
    table = rng.standard_normal((v, d)).astype(np.float32)
    w = rng.uniform(0.1, 2.0, (b, c)).astype(np.float32) if weighted else None
    ref = P.embedding_using_sparse_batch_segment_ids(lambda i: table[np.asarray(i)], slots, target, ids, weights=w,
                                                     method=method, use_unique=False)
    tt = torch.tensor(table, device="cuda", requires_grad=True)
    tw = None if w is None else torch.tensor(w, device="cuda", requires_grad=True)
    out = EU.segment_pool(tt, torch.tensor(slots, device="cuda"), target, torch.tensor(ids, device="cuda"), tw, method)
    got = out.detach().cpu().numpy()
    assert got.shape == ref.shape
    assert np.array_equal(got, ref), f"max diff {np.abs(got - ref).max()}"          # bit-exact: same order, same roundings
    # gradients of sum(out * G) in float64
    G = rng.standard_normal(ref.shape).astype(np.float32)
    (out * torch.tensor(G, device="cuda")).sum().backward()
    lut = {s: k for k, s in enumerate(target)}
    d_table = np.zeros((v, d), np.float64)
    d_w = np.zeros((b, c), np.float64)
    for bi in range(b):
        cnts = {}
        if method == "mean":
            for ci in range(c):
                if slots[bi, ci] in lut:
                    cnts[lut[slots[bi, ci]]] = cnts.get(lut[slots[bi, ci]], 0) + 1
        for ci in range(c):
            k = lut.get(slots[bi, ci])
            if k is None:
                continue
            g = G[bi, k].astype(np.float64) / (cnts[k] if method == "mean" else 1)
            d_table[ids[bi, ci]] += g * (1.0 if w is None else w[bi, ci])
            d_w[bi, ci] = float(table[ids[bi, ci]].astype(np.float64) @ g)
    scale = max(1.0, np.abs(d_table).max())
    assert np.abs(tt.grad.cpu().numpy() - d_table).max() <= 2e-5 * scale
    if weighted:
        assert np.abs(tw.grad.cpu().numpy() - d_w).max() <= 2e-5 * max(1.0, np.abs(d_w).max())


def test_callable_embedding_func_gets_unique_target_ids_once():
    from rec_now_b200.rec_block import embedding_util as EU
    rng = np.random.default_rng(5)
    b, c, v, d = 200, 20, 300, 8
    slots = torch.tensor(rng.integers(0, 6, (b, c)).astype(np.int32), device="cuda")
    ids = torch.tensor(rng.integers(0, v, (b, c)), device="cuda")
    table = torch.tensor(rng.standard_normal((v, d)).astype(np.float32), device="cuda", requires_grad=True)
    seen = []

    def embedding_func(i):
        seen.append(i.detach().cpu().numpy())
        return table[i] * 2.0                                   # (not a plain lookup: autograd continues through it)
    out = EU.embedding_using_sparse_batch_segment_ids(embedding_func, slots, [1, 4], ids)
    assert len(seen) == 1
    want_ids = np.unique(ids.cpu().numpy()[np.isin(slots.cpu().numpy(), [1, 4])])
    assert np.array_equal(np.sort(seen[0]), want_ids)            # only ids of the target slots, each once
    ref = EU.segment_pool(table.detach() * 2.0, slots, [1, 4], ids)
    assert torch.equal(out.detach(), ref)
    out.sum().backward()
    assert table.grad is not None and float(table.grad.abs().sum()) > 0
    # C-ABI argument validation
    import ctypes as C
    from rec_now_b200 import _lib
    assert _lib.lib().rn_segment_pool_fwd(C.byref(_lib.PoolArgs()), None, None, None) == 1
