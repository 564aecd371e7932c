"""GPU parity: rn_pairwise_fwd_bwd (through the C ABI) vs the float64 segmented oracle.

Bars (BASELINE.json north_star): pair counts bit-exact, fp32 loss and gradient within 1e-5 relative.
"""
import numpy as np
import pytest

from oracle import generators as G
from oracle import seg_ref as S
from tests.util import check_pairwise, dev, run_pairwise

pytestmark = pytest.mark.gpu


def col(v, dt=np.float32):
    return np.asarray([v], dtype=dt).T


def test_reference_known_answers():
    """tests/rec_block/test_pairwise_loss_from_batch.py:33-74 of the reference, float32 ids."""
    g, s, y = col([1, 1, 2, 2, 2]), col([0, 1, 2, 3, 4]), col([1.1, 0, 0, 1, 1])
    out = run_pairwise(s, y, g, S.PairSpec(power=-0.5))
    assert int(out["n_pair"].item()) == 3
    assert abs(float(out["loss"].item()) - 0.5415076) < 1e-5
    out = run_pairwise(s, y, g, S.PairSpec(power=-0.5, rw_pos=np.ones(5, np.float32)))
    assert abs(float(out["loss"].item()) - 0.5415076) < 1e-5
    out = run_pairwise(s, y, g, S.PairSpec(power=-0.5), mask=[True, True, False, False, False])
    assert int(out["n_pair"].item()) == 1
    assert abs(float(out["loss"].item()) - 1.3132617) < 1e-5


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_cfg1(seed):
    d = G.cfg1(seed)
    for ids in (d["g"], d["g_f32"]):
        out = run_pairwise(d["s"], d["y"], ids)
        check_pairwise(out, S.pairwise(d["s"], d["y"], ids), ctx=f"cfg1 seed{seed}")


SPECS = {
    "default": dict(),
    "power": dict(power=-0.5),
    "power1": dict(power=1.0),
    "factor": dict(factor=2.5),
    "sum": dict(reduce_mean=False),
    "wrong": dict(only_wrong=True),
    "wrong_power": dict(only_wrong=True, power=-1.0),
    "diff": dict(label_func="diff"),
    "diff_rwp_power": dict(label_func="diff", rw_pos="w", power=-0.5),
    "rwp": dict(rw_pos="w"),
    "rwn": dict(rw_neg="w"),
    "diff_rwn_wrong": dict(label_func="diff", rw_neg="w", rw_pos="w2", only_wrong=True, power=0.5),
}


@pytest.mark.parametrize("name", list(SPECS))
@pytest.mark.parametrize("graded", [False, True])
def test_options_small(name, graded):
    rng = np.random.default_rng(hash(name) % 1000 + graded)
    b = 3000
    gidx = G.zipf_groups(rng, b, 97)
    s = rng.standard_normal(b).astype(np.float32) * 2
    y = (rng.integers(0, 5, b) if graded else (rng.random(b) < 0.3)).astype(np.float32)
    w = rng.uniform(0.5, 1.5, b).astype(np.float32)
    w2 = rng.uniform(0.1, 2.0, b).astype(np.float32)
    w[rng.integers(0, b, 40)] = 0.0           # non-positive factors remove pairs (W > 0 rule)
    w2[rng.integers(0, b, 10)] = -1.0
    kw = {k: (w if v == "w" else w2 if v == "w2" else v) for k, v in SPECS[name].items()}
    spec = S.PairSpec(**kw)
    mask = rng.random(b) < 0.9 if name in ("power", "diff") else None
    ids = (gidx.astype(np.int64) * 7919 + 13)
    out = run_pairwise(s, y, ids, spec, mask=mask)
    check_pairwise(out, S.pairwise(s, y, ids, spec, mask=mask), ctx=name)


def test_multi_key_and_float_ids():
    rng = np.random.default_rng(5)
    b = 2048
    k0 = rng.integers(0, 30, b).astype(np.float32)
    k1 = rng.integers(0, 3, b).astype(np.float64) - 1.0     # contains -1, 0, 1
    k0[rng.integers(0, b, 20)] = np.nan
    k0[rng.integers(0, b, 20)] = np.inf
    k1[k1 == 0] = np.where(rng.random((k1 == 0).sum()) < 0.5, 0.0, -0.0)
    s = rng.standard_normal(b).astype(np.float32)
    y = rng.integers(0, 3, b).astype(np.float32)
    spec = S.PairSpec(power=-0.5)
    out = run_pairwise(s, y, [k0, k1], spec)
    check_pairwise(out, S.pairwise(s, y, [k0, k1], spec), ctx="multikey")


def test_degenerate():
    # no pair at all -> loss 0, zero gradient, not NaN (PW:13, PW:126)
    s = np.arange(8, dtype=np.float32)
    out = run_pairwise(s, np.ones(8, np.float32), np.arange(8, dtype=np.int64))
    assert int(out["n_pair"].item()) == 0 and float(out["loss"].item()) == 0.0
    assert not out["dlogits"].cpu().numpy().any()
    # one row
    out = run_pairwise(s[:1], np.ones(1, np.float32), np.zeros(1, np.int64))
    assert int(out["n_pair"].item()) == 0
    # one big group, all distinct labels: n = B(B-1)/2; extreme logits stay finite
    b = 777
    rng = np.random.default_rng(1)
    s = (rng.standard_normal(b) * 40).astype(np.float32)
    y = rng.permutation(b).astype(np.float32)
    ids = np.zeros(b, np.int64)
    out = run_pairwise(s, y, ids)
    assert int(out["n_pair"].item()) == b * (b - 1) // 2
    check_pairwise(out, S.pairwise(s, y, ids), ctx="one-group")
    # NaN labels pair with nothing
    y2 = y.copy(); y2[::7] = np.nan
    out = run_pairwise(s, y2, ids)
    check_pairwise(out, S.pairwise(s, y2, ids), ctx="nan-labels")


def test_cfg2():
    d = G.cfg2(0)
    out = run_pairwise(d["s"], d["y"], d["g"])
    r = check_pairwise(out, S.pairwise(d["s"], d["y"], d["g"]), ctx="cfg2")
    print("cfg2 parity", r, "n_pair", int(out["n_pair"].item()))


@pytest.mark.parametrize("parts", [2, 3, 8])
def test_partition_partials_sum_to_full(parts):
    """Multi-GPU work split (part_rank/part_count): partial loss / gradients add up to the single-call result,
    counts are global in every part."""
    d = G.cfg2(1)
    rng = np.random.default_rng(0)
    y = rng.integers(0, 5, d["s"].size).astype(np.float32)
    spec = S.PairSpec(power=-0.5, label_func="diff", rw_pos=rng.uniform(0.5, 1.5, y.size).astype(np.float32))
    full = run_pairwise(d["s"], y, d["g"], spec)
    loss, grad = 0.0, 0.0
    for r in range(parts):
        out = run_pairwise(d["s"], y, d["g"], spec, part=(r, parts))
        assert int(out["n_pair"].item()) == int(full["n_pair"].item())
        loss += float(out["loss"].item())
        grad = grad + out["dlogits"].double()
    assert abs(loss - float(full["loss"].item())) <= 2e-6 * abs(loss)
    ref = S.pairwise(d["s"], y, d["g"], spec)
    err = np.abs(grad.cpu().numpy() - ref["grad"])
    assert (err <= 1e-5 * ref["grad_abs"] + 1e-12).all()


@pytest.mark.parametrize("world,kk,with_ok", [(2, 1, False), (4, 2, True)])
def test_blocked_rows_match_contiguous(world, kk, with_ok):
    """Global-mode layouts of the C ABI: rows as packed per-rank blocks (block_rows / block_stride) and outputs laid
    out for one reduce-scatter (out_chunk) give the same numbers as the contiguous call on the concatenated rows."""
    import torch
    from rec_now_b200 import ops
    b_loc = 2048
    d = G.cfg5(world, seed=5, rows_per_rank=b_loc, groups_per_rank=64)
    rng = np.random.default_rng(2)
    b = world * b_loc
    cols = [d["g"]] + ([rng.integers(0, 3, b).astype(np.int64)] if kk > 1 else [])
    ok = (rng.random(b) > 0.1) if with_ok else None
    spec = S.PairSpec(power=-0.5, label_func="diff", rw_pos=d["w"])
    ref = S.pairwise(d["s"], d["y"], cols if kk > 1 else cols[0], spec, mask=ok)
    lay = ops.packed_block_layout(b_loc, kk, True, with_ok)
    blocks = []
    for r in range(world):
        sl = slice(r * b_loc, (r + 1) * b_loc)
        parts = [np.ascontiguousarray(c[sl]).view(np.uint8) for c in cols]
        parts += [np.ascontiguousarray(d[k][sl]).view(np.uint8) for k in ("s", "y", "w")]
        if with_ok:
            parts.append(ok[sl].astype(np.uint8))
        blk = np.concatenate(parts)
        blocks.append(np.concatenate([blk, np.zeros(lay["stride"] - blk.size, np.uint8)]))
    gbuf = torch.tensor(np.concatenate(blocks)).cuda()
    loss, grad = 0.0, np.zeros(b)
    nparts = 3
    for r in range(nparts):
        res = ops.pairwise_fwd_bwd_blocked(gbuf, world, b_loc, kk, True, with_ok, label_func="diff", power=-0.5,
                                           part=(r, nparts))
        o = res["out"].cpu().numpy().reshape(world, res["chunk"])
        assert int(res["n_pair"].item()) == ref["n_pair"]
        assert np.all(o[:, b_loc] == o[0, b_loc]) and np.all(o[:, b_loc + 1:] == 0)       # loss slot + zeroed pad
        loss += float(o[0, b_loc])
        grad += o[:, :b_loc].reshape(-1).astype(np.float64)
    assert abs(loss - ref["loss"]) <= 1e-5 * abs(ref["loss"])
    err = np.abs(grad - ref["grad"])
    assert (err <= 1e-5 * ref["grad_abs"] + 1e-12).all(), err.max()


def test_cfg3_full_size():
    """BASELINE.json config 3 at full size (B = 65536, graded labels, per-sample weights, power -0.5) against the
    float64 segmented oracle: exact counts, 1e-5 loss / gradient."""
    d = G.cfg3(0)
    spec = S.PairSpec(power=-0.5, label_func="diff", rw_pos=d["w"])
    out = run_pairwise(d["s"], d["y"], d["g"], spec)
    r = check_pairwise(out, S.pairwise(d["s"], d["y"], d["g"], spec), ctx="cfg3")
    print("cfg3 parity", r, "n_pair", int(out["n_pair"].item()))


def test_size_independent_properties_full_size():
    """Properties that need no oracle, at B = 65536: (1) the gradient sums to zero (every pair adds +d to one row
    and -d to another); (2) n_pair == sum(row_pairs); (3) permuting the rows and relabelling the group ids permutes
    the gradient and leaves n_pair / loss unchanged; (4) the loss of reduce_mean=False equals n_pair * mean loss."""
    d = G.cfg3(1)
    rng = np.random.default_rng(7)
    spec = S.PairSpec(power=0.0, label_func="diff", rw_pos=d["w"])
    out = run_pairwise(d["s"], d["y"], d["g"], spec)
    g = out["dlogits"].double().cpu().numpy()
    n = int(out["n_pair"].item())
    assert abs(g.sum()) <= 1e-6 * np.abs(g).sum()
    assert int(out["row_pairs"].sum().item()) == n
    perm = rng.permutation(g.size)
    ids2 = (d["g"] ^ np.int64(0x5DEECE66D))[perm]                 # bijective relabelling
    w2 = d["w"][perm]
    out2 = run_pairwise(d["s"][perm], d["y"][perm], ids2, S.PairSpec(power=0.0, label_func="diff", rw_pos=w2))
    assert int(out2["n_pair"].item()) == n
    l1, l2 = float(out["loss"].item()), float(out2["loss"].item())
    assert abs(l1 - l2) <= 2e-6 * abs(l1)
    g2 = out2["dlogits"].double().cpu().numpy()
    scale = np.abs(g).max()
    assert np.abs(g2 - g[perm]).max() <= 2e-6 * scale
    assert np.array_equal(out2["row_pairs"].cpu().numpy(), out["row_pairs"].cpu().numpy()[perm])
    out3 = run_pairwise(d["s"], d["y"], d["g"], S.PairSpec(power=0.0, label_func="diff", rw_pos=d["w"], reduce_mean=False))
    assert abs(float(out3["loss"].item()) - l1 * n) <= 2e-6 * abs(l1 * n)


def test_edge_tiles_and_wide_logits():
    """Groups whose level runs straddle the 32-row J blocks and 64-row I blocks in every way (sizes 1..200, 1-7
    label levels), logits spread over +-60 (softplus from ~1e-26 to ~60): exercises the sentinel / zero-weight fast
    tile and the general tile against the oracle."""
    rng = np.random.default_rng(11)
    sizes = np.r_[np.arange(1, 201), rng.integers(1, 200, 150)]
    gidx = np.repeat(np.arange(sizes.size), sizes)
    b = gidx.size
    perm = rng.permutation(b)
    gidx = gidx[perm]
    levels = rng.integers(1, 8, sizes.size)
    y = (rng.integers(0, 1 << 30, b) % levels[gidx]).astype(np.float32)
    s = (rng.standard_normal(b) * 20).astype(np.float32)
    w = rng.uniform(0.5, 1.5, b).astype(np.float32)
    ids = gidx.astype(np.int64) * 1000003 + 17
    for spec in (S.PairSpec(), S.PairSpec(power=-0.5, label_func="diff", rw_pos=w), S.PairSpec(rw_pos=w, factor=0.3)):
        out = run_pairwise(s, y, ids, spec)
        check_pairwise(out, S.pairwise(s, y, ids, spec), ctx=f"edge {spec.label_func}")


@pytest.mark.parametrize("b", [200_000, 300_000, 600_000])
def test_large_batch_wide_tiles(b):
    """B > 148 * 1024 switches the segmentation kernel to 4096-row sort tiles (several tiles per CTA) and needs
    18+ group bits: the size of the all-gathered batch in global mode at 4+ GPUs.  Up to B = 524288 (8 GPUs) the pair
    kernel partitions the tile staircase statically (cost prefix in shared memory); B = 600000 takes the explicit
    work-unit records written by the segmentation kernel.  For the two larger sizes the 2-way partition of the global
    mode is checked too."""
    rng = np.random.default_rng(21)
    gidx = np.r_[np.repeat([0, 1], 3000), 2 + rng.integers(0, 4000, b - 6000)][rng.permutation(b)]
    ids = (gidx.astype(np.int64) * 2654435761 + 12345) ^ np.int64(0x0123456789ABCDEF)
    y = rng.integers(0, 5, b).astype(np.float32)
    s = rng.standard_normal(b).astype(np.float32)
    w = rng.uniform(0.5, 1.5, b).astype(np.float32)
    spec = S.PairSpec(power=-0.5, label_func="diff", rw_pos=w)
    ref = S.pairwise(s, y, ids, spec)
    out = run_pairwise(s, y, ids, spec)
    check_pairwise(out, ref, ctx=f"B={b}")
    if b > 262144:
        loss, grad = 0.0, 0.0
        for r in range(2):
            o = run_pairwise(s, y, ids, spec, part=(r, 2))
            assert int(o["n_pair"].item()) == ref["n_pair"]
            loss += float(o["loss"].item()); grad = grad + o["dlogits"].double()
        assert abs(loss - ref["loss"]) <= 1e-5 * abs(ref["loss"])
        assert (np.abs(grad.cpu().numpy() - ref["grad"]) <= 1e-5 * ref["grad_abs"] + 1e-12).all()


@pytest.mark.parametrize("scale", [0.2, 1.0, 4.0, 9.0, 30.0, 300.0])
@pytest.mark.parametrize("shift", [0.0, -40.0])
def test_score_range_product_and_exp_tiles(scale, shift):
    """The fast tile has two forms: the product form (exp(-x) = E_i * F_j, scores within 2^+-12 of the tile's reference
    score) and exp-per-pair.  Wide and shifted score distributions exercise both and the switch between them."""
    rng = np.random.default_rng(int(scale * 10) + int(-shift))
    b = 12000
    g = np.minimum(rng.zipf(1.3, b), 40).astype(np.int64)          # a few big groups -> many fast tiles
    s = (rng.standard_normal(b) * scale + shift).astype(np.float32)
    y = rng.integers(0, 5, b).astype(np.float32)
    w = rng.uniform(0.5, 1.5, b).astype(np.float32)
    for spec in (S.PairSpec(), S.PairSpec(label_func="diff", rw_pos=w, power=-0.5)):
        out = run_pairwise(s, y, g, spec)
        check_pairwise(out, S.pairwise(s, y, g, spec), ctx=f"scale {scale} shift {shift}")
