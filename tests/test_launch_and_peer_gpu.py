"""GPU tests of the launch plumbing and the peer-memory pieces of the global mode, all on one GPU:
the cached-CUDA-graph launch of rn_pairwise_fwd_bwd across changing shapes / modes / pointers, the plain-launch path
under the caller's own graph capture, rn_pack_row_block against the documented block layout, the gather done by the
first kernel of a call (rn_pairwise_args.peer_blocks, here with "peers" in local memory) and rn_reduce_peer_chunks."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import generators as G
from oracle import seg_ref as S
from tests.util import check_pairwise, dev, run_pairwise

pytestmark = pytest.mark.gpu


def _case(b, groups, seed, graded, weights):
    rng = np.random.default_rng(seed)
    g = (rng.integers(0, groups, b).astype(np.int64) * 1000003) ^ 0x5DEECE66D
    y = (rng.integers(0, 5, b) if graded else (rng.random(b) < 0.3)).astype(np.float32)
    s = rng.standard_normal(b).astype(np.float32)
    w = rng.uniform(0.5, 1.5, b).astype(np.float32) if weights else None
    return s, y, g, S.PairSpec(power=-0.5 if weights else 0.0, label_func="diff" if graded else "step", rw_pos=w)


def test_graph_launch_across_shapes_and_modes():
    """Consecutive calls with different batch sizes, kernel variants and fresh tensors: every call is parity-exact and
    (after the first call of a variant) goes out as one launch of the cached graph with updated node parameters."""
    from rec_now_b200 import _lib
    lib = _lib.lib()
    before = int(lib.rn_debug_graph_launches())
    assert before >= 0, "graph launches were switched off by an earlier failure"
    cases = [(3000, 40, 0, True, True), (1024, 64, 1, False, False), (20000, 300, 2, True, True),
             (3000, 40, 3, True, True), (777, 5, 4, False, False), (20000, 300, 5, True, False)]
    for rep in range(2):
        for b, groups, seed, graded, weights in cases:
            s, y, g, spec = _case(b, groups, seed + 10 * rep, graded, weights)
            check_pairwise(run_pairwise(s, y, g, spec), S.pairwise(s, y, g, spec), ctx=f"B={b} rep={rep}")
    after = int(lib.rn_debug_graph_launches())
    # (batches of up to 1024 rows are one ordinary launch of the one-CTA kernel, not a graph)
    assert after - before == 2 * sum(1 for c in cases if c[0] > 1024), (before, after)


def test_plain_launches_under_callers_graph_capture():
    """Inside the caller's stream capture the library must not launch its own graph: the three kernels are captured
    into the caller's graph, and replaying that graph recomputes the outputs in place."""
    from rec_now_b200 import ops
    s, y, g, spec = _case(6000, 80, 7, True, True)
    ref = S.pairwise(s, y, g, spec)
    ds, dy, dw = dev(s), dev(y), dev(spec.rw_pos)
    keys, _ = ops.canon_keys([dev(g)])
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        for _ in range(2):
            ops.pairwise_fwd_bwd(ds, dy, keys, rw_pos=dw, label_func="diff", power=-0.5)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        out = ops.pairwise_fwd_bwd(ds, dy, keys, rw_pos=dw, label_func="diff", power=-0.5, want_row_pairs=True)
    out["dlogits"].zero_()
    graph.replay()
    torch.cuda.synchronize()
    check_pairwise(out, ref, ctx="replayed capture")


@pytest.mark.parametrize("kk,with_w,with_ok", [(1, True, False), (2, False, True), (3, True, True)])
def test_pack_row_block_layout(kk, with_w, with_ok):
    from rec_now_b200 import ops
    b_loc = 4096
    rng = np.random.default_rng(kk)
    keys = rng.integers(-2**62, 2**62, (kk, b_loc), dtype=np.int64)
    s, y, w = (rng.standard_normal(b_loc).astype(np.float32) for _ in range(3))
    ok = (rng.random(b_loc) > 0.2).astype(np.uint8)
    lay = ops.packed_block_layout(b_loc, kk, with_w, with_ok)
    blk = ops.pack_row_block(dev(keys), dev(s), dev(y), dev(w) if with_w else None, dev(ok) if with_ok else None,
                             lay["stride"]).cpu().numpy()
    assert blk.size == lay["stride"]
    assert np.array_equal(blk[lay["keys"]:lay["keys"] + 8 * kk * b_loc].view(np.int64).reshape(kk, b_loc), keys)
    assert np.array_equal(blk[lay["logits"]:lay["logits"] + 4 * b_loc].view(np.float32), s)
    assert np.array_equal(blk[lay["labels"]:lay["labels"] + 4 * b_loc].view(np.float32), y)
    end = lay["labels"] + 4 * b_loc
    if with_w:
        assert np.array_equal(blk[lay["w"]:lay["w"] + 4 * b_loc].view(np.float32), w)
        end = lay["w"] + 4 * b_loc
    if with_ok:
        assert np.array_equal(blk[lay["ok"]:lay["ok"] + b_loc], ok)
        end = lay["ok"] + b_loc
    assert not blk[end:].any()                      # zero padding up to the stride


def test_gather_in_first_kernel_and_peer_reduce():
    """The all-gather done by k_init (peer_blocks / gather_dst) and the reduction over the peers' chunked gradient
    buffers, with every "peer" buffer in local memory: `world` separately allocated blocks give the same result as the
    single-batch oracle, and the reduced chunk of every rank is its slice of the gradient."""
    from rec_now_b200 import ops
    world, b_loc = 4, 4096
    d = G.cfg5(world, seed=11, rows_per_rank=b_loc, groups_per_rank=128)
    spec = S.PairSpec(power=-0.5, label_func="diff", rw_pos=d["w"])
    ref = S.pairwise(d["s"], d["y"], d["g"], spec)
    lay = ops.packed_block_layout(b_loc, 1, True, False)
    blocks = []
    for r in range(world):
        sl = slice(r * b_loc, (r + 1) * b_loc)
        blocks.append(ops.pack_row_block(dev(d["g"][sl]).reshape(1, -1), dev(d["s"][sl]), dev(d["y"][sl]), dev(d["w"][sl]),
                                         None, lay["stride"]))
    outs = []
    for r in range(world):                          # what rank r computes: its share of the tiles over all rows
        gbuf = torch.full((world * lay["stride"],), 0xAB, dtype=torch.uint8, device="cuda")      # garbage until gathered
        res = ops.pairwise_fwd_bwd_blocked(gbuf, world, b_loc, 1, True, False, label_func="diff", power=-0.5,
                                           part=(r, world), peer_blocks=[b.data_ptr() for b in blocks])
        assert int(res["n_pair"].item()) == ref["n_pair"]
        assert np.array_equal(gbuf.cpu().numpy(), torch.cat(blocks).cpu().numpy())               # the gather itself
        outs.append(res["out"])
    chunk = b_loc + 4
    grad, losses = [], []
    for r in range(world):
        mine = ops.reduce_peer_chunks([o.data_ptr() for o in outs], r, chunk, torch.device("cuda", torch.cuda.current_device()))
        grad.append(mine[:b_loc].cpu().numpy().astype(np.float64))
        losses.append(float(mine[b_loc].item()))
    assert all(abs(l - ref["loss"]) <= 1e-5 * abs(ref["loss"]) for l in losses), (losses, ref["loss"])
    err = np.abs(np.concatenate(grad) - ref["grad"])
    assert (err <= 1e-5 * ref["grad_abs"] + 1e-12).all(), err.max()


@pytest.mark.parametrize("in_graph", [1, 0])
def test_pair_kernel_timing_aid(in_graph):
    """rn_profile_enable_ex: the pair kernel's time per call, from event-record nodes inside the call's CUDA graph
    (in_graph = 1) or from stream events around a plain launch (0); results stay correct while it is on."""
    from rec_now_b200 import _lib
    lib = _lib.lib()
    d = G.cfg2()
    ref = S.pairwise(d["s"], d["y"], d["g"])
    assert lib.rn_profile_enable_ex(4, in_graph) == 0
    try:
        for _ in range(6):                     # two more calls than slots: the extra ones are simply not timed
            out = run_pairwise(d["s"], d["y"], d["g"])
        torch.cuda.synchronize()
        ms = (C.c_float * 4)(); n = C.c_int32(0)
        assert lib.rn_profile_collect(ms, 4, C.byref(n)) == 0
    finally:
        lib.rn_profile_disable()
    assert n.value == 4
    assert all(0.002 < t < 5.0 for t in ms), list(ms)
    assert lib.rn_debug_graph_launches() >= 0, "a graph API call failed: the thread fell back to plain launches"
    check_pairwise(out, ref, ctx=f"timing aid in_graph={in_graph}")
