"""GPU parity of the sort-free counting segmentation (csrc/group_count.cuh, HeadsTail::count_run) and of the
persistent-arena contract behind it (rn_pairwise_args.scratch_persistent).

The counting path replaces the same reference lines as the radix path (pairwise_loss_from_batch.py:33-37, 68-73,
187-190): same oracle, same bars (exact counts, 1e-5 loss / gradient).  What is specific here: which path ran, the
in-kernel fallback, and that every call leaves the arena clean for the next one whatever it was.
"""
import numpy as np
import pytest
import torch

from oracle import generators as G
from oracle import seg_ref as S
from tests.util import check_pairwise, dev, run_pairwise

pytestmark = pytest.mark.gpu


def _path(out):
    from rec_now_b200 import ops
    return ops.last_segmentation_path(out["_scratch"])


def test_counting_path_taken_on_baseline_configs():
    for d in (G.cfg1(0), G.cfg2(0)):
        out = run_pairwise(d["s"], d["y"], d["g"])
        check_pairwise(out, S.pairwise(d["s"], d["y"], d["g"]), ctx=d["name"])
        assert _path(out) == (3 if d["s"].size <= 1024 else 1), d["name"]      # (3: the one-CTA kernel of small batches)


def test_cfg3_counting_path_full_size():
    d = G.cfg3(0)
    spec = S.PairSpec(power=-0.5, label_func="diff", rw_pos=d["w"])
    out = run_pairwise(d["s"], d["y"], d["g"], spec)
    check_pairwise(out, S.pairwise(d["s"], d["y"], d["g"], spec), ctx="cfg3")
    assert _path(out) == 1


@pytest.mark.parametrize("labels", ["frac", "many", "neg2", "level8", "wzero"])
def test_fallback_to_radix(labels):
    """Labels outside the level menu (or non-positive row weights) are detected on the device; the same call goes on
    with the radix path and stays exact."""
    rng = np.random.default_rng(11)
    b = 5000
    ids = (G.zipf_groups(rng, b, 61).astype(np.int64) * 104729 + 7)
    s = rng.standard_normal(b).astype(np.float32)
    y = rng.integers(0, 4, b).astype(np.float32)
    w = None
    if labels == "frac":
        y[17] = 1.5
    elif labels == "many":
        y = rng.integers(0, 40, b).astype(np.float32)
    elif labels == "neg2":
        y[100] = -2.0
    elif labels == "level8":
        y[3] = 7.0                      # -1 .. 6 are the eight levels; 7 is outside
    else:
        w = rng.uniform(0.5, 1.5, b).astype(np.float32)
        w[5] = 0.0
    spec = S.PairSpec(power=-0.5, rw_pos=w)
    out = run_pairwise(s, y, ids, spec)
    check_pairwise(out, S.pairwise(s, y, ids, spec), ctx=labels)
    assert _path(out) == 2


def test_levels_minus_one_to_six_and_unpairable_rows():
    rng = np.random.default_rng(3)
    b = 7000
    ids = (G.zipf_groups(rng, b, 200).astype(np.int64) << 33) + 5
    s = rng.standard_normal(b).astype(np.float32)
    y = rng.integers(-1, 7, b).astype(np.float32)
    y[::97] = np.nan                     # NaN labels pair with nothing
    mask = rng.random(b) < 0.8
    spec = S.PairSpec(power=-1.0, label_func="diff")
    out = run_pairwise(s, y, ids, spec, mask=mask)
    check_pairwise(out, S.pairwise(s, y, ids, spec, mask=mask), ctx="levels")
    assert _path(out) == 1


def test_arena_stays_clean_across_mixed_calls():
    """One persistent arena, many calls of different kinds: counting, fallback, score-dependent pair sets, different
    data.  Every call must find the arena as the previous one left it."""
    from rec_now_b200 import ops
    rng = np.random.default_rng(9)
    b = 4096
    cases = []
    for k in range(8):
        ids = (G.zipf_groups(rng, b, 50 + 40 * k).astype(np.int64) * 7919 + k)
        s = rng.standard_normal(b).astype(np.float32)
        y = rng.integers(0, 5, b).astype(np.float32)
        kind = ["plain", "frac", "wrong", "rwn", "plain", "power", "frac", "plain"][k]
        if kind == "frac":
            y[k] = 0.25
        w = rng.uniform(0.5, 1.5, b).astype(np.float32)
        spec = {"plain": S.PairSpec(), "frac": S.PairSpec(power=0.5), "wrong": S.PairSpec(only_wrong=True, power=-0.5),
                "rwn": S.PairSpec(rw_neg=w, label_func="diff"), "power": S.PairSpec(power=-0.5, label_func="diff", rw_pos=w)}[kind]
        cases.append((kind, s, y, ids, spec))
    arenas = set()
    for rep in range(2):
        for kind, s, y, ids, spec in cases:
            out = run_pairwise(s, y, ids, spec)
            check_pairwise(out, S.pairwise(s, y, ids, spec), ctx=f"{kind} rep{rep}")
            assert _path(out) == (2 if kind == "frac" else 1), kind
            assert ops.device_error(out["_scratch"]) == 0
            arenas.add(out["_scratch"].data_ptr())
    assert len(arenas) == 1                      # all of them shared one arena


def test_many_tiles_per_cta_and_scratch_rows():
    """B > 148 x 512 rows: several tiles per CTA (rows spilled between the phases); and one arena sized for a larger
    capacity shared by batches of different sizes (scratch_rows)."""
    import ctypes as C
    from rec_now_b200 import _lib, ops
    d = G.cfg3(1, b=200_000, n_groups=9000)
    spec = S.PairSpec(power=-0.5, label_func="diff", rw_pos=d["w"])
    out = run_pairwise(d["s"], d["y"], d["g"], spec)
    check_pairwise(out, S.pairwise(d["s"], d["y"], d["g"], spec), ctx="200k")
    assert _path(out) == 1
    # capacity arena through the raw ABI
    lib = _lib.lib()
    cap = 50_000
    nbytes = lib.rn_pairwise_scratch_bytes(cap, 1)
    scratch = torch.empty(nbytes, dtype=torch.uint8, device="cuda").fill_(0xA5)          # garbage first ...
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    assert lib.rn_pairwise_scratch_init(scratch.data_ptr(), nbytes, st) == 0              # ... then the one-time init
    rng = np.random.default_rng(4)
    for b in (50_000, 1234, 33_333, 7, 50_000):
        ids = (G.zipf_groups(rng, b, max(2, b // 20)).astype(np.int64) * 31 + 1)
        s = rng.standard_normal(b).astype(np.float32)
        y = (rng.random(b) < 0.3).astype(np.float32)
        ts, ty, tk = dev(s), dev(y), dev(ids)
        outv = torch.empty(4, dtype=torch.float32, device="cuda")
        dl = torch.empty(b, dtype=torch.float32, device="cuda")
        a = _lib.PairwiseArgs(B=b, K=1, label_func=0, keys=tk.data_ptr(), logits=ts.data_ptr(), labels=ty.data_ptr(),
                              factor=1.0, power=-0.5, only_wrong=0, reduce_mean=1, part_rank=0, part_count=1,
                              loss=outv.data_ptr(), n_pair_f32=outv.data_ptr() + 4, n_pair=outv.data_ptr() + 8,
                              dlogits=dl.data_ptr(), scratch_persistent=1, scratch_rows=cap)
        assert lib.rn_pairwise_fwd_bwd(C.byref(a), scratch.data_ptr(), nbytes, st) == 0
        res = dict(loss=outv[0], n_pair_f32=outv[1], n_pair=outv[2:4].view(torch.int64)[0], dlogits=dl)
        check_pairwise(res, S.pairwise(s, y, ids, S.PairSpec(power=-0.5)), ctx=f"cap b={b}")
        assert ops.last_segmentation_path(scratch) == (3 if b <= 1024 else 1)
        assert ops.device_error(scratch) == 0


def test_non_persistent_arena_still_takes_the_radix_path():
    """scratch_persistent = 0 (the default of a zero-initialised struct): the arena may hold anything, k_init runs."""
    import ctypes as C
    from rec_now_b200 import _lib, ops
    lib = _lib.lib()
    d = G.cfg1(2, b=3000, n_groups=150)          # (above the 1024 rows the one-CTA kernel takes)
    b = d["s"].size
    nbytes = lib.rn_pairwise_scratch_bytes(b, 1)
    scratch = torch.empty(nbytes, dtype=torch.uint8, device="cuda").fill_(0x5A)
    ts, ty, tk = dev(d["s"]), dev(d["y"]), dev(d["g"])
    outv = torch.empty(4, dtype=torch.float32, device="cuda")
    dl = torch.empty(b, dtype=torch.float32, device="cuda")
    a = _lib.PairwiseArgs(B=b, K=1, label_func=0, keys=tk.data_ptr(), logits=ts.data_ptr(), labels=ty.data_ptr(),
                          factor=1.0, power=0.0, only_wrong=0, reduce_mean=1, part_rank=0, part_count=1,
                          loss=outv.data_ptr(), n_pair_f32=outv.data_ptr() + 4, n_pair=outv.data_ptr() + 8,
                          dlogits=dl.data_ptr())
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for _ in range(3):
        assert lib.rn_pairwise_fwd_bwd(C.byref(a), scratch.data_ptr(), nbytes, st) == 0
    res = dict(loss=outv[0], n_pair_f32=outv[1], n_pair=outv[2:4].view(torch.int64)[0], dlogits=dl)
    check_pairwise(res, S.pairwise(d["s"], d["y"], d["g"]), ctx="non-persistent")
    assert ops.last_segmentation_path(scratch) == 2


def test_device_side_failure_returns_nan_not_a_result():
    """A call whose control block carries a device-side error flag (a grid barrier that timed out, a full group table --
    here planted) must not look like a result: loss = NaN and rn_last_device_error != 0; the call after it is clean."""
    import ctypes as C
    from rec_now_b200 import _lib, ops
    lib = _lib.lib()
    d = G.cfg1(3, b=5000, n_groups=200)
    b = d["s"].size
    nbytes = lib.rn_pairwise_scratch_bytes(b, 1)
    scratch = torch.zeros(nbytes, dtype=torch.uint8, device="cuda")
    ts, ty, tk = dev(d["s"]), dev(d["y"]), dev(d["g"])
    outv = torch.empty(4, dtype=torch.float32, device="cuda")
    dl = torch.empty(b, dtype=torch.float32, device="cuda")
    a = _lib.PairwiseArgs(B=b, K=1, label_func=0, keys=tk.data_ptr(), logits=ts.data_ptr(), labels=ty.data_ptr(),
                          factor=1.0, power=0.0, only_wrong=0, reduce_mean=1, part_rank=0, part_count=1,
                          loss=outv.data_ptr(), n_pair_f32=outv.data_ptr() + 4, n_pair=outv.data_ptr() + 8,
                          dlogits=dl.data_ptr(), scratch_persistent=1)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    scratch[48:52].view(torch.int32).fill_(1)               # Ctl::err (the 13th word of the control block)
    assert lib.rn_pairwise_fwd_bwd(C.byref(a), scratch.data_ptr(), nbytes, st) == 0
    assert torch.isnan(outv[0]).item()
    assert ops.device_error(scratch) != 0
    assert lib.rn_pairwise_fwd_bwd(C.byref(a), scratch.data_ptr(), nbytes, st) == 0      # (the report is filed, the flag reset)
    res = dict(loss=outv[0], n_pair_f32=outv[1], n_pair=outv[2:4].view(torch.int64)[0], dlogits=dl)
    check_pairwise(res, S.pairwise(d["s"], d["y"], d["g"]), ctx="after a failed call")
    assert ops.device_error(scratch) == 0
