"""GPU parity of the one-CTA kernel that takes batches of up to 1024 rows (csrc/small.cu; BASELINE.json configs[0], the
reference's own CPU-runnable case: B = 1024, 64 groups, and the sizes of its unit tests) against the float64 oracle,
over the whole option menu -- and the general path on the same small batches (RN_SMALL=0), which it must agree with."""
import numpy as np
import pytest

from oracle import generators as G
from oracle import seg_ref as S
from tests.util import check_pairwise, run_pairwise

pytestmark = pytest.mark.gpu

SPECS = {
    "default": dict(),
    "power": dict(power=-0.5),
    "power1_sum": dict(power=1.0, reduce_mean=False),
    "factor": dict(factor=2.5),
    "wrong_power": dict(only_wrong=True, power=-1.0),
    "diff_rwp_power": dict(label_func="diff", rw_pos="w", power=-0.5),
    "rwp": dict(rw_pos="w"),
    "rwn": dict(rw_neg="w"),
    "diff_rwn_wrong": dict(label_func="diff", rw_neg="w", rw_pos="w2", only_wrong=True, power=0.5),
    "gain2": dict(label_func="gain2", power=-0.5),
    "hinge": dict(pair_loss="hinge", margin=0.5, power=-0.5),
    "hinge_diff_wrong": dict(pair_loss="hinge", label_func="diff", only_wrong=True, rw_neg="w"),
}


def _path(out):
    from rec_now_b200 import ops
    return ops.last_segmentation_path(out["_scratch"])


def _case(name, b, ng, seed):
    rng = np.random.default_rng(seed)
    gidx = G.zipf_groups(rng, b, ng)
    s = rng.standard_normal(b).astype(np.float32) * 2
    s[rng.integers(0, b, max(1, b // 8))] = 0.5               # tied scores
    y = rng.integers(0, 5, b).astype(np.float32)
    if b > 20:
        y[rng.integers(0, b, 3)] = np.nan
    w = rng.uniform(0.5, 1.5, b).astype(np.float32)
    w2 = rng.uniform(0.1, 2.0, b).astype(np.float32)
    w[rng.integers(0, b, max(1, b // 50))] = 0.0
    w2[rng.integers(0, b, max(1, b // 100))] = -1.0
    kw = {k: (w if v == "w" else w2 if v == "w2" else v) for k, v in SPECS[name].items()}
    mask = rng.random(b) < 0.9 if name in ("power", "rwn", "hinge") else None
    return s, y, gidx.astype(np.int64) * 7919 + 13, S.PairSpec(**kw), mask


@pytest.mark.parametrize("name", list(SPECS))
@pytest.mark.parametrize("b,ng", [(1, 1), (5, 2), (33, 4), (700, 23), (1024, 64), (1024, 3)])
def test_small_kernel_matches_oracle(name, b, ng):
    s, y, ids, spec, mask = _case(name, b, ng, 1000 * b + len(name))
    out = run_pairwise(s, y, ids, spec, mask=mask)
    check_pairwise(out, S.pairwise(s, y, ids, spec, mask=mask), ctx=f"{name} B={b}")
    assert _path(out) == 3


def test_general_path_on_small_batches(monkeypatch):
    """RN_SMALL=0: the same batches through k_seg / k_pair, same results (the two paths share nothing but the oracle)."""
    monkeypatch.setenv("RN_SMALL", "0")
    for name in ("default", "diff_rwp_power", "wrong_power", "hinge"):
        for b, ng in ((5, 2), (700, 23), (1024, 64)):
            s, y, ids, spec, mask = _case(name, b, ng, 77 * b + len(name))
            out = run_pairwise(s, y, ids, spec, mask=mask)
            check_pairwise(out, S.pairwise(s, y, ids, spec, mask=mask), ctx=f"general {name} B={b}")
            assert _path(out) in (1, 2)


def test_cfg1_and_extremes():
    for seed in range(3):
        d = G.cfg1(seed)
        for ids in (d["g"], d["g_f32"]):
            out = run_pairwise(d["s"], d["y"], ids)
            check_pairwise(out, S.pairwise(d["s"], d["y"], ids), ctx=f"cfg1 seed{seed}")
            assert _path(out) == 3
    # one group of 1024 rows, all labels distinct: n = B (B - 1) / 2, the longest loops the kernel can see
    rng = np.random.default_rng(0)
    b = 1024
    s = (rng.standard_normal(b) * 30).astype(np.float32)
    y = rng.permutation(b).astype(np.float32)
    out = run_pairwise(s, y, np.zeros(b, np.int64))
    assert int(out["n_pair"].item()) == b * (b - 1) // 2
    check_pairwise(out, S.pairwise(s, y, np.zeros(b, np.int64)), ctx="one group")
    # every row its own group; all rows masked
    out = run_pairwise(s, y, np.arange(b, dtype=np.int64))
    assert int(out["n_pair"].item()) == 0 and float(out["loss"].item()) == 0.0 and not out["dlogits"].cpu().numpy().any()
    out = run_pairwise(s, y, np.zeros(b, np.int64), mask=np.zeros(b, bool))
    assert int(out["n_pair"].item()) == 0 and float(out["loss"].item()) == 0.0


def test_small_kernel_fused_focal():
    """The fused joint objective on the small kernel: pairwise + focal_weight * focal (rn_pairwise_args.focal_*)."""
    import torch
    from rec_now_b200 import ops
    from tests.util import dev
    d = G.cfg1(1)
    focal = (0.3, 0.25, 2.0, False)
    keys, ok = ops.canon_keys([dev(d["g"])])
    out = ops.pairwise_fwd_bwd(dev(d["s"]), dev(d["y"]), keys, power=-0.5, focal=focal)
    rp = S.pairwise(d["s"], d["y"], d["g"], S.PairSpec(power=-0.5))
    rf = S.focal(d["y"], d["s"], alpha=0.25, gamma=2.0)
    want = rp["loss"] + 0.3 * rf["loss"]
    assert abs(float(out["loss"].item()) - want) <= 1e-5 * abs(want)
    g = out["dlogits"].cpu().numpy().astype(np.float64)
    gw = rp["grad"] + 0.3 * rf["grad"]
    assert np.abs(g - gw).max() <= 1e-5 * max(np.abs(gw).max(), rp["grad_abs"].max())
    assert _path(out) == 3
