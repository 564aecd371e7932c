"""Deterministic mode of rn_pairwise_fwd_bwd (rn_pairwise_args.deterministic): bit-identical loss and gradient across
runs, still within the parity bars of the float64 oracle.  (The default path accumulates d loss / d logits with
floating-point atomics -- SURVEY.md hard part 4 asks for a deterministic formulation.)"""
import numpy as np
import pytest
import torch

from oracle import generators as G
from oracle import seg_ref as S
from tests.util import check_pairwise, dev

pytestmark = pytest.mark.gpu


def _run(d, spec_kw, **kw):
    from rec_now_b200 import ops
    keys = dev(d["g"]).reshape(1, -1)
    return ops.pairwise_fwd_bwd(dev(d["s"]), dev(d["y"]), keys, rw_pos=None if "w" not in spec_kw else dev(d["w"]),
                                label_func=spec_kw.get("label_func", "step"), power=spec_kw.get("power", 0.0),
                                want_row_pairs=True, **kw)


@pytest.mark.parametrize("cfg", ["cfg2", "cfg3", "cfg3_small_weights"])
def test_bit_identical_across_runs(cfg):
    from rec_now_b200 import ops
    if cfg == "cfg2":
        d, kw = G.cfg2(0), dict()
    else:
        d, kw = G.cfg3(0), dict(label_func="diff", power=-0.5, w=True)
        if cfg == "cfg3_small_weights":
            d = dict(d); d["w"] = (d["w"] * 1e-6).astype(np.float32)      # the accumulator scale follows the weights
    spec = S.PairSpec(power=kw.get("power", 0.0), label_func=kw.get("label_func", "step"),
                      rw_pos=d["w"] if "w" in kw else None)
    ref = S.pairwise(d["s"], d["y"], d["g"], spec)
    outs = []
    for _ in range(5):
        out = _run(d, kw, deterministic=True)
        assert ops.last_segmentation_path(out["_scratch"]) == 2          # (the sort: groups and rows in a fixed order)
        assert ops.device_error(out["_scratch"]) == 0
        outs.append((out["loss"].cpu().numpy().tobytes(), out["dlogits"].cpu().numpy().tobytes(), int(out["n_pair"].item())))
        _run(d, kw)                                                      # a default-mode call in between, same arena
    assert len(set(outs)) == 1, "deterministic mode produced different bits across runs"
    check_pairwise(_run(d, kw, deterministic=True), ref, ctx=cfg)


def test_unsupported_combinations():
    from rec_now_b200 import ops
    from rec_now_b200._lib import RnError
    d = G.cfg1(0)
    with pytest.raises(RnError):
        ops.pairwise_fwd_bwd(dev(d["s"]), dev(d["y"]), dev(d["g"]).reshape(1, -1), only_wrong=True, deterministic=True)
