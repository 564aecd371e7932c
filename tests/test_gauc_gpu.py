"""GPU parity of the GAUC metric (rn_gauc, csrc/gauc.cu) against the float64 restatement oracle/seg_ref.py::gauc.
The reference names the metric (README.md:5, 8) and ships no implementation: the definition is this repo's
(include/recnow_b200.h), pinned by a hand-computed known answer below.  Integer counts must be exact."""
import numpy as np
import pytest
import torch

from oracle import generators as G
from oracle import seg_ref as S
from tests.util import dev

pytestmark = pytest.mark.gpu


def run(s, y, g, mask=None):
    from rec_now_b200 import metrics, ops
    cols = [dev(c) for c in g] if isinstance(g, list) else dev(g)
    out = metrics.gauc(dev(s), dev(y), cols, mask=None if mask is None else dev(np.asarray(mask, bool)), return_details=True)
    assert ops.device_error(out["_scratch"]) == 0
    return out


def check(out, ref):
    assert int(out["n_pair"].item()) == ref["n_pair"]
    assert int(out["concordant2"].item()) == ref["concordant2"]
    assert int(out["n_valid_groups"].item()) == ref["n_valid_groups"]
    assert abs(float(out["gauc"].item()) - ref["gauc"]) <= 1e-6
    assert abs(float(out["auc_mean"].item()) - ref["auc_mean"]) <= 1e-6


def test_known_answer():
    # group 1: the positive on top of both negatives -> AUC 1; group 2: the positive below the negative -> AUC 0;
    # group 3: a tie -> 1/2; group 4: no negative -> no pairs, not counted.  Weights = rows: (3*1 + 2*0 + 2*0.5) / 7
    g = np.array([1, 1, 1, 2, 2, 3, 3, 4, 4], np.float32)
    y = np.array([1, 0, 0, 1, 0, 1, 0, 1, 1], np.float32)
    s = np.array([.9, .1, .5, .2, .8, .3, .3, .7, .6], np.float32)
    ref = S.gauc(s, y, g)
    assert ref["n_valid_groups"] == 3 and ref["n_pair"] == 4 and ref["concordant2"] == 5
    assert abs(ref["gauc"] - 4.0 / 7.0) < 1e-12
    out = run(s, y, g)
    check(out, ref)
    assert abs(float(out["gauc"].item()) - 4.0 / 7.0) < 1e-6


@pytest.mark.parametrize("cfg", ["cfg1", "cfg2", "cfg3"])
def test_baseline_configs(cfg):
    from rec_now_b200 import ops
    d = getattr(G, cfg)(0)
    s = np.round(d["s"], 2)                     # (rounded scores: plenty of ties)
    out = run(s, d["y"], d["g"])
    check(out, S.gauc(s, d["y"], d["g"]))
    assert ops.last_segmentation_path(out["_scratch"]) == 1
    out = run(s, d["y"], d["g"])                # the arena was left clean
    check(out, S.gauc(s, d["y"], d["g"]))


def test_radix_fallback_multi_key_mask_and_nan():
    from rec_now_b200 import ops
    rng = np.random.default_rng(7)
    b = 6000
    k0 = rng.integers(0, 40, b).astype(np.float32)
    k1 = rng.integers(0, 3, b).astype(np.float32)
    k0[rng.integers(0, b, 10)] = np.nan
    y = rng.integers(0, 30, b).astype(np.float32) / 4           # outside the level menu -> radix path
    y[::53] = np.nan
    s = np.round(rng.standard_normal(b), 1).astype(np.float32)
    mask = rng.random(b) < 0.85
    out = run(s, y, [k0, k1], mask)
    check(out, S.gauc(s, y, [k0, k1], mask))
    assert ops.last_segmentation_path(out["_scratch"]) == 2
    out = run(s, y, k0, mask)                    # one key, labels outside the menu: in-kernel fallback
    check(out, S.gauc(s, y, k0, mask))
    assert ops.last_segmentation_path(out["_scratch"]) == 2
    # perfect and inverted rankings
    yb = (rng.random(b) < 0.3).astype(np.float32)
    assert abs(float(run(yb, yb, k0)["gauc"].item()) - 1.0) < 1e-7
    assert abs(float(run(-yb, yb, k0)["gauc"].item())) < 1e-7
