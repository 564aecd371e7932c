"""Shared helpers of the GPU parity tests."""
import numpy as np
import torch

from oracle import seg_ref as S


def dev(x, dtype=None):
    t = torch.as_tensor(np.ascontiguousarray(x))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


def check_pairwise(out, ref, tol=1e-5, ctx=""):
    """Parity gates of SURVEY 8d: exact integers, 1e-5 relative loss, gradient within 1e-5 of the
    per-row scale A_i (sum of absolute pair terms)."""
    n = int(out["n_pair"].item())
    assert n == ref["n_pair"], f"{ctx} n_pair {n} != {ref['n_pair']}"
    assert float(out["n_pair_f32"].item()) == float(np.float32(ref["n_pair"])), ctx
    if out.get("row_pairs") is not None:
        rp = out["row_pairs"].cpu().numpy()
        assert np.array_equal(rp, ref["row_pairs"]), f"{ctx} row_pairs differ at {np.flatnonzero(rp != ref['row_pairs'])[:8]}"
    loss = float(out["loss"].item())
    assert np.isfinite(loss), ctx
    assert abs(loss - ref["loss"]) <= tol * max(abs(ref["loss"]), 1e-30) + 1e-12, \
        f"{ctx} loss {loss} vs {ref['loss']} rel {abs(loss - ref['loss']) / max(abs(ref['loss']), 1e-30):.3e}"
    g = out["dlogits"].cpu().numpy().astype(np.float64)
    err = np.abs(g - ref["grad"])
    scale = ref["grad_abs"]
    bad = err > tol * scale + 1e-12
    assert not bad.any(), (f"{ctx} grad: {bad.sum()} rows off, worst rel-to-A "
                           f"{(err / np.maximum(scale, 1e-30))[bad].max():.3e}")
    return dict(loss_rel=abs(loss - ref["loss"]) / max(abs(ref["loss"]), 1e-30),
                grad_rel_A=float((err / np.maximum(scale, 1e-30))[scale > 0].max()) if (scale > 0).any() else 0.0)


def run_pairwise(s, y, groups, spec=S.PairSpec(), mask=None, part=(0, 1), want_row_pairs=True):
    """Product path through the C ABI for numpy inputs (groups: array or list of arrays)."""
    from rec_now_b200 import ops
    cols = groups if isinstance(groups, list) else [groups]
    keys, ok = ops.canon_keys([dev(c) for c in cols], None if mask is None else dev(np.asarray(mask, bool)))
    return ops.pairwise_fwd_bwd(
        dev(s), dev(y), keys, row_ok=ok,
        rw_pos=None if spec.rw_pos is None else dev(spec.rw_pos),
        rw_neg=None if spec.rw_neg is None else dev(spec.rw_neg),
        label_func=spec.label_func, factor=spec.factor, power=spec.power, only_wrong=spec.only_wrong,
        reduce_mean=spec.reduce_mean, part=part, want_row_pairs=want_row_pairs,
        pair_loss=spec.pair_loss, margin=spec.margin)
