"""Seeded synthetic workloads.  TEST / BENCH INFRASTRUCTURE.

Pure NumPy, no dependency on the product.  Every workload is a dict with
    g      int64[B]   sparse 62-bit group ids      (product / C-ABI form)
    g_f32  float32[B] dense ids as float32         (reference-compatible form, tests/rec_block style)
    s      float32[B] logits
    y      float32[B] labels
    w      float32[B] per-sample weights (when the config has them)
plus the option fields the config fixes.  Rows are NOT pre-sorted by group.
"""
from __future__ import annotations

import numpy as np


def _sparse_ids(rng, n_groups):
    ids = rng.integers(1, 2**62, size=n_groups, dtype=np.int64)
    ids = np.unique(ids)
    while ids.size < n_groups:                       # astronomically unlikely
        ids = np.unique(np.concatenate([ids, rng.integers(1, 2**62, size=n_groups, dtype=np.int64)]))
    ids = ids[:n_groups]
    rng.shuffle(ids)
    return ids


def zipf_groups(rng, b, n_groups):
    """p_k ~ 1/k for k = 1..G."""
    p = 1.0 / np.arange(1, n_groups + 1)
    p /= p.sum()
    return rng.choice(n_groups, size=b, p=p)


def _pack(rng, gidx, n_groups, labels, with_w):
    b = gidx.size
    ids = _sparse_ids(rng, n_groups)
    out = dict(g=ids[gidx], g_f32=gidx.astype(np.float32),
               s=rng.standard_normal(b).astype(np.float32), y=labels)
    if with_w:
        out["w"] = rng.uniform(0.5, 1.5, b).astype(np.float32)
    return out


def cfg1(seed=0, b=1024, n_groups=64):
    """pairwise, TF2-CPU style: uniform groups, binary labels, defaults."""
    rng = np.random.default_rng(seed)
    gidx = rng.integers(0, n_groups, b)
    y = (rng.random(b) < 0.25).astype(np.float32)
    return dict(_pack(rng, gidx, n_groups, y, False), name="cfg1", power=0.0, label_func="step")


def cfg2(seed=0, b=16384, n_groups=1024):
    """pairwise fwd+bwd: Zipf groups, binary labels, power 0, no mask."""
    rng = np.random.default_rng(seed)
    gidx = zipf_groups(rng, b, n_groups)
    y = (rng.random(b) < 0.25).astype(np.float32)
    return dict(_pack(rng, gidx, n_groups, y, False), name="cfg2", power=0.0, label_func="step")


def cfg3(seed=0, b=65536, n_groups=4096):
    """pairwise graded labels 0-4 + per-sample weights: W_ij = (y_i-y_j)[y_i>y_j] w_i, power -0.5."""
    rng = np.random.default_rng(seed)
    gidx = zipf_groups(rng, b, n_groups)
    y = rng.integers(0, 5, b).astype(np.float32)
    return dict(_pack(rng, gidx, n_groups, y, True), name="cfg3", power=-0.5, label_func="diff")


def cfg4(seed=0, b=65536, cap=512):
    """listwise: Zipf(1.5)-sized lists capped at `cap`, binary labels, one random row permutation."""
    rng = np.random.default_rng(seed)
    sizes = []
    tot = 0
    while tot < b:
        z = int(min(rng.zipf(1.5), cap))
        z = min(z, b - tot)
        sizes.append(z)
        tot += z
    gidx = np.repeat(np.arange(len(sizes)), sizes)
    gidx = gidx[rng.permutation(b)]
    y = (rng.random(b) < 0.25).astype(np.float32)
    return dict(_pack(rng, gidx, len(sizes), y, False), name="cfg4", n_lists=len(sizes))


def cfg5(world, seed=0, rows_per_rank=65536, groups_per_rank=4096):
    """global in-batch pairwise: one global draw of world*rows rows, G = groups_per_rank*world Zipf groups,
    graded labels + per-sample weights (same options as cfg3); rank r owns rows [r*rows, (r+1)*rows)."""
    d = cfg3(seed=seed, b=world * rows_per_rank, n_groups=world * groups_per_rank)
    d["name"] = f"cfg5_w{world}"
    d["rows_per_rank"] = rows_per_rank
    return d
