"""Op-for-op NumPy restatement of the reference's slot-embedding pooling (rec_block/embedding_util.py).

TEST INFRASTRUCTURE (see oracle/__init__.py).  EU:n = /root/reference/rec_now/rec_block/embedding_util.py:n.
Pinned against the reference's own test literals (tests/rec_block/test_embedding_util.py:55-106, committed as
tests/golden/reference_known_answers.json by scripts/make_golden.py).
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


def sparse_batch_segment_ids_of_targets(slots, target_slots):
    """EU:127-198.  Returns (mask bool[B,C], sp_segment_ids int32[kept], num_rows, num_ids, num_segments)."""
    slots = np.asarray(slots)
    target_slots = list(target_slots)
    lut = {k: v for v, k in enumerate(target_slots)}                       # StaticHashTable, default -1 (EU:181-186)
    segment_ids = np.vectorize(lambda s: lut.get(s, -1), otypes=[np.int32])(slots)
    mask = segment_ids >= 0                                                # EU:188
    num_rows, num_ids = slots.shape[0], len(target_slots)
    sp = segment_ids[mask]                                                 # boolean_mask: row-major order (EU:192)
    row = np.nonzero(mask)[0].astype(sp.dtype)                             # tf.where(mask)[:, 0] (EU:193)
    return mask, row * num_ids + sp, num_rows, num_ids, num_rows * num_ids  # EU:195-197


def embedding_using_sparse_batch_segment_ids(embedding_func, slots, target_slots, ids, weights=None, method="sum",
                                             use_unique=True):
    """EU:254-324: float32, the segment sum accumulated in flat (row-major) order as TF's CPU kernel does."""
    mask, seg, num_rows, num_ids, num_segments = sparse_batch_segment_ids_of_targets(slots, target_slots)
    sp_ids = np.asarray(ids)[mask]                                         # EU:303
    if use_unique:
        _, first = np.unique(sp_ids, return_index=True)                    # tf.unique: first-occurrence order (EU:305)
        uniq = sp_ids[np.sort(first)]
        inv = np.searchsorted(np.sort(uniq), sp_ids)
        inv = np.argsort(np.argsort(uniq))[inv] if uniq.size else inv
        emb = np.asarray(embedding_func(uniq), F32)[np.array([np.flatnonzero(uniq == v)[0] for v in sp_ids], int)] \
            if sp_ids.size else np.zeros((0, 1), F32)                      # gather (EU:310)
    else:
        emb = np.asarray(embedding_func(sp_ids), F32)                      # EU:312
    if weights is not None:
        emb = (emb * np.asarray(weights, F32)[mask][:, None]).astype(F32)  # EU:314-316
    d = emb.shape[1] if emb.ndim == 2 else 1
    out = np.zeros((num_segments, d), F32)
    cnt = np.zeros(num_segments, np.int64)
    for k in range(seg.size):                                              # unsorted_segment_sum, sequential (EU:318-321)
        out[seg[k]] = (out[seg[k]] + emb[k]).astype(F32)
        cnt[seg[k]] += 1
    if method == "mean":                                                   # unsorted_segment_mean: sum / max(count, 1)
        out = (out / np.maximum(cnt, 1).astype(F32)[:, None]).astype(F32)
    return out.reshape(num_rows, num_ids, -1)                              # EU:322
