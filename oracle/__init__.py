"""CPU oracle for the rec_now in-batch ranking-loss hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and only as the checker (or as the
timed CPU baseline), never as a fallback for the CUDA path.

Contents
--------
dense_ref.py    op-for-op float32 NumPy restatement of the reference's dense
                (B,B) / (G,B) algorithm (pairwise_loss_from_batch.py and
                listwise_loss_from_batch.py).  Bit-exact masks, counts and
                row-major pair order.  Usable to B ~ 8K.
seg_ref.py      float64 segmented restatement (truth for loss / gradient at
                any B), NumPy.
generators.py   the seeded synthetic generators of SURVEY.md section 8d.
torch_dense.py  op-for-op torch-CPU float32 restatement with autograd backward:
                the timed "reference restatement (torch CPU), not TensorFlow"
                baseline (TensorFlow is absent from this image).

Parity pinning: the reference itself (TensorFlow 2) cannot run in this image,
so the oracle is pinned against the 8 golden numbers of the reference's own
unit tests (tests/rec_block/test_pairwise_loss_from_batch.py:20-74,
tests/rec_block/test_listwise_loss_from_batch.py:22-51); see
tests/test_oracle_golden.py.  Everything those tests do not cover (gradients,
wrong-order filter, multi-key groups, ...) is pinned only to the source lines
cited in each function: "parity unpinned" for those branches.
"""
