"""The reference's own known-answer tests, restated for the oracle (no GPU).

Source of every number: /root/reference/tests/rec_block/test_pairwise_loss_from_batch.py:19-74 (TPW)
and /root/reference/tests/rec_block/test_listwise_loss_from_batch.py:18-51 (TLW).  These 8 values are
the only golden vectors the reference holds for the hot path; they pin oracle/dense_ref.py (op-for-op
float32) and oracle/seg_ref.py (float64 segmented).  Layout mirrors the reference tests.
"""
import unittest

import numpy as np

from oracle import dense_ref as D
from oracle import seg_ref as S


def col(v, dt=np.float32):
    return np.asarray([v], dtype=dt).T


class TestPairwiseLossFromBatch(unittest.TestCase):
    def test_occurance_power_weight(self):                      # TPW:19-31
        group_id = [1, 1, 2, 4, 4, 4]
        w1 = D.occurance_power_weight(group_id, power=-1)
        w2 = D.occurance_power_weight(group_id, power=2)
        for e, r in zip([0.5, 0.5, 1., 0.33333334, 0.33333334, 0.33333334], w1):
            self.assertAlmostEqual(e, r, delta=0.0001)
        for e, r in zip([4., 4., 1., 9., 9., 9.], w2):
            self.assertAlmostEqual(e, r, delta=0.0001)

    def test_pairwise_loss(self):                               # TPW:33-74
        g = col([1, 1, 2, 2, 2])
        logits = col([0, 1, 2, 3, 4])
        label = col([1.1, 0, 0, 1, 1])

        def pairwise_loss_func(outputs_pos, outputs_neg, weights):
            return D.bpr_loss_func(outputs_pos, outputs_neg, weights, 1.0)

        loss = D.pairwise_loss(logits, label, g, pairwise_loss_func,
                               only_use_wrong_order_pair=False, click_occurance_power=-0.5)
        self.assertAlmostEqual(float(loss), 0.5415076, delta=1e-4)

        def _label_pair_to_weight_func(label_matrix, label_matrix_transpose, **kwargs):
            return (label_matrix > label_matrix_transpose).astype(np.float32)

        loss_w = D.pairwise_loss(logits, label, g, pairwise_loss_func,
                                 only_use_wrong_order_pair=False, click_occurance_power=-0.5,
                                 label_pair_to_weight_func=_label_pair_to_weight_func)
        self.assertAlmostEqual(float(loss_w), 0.5415076, delta=1e-4)

        mask = col([True, True, False, False, False], bool)
        loss_m = D.pairwise_loss(logits, label, g, pairwise_loss_func,
                                 only_use_wrong_order_pair=False, click_occurance_power=-0.5, mask=mask)
        self.assertAlmostEqual(float(loss_m), 1.3132617, delta=1e-4)

    def test_pairwise_loss_segmented(self):                     # same three cases through seg_ref
        g, logits, label = col([1, 1, 2, 2, 2]), col([0, 1, 2, 3, 4]), col([1.1, 0, 0, 1, 1])
        r = S.pairwise(logits, label, g, S.PairSpec(power=-0.5))
        self.assertEqual(r["n_pair"], 3)
        self.assertAlmostEqual(r["loss"], 0.5415076, delta=1e-6)
        r = S.pairwise(logits, label, g, S.PairSpec(power=-0.5, rw_pos=np.ones(5, np.float32)))
        self.assertAlmostEqual(r["loss"], 0.5415076, delta=1e-6)
        r = S.pairwise(logits, label, g, S.PairSpec(power=-0.5), mask=[True, True, False, False, False])
        self.assertEqual(r["n_pair"], 1)
        self.assertAlmostEqual(r["loss"], 1.3132617, delta=1e-6)

    def test_pair_order_and_count(self):
        g, logits, label = col([1, 1, 2, 2, 2]), col([0, 1, 2, 3, 4]), col([1.1, 0, 0, 1, 1])
        full = D.pairwise_full(logits, label, g, click_occurance_power=-0.5)
        self.assertEqual(full["n_pair"], 3)
        self.assertEqual(full["pos_idx"].tolist(), [0, 3, 4])   # row-major (PW:217)
        self.assertEqual(full["neg_idx"].tolist(), [1, 2, 2])
        loss, n = D.pairwise_loss(logits, label, g, return_num_pair=True)
        self.assertEqual(float(n), 3.0)


class TestListwiseLoss(unittest.TestCase):
    def _run(self, g, labels, logits):
        sample_mask, lab, logit = D.to_listwise_sample(g, labels, logits)
        n_valid_list = lab.shape[0]
        with np.errstate(invalid="ignore"):
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                n_per = np.float32(np.mean(np.sum(sample_mask.astype(np.float32), axis=-1)))
        n_per = D.nan_to_zero(n_per)
        loss = D.listwise_loss_via_softmax_cross_entropy_with_logits(labels_for_softmax=lab,
                                                                     logits_for_softmax=logit)
        return n_valid_list, n_per, loss

    def test_listwise_loss(self):                               # TLW:18-34
        g = col([1, 1, 2, 1, 2, 2, 3, 4])
        labels = col([1, 1, 1, 0, 0, 0, 1, 0])
        logits = col([0.1, 0.01, 0.2, 0.001, 0.02, 0.002, 0.3, 0.4])
        n_valid_list, _, loss = self._run(g, labels, logits)
        self.assertEqual(n_valid_list, 2)
        self.assertAlmostEqual(float(loss), 1.0291535, delta=1e-4)
        r = S.listwise(g, labels, logits)
        self.assertEqual(r["n_valid"], 2)
        self.assertAlmostEqual(r["loss"], 1.0291535, delta=1e-6)

    def test_listwise_loss_case2(self):                         # TLW:36-51
        g, labels, logits = col([3, 4]), col([1, 0]), col([0.3, 0.4])
        n_valid_list, n_per, loss = self._run(g, labels, logits)
        self.assertEqual(n_valid_list, 0)
        self.assertAlmostEqual(float(loss), 0.0, delta=1e-4)
        self.assertEqual(float(n_per), 0.0)
        r = S.listwise(g, labels, logits)
        self.assertEqual(r["n_valid"], 0)
        self.assertEqual(r["loss"], 0.0)

    def test_nan_to_zero_rank(self):                            # LW:83-85
        with self.assertRaises(ValueError):
            D.nan_to_zero(np.zeros(3, np.float32))


if __name__ == "__main__":
    unittest.main()


def test_focal_loss_known_answers():
    """tests/rec_block/test_focal_loss.py of the reference: the three golden values pin the focal restatements (float32
    op-for-op and float64) that check the fused joint loss."""
    import json, os
    from oracle import dense_ref as D
    from oracle import seg_ref as S
    here = os.path.dirname(os.path.abspath(__file__))
    g = json.load(open(os.path.join(here, "golden", "reference_known_answers.json")))
    c = [x for x in g["cases"] if x["name"] == "focal_crossentropy_loss"][0]
    y, z = np.asarray(c["labels"], np.float32), np.asarray(c["logits"], np.float32)
    for key, kw in (("expected_alpha_none_gamma_none", dict(alpha=None, gamma=None)),
                    ("expected_alpha_0.25_gamma_none", dict(alpha=0.25, gamma=None)),
                    ("expected_alpha_none_gamma_1", dict(alpha=None, gamma=1))):
        assert abs(float(D.focal_crossentropy_loss(y, z, **kw)) - c[key]) < c["tolerance"], key
        assert abs(S.focal(y, z, alpha=kw["alpha"] or 0, gamma=kw["gamma"] or 0)["loss"] - c[key]) < c["tolerance"], key
    # the float64 gradient against central differences
    rng = np.random.default_rng(0)
    y = (rng.random(50) < 0.4).astype(np.float32); z = rng.standard_normal(50).astype(np.float32)
    for kw in (dict(alpha=0.25, gamma=2.0), dict(alpha=0, gamma=1.5), dict(alpha=0.7, gamma=0), dict(alpha=0.25, gamma=2.0, stop_weight_gradient=True)):
        r = S.focal(y, z, **kw)
        if kw.get("stop_weight_gradient"):
            continue
        eps = 1e-3
        for i in (0, 7, 23):
            zp, zm = z.astype(np.float64).copy(), z.astype(np.float64).copy()
            zp[i] += eps; zm[i] -= eps
            fd = (S.focal(y, zp.astype(np.float32), **kw)["loss"] - S.focal(y, zm.astype(np.float32), **kw)["loss"]) / (2 * eps)
            assert abs(fd - r["grad"][i]) < 2e-4 * max(1.0, abs(fd)), (kw, i, fd, r["grad"][i])


class TestEmbeddingUtil(unittest.TestCase):
    """tests/rec_block/test_embedding_util.py:55-106 of the reference (TEU), against oracle/pool_ref.py -- the checker of
    the segment-pooling kernel."""

    def _golden(self, name):
        import json
        import os
        g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_known_answers.json")))
        return [c for c in g["cases"] if c["name"] == name][0]

    def test_sparse_batch_segment_ids_of_targets(self):         # TEU:55-69
        from oracle import pool_ref as P
        c = self._golden("sparse_batch_segment_ids_of_targets")
        mask, sp, num_rows, num_ids, num_segments = P.sparse_batch_segment_ids_of_targets(c["slots"], c["target_slots"])
        self.assertEqual(mask.tolist(), c["expected_mask"])
        self.assertEqual(sp.tolist(), c["expected_sp_segment_ids"])
        self.assertEqual((num_rows, num_ids, num_segments), (c["num_rows"], c["num_ids"], c["num_segments"]))

    def test_embedding_using_sparse_batch_segment_ids(self):    # TEU:71-109
        from oracle import pool_ref as P
        c = self._golden("embedding_using_sparse_batch_segment_ids")
        params = np.array([[i, -i] for i in range(c["num_slots"] * c["num_keys_per_slot"])], np.float32)   # TEU:77-78
        ids = np.array(c["ids"])
        slots = ((ids.astype(np.float64) + 0.5) / 10.0).astype(np.int32)                                  # TEU:88-89
        weights = ids.astype(np.float32) * 10.0                                                            # TEU:91
        for uu in (True, False):
            out = P.embedding_using_sparse_batch_segment_ids(lambda i: params[np.asarray(i)], slots, c["target_slots"], ids,
                                                             weights=weights, use_unique=uu)
            self.assertEqual(out.tolist(), c["expected_with_weights"])
            out = P.embedding_using_sparse_batch_segment_ids(lambda i: params[np.asarray(i)], slots, c["target_slots"], ids,
                                                             use_unique=uu)
            self.assertEqual(out.tolist(), c["expected_without_weights"])
