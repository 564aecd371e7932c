#!/usr/bin/env python
"""Generate tests/golden/reference_known_answers.json from the reference's OWN test files.

The reference (TensorFlow 2) cannot be imported in this image, so nothing is executed: the known-answer inputs
and expected values are parsed out of the literals in
    /root/reference/tests/rec_block/test_pairwise_loss_from_batch.py   (TPW)
    /root/reference/tests/rec_block/test_listwise_loss_from_batch.py   (TLW)
with `ast`, and written with the file:line each one came from.  Run it in the build container only
(/root/reference does not exist on the GPU box); the JSON it writes is committed.
"""
import ast
import json
import os
import re

REF = "/root/reference/tests/rec_block"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden",
                   "reference_known_answers.json")


def consts(path):
    """{line: literal} for every list / number literal that is the argument of tf.constant or an assignment."""
    src = open(path).read()
    tree = ast.parse(src)
    out = {}
    for node in ast.walk(tree):
        if isinstance(node, ast.Call) and getattr(node.func, "attr", "") == "constant" and node.args:
            try:
                out[node.lineno] = ast.literal_eval(node.args[0])
            except ValueError:
                pass
        if isinstance(node, ast.Assign) and isinstance(node.value, (ast.List, ast.Tuple)):
            try:
                out[node.lineno] = ast.literal_eval(node.value)
            except ValueError:
                pass
        if isinstance(node, ast.Call) and getattr(node.func, "attr", "") in ("assertAlmostEqual", "assertEquals"):
            for k in (1, 0):              # (expected value second -- TPW, TLW -- or first -- the focal test)
                try:
                    out[node.lineno] = ast.literal_eval(node.args[k])
                    break
                except (ValueError, IndexError):
                    pass
    return out, src.splitlines()


def main():
    pw, pw_src = consts(os.path.join(REF, "test_pairwise_loss_from_batch.py"))
    lw, lw_src = consts(os.path.join(REF, "test_listwise_loss_from_batch.py"))

    def line_of(src, pattern, start=0):
        for i, l in enumerate(src[start:], start + 1):
            if re.search(pattern, l):
                return i
        raise KeyError(pattern)

    g = {"source": {"TPW": "tests/rec_block/test_pairwise_loss_from_batch.py",
                    "TLW": "tests/rec_block/test_listwise_loss_from_batch.py"},
         "tolerance": 1e-4, "cases": []}
    l_gid = line_of(pw_src, r"group_id = \[")
    g["cases"].append({"name": "occurance_power_weight", "cite": f"TPW:{l_gid}-{l_gid + 5}", "group_id": pw[l_gid],
                       "power_-1": pw[line_of(pw_src, r"result1 = ")], "power_2": pw[line_of(pw_src, r"result2 = ")]})
    l_g = line_of(pw_src, r"sample_group_idx_var = ")
    exp = [pw[i + 1] for i, l in enumerate(pw_src) if "assertAlmostEqual(pairloss" in l]
    g["cases"].append({"name": "pairwise_loss", "cite": f"TPW:{l_g}-74", "groups": pw[l_g][0], "logits": pw[l_g + 1][0],
                       "labels": pw[l_g + 2][0], "click_occurance_power": -0.5,
                       "mask": pw[line_of(pw_src, r"mask = tf")][0],
                       "expected_plain": exp[0], "expected_with_weight_func": exp[1], "expected_with_mask": exp[2]})
    l1 = line_of(lw_src, r"sample_group_idx_var = ")
    l2 = line_of(lw_src, r"sample_group_idx_var = ", l1)
    exp_lw = [lw[i + 1] for i, l in enumerate(lw_src) if "assertAlmostEqual(listwise_loss" in l]
    exp_nv = [lw[i + 1] for i, l in enumerate(lw_src) if "assertEquals(n_valid_list" in l]
    for name, l, e, nv in (("listwise_loss", l1, exp_lw[0], exp_nv[0]), ("listwise_loss_case2", l2, exp_lw[1], exp_nv[1])):
        g["cases"].append({"name": name, "cite": f"TLW:{l}-{l + 12}", "groups": lw[l][0], "labels": lw[l + 1][0],
                           "logits": lw[l + 2][0], "expected_n_valid_list": nv, "expected_loss": e})
    # focal loss (fused by the product into the pairwise call): tests/rec_block/test_focal_loss.py
    fl, fl_src = consts(os.path.join(REF, "test_focal_loss.py"))
    g["source"]["TFL"] = "tests/rec_block/test_focal_loss.py"
    l_lab = line_of(fl_src, r"labels = tf.constant")
    l_log = line_of(fl_src, r"logits = tf.constant")
    exp_fl = [fl[i + 1] for i, l in enumerate(fl_src) if "assertAlmostEqual(" in l]
    g["cases"].append({"name": "focal_crossentropy_loss", "cite": f"TFL:{l_lab}-{len(fl_src)}", "labels": fl[l_lab],
                       "logits": fl[l_log], "expected_alpha_none_gamma_none": exp_fl[0],
                       "expected_alpha_0.25_gamma_none": exp_fl[1], "expected_alpha_none_gamma_1": exp_fl[2],
                       "tolerance": 1e-5})
    # slot pooling: tests/rec_block/test_embedding_util.py:55-106
    eu, eu_src = consts(os.path.join(REF, "test_embedding_util.py"))
    g["source"]["TEU"] = "tests/rec_block/test_embedding_util.py"
    l_t = line_of(eu_src, r"def test_sparse_batch_segment_ids_of_targets")
    l_sl = line_of(eu_src, r"slots = \[\[", l_t)
    g["cases"].append({"name": "sparse_batch_segment_ids_of_targets", "cite": f"TEU:{l_t}-{l_t + 15}",
                       "slots": eu[l_sl], "target_slots": eu[line_of(eu_src, r"target_slots = ", l_t)],
                       "expected_mask": eu[line_of(eu_src, r"expected_mask = ", l_t)],
                       "expected_sp_segment_ids": eu[line_of(eu_src, r"expected_sp_segment_ids = ", l_t)],
                       "num_rows": 2, "num_ids": 3, "num_segments": 6})
    l_p = line_of(eu_src, r"def test_embedding_using_sparse_batch_segment_ids")
    l_e1 = line_of(eu_src, r"expected_results = ", l_p)
    l_e2 = line_of(eu_src, r"expected_results = ", l_e1)
    # (the test derives its inputs in code -- params = [[i, -i]] for 40 keys, slots = int((ids + 0.5) / 10), weights =
    # ids * 10, TEU:72-91 -- only the ids and the expected values are literals)
    g["cases"].append({"name": "embedding_using_sparse_batch_segment_ids", "cite": f"TEU:{l_p}-{l_e2 + 6}",
                       "num_slots": 4, "num_keys_per_slot": 10, "target_slots": eu[line_of(eu_src, r"target_slots = ", l_p)],
                       "ids": eu[line_of(eu_src, r"ids = \[\[", l_p)],
                       "expected_with_weights": eu[l_e1],
                       # (TEU:104-108 ends in a stray comma: a 1-tuple holding the list)
                       "expected_without_weights": eu[l_e2][0] if isinstance(eu[l_e2], tuple) else eu[l_e2]})
    with open(OUT, "w") as f:
        json.dump(g, f, indent=1)
    print(json.dumps(g, indent=1))


if __name__ == "__main__":
    main()
