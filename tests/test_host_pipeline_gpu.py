"""GPU parity of the host-buffer front end (rn_host_pairwise_*): host arrays in, host arrays out, several batches in
flight; results against the float64 segmented oracle and bit-identical pair counts."""
import numpy as np
import pytest
import torch

from oracle import generators as G
from oracle import seg_ref as S
from tests.util import check_pairwise

pytestmark = pytest.mark.gpu


def _outs(B, pinned):
    mk = (lambda n, dt: torch.empty(n, dtype=dt).pin_memory()) if pinned else (lambda n, dt: torch.empty(n, dtype=dt))
    return dict(loss=mk(1, torch.float32), n_pair_f32=mk(1, torch.float32), n_pair=mk(1, torch.int64),
                dlogits=mk(B, torch.float32), row_pairs=mk(B, torch.int64))


def _as_out(o):
    return {k: v.clone() for k, v in o.items()}


@pytest.mark.parametrize("pinned", [False, True])
def test_batches_in_flight(pinned):
    from rec_now_b200.host import HostPairwise
    B = 20000
    hp = HostPairwise(B, K=1, depth=3)
    batches, outs, tickets = [], [], []
    for seed in range(5):                      # more batches than slots: submit has to recycle them
        rng = np.random.default_rng(seed)
        n = B - 1000 * seed                    # ragged sizes below B_max
        g = rng.integers(0, 300, n).astype(np.int64)
        s = rng.standard_normal(n).astype(np.float32)
        y = rng.integers(0, 5, n).astype(np.float32)
        w = rng.uniform(0.5, 1.5, n).astype(np.float32)
        o = _outs(n, pinned)
        if pinned:
            g_, s_, y_, w_ = (torch.from_numpy(x).pin_memory() for x in (g, s, y, w))
        else:
            g_, s_, y_, w_ = g, s, y, w
        t = hp.submit(g_, s_, y_, rw_pos=w_, label_func="diff", power=-0.5, **o)
        batches.append((g, s, y, w)); outs.append(o); tickets.append(t)
        if seed >= 2:                          # results of an older batch while newer ones are in flight
            hp.wait(tickets[seed - 2])
    for t in tickets:
        hp.wait(t)
    hp.close()
    for (g, s, y, w), o in zip(batches, outs):
        ref = S.pairwise(s, y, g, S.PairSpec(label_func="diff", rw_pos=w, power=-0.5))
        check_pairwise(_as_out(o), ref, ctx="host pipeline")


def test_matches_device_path_cfg3():
    """Same batch through the device-pointer ABI and through the host front end: identical counts, loss and gradient
    equal up to the run-to-run jitter of the floating-point atomics."""
    from rec_now_b200 import ops
    from rec_now_b200.host import HostPairwise
    d = G.cfg3()
    B = d["s"].size
    dev_out = ops.pairwise_fwd_bwd(torch.tensor(d["s"]).cuda(), torch.tensor(d["y"]).cuda(),
                                   torch.tensor(d["g"]).cuda().reshape(1, -1), rw_pos=torch.tensor(d["w"]).cuda(),
                                   label_func="diff", power=-0.5)
    o = _outs(B, True); o.pop("row_pairs")
    hp = HostPairwise(B)
    hp.wait(hp.submit(d["g"], d["s"], d["y"], rw_pos=d["w"], label_func="diff", power=-0.5, **o))
    hp.close()
    assert int(o["n_pair"]) == int(dev_out["n_pair"].item())
    assert abs(float(o["loss"]) - float(dev_out["loss"].item())) <= 1e-6 * abs(float(o["loss"]))
    gd = dev_out["dlogits"].cpu().numpy()
    assert np.abs(o["dlogits"].numpy() - gd).max() <= 1e-6 * np.abs(gd).max()


def test_level_table_and_lambdarank_through_host_buffers():
    """RN_LABEL_LUT with the table in HOST memory (it is copied in with the batch's columns) and RN_LABEL_LAMBDA through the
    host front end, against the oracle with the callable itself."""
    from rec_now_b200.host import HostPairwise
    rng = np.random.default_rng(21)
    B = 12000
    g = rng.integers(0, 90, B).astype(np.int64)
    s = rng.standard_normal(B).astype(np.float32)
    y = rng.integers(0, 5, B).astype(np.float32)
    w = rng.uniform(0.5, 1.5, B).astype(np.float32)
    f = lambda a, b: (((a - b) ** 2 + 0.5 * a + 1.0) * (a > b)).astype(np.float32)
    lev = np.arange(-1, 7, dtype=np.float32)
    table = np.ascontiguousarray(f(np.broadcast_to(lev[:, None], (8, 8)), np.broadcast_to(lev[None, :], (8, 8))))
    hp = HostPairwise(B, depth=2)
    o1, o2 = _outs(B, False), _outs(B, False)
    t1 = hp.submit(g, s, y, rw_pos=w, label_func="lut", weight_lut=table, power=-0.5, **o1)
    t2 = hp.submit(g, s, y, rw_pos=w, label_func="lambda", **o2)
    hp.wait(t1); hp.wait(t2)
    hp.close()
    check_pairwise(_as_out(o1), S.pairwise(s, y, g, S.PairSpec(label_func="callable", weight_func=f, rw_pos=w, power=-0.5)),
                   ctx="host lut")
    check_pairwise(_as_out(o2), S.pairwise(s, y, g, S.PairSpec(label_func="lambda", rw_pos=w)), ctx="host lambda")


def test_argument_errors():
    from rec_now_b200 import _lib
    from rec_now_b200.host import HostPairwise
    hp = HostPairwise(1000)
    o = _outs(2000, False)
    with pytest.raises(_lib.RnError):          # batch larger than B_max
        hp.submit(np.zeros(2000, np.int64), np.zeros(2000, np.float32), np.zeros(2000, np.float32), **o)
    with pytest.raises(ValueError):            # device tensor where a host buffer is expected
        hp.submit(torch.zeros(10, dtype=torch.int64).cuda(), np.zeros(10, np.float32), np.zeros(10, np.float32), **_outs(10, False))
    hp.close()


@pytest.mark.parametrize("step_graph", ["1", "0"])
def test_bound_buffers_are_read_at_submit(monkeypatch, step_graph):
    """bind() validates the staging buffers once; every submit() copies what they hold at that moment -- on the eager path
    and when the slots launch their whole-step graphs (RN_HOST_STEP_GRAPH, read when the object is created)."""
    from rec_now_b200.host import HostPairwise
    monkeypatch.setenv("RN_HOST_STEP_GRAPH", step_graph)
    B = 5000
    rng = np.random.default_rng(7)
    g = torch.from_numpy(rng.integers(0, 50, B).astype(np.int64)).pin_memory()
    s = torch.from_numpy(rng.standard_normal(B).astype(np.float32)).pin_memory()
    y = torch.from_numpy(rng.integers(0, 2, B).astype(np.float32)).pin_memory()
    o = _outs(B, True)
    hp = HostPairwise(B)
    batch = hp.bind(g, s, y, **o)
    for rep in range(7):
        hp.wait(batch.submit())
        check_pairwise(_as_out(o), S.pairwise(s.numpy(), y.numpy(), g.numpy()), ctx=f"bound rep {rep}")
        s.copy_(torch.from_numpy(rng.standard_normal(B).astype(np.float32)))       # the loader refills its buffer
        y.copy_(torch.from_numpy(rng.integers(0, 2, B).astype(np.float32)))
    # two slots, the same pinned buffers every step: from its second submit on a slot launches its whole-step graph
    # (copy-in, kernels, copy-out as one launch) -- submits 2 .. 6 here
    want = 5 if step_graph == "1" else 0
    assert hp.graph_steps() == want, hp.graph_steps()
    # pageable buffers stay on the eager path
    o2 = _outs(B, False)
    b2 = hp.bind(g.numpy().copy(), s.numpy().copy(), y.numpy().copy(), **o2)
    for rep in range(4):
        hp.wait(b2.submit())
    check_pairwise(_as_out(o2), S.pairwise(s.numpy(), y.numpy(), g.numpy()), ctx="pageable")
    assert hp.graph_steps() == want
    hp.close()


@pytest.mark.parametrize("step_graph", ["0", "1"])
def test_soak_slots_on_their_own_streams(monkeypatch, step_graph):
    """Every slot enqueues on its own compute stream with its own instance of the cached graph, so the cooperative kernels
    of consecutive batches are in the launch queues at the same time.  4000 batches, three in flight: every result must
    be the first one's (exact pair count, loss and gradient up to the jitter of the float atomics), no device error."""
    from rec_now_b200.host import HostPairwise
    monkeypatch.setenv("RN_HOST_STEP_GRAPH", step_graph)          # (eager submits / whole-step graphs from each slot's second submit on)
    d = G.cfg3(3, b=30000, n_groups=1500)
    B = d["s"].size
    hin = {k: torch.from_numpy(np.ascontiguousarray(d[k])).pin_memory() for k in ("g", "s", "y", "w")}
    depth = 3
    hp = HostPairwise(B, depth=depth)
    outs = [dict(loss=torch.empty(1).pin_memory(), n_pair_f32=torch.empty(1).pin_memory(),
                 n_pair=torch.empty(1, dtype=torch.int64).pin_memory(), dlogits=torch.empty(B).pin_memory())
            for _ in range(depth)]
    bound = [hp.bind(hin["g"], hin["s"], hin["y"], rw_pos=hin["w"], label_func="diff", power=-0.5, **o) for o in outs]
    ref = S.pairwise(d["s"], d["y"], d["g"], S.PairSpec(label_func="diff", rw_pos=d["w"], power=-0.5))
    pending, seen = [], 0
    for k in range(4000):
        pending.append((bound[k % depth].submit(), k % depth))
        if len(pending) == depth:
            t, q = pending.pop(0)
            hp.wait(t)
            assert int(outs[q]["n_pair"]) == ref["n_pair"], f"batch {seen}"
            assert abs(float(outs[q]["loss"]) - ref["loss"]) <= 1e-5 * abs(ref["loss"]), f"batch {seen}"
            if seen % 500 == 0:
                check_pairwise(_as_out(outs[q]), ref, ctx=f"soak batch {seen}")
            seen += 1
    for t, q in pending:
        hp.wait(t)
        assert int(outs[q]["n_pair"]) == ref["n_pair"]
    assert hp.graph_steps() == (4000 - depth if step_graph == "1" else 0)
    hp.close()
