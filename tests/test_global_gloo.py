"""world_size-2 gloo test (CPU) of the global in-batch mode's collective plumbing: all-gather order,
partition arguments, reduce-scatter slices and the loss all-reduce.  The CUDA call is replaced by an
oracle-backed stand-in that returns 1/world of the global result (so the summed partials must equal it)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import generators as G
from oracle import seg_ref as S


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, b_loc, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from rec_now_b200 import global_mode
    d = G.cfg5(world, seed=1, rows_per_rank=b_loc, groups_per_rank=16)
    lo, hi = rank * b_loc, (rank + 1) * b_loc
    seen = {}

    def fake_compute(gs, gy, gkeys, row_ok=None, rw_pos=None, label_func="step", factor=1.0, power=0.0,
                     reduce_mean=True, part=(0, 1)):
        seen["part"] = part
        seen["rows"] = gs.numel()
        r = S.pairwise(gs.numpy(), gy.numpy(), gkeys[0].numpy(),
                       S.PairSpec(factor=factor, power=power, label_func=label_func,
                                  rw_pos=None if rw_pos is None else rw_pos.numpy()))
        return dict(loss=torch.tensor(r["loss"] / part[1], dtype=torch.float32),
                    n_pair=torch.tensor(r["n_pair"]), dlogits=torch.tensor(r["grad"] / part[1], dtype=torch.float32))

    out = global_mode.global_pairwise_fwd_bwd(
        torch.tensor(d["s"][lo:hi]), torch.tensor(d["y"][lo:hi]), torch.tensor(d["g"][lo:hi]).reshape(1, -1),
        rw_pos=torch.tensor(d["w"][lo:hi]), label_func="diff", power=-0.5, _compute=fake_compute)
    assert seen["part"] == (rank, world) and seen["rows"] == world * b_loc
    ref = S.pairwise(d["s"], d["y"], d["g"], S.PairSpec(power=-0.5, label_func="diff", rw_pos=d["w"]))
    ok = (abs(float(out["loss"]) - ref["loss"]) < 1e-6 * abs(ref["loss"]) + 1e-9
          and int(out["n_pair"]) == ref["n_pair"]
          and np.abs(out["dlogits"].numpy() - ref["grad"][lo:hi]).max() < 1e-6 * np.abs(ref["grad"]).max() + 1e-9)
    # the two-collective path (packed blocks -> one all-gather; chunked outputs + loss slot -> one reduce-scatter)
    from rec_now_b200 import ops

    def fake_blocked(gbuf, world_, b_loc_, kk, has_w, has_ok, label_func="step", factor=1.0, power=0.0,
                     reduce_mean=True, part=(0, 1)):
        lay = ops.packed_block_layout(b_loc_, kk, has_w, has_ok)
        blk = gbuf.numpy().reshape(world_, lay["stride"])
        col = lambda name, dt, n: np.concatenate([blk[r, lay[name]:lay[name] + n].copy().view(dt) for r in range(world_)])
        gk, gs, gy = col("keys", np.int64, 8 * b_loc_), col("logits", np.float32, 4 * b_loc_), col("labels", np.float32, 4 * b_loc_)
        gw = col("w", np.float32, 4 * b_loc_) if has_w else None
        r = S.pairwise(gs, gy, gk, S.PairSpec(factor=factor, power=power, label_func=label_func, rw_pos=gw))
        chunk = b_loc_ + 4
        o = np.zeros((world_, chunk), np.float32)
        o[:, :b_loc_] = (r["grad"] / part[1]).reshape(world_, b_loc_)
        o[:, b_loc_] = r["loss"] / part[1]
        seen["blocked_part"] = part
        return dict(out=torch.tensor(o.reshape(-1)), n_pair=torch.tensor(r["n_pair"]), chunk=chunk)

    out2 = global_mode._packed_global(
        torch.tensor(d["s"][lo:hi]), torch.tensor(d["y"][lo:hi]), torch.tensor(d["g"][lo:hi]).reshape(1, -1),
        torch.tensor(d["w"][lo:hi]), None, "diff", 1.0, -0.5, True, None, _compute_blocked=fake_blocked)
    ok2 = (seen["blocked_part"] == (rank, world)
           and abs(float(out2["loss"]) - ref["loss"]) < 1e-6 * abs(ref["loss"]) + 1e-9
           and int(out2["n_pair"]) == ref["n_pair"]
           and np.abs(out2["dlogits"].numpy() - ref["grad"][lo:hi]).max() < 1e-6 * np.abs(ref["grad"]).max() + 1e-9)
    ret[rank] = bool(ok and ok2)
    dist.destroy_process_group()


def test_global_mode_plumbing_world2():
    world, b_loc = 2, 256
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, b_loc, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    assert all(ret.get(r) for r in range(world)), dict(ret)
