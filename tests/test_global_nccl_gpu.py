"""Global in-batch mode (needs >= 2 GPUs on the box; skipped otherwise): every rank's loss / n_pair / gradient slice
equals the single-batch oracle on the concatenated rows -- over NVLink peer mappings (the default: gather inside
the first kernel, peer-read reduction, three consecutive steps to exercise the alternating buffers) and over the NCCL
collectives (RN_GLOBAL_P2P=0)."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, b_loc, q, p2p, only_wrong=False):
    import torch
    import torch.distributed as dist
    from oracle import generators as G
    from rec_now_b200 import global_mode
    os.environ["RN_GLOBAL_P2P"] = "1" if p2p else "0"
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        d = G.cfg5(world, seed=3, rows_per_rank=b_loc, groups_per_rank=256)
        lo, hi = rank * b_loc, (rank + 1) * b_loc
        t = lambda k: torch.tensor(np.ascontiguousarray(d[k][lo:hi]), device="cuda")
        for _ in range(3 if p2p else 1):
            out = global_mode.global_pairwise_fwd_bwd(t("s"), t("y"), t("g").reshape(1, -1), rw_pos=t("w"),
                                                      label_func="diff", power=-0.5, only_wrong=only_wrong)
        torch.cuda.synchronize()
        if p2p:
            assert global_mode._peer_states and not global_mode._peer_broken, "peer-memory path was not taken"
            assert global_mode.last_segmentation_path() == 1, "the global step did not take the counting segmentation"
        q.put((rank, float(out["loss"].item()), int(out["n_pair"].item()), out["dlogits"].cpu().numpy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("p2p", [True, False])
@pytest.mark.parametrize("world", [2])
def test_global_pairwise_matches_oracle(world, p2p):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp
    from oracle import generators as G
    from oracle import seg_ref as S
    b_loc = 8192
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, b_loc, q, p2p)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    d = G.cfg5(world, seed=3, rows_per_rank=b_loc, groups_per_rank=256)
    ref = S.pairwise(d["s"], d["y"], d["g"], S.PairSpec(power=-0.5, label_func="diff", rw_pos=d["w"]))
    grad = np.concatenate([r[3] for r in res])
    for rank, loss, n, _ in res:
        assert n == ref["n_pair"]
        assert abs(loss - ref["loss"]) <= 1e-5 * abs(ref["loss"])
    err = np.abs(grad - ref["grad"])
    assert (err <= 1e-5 * ref["grad_abs"] + 1e-12).all(), err.max()


def test_global_wrong_order_filter_matches_oracle():
    """only_use_wrong_order_pair in the global mode (pairwise_loss_from_batch.py:197-203, counts at :282-291): the pair
    set depends on the scores, so n and c_h are known only after every rank has counted -- the two-stage call sums the
    ranks' per-row counts over peer memory between counting and weighting.  Against the oracle on the concatenated rows."""
    import torch
    world = 2
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp
    from oracle import generators as G
    from oracle import seg_ref as S
    b_loc = 8192
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, b_loc, q, True, True)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    d = G.cfg5(world, seed=3, rows_per_rank=b_loc, groups_per_rank=256)
    ref = S.pairwise(d["s"], d["y"], d["g"], S.PairSpec(power=-0.5, label_func="diff", rw_pos=d["w"], only_wrong=True))
    grad = np.concatenate([r[3] for r in res])
    for rank, loss, n, _ in res:
        assert n == ref["n_pair"]
        assert abs(loss - ref["loss"]) <= 1e-5 * abs(ref["loss"])
    err = np.abs(grad - ref["grad"])
    assert (err <= 1e-5 * ref["grad_abs"] + 1e-12).all(), err.max()
