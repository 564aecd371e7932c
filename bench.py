#!/usr/bin/env python
"""bench.py -- in-batch pairwise ranking loss, forward + backward, on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Metric (BASELINE.json): in-batch pairwise loss fwd+bwd throughput at B = 65536 rows per GPU, reported as
pairs/s (kept ordered pairs scored per second; `value`) and samples/s (`samples_per_s`).
  N = 1 : cfg3  -- 65536 rows, 4096 Zipf groups, graded labels 0-4, W_ij = (y_i-y_j)[y_i>y_j] w_i,
          click_occurance_power = -0.5  (the largest single-GPU configuration of BASELINE.json).
  N > 1 : cfg5  -- global in-batch mode, 65536 rows per GPU all-gathered over NVLink, same options
          (weak scaling: per-GPU rows fixed; pairs grow with the global batch).
A "step" is one full pass of the hot path over one batch: segmentation + pair kernel + finalisation
(+ all-gather / reduce-scatter for N > 1), producing loss, n_pair and d loss / d logits.

`value` is timed with inputs resident in HBM; `e2e` goes through the public drop-in API
(rec_now_b200.rec_block.pairwise_loss_from_batch.pairwise_loss + backward) with pinned HOST buffers, H2D and
D2H copies inside the timed region.  `--impl reference` times the reference's dense (B,B) algorithm on the host
cores (torch-CPU op-for-op restatement from oracle/torch_dense.py -- TensorFlow is not in this image) on a
bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ROWS_PER_GPU = 65536
NOMINAL_MUFU_PER_S = 148 * 16 * 1.965e9          # 148 SMs x 16 SFU lanes x max SM clock
MUFU_PER_PAIR = 3                                # ex2 + lg2 + rcp  (algorithmic work per kept pair)
HBM_BYTES_PER_SAMPLE = 24                        # g 8 + s 4 + y 4 + w 4 read, dlogits 4 written (cfg3)


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


def kpair_capture(world):
    """Figures of one k_pair launch from the committed `ncu --set full` capture (profiles/kpair_traffic.json, written
    from scripts/ncu_summary.py's output): DRAM bytes, MUFU operations actually issued per pair, pipe utilisations."""
    p = os.path.join(ROOT, "profiles", "kpair_traffic.json")
    if world != 1 or not os.path.exists(p):
        return {}
    with open(p) as f:
        return json.load(f)


def kpair_traffic(world):
    d = kpair_capture(world)
    return (d.get("dram_bytes_read", 0) + d.get("dram_bytes_write", 0)) if d else None


# ----------------------------------------------------------------------------------------------------------
# clocks sampler (NVML) -- runs in a thread during warm-up + timed region
# ----------------------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown",
               0x4: "sw_power_cap", 0x80: "hw_power_brake", 0x2: "applications_clocks_setting"}

    def __init__(self, index: int):
        self.samples, self.mask, self.sm_max, self._stop = [], 0, None, threading.Event()
        self.paused = False
        self.timed = [None, None]
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.t = threading.Thread(target=self._run, daemon=True)
            self.t.start()
        except Exception as e:          # NVML missing: report that, do not fake numbers
            self.nv = None
            self.err = repr(e)

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            if self.paused:                 # (host-timed regions: an NVML query holds driver locks for milliseconds)
                time.sleep(0.002)
                continue
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                self.samples.append((time.perf_counter(), mhz, r))
            except Exception:
                pass
            time.sleep(0.02)

    def summary(self):
        if self.nv is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "error": self.err}
        self._stop.set()
        self.t.join(timeout=1)
        t0, t1 = self.timed
        inside = [s for s in self.samples if t0 is not None and t0 <= s[0] <= t1] or self.samples[-5:]
        mask = 0
        for s in inside:
            mask |= s[2]
        reasons = [n for b, n in self.REASONS.items() if mask & b]
        return {"sm_mhz": float(np.median([s[1] for s in inside])) if inside else None,
                "sm_max_mhz": float(self.sm_max), "reasons": reasons, "samples": len(inside)}


# ----------------------------------------------------------------------------------------------------------
# workloads
# ----------------------------------------------------------------------------------------------------------
def bind_to_gpu_numa(index: int):
    """Run this process on the CPUs of the NUMA node its GPU hangs off (what `numactl --cpunodebind` / NCCL's affinity do for
    a production job): page-locked host buffers are then allocated next to the GPU's PCIe root.  Host-side e2e figures
    of this bench were bimodal across boxes (46 vs 76 us per step with the same device-timed step) before this.  Returns a
    record for the JSON line and the previous affinity (restored in front of the CPU-baseline legs, which use every core
    the process may use)."""
    rec = {"gpu_numa_node": None, "bound": False}
    before = None
    try:
        before = os.sched_getaffinity(0)
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(index)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        dom, rest = bus.split(":", 1)
        sysid = f"{int(dom, 16):04x}:{rest.lower()}"
        node = int(open(f"/sys/bus/pci/devices/{sysid}/numa_node").read())
        rec["gpu_numa_node"] = node
        cpu_now = os.sched_getcpu() if hasattr(os, "sched_getcpu") else -1
        rec["cpu_at_start"] = cpu_now
        if node < 0:
            return rec, before
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        rec["cpu_at_start_on_gpu_node"] = cpu_now in cpus
        target = cpus & before
        if target:
            os.sched_setaffinity(0, target)
            rec["bound"] = True
            rec["cpus"] = len(target)
    except Exception as e:          # (no sysfs / NVML: run as started)
        rec["note"] = f"{type(e).__name__}: {e}"[:120]
    return rec, before


def make_workload(world: int):
    from oracle import generators as G        # generators only (pure numpy); nothing of the oracle is timed here
    d = G.cfg3(0) if world == 1 else G.cfg5(world, 0)
    name = ("cfg3: pairwise B=65536, 4096 Zipf groups, graded labels 0-4, W=(yi-yj)[yi>yj]*w_i, power=-0.5"
            if world == 1 else
            f"cfg5: global in-batch pairwise, {ROWS_PER_GPU} rows/GPU x {world} GPUs, {4096 * world} Zipf groups, "
            "graded labels, W=(yi-yj)[yi>yj]*w_i, power=-0.5")
    return d, name


# ----------------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: dense (B,B) algorithm on the host cores
# ----------------------------------------------------------------------------------------------------------
def cpu_dense_step_fn(d, rows):
    import torch
    from oracle import torch_dense as T
    s = torch.tensor(d["s"][:rows]); y = torch.tensor(d["y"][:rows])
    g = torch.tensor(d["g"][:rows]).to(torch.float64)       # float ids, as the reference requires (PW:35)
    w = torch.tensor(d["w"][:rows]).reshape(-1, 1)
    wf = lambda a, b, sample_weight: (a - b) * (a > b).to(torch.float32) * sample_weight

    def step():
        loss, n, grad = T.pairwise_fwd_bwd(s, y, g, power=-0.5, weight_func=wf, sample_weight=w)
        return loss, n
    return step


def pick_cpu_rows(d, budget_s, steps):
    """Largest power-of-two sample whose (steps) dense fwd+bwd passes fit the time budget (cost ~ rows^2)."""
    step = cpu_dense_step_fn(d, 1024)
    step()
    t = time.perf_counter(); step(); t1k = time.perf_counter() - t
    rows = 1024
    while rows < 8192 and t1k * ((2 * rows) / 1024) ** 2 * steps <= budget_s:
        rows *= 2
    return rows


def run_cpu_dense(d, steps, warmup, budget_s):
    import torch
    # all host threads this process may use (torchrun exports OMP_NUM_THREADS=1 for its workers: undo that here)
    try:
        ncpu = len(os.sched_getaffinity(0))
    except AttributeError:
        ncpu = os.cpu_count() or 1
    if ncpu > torch.get_num_threads():
        torch.set_num_threads(ncpu)
    rows = pick_cpu_rows(d, budget_s, steps + warmup)
    step = cpu_dense_step_fn(d, rows)
    for _ in range(warmup):
        step()
    t = time.perf_counter()
    n = 0
    for _ in range(steps):
        _, n = step()
    dt = time.perf_counter() - t
    return dict(rows=rows, n_pair=n, sec_per_step=dt / steps, pairs_per_s=n * steps / dt,
                samples_per_s=rows * steps / dt, cores=torch.get_num_threads())


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    d, name = make_workload(max(1, args.gpus))
    r = run_cpu_dense(d, args.steps, args.warmup, budget_s=150.0)
    sample = (f"first {r['rows']} rows of the workload batch, dense (B,B) fwd+bwd (the reference algorithm needs "
              f">= 36*B^2 bytes: B=65536 is infeasible), {r['n_pair']} pairs/step")
    line = {
        "impl": "reference", "metric": "pairwise_loss_fwd_bwd_pairs_per_s", "value": r["pairs_per_s"],
        "unit": "pairs/s", "samples_per_s": r["samples_per_s"], "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": r["sec_per_step"] * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic (seeded Zipf groups)",
        "config": {"workload": name, "device": "host CPU", "implementation":
                   "op-for-op torch-CPU float32 restatement of the reference's dense TF graph (TensorFlow absent)"},
        "cpu_baseline": {"value": r["pairs_per_s"], "unit": "pairs/s", "cores": r["cores"], "kind": "port",
                         "sample": sample},
        "e2e": {"value": r["pairs_per_s"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit_line(line)



# ----------------------------------------------------------------------------------------------------------
# parity gates (run with every benchmark, outside the timed region) and the other BASELINE configs
# ----------------------------------------------------------------------------------------------------------
def parity_pairwise(out, d, spec_kw):
    """Exact pair count, 1e-5 loss, gradient within 1e-5 of the per-row scale A_i -- against the float64 segmented
    restatement of the reference (oracle/seg_ref.py: the checker, never the thing measured)."""
    from oracle import seg_ref as S
    ref = S.pairwise(d["s"], d["y"], d["g"], S.PairSpec(**spec_kw))
    n = int(out["n_pair"].item())
    loss = float(out["loss"].item())
    g = out["dlogits"].cpu().numpy().astype(np.float64)
    err = np.abs(g - ref["grad"])
    scale = ref["grad_abs"]
    loss_rel = abs(loss - ref["loss"]) / max(abs(ref["loss"]), 1e-30)
    grad_rel = float((err / np.maximum(scale, 1e-30))[scale > 0].max()) if (scale > 0).any() else 0.0
    rec = {"oracle": "oracle/seg_ref.py (float64, segmented)", "n_pair": n, "n_pair_exact": n == ref["n_pair"],
           "n_pair_f32_exact": float(out["n_pair_f32"].item()) == float(np.float32(ref["n_pair"])),
           "loss_rel": loss_rel, "grad_max_abs_err_over_A": grad_rel,
           "grad_max_abs_err_over_max_abs_grad": float(err.max() / max(np.abs(ref["grad"]).max(), 1e-30)),
           "tolerance": 1e-5}
    rec["ok"] = bool(rec["n_pair_exact"] and rec["n_pair_f32_exact"] and loss_rel <= 1e-5 and
                     bool((err <= 1e-5 * scale + 1e-12).all()))
    return rec


def parity_listwise(out, d):
    from oracle import seg_ref as S
    ref = S.listwise(d["g"], d["y"], d["s"])
    v = int(out["n_valid"].item())
    loss = float(out["loss"].item())
    g = out["dlogits"].cpu().numpy().astype(np.float64)
    loss_rel = abs(loss - ref["loss"]) / max(abs(ref["loss"]), 1e-30)
    grad_rel = float(np.abs(g - ref["grad"]).max() / max(np.abs(ref["grad"]).max(), 1e-30))
    return {"oracle": "oracle/seg_ref.py (float64, segmented)", "n_valid": v, "n_valid_exact": v == ref["n_valid"],
            "loss_rel": loss_rel, "grad_max_abs_err_over_max_abs_grad": grad_rel, "tolerance": 1e-5,
            "ok": bool(v == ref["n_valid"] and loss_rel <= 1e-5 and grad_rel <= 1e-5)}


def time_device(step, steps, warmup, flush):
    """Per-step CUDA events with the L2 flushed before every step (single-call latency), then the same steps back to
    back without the flush (streamed: what a training loop sees)."""
    import torch
    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for k in range(steps):
        flush.fill_(k & 0xFF)
        evs[k][0].record(); step(); evs[k][1].record()
    torch.cuda.synchronize()
    ms = [a.elapsed_time(b) for a, b in evs]
    b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    b0.record()
    for _ in range(steps):
        step()
    b1.record(); torch.cuda.synchronize()
    return {"single_call_us": float(np.mean(ms)) * 1e3, "single_call_median_us": float(np.median(ms)) * 1e3,
            "streamed_us": b0.elapsed_time(b1) / steps * 1e3}


def cpu_dense_full(d, kind, budget_s=60.0):
    """The reference's dense algorithm (torch-CPU port) at the FULL size of a config, all host threads: the one
    like-for-like CPU comparison (cfg1, cfg2 pairwise; cfg4 listwise).  None if the host lacks the memory."""
    import torch
    from oracle import torch_dense as T
    b = d["s"].size
    try:
        import psutil
        avail = psutil.virtual_memory().available
    except Exception:
        avail = 0
    need = (60 if kind == "pairwise" else 0) * b * b + (0 if kind == "pairwise" else 40 * b * d.get("n_lists", 1))
    if avail and need > 0.6 * avail:
        return {"skipped": f"needs ~{need / 2**30:.0f} GiB of host memory, {avail / 2**30:.0f} GiB available"}
    try:
        ncpu = len(os.sched_getaffinity(0))
    except AttributeError:
        ncpu = os.cpu_count() or 1
    if ncpu > torch.get_num_threads():
        torch.set_num_threads(ncpu)
    s, y = torch.tensor(d["s"]), torch.tensor(d["y"])
    g = torch.tensor(d["g"]).to(torch.float64)
    if kind == "pairwise":
        fn = lambda: T.pairwise_fwd_bwd(s, y, g, power=d["power"])
    else:
        fn = lambda: T.listwise_fwd_bwd(g, y, s)
    t = time.perf_counter(); r = fn(); first = time.perf_counter() - t           # (warm-up; also sizes the run)
    reps = int(max(1, min(5, budget_s / max(first, 1e-3) - 1)))
    t = time.perf_counter()
    for _ in range(reps):
        r = fn()
    dt = (time.perf_counter() - t) / reps
    return {"sec_per_step": dt, "steps": reps, "cores": torch.get_num_threads(), "same_config": True,
            "kind": "port", "implementation": "oracle/torch_dense.py (op-for-op dense torch-CPU port of the reference)",
            "result": {"loss": r[0], "count": r[1]}}


def other_configs(lib, dev, flush, mufu_peak, hbm_gbs, steps, with_cpu):
    """cfg1, cfg2 (pairwise, binary labels) and cfg4 (listwise): single-call and streamed step time, roofline
    fraction, parity record, and the dense CPU port at the full size of the config."""
    import torch
    from oracle import generators as G
    from rec_now_b200 import ops
    recs = {}
    for name in ("cfg1", "cfg2"):
        d = getattr(G, name)(0)
        s, y = torch.tensor(d["s"], device=dev), torch.tensor(d["y"], device=dev)
        keys = torch.tensor(d["g"], device=dev).reshape(1, -1)
        step = lambda: ops.pairwise_fwd_bwd(s, y, keys, label_func="step", power=0.0)
        t = time_device(step, steps, 5, flush)
        out = step()
        n = int(out["n_pair"].item())
        rec = {"workload": f"{name}: pairwise B={s.numel()}, binary labels", "rows": s.numel(), "n_pair": n, **t,
               "pairs_per_s": n / (t["single_call_us"] * 1e-6), "samples_per_s": s.numel() / (t["single_call_us"] * 1e-6),
               "sfu_roofline_frac_step": MUFU_PER_PAIR * n / (t["single_call_us"] * 1e-6) / mufu_peak,
               "hbm_roofline_frac_step": 20 * s.numel() / (t["single_call_us"] * 1e-6) / 1e9 / hbm_gbs,
               "segmentation_path": {1: "counting", 2: "radix", 3: "one-CTA kernel (small batch, csrc/small.cu)"}.get(
                   ops.last_segmentation_path(out["_scratch"]), "?"),
               "parity": parity_pairwise(out, d, dict(power=0.0))}
        if with_cpu:
            c = cpu_dense_full(d, "pairwise")
            if "sec_per_step" in c:
                c["pairs_per_s"] = n / c["sec_per_step"]
                c["n_pair_matches"] = c["result"]["count"] == n
                c["speedup_single_call"] = c["sec_per_step"] / (t["single_call_us"] * 1e-6)
            rec["cpu_full_size"] = c
        recs[name] = rec
    d = G.cfg4(0)
    s, y = torch.tensor(d["s"], device=dev), torch.tensor(d["y"], device=dev)
    keys = torch.tensor(d["g"], device=dev)
    step = lambda: ops.listwise_fwd_bwd(keys, y, s)
    t = time_device(step, steps, 5, flush)
    out = step()
    rec = {"workload": f"cfg4: listwise segmented softmax-CE B={s.numel()}, {d['n_lists']} Zipf-size lists (cap 512)",
           "rows": s.numel(), "n_valid_lists": int(out["n_valid"].item()), **t,
           "samples_per_s": s.numel() / (t["single_call_us"] * 1e-6),
           "hbm_roofline": {"bound": "hbm", "algorithmic_bytes_per_sample": 20, "peak_gbs": hbm_gbs,
                            "achieved_gbs_single_call": 20 * s.numel() / (t["single_call_us"] * 1e-6) / 1e9,
                            "frac_single_call": 20 * s.numel() / (t["single_call_us"] * 1e-6) / 1e9 / hbm_gbs,
                            "achieved_gbs_streamed": 20 * s.numel() / (t["streamed_us"] * 1e-6) / 1e9,
                            "frac_streamed": 20 * s.numel() / (t["streamed_us"] * 1e-6) / 1e9 / hbm_gbs,
                            "note": "1.3 MB of algorithmic traffic = 0.2 us of HBM time: the path is launch / latency "
                                    "bound at this size, the fraction says how far"},
           "parity": parity_listwise(out, d)}
    if with_cpu:
        c = cpu_dense_full(d, "listwise")
        if "sec_per_step" in c:
            c["samples_per_s"] = s.numel() / c["sec_per_step"]
            c["speedup_single_call"] = c["sec_per_step"] / (t["single_call_us"] * 1e-6)
        rec["cpu_full_size"] = c
    recs["cfg4"] = rec
    # GAUC: the evaluation metric on the same segmentation, cfg3's batch
    from oracle import seg_ref as S
    from rec_now_b200 import metrics
    d = G.cfg3(0)
    s, y, g = torch.tensor(d["s"], device=dev), torch.tensor(d["y"], device=dev), torch.tensor(d["g"], device=dev)
    step = lambda: metrics.gauc(s, y, g, return_details=True)
    t = time_device(step, steps, 5, flush)
    out = step()
    ref = S.gauc(d["s"], d["y"], d["g"])
    n = int(out["n_pair"].item())
    recs["gauc_cfg3"] = {"workload": "GAUC (group AUC over label-ordered pairs, ties 1/2, weight = rows of the group) on cfg3's batch",
                         "rows": s.numel(), "n_pair": n, **t, "pairs_per_s": n / (t["single_call_us"] * 1e-6),
                         "gauc": float(out["gauc"].item()),
                         "parity": {"oracle": "oracle/seg_ref.py::gauc (float64, exact integer counts)",
                                    "n_pair_exact": n == ref["n_pair"],
                                    "concordant2_exact": int(out["concordant2"].item()) == ref["concordant2"],
                                    "n_valid_groups_exact": int(out["n_valid_groups"].item()) == ref["n_valid_groups"],
                                    "gauc_abs_err": abs(float(out["gauc"].item()) - ref["gauc"]),
                                    "ok": bool(n == ref["n_pair"] and int(out["concordant2"].item()) == ref["concordant2"]
                                               and abs(float(out["gauc"].item()) - ref["gauc"]) <= 1e-6)}}
    # hinge (margin) pair loss on cfg3's batch: same pair set / weights, no SFU operation per pair
    w = torch.tensor(d["w"], device=dev)
    keys = g.reshape(1, -1)
    step = lambda: ops.pairwise_fwd_bwd(s, y, keys, rw_pos=w, label_func="diff", power=-0.5, pair_loss="hinge", margin=1.0)
    t = time_device(step, steps, 5, flush)
    out = step()
    spec = S.PairSpec(power=-0.5, label_func="diff", rw_pos=d["w"], pair_loss="hinge", margin=1.0)
    ref = S.pairwise(d["s"], d["y"], d["g"], spec)
    gerr = np.abs(out["dlogits"].cpu().numpy().astype(np.float64) - ref["grad"])
    hn = int(out["n_pair"].item())
    recs["hinge_cfg3"] = {"workload": "hinge pair loss max(0, 1 - x) on cfg3's batch (graded labels, label-gain x sample weights, power -0.5)",
                          "rows": s.numel(), "n_pair": hn, **t, "pairs_per_s": hn / (t["single_call_us"] * 1e-6),
                          "parity": {"oracle": "oracle/seg_ref.py (float64, pair_loss='hinge')", "n_pair_exact": hn == ref["n_pair"],
                                     "loss_rel": abs(float(out["loss"].item()) - ref["loss"]) / max(abs(ref["loss"]), 1e-30),
                                     "grad_max_abs_err_over_A": float((gerr / np.maximum(ref["grad_abs"], 1e-30))[ref["grad_abs"] > 0].max()),
                                     "tolerance": 1e-5,
                                     "ok": bool(hn == ref["n_pair"] and abs(float(out["loss"].item()) - ref["loss"]) <= 1e-5 * abs(ref["loss"])
                                                and (gerr <= 1e-5 * ref["grad_abs"] + 1e-12).all())}}
    # the rarer kernel variants on cfg3's batch: score- / weight-dependent pair sets (general tiles, counts from the kernel),
    # the level weight table (any label-only weight function: the label-gain tiles with a lookup) and LambdaRank weights
    # (rank pre-pass inside the pair kernel, general tiles)
    def lut_fn(a, b):
        return (((a - b) ** 2 + 0.5 * a + 1.0) * (a > b)).astype(np.float32)
    lev = np.arange(-1, 7, dtype=np.float32)
    lut_t = torch.tensor(lut_fn(np.broadcast_to(lev[:, None], (8, 8)), np.broadcast_to(lev[None, :], (8, 8))), device=dev)
    variants = (
        ("wrong_order_cfg3", dict(label_func="diff", only_wrong=True), dict(label_func="diff", only_wrong=True),
         "only_use_wrong_order_pair (k_pair variant with general tiles and kernel-side pair counts)"),
        ("rw_neg_cfg3", dict(label_func="diff", rw_neg=w), dict(label_func="diff", rw_neg=d["w"]),
         "a negative-side weight column (k_pair variant with general tiles and kernel-side pair counts)"),
        ("weight_lut_cfg3", dict(label_func="lut", weight_lut=lut_t), dict(label_func="callable", weight_func=lut_fn),
         "a label-only weight function that is none of the closed forms, W = ((y_i - y_j)^2 + y_i / 2 + 1) [y_i > y_j], as the "
         "8 x 8 level table RN_LABEL_LUT (fast tiles with a table lookup)"),
        ("lambdarank_cfg3", dict(label_func="lambda"), dict(label_func="lambda"),
         "LambdaRank |delta NDCG| pair weights RN_LABEL_LAMBDA (score ranks inside the groups worked out by a pre-pass of "
         "the pair kernel, general tiles)"),
    )
    for vname, vkw, skw, what in variants:
        step = lambda: ops.pairwise_fwd_bwd(s, y, keys, rw_pos=w, power=-0.5, **vkw)
        t = time_device(step, steps, 5, flush)
        out = step()
        spec = S.PairSpec(power=-0.5, rw_pos=d["w"], **skw)
        ref = S.pairwise(d["s"], d["y"], d["g"], spec)
        gerr = np.abs(out["dlogits"].cpu().numpy().astype(np.float64) - ref["grad"])
        vn = int(out["n_pair"].item())
        recs[vname] = {"workload": f"cfg3's batch with {what}",
                       "rows": s.numel(), "n_pair": vn, **t, "pairs_per_s": vn / (t["single_call_us"] * 1e-6),
                       "parity": {"oracle": "oracle/seg_ref.py (float64)", "n_pair_exact": vn == ref["n_pair"],
                                  "loss_rel": abs(float(out["loss"].item()) - ref["loss"]) / max(abs(ref["loss"]), 1e-30),
                                  "grad_max_abs_err_over_A": float((gerr / np.maximum(ref["grad_abs"], 1e-30))[ref["grad_abs"] > 0].max()),
                                  "tolerance": 1e-5,
                                  "ok": bool(vn == ref["n_pair"] and abs(float(out["loss"].item()) - ref["loss"]) <= 1e-5 * abs(ref["loss"])
                                             and (gerr <= 1e-5 * ref["grad_abs"] + 1e-12).all())}}
    # segment pooling of slot embeddings: HBM-bound gather + segment sum, one kernel
    from oracle import pool_ref as PR
    from rec_now_b200.rec_block import embedding_util as EU
    rng = np.random.default_rng(0)
    pb, pc, pt, pd, pv = 65536, 32, 8, 64, 1 << 20
    slots_h = rng.integers(0, 16, (pb, pc)).astype(np.int32)
    ids_h = rng.integers(0, pv, (pb, pc)).astype(np.int64)
    w_h = rng.uniform(0.5, 1.5, (pb, pc)).astype(np.float32)
    targets = [1, 3, 4, 7, 8, 10, 13, 15]
    table = torch.randn((pv, pd), device=dev)
    slots_t, ids_t, w_t = (torch.tensor(v, device=dev) for v in (slots_h, ids_h, w_h))
    step = lambda: EU.segment_pool(table, slots_t, targets, ids_t, w_t)
    t = time_device(step, steps, 5, flush)
    out = step()
    # forward + backward (gradient into the table: vector reductions; includes zeroing the 256 MiB gradient buffer)
    table_g = table.clone().requires_grad_(True)
    g_out = torch.randn_like(out)

    def step_fb():
        table_g.grad = None
        EU.segment_pool(table_g, slots_t, targets, ids_t, w_t).backward(g_out)
    t_fb = time_device(step_fb, max(10, steps // 4), 3, flush)
    kept = int(np.isin(slots_h, targets).sum())
    bytes_alg = kept * pd * 4 + pb * pt * pd * 4 + pb * pc * 16          # table rows in, pooled rows out, slots + ids + weights
    nchk = 256
    th = table.cpu().numpy()
    refp = PR.embedding_using_sparse_batch_segment_ids(lambda i: th[np.asarray(i)], slots_h[:nchk], targets, ids_h[:nchk],
                                                       weights=w_h[:nchk], use_unique=False)
    recs["segment_pool"] = {"workload": f"embedding_using_sparse_batch_segment_ids: B={pb} rows x {pc} columns, {pt} target slots of 16, "
                                        f"D={pd}, table {pv} x {pd} f32 (256 MiB), weights, method sum",
                            "rows": pb, "kept_ids": kept, **t,
                            "fwd_bwd_single_call_us": t_fb["single_call_us"], "fwd_bwd_streamed_us": t_fb["streamed_us"],
                            "hbm_roofline": {"bound": "hbm", "algorithmic_bytes": bytes_alg, "peak_gbs": hbm_gbs,
                                             "achieved_gbs_single_call": bytes_alg / (t["single_call_us"] * 1e-6) / 1e9,
                                             "frac_single_call": bytes_alg / (t["single_call_us"] * 1e-6) / 1e9 / hbm_gbs,
                                             "achieved_gbs_streamed": bytes_alg / (t["streamed_us"] * 1e-6) / 1e9,
                                             "frac_streamed": bytes_alg / (t["streamed_us"] * 1e-6) / 1e9 / hbm_gbs,
                                             "note": "random 256-byte table rows (sector-granular gathers), bytes = D*4 per kept id "
                                                     "+ the pooled output + the slot / id / weight columns"},
                            "parity": {"oracle": "oracle/pool_ref.py (op-for-op restatement of embedding_util.py:127-324) on the first "
                                                 f"{nchk} rows", "bit_exact": bool(np.array_equal(out[:nchk].cpu().numpy(), refp)),
                                       "ok": bool(np.array_equal(out[:nchk].cpu().numpy(), refp))}}
    return recs

# ----------------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------------
def ours(args):
    import torch
    import torch.distributed as dist
    from rec_now_b200 import _lib, global_mode, ops
    from rec_now_b200.rec_block import pairwise_loss_from_batch as PW

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != max(1, args.gpus):
        log(f"warning: --gpus {args.gpus} but WORLD_SIZE={world}; using WORLD_SIZE")
    numa_rec, affinity_before = bind_to_gpu_numa(local)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"          # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.lib()
    K, W = args.steps, max(3, args.warmup)

    d, name = make_workload(world)
    lo, hi = rank * ROWS_PER_GPU, (rank + 1) * ROWS_PER_GPU
    host = {k: np.ascontiguousarray(d[k][lo:hi]) for k in ("g", "s", "y", "w")}
    s, y, w = (torch.tensor(host[k], device=dev) for k in ("s", "y", "w"))
    keys = torch.tensor(host["g"], device=dev).reshape(1, -1)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2

    def step():
        if world == 1:
            return ops.pairwise_fwd_bwd(s, y, keys, rw_pos=w, label_func="diff", power=-0.5)
        return global_mode.global_pairwise_fwd_bwd(s, y, keys, rw_pos=w, label_func="diff", power=-0.5)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- clock ramp + measured SFU peak (same run, same clocks) ---------------------------------------
    sampler = ClockSampler(local) if rank == 0 else None
    sink = torch.zeros(4, dtype=torch.float32, device=dev)
    nops = C.c_int64(0)
    t_end = time.perf_counter() + 0.4
    while time.perf_counter() < t_end:
        lib.rn_bench_mufu(2000, sink.data_ptr(), C.byref(nops), C.c_void_p(torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 0.0
    for _ in range(5):
        e0.record()
        lib.rn_bench_mufu(4000, sink.data_ptr(), C.byref(nops), C.c_void_p(torch.cuda.current_stream().cuda_stream))
        e1.record(); torch.cuda.synchronize()
        best = max(best, nops.value / (e0.elapsed_time(e1) * 1e-3))
    mufu_peak = best

    # ---- warm-up, then the timed region (per-step events, L2 flushed between steps) --------------------
    for _ in range(W):
        out = step()
    barrier()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    g0 = int(lib.rn_debug_graph_launches())
    barrier()
    if sampler:
        sampler.timed[0] = time.perf_counter()
    for k in range(K):
        flush.fill_(k & 0xFF)
        evs[k][0].record()
        out = step()
        evs[k][1].record()
    barrier()
    if sampler:
        sampler.timed[1] = time.perf_counter()
    step_ms = [a.elapsed_time(b) for a, b in evs]
    t_ms = sum(step_ms)
    # the same K steps back to back, no L2 flush between them (what a training loop sees; reported beside `value`)
    b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    b0.record()
    for k in range(K):
        out = step()
    b1.record()
    barrier()
    b2b_ms = b0.elapsed_time(b1) / K
    graph_calls = int(lib.rn_debug_graph_launches()) - g0       # steps that went out as one CUDA-graph launch
    # ---- the dominant kernel alone: the same steps again with CUDA events around the k_pair launch (the events sit
    #      between the kernels of a call, so this pass uses plain launches instead of the cached graph) -----------
    KP = min(K, 100)

    def time_pair_kernel(in_graph):
        lib.rn_profile_enable_ex(KP, in_graph)
        for k in range(KP):
            flush.fill_(k & 0xFF)
            o = step()
        barrier()
        ms = (C.c_float * KP)(); nget = C.c_int32(0)
        lib.rn_profile_collect(ms, KP, C.byref(nget))
        lib.rn_profile_disable()
        return (float(np.mean(list(ms)[:nget.value])) if nget.value else float("nan")), o

    # on the default path (the step's cached CUDA graph with two event-record nodes around the pair kernel) ...
    pair_ms, out = time_pair_kernel(1)
    # ... and with plain launches and stream events between the kernels (includes the cooperative launch's latency)
    pair_ms_plain, out = time_pair_kernel(0)
    n_pair = int(out["n_pair"].item())
    err = ops.device_error(out["_scratch"]) if world == 1 else 0
    # ... and by the kernels' own %globaltimer stamps (CTA 0: first instruction behind the dependency wait -> last store of
    # the final pass): the kernel without its launch latency, cold L2 as above.  Supplementary: the roofline uses the events.
    pair_stamp_us = seg_stamp_us = None
    if world == 1:
        ts = (C.c_uint64 * 24)()
        dp, ds = [], []
        for k in range(20):
            flush.fill_(k & 0xFF)
            o = step()
            lib.rn_debug_timestamps(o["_scratch"].data_ptr(), ts, 24, C.c_void_p(torch.cuda.current_stream().cuda_stream))
            if ts[23] > ts[20] > ts[0] > 0:
                dp.append((ts[23] - ts[20]) / 1e3); ds.append((ts[19] - ts[0]) / 1e3 if ts[19] > ts[0] else float("nan"))
        if dp:
            pair_stamp_us, seg_stamp_us = float(np.median(dp)), float(np.nanmedian(ds))

    # ---- parity gate of the benchmarked workload (outside every timed region) ---------------------------
    spec_kw = dict(power=-0.5, label_func="diff")
    strong = None
    if world == 1:
        parity = parity_pairwise(out, d, dict(spec_kw, rw_pos=d["w"]))
        seg_path = {1: "counting (sort-free)", 2: "radix sort"}.get(ops.last_segmentation_path(out["_scratch"]), "?")
    else:
        # N > 1: the same global batch on ONE GPU (rank 0), through the single-GPU product path: exact pair count, loss,
        # this rank's gradient rows; the time of that run is the strong-scaling reference of the global mode
        seg_path = {1: "counting (sort-free), replicated on every rank; each rank scores the pairs whose positive row it owns",
                    2: "radix sort, replicated on every rank; the cost line is split across the ranks"}.get(
                        global_mode.last_segmentation_path(), "?")
        parity = None
        if rank == 0:
            gs, gy, gw = (torch.tensor(d[k], device=dev) for k in ("s", "y", "w"))
            gk = torch.tensor(d["g"], device=dev).reshape(1, -1)
            one = lambda: ops.pairwise_fwd_bwd(gs, gy, gk, rw_pos=gw, label_func="diff", power=-0.5)
            t1 = time_device(one, min(K, 20), 3, flush)
            ref = one()
            torch.cuda.synchronize()
            n1 = int(ref["n_pair"].item())
            l1, lN = float(ref["loss"].item()), float(out["loss"].item())
            g1 = ref["dlogits"][lo:hi].double().cpu().numpy(); gN = out["dlogits"].double().cpu().numpy()
            gerr = float(np.abs(g1 - gN).max() / max(np.abs(g1).max(), 1e-30))
            parity = {"against": "the same global batch on one GPU through rn_pairwise_fwd_bwd (rank 0)",
                      "n_pair": n_pair, "n_pair_exact": n1 == n_pair, "loss_rel": abs(l1 - lN) / max(abs(l1), 1e-30),
                      "grad_max_abs_err_over_max_abs_grad": gerr, "tolerance": 1e-5}
            parity["ok"] = bool(parity["n_pair_exact"] and parity["loss_rel"] <= 1e-5 and gerr <= 1e-5)
            strong = {"one_gpu_same_global_batch_us": t1["single_call_us"],
                      "segmentation_path_one_gpu": {1: "counting", 2: "radix"}.get(ops.last_segmentation_path(ref["_scratch"]), "?")}
            del gs, gy, gw, gk, ref

    # ---- e2e through the public API: pinned host buffers, H2D + loss.backward() + D2H every step -------
    # the step's inputs arrive as ONE pinned host buffer [g int64 | s f32 | y f32 | w f32] (what a data loader hands
    # over); it is copied to the device every step and the column tensors are views of the device buffer
    sizes = {"g": 8 * ROWS_PER_GPU, "s": 4 * ROWS_PER_GPU, "y": 4 * ROWS_PER_GPU, "w": 4 * ROWS_PER_GPU}
    pin_all = torch.empty(sum(sizes.values()), dtype=torch.uint8).pin_memory()
    o = 0
    for k, dt in (("g", torch.int64), ("s", torch.float32), ("y", torch.float32), ("w", torch.float32)):
        pin_all[o:o + sizes[k]].view(dt).copy_(torch.tensor(host[k]))
        o += sizes[k]
    # two device buffers: the copy of step k+1's inputs (copy stream) overlaps the compute of step k -- what an input
    # prefetcher does; every step still copies its own inputs from pinned host memory inside the timed region
    dev_all = [torch.empty_like(pin_all, device=dev) for _ in range(2)]
    dbuf = []
    for q in range(2):
        cols, o = {}, 0
        for k, dt in (("g", torch.int64), ("s", torch.float32), ("y", torch.float32), ("w", torch.float32)):
            cols[k] = dev_all[q][o:o + sizes[k]].view(dt)
            o += sizes[k]
        dbuf.append(cols)
    h_loss = torch.empty(1, dtype=torch.float32).pin_memory()
    h_grad = torch.empty(ROWS_PER_GPU, dtype=torch.float32).pin_memory()
    main = torch.cuda.current_stream(dev)
    cs = torch.cuda.Stream(dev)
    copied = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]

    # (cudaMemcpyAsync on the copy stream through ctypes: torch's stream context manager costs ~10 us of host time)
    try:
        rt = C.CDLL("libcudart.so.12")
        rt.cudaMemcpyAsync.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
        rt.cudaMemcpyAsync.restype = C.c_int
    except OSError:
        rt = None
    nbytes_in, cs_ptr, pin_ptr = pin_all.numel(), C.c_void_p(cs.cuda_stream), pin_all.data_ptr()
    dev_ptr = [t.data_ptr() for t in dev_all]
    h_loss_ptr, h_grad_ptr, main_ptr = h_loss.data_ptr(), h_grad.data_ptr(), C.c_void_p(main.cuda_stream)
    outs = torch.cuda.Stream(dev)
    outs_ptr = C.c_void_p(outs.cuda_stream)
    done_ev = [torch.cuda.Event(), torch.cuda.Event()]

    def enqueue_copy(q):
        cs.wait_event(consumed[q])                       # the step that last read this buffer is done with it
        if rt is not None:
            if rt.cudaMemcpyAsync(dev_ptr[q], pin_ptr, nbytes_in, 1, cs_ptr) != 0:
                raise RuntimeError("cudaMemcpyAsync failed")
        else:
            with torch.cuda.stream(cs):
                dev_all[q].copy_(pin_all, non_blocking=True)
        copied[q].record(cs)

    def e2e_steps(n):
        for q in range(2):
            consumed[q].record(main)
        enqueue_copy(0)
        for k in range(n):
            q = k & 1
            if k + 1 < n:
                enqueue_copy(1 - q)
            main.wait_event(copied[q])
            cols = dbuf[q]
            logits = cols["s"].requires_grad_(True)
            if world == 1:
                loss = PW.pairwise_loss(logits, cols["y"], cols["g"], click_occurance_power=-0.5,
                                        label_pair_to_weight_func=PW.label_gain_times_sample_weight,
                                        sample_weight=cols["w"])
            else:
                # global mode: the forward + backward entry point (ONE C-ABI call per step and rank,
                # rn_global_pairwise_fwd_bwd: loss and d loss / d logits of this rank's rows come back together)
                res = global_mode.global_pairwise_fwd_bwd(cols["s"], cols["y"], cols["g"].reshape(1, -1), rw_pos=cols["w"],
                                                          label_func="diff", power=-0.5)
                loss, grad_t = res["loss"], res["dlogits"]
            if world == 1:
                loss.backward()
                grad_t = logits.grad
            if rt is not None and world > 1:
                # D2H of the step's results on a copy-out stream (behind an event of the compute stream), so that the
                # next step's kernels do not queue behind the copy; the allocator is told the stream uses the tensors
                done_ev[q].record(main)
                outs.wait_event(done_ev[q])
                rt.cudaMemcpyAsync(h_loss_ptr, loss.data_ptr(), 4, 2, outs_ptr)
                rt.cudaMemcpyAsync(h_grad_ptr, grad_t.data_ptr(), 4 * ROWS_PER_GPU, 2, outs_ptr)
                loss.record_stream(outs); grad_t.record_stream(outs)
            elif rt is not None:                         # D2H of the step's results on the compute stream
                rt.cudaMemcpyAsync(h_loss_ptr, loss.data_ptr(), 4, 2, main_ptr)
                rt.cudaMemcpyAsync(h_grad_ptr, grad_t.data_ptr(), 4 * ROWS_PER_GPU, 2, main_ptr)
            else:
                h_loss.copy_(loss.detach().reshape(1), non_blocking=True)
                h_grad.copy_(grad_t, non_blocking=True)
            consumed[q].record(main)
            logits.grad = None
            cols["s"].requires_grad_(False)

    # the backward graph is one node: running the autograd engine on the calling thread saves its thread hand-off
    with torch.autograd.set_multithreading_enabled(False):
        e2e_steps(W)
        # (the region is enqueued by a Python loop: one host hiccup -- an NVML query of the sampler thread, a GC pause --
        #  starves the GPU for a millisecond.  Five repetitions of the K steps, the median reported, the sampler resting)
        if sampler:
            sampler.paused = True
        rep_ms = []
        for _ in range(5):
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            e2e_steps(K)
            cs.synchronize()
            main.wait_stream(outs)                       # (the last results are in host memory when the clock stops)
            b.record()
            barrier()
            rep_ms.append(a.elapsed_time(b))
        if sampler:
            sampler.paused = False
    e2e_ms = float(np.median(rep_ms))
    e2e_reps = rep_ms if world > 1 else []
    h2d = pin_all.numel()
    d2h = h_loss.numel() * 4 + h_grad.numel() * 4
    e2e_api_ms = e2e_ms
    e2e_api = ("rec_block.pairwise_loss_from_batch.pairwise_loss + backward" if world == 1
               else "global_mode.global_pairwise_fwd_bwd (torch device tensors in, loss + d loss / d logits out: one "
                    "rn_global_pairwise_fwd_bwd call per step and rank)")
    e2e_pipeline = ("inputs of step k+1 copied H2D on a copy stream (two device buffers) while step k computes; loss and "
                    "gradient copied D2H every step" + ("; autograd engine single-threaded "
                    "(torch.autograd.set_multithreading_enabled(False))" if world == 1 else " on a copy-out stream"))
    e2e_timing = "CUDA events on the compute stream around the K steps; median of five repetitions (reps_ms_per_step)"
    if world == 1:
        # ---- e2e through the C ABI with HOST buffers (rn_host_pairwise_*): every step copies its pinned host columns
        #      to the device, runs the three kernels and copies loss, pair count and gradient back; two slots in flight
        from rec_now_b200.host import HostPairwise
        DEPTH = 3
        hp = HostPairwise(ROWS_PER_GPU, K=1, depth=DEPTH)
        hin = {k: torch.from_numpy(host[k]).pin_memory() for k in ("g", "s", "y", "w")}
        houts = [dict(loss=torch.empty(1, dtype=torch.float32).pin_memory(),
                      n_pair_f32=torch.empty(1, dtype=torch.float32).pin_memory(),
                      n_pair=torch.empty(1, dtype=torch.int64).pin_memory(),
                      dlogits=torch.empty(ROWS_PER_GPU, dtype=torch.float32).pin_memory()) for _ in range(DEPTH)]

        # (the loader's staging buffers are bound once; every submit copies what they hold at that moment)
        bound = [hp.bind(hin["g"], hin["s"], hin["y"], rw_pos=hin["w"], label_func="diff", power=-0.5, **houts[q])
                 for q in range(DEPTH)]

        def host_steps(n):
            pending = []
            for k in range(n):
                pending.append(bound[k % DEPTH].submit())
                if len(pending) == DEPTH:
                    hp.wait(pending.pop(0))      # the results of step k-2 are in host memory
            for t in pending:
                hp.wait(t)

        host_steps(W)
        torch.cuda.synchronize()
        # host wall clock over K steps is ~10 ms: one scheduling hiccup of the host doubles it.  The region is therefore
        # run five times and the MEDIAN repetition reported (all five are in the line); the NVML sampler thread rests
        # meanwhile (the clocks line covers the device-timed region above)
        if sampler:
            sampler.paused = True
        e2e_reps = []
        for _ in range(5):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            host_steps(K)
            e2e_reps.append((time.perf_counter() - t0) * 1e3)
        if sampler:
            sampler.paused = False
        e2e_ms = float(np.median(e2e_reps))
        assert int(houts[(K - 1) % DEPTH]["n_pair"]) == n_pair
        hp.close()
        h2d = sum(int(t.numel() * t.element_size()) for t in hin.values())
        d2h = 4 + 4 + 8 + 4 * ROWS_PER_GPU
        e2e_api = "C ABI with host buffers: rn_host_pairwise_submit / rn_host_pairwise_wait (rec_now_b200.host.HostPairwise)"
        e2e_pipeline = ("every step: H2D of its pinned host columns (copy-in stream) -> k_init, k_seg, k_pair (compute "
                        "stream) -> D2H of loss, n_pair and the gradient (copy-out stream); three device slots, the host "
                        "waits for step k-2 after submitting step k")
        e2e_timing = ("host wall clock from the first submit to the last wait (device idle before, results in host memory "
                      "after); median of five repetitions of the K steps, all listed in reps_ms_per_step")

    # ---- max over ranks ------------------------------------------------------------------------------
    times = torch.tensor([t_ms, e2e_ms, pair_ms, e2e_api_ms, pair_ms_plain], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    t_ms, e2e_ms, pair_ms, e2e_api_ms, pair_ms_plain = (float(x) for x in times.tolist())

    if rank == 0:
        peaks, peak_src = measured_peaks()
        clocks = sampler.summary()
        if world == 1:
            kernels_per_step, kernel_names = 2, ["k_seg<HeadsTail> (count / offsets / scatter)", "k_pair"]
            if not seg_path.startswith("counting"):
                kernels_per_step, kernel_names = 3, ["k_init", "k_seg<HeadsTail>", "k_pair"]
        else:
            kernels_per_step = 7
            kernel_names = ["k_pack", "k_xbar", "k_init (+ peer gather)", "k_seg<HeadsTail>", "k_pair", "k_xbar", "k_reduce_out"]
        value = n_pair * K / (t_ms * 1e-3)
        rows_total = ROWS_PER_GPU * world
        # roofline of the dominant kernel (k_pair): SFU-bound -- 3 MUFU per kept pair; this rank scored 1/world
        # of the pairs per launch
        pairs_per_launch = n_pair / world
        achieved = MUFU_PER_PAIR * pairs_per_launch / (pair_ms * 1e-3)
        line = {
            "metric": "pairwise_loss_fwd_bwd_pairs_per_s", "value": value, "unit": "pairs/s",
            "samples_per_s": rows_total * K / (t_ms * 1e-3), "n_pair_per_step": n_pair,
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": t_ms / K, "higher_is_better": True,
            "step_us": {"median": float(np.median(step_ms)) * 1e3, "p10": float(np.percentile(step_ms, 10)) * 1e3,
                        "p90": float(np.percentile(step_ms, 90)) * 1e3, "back_to_back_no_flush": b2b_ms * 1e3,
                        "note": "rank 0's per-step CUDA events (single-call latency, cold L2); back to back = the K "
                                "steps enqueued without the flush, one event pair around all of them"},
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic (seeded Zipf groups)",
            "config": {"workload": name, "rows_per_gpu": ROWS_PER_GPU, "seed": 0,
                       "l2": "flushed between timed steps (256 MiB write); inputs ~1.6 MB/GPU",
                       "timing": "CUDA events per step on the launch stream, summed over steps, max over ranks",
                       "multi_gpu": None if world == 1 else (
                           "ONE C-ABI call per step (rn_global_pairwise_fwd_bwd) over NVLink peer mappings, no collective "
                           "calls: pack kernel -> device-side flag barrier -> ONE graph launch whose first kernel gathers every "
                           "rank's row block with peer loads -> replicated sort-free segmentation -> each rank scores the "
                           "pairs whose positive row it owns -> flag barrier -> one kernel sums this rank's gradient chunk "
                           "from the peers' buffers"
                           if global_mode.exchange_path() == "peer" else
                           "ONE NCCL all-gather of packed per-rank row blocks -> replicated segmentation on the blocked "
                           "rows -> even work-unit split -> ONE NCCL reduce-scatter (gradient chunks, partial loss in a "
                           "spare slot)")},
            "roofline": {"bound": "sfu", "kernel": "k_pair", "achieved": achieved / 1e9, "peak": mufu_peak / 1e9,
                         "unit": "GMUFU/s", "frac": achieved / mufu_peak,
                         "peak_source": "measured in this run (rn_bench_mufu ex2/lg2/rcp chains)",
                         "nominal_peak": NOMINAL_MUFU_PER_S / 1e9, "frac_of_nominal": achieved / NOMINAL_MUFU_PER_S,
                         "kernel_ms": pair_ms, "kernel_share_of_step": pair_ms / (t_ms / K),
                         "kernel_us_by_device_stamps": pair_stamp_us, "k_seg_us_by_device_stamps": seg_stamp_us,
                         "frac_by_device_stamps": (MUFU_PER_PAIR * pairs_per_launch / (pair_stamp_us * 1e-6) / mufu_peak
                                                   if pair_stamp_us else None),
                         "kernel_timing": "CUDA event-record nodes around the k_pair node of the step's CUDA graph (the "
                                          "default launch path), L2 flushed before every step",
                         "kernel_ms_plain_launches": pair_ms_plain,
                         "frac_plain_launches": MUFU_PER_PAIR * pairs_per_launch / (pair_ms_plain * 1e-3) / mufu_peak,
                         "algorithmic_mufu_per_pair": MUFU_PER_PAIR, "pairs_per_launch": pairs_per_launch,
                         "traffic": kpair_traffic(world),
                         # what the kernel really issues (ncu capture): the product-form tiles need ~1.5 MUFU per pair, so
                         # the XU pipe itself is far from full and the kernel is bound by issue slots (FP32x2, SHFL, MUFU
                         # sharing the schedulers); `frac` above is the 3-MUFU algorithmic convention of SURVEY 8d
                         "issued": {k: kpair_capture(world).get(k) for k in
                                    ("mufu_lane_ops_per_pair_issued", "xu_pipe_pct_of_peak_active", "issue_active_pct",
                                     "fma_pipe_pct_of_peak_active", "alu_pipe_pct_of_peak_active",
                                     "thread_instructions_per_pair", "source")},
                         "binding_resource_measured": "issue slots (see `issued`); the SFU roofline is the algorithmic yardstick",
                         "hbm_view": {"bound": "hbm", "achieved": HBM_BYTES_PER_SAMPLE * rows_total / (pair_ms * 1e-3) / 1e9,
                                      "peak": peaks["hbm_gbs"], "unit": "GB/s", "peak_source": peak_src,
                                      "frac": HBM_BYTES_PER_SAMPLE * rows_total / (pair_ms * 1e-3) / 1e9 / peaks["hbm_gbs"]}},
            "e2e": {"value": n_pair * K / (e2e_ms * 1e-3), "unit": "pairs/s", "ms_per_step": e2e_ms / K,
                    "samples_per_s": rows_total * K / (e2e_ms * 1e-3),
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": e2e_api, "pipeline": e2e_pipeline, "timing": e2e_timing,
                    "reps_ms_per_step": [x / K for x in e2e_reps],
                    "dropin_api": {"value": n_pair * K / (e2e_api_ms * 1e-3), "unit": "pairs/s",
                                   "ms_per_step": e2e_api_ms / K,
                                   "api": "rec_block.pairwise_loss_from_batch.pairwise_loss + backward on torch tensors "
                                          "(H2D / D2H every step, Python + autograd-engine host time included)"}
                                  if world == 1 else None},
            "gpu_launches": kernels_per_step * K,
            "launch_mode": {"kernels_per_step": kernels_per_step, "kernels": kernel_names,
                            "steps_enqueued_as_one_cuda_graph_launch": graph_calls},
            "segmentation_path": seg_path,
            "parity": parity,
            "step_breakdown_us": {"step": t_ms / K * 1e3, "k_pair": pair_ms * 1e3, "non_pair": (t_ms / K - pair_ms) * 1e3,
                                  "step_sfu_roofline_frac": MUFU_PER_PAIR * pairs_per_launch / (t_ms / K * 1e-3) / mufu_peak},
            "clocks": clocks, "device_error": err,
            "host_placement": numa_rec,
        }
        if affinity_before:
            # the CPU legs below use every core the process may use: the affinity goes back on EVERY thread of the process
            # (worker threads created meanwhile inherited the narrow mask)
            for tid in [0] + [int(t) for t in os.listdir("/proc/self/task")]:
                try:
                    os.sched_setaffinity(tid, affinity_before)
                except OSError:
                    pass
        if world == 1 and not args.no_cpu:
            r = run_cpu_dense(d, steps=3, warmup=1, budget_s=25.0)
            line["cpu_baseline"] = {
                "value": r["pairs_per_s"], "unit": "pairs/s", "cores": r["cores"], "kind": "port",
                "samples_per_s": r["samples_per_s"],
                "sample": f"first {r['rows']} rows of the cfg3 batch through the dense (B,B) torch-CPU restatement of "
                          f"the reference (fwd+bwd, {r['n_pair']} pairs/step, {r['sec_per_step']:.2f} s/step); the "
                          "full B=65536 needs >= 155 GB of dense temporaries"}
        if world == 1 and not args.no_configs:
            # the other BASELINE.json configurations (cfg1, cfg2 binary pairwise; cfg4 listwise) + this one, one record each
            cfgs = other_configs(lib, dev, flush, mufu_peak, peaks["hbm_gbs"], min(K, 100), not args.no_cpu)
            cfgs["cfg3"] = {"workload": name, "rows": ROWS_PER_GPU, "n_pair": n_pair,
                            "single_call_us": t_ms / K * 1e3, "streamed_us": b2b_ms * 1e3,
                            "pairs_per_s": value, "samples_per_s": line["samples_per_s"],
                            "sfu_roofline_frac_step": line["step_breakdown_us"]["step_sfu_roofline_frac"],
                            "sfu_roofline_frac_k_pair": line["roofline"]["frac"], "segmentation_path": seg_path,
                            "parity": parity, "cpu_full_size": {"skipped": "the dense reference algorithm needs >= 155 GB at "
                                                                "B=65536 (see cpu_baseline for the bounded sample)"}}
            line["configs"] = cfgs
        if world > 1:
            # what the pairs/s figure hides (pairs grow with the square of the group sizes): samples/s, the share of the
            # step that is not pair scoring, and the strong-scaling view -- the same global batch on one GPU
            gm = {"rows_total": rows_total, "samples_per_s": line["samples_per_s"],
                  "k_pair_us": pair_ms * 1e3, "non_pair_us": (t_ms / K - pair_ms) * 1e3}
            if strong:
                gm.update(strong)
                gm["speedup_over_one_gpu_same_batch"] = strong["one_gpu_same_global_batch_us"] / (t_ms / K * 1e3)
                gm["parallel_efficiency_same_batch"] = gm["speedup_over_one_gpu_same_batch"] / world
            line["global_mode"] = gm
        emit_line(line)
    if world > 1:
        dist.destroy_process_group()


def main():
    # stdout carries exactly ONE line, the JSON: everything else that a library prints there (the NCCL version banner
    # at communicator creation, for one) is sent to stderr at the file-descriptor level
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    global emit_line

    def emit_line(obj):
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(obj) + "\n").encode())

    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-configs", action="store_true", help="skip the per-config records (cfg1, cfg2, cfg4)")
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
    else:
        ours(args)


if __name__ == "__main__":
    main()
