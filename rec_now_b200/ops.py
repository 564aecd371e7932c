"""Thin torch-tensor wrappers over the C ABI (torch is only the allocator / stream provider here).

Every function requires CUDA tensors and enqueues on torch's current stream; nothing synchronises.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import ListwiseArgs, PairwiseArgs, check, lib


def _stream(dev=None) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


class _on_device:
    """`with torch.cuda.device(dev)` only when dev is not already current (the context manager costs ~10 us)."""
    __slots__ = ("ctx",)

    def __init__(self, dev):
        self.ctx = None if dev.index is None or dev.index == torch.cuda.current_device() else torch.cuda.device(dev)

    def __enter__(self):
        if self.ctx is not None:
            self.ctx.__enter__()

    def __exit__(self, *a):
        if self.ctx is not None:
            self.ctx.__exit__(*a)


# Scratch arenas are reused per (device, stream, size): calls on one stream are ordered, so the next call may
# overwrite the arena of the previous one (its outputs live in their own tensors).  Set to False to get a fresh
# arena per call (needed when the control block of an earlier call is inspected after a later one was enqueued).
REUSE_SCRATCH = True
_scratch_cache: dict = {}


def _scratch(nbytes: int, dev: torch.device, stream_ptr: int, tag=None) -> torch.Tensor:
    """The arena of (device, stream, size[, tag]).  Arenas are created ZEROED and then written by nothing but the
    library's calls of one kind (tag), which is what rn_pairwise_args.scratch_persistent promises."""
    if not REUSE_SCRATCH:
        return torch.zeros(nbytes, dtype=torch.uint8, device=dev)
    key = (dev.index, stream_ptr, nbytes, tag)
    t = _scratch_cache.get(key)
    if t is None:
        if len(_scratch_cache) > 64:
            _scratch_cache.clear()
        t = _scratch_cache[key] = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
    return t


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("rec_now_b200 ops need CUDA tensors: there is no CPU fallback")


def _f32(t: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    if t is None:
        return None
    if t.dtype is torch.float32 and t.is_contiguous():
        if t.data_ptr() & 15:          # (a legal view with an offset, e.g. logits[1:]: the ABI wants 16-byte aligned columns)
            return t.detach().clone()
        return t                       # only the data pointer and numel() are used: no detach / view (each costs ~1.5 us)
    return t.detach().reshape(-1).to(torch.float32).contiguous()


def canon_keys(groups, row_ok: Optional[torch.Tensor] = None, inf_is_id: bool = False):
    """Group id tensor(s) -> (int64 keys [K,B], row_ok uint8[B] or None).  inf_is_id: +-inf ids equal themselves (the
    listwise path's tf.unique semantics) instead of matching nothing (the pairwise path's g_i - g_j == 0).

    Float ids (what the reference passes, pairwise_loss_from_batch.py:33-35) are canonicalised on the device:
    value equality, -0.0 == +0.0, NaN/inf match nothing.  Integer ids are used as they are.
    """
    cols = list(groups) if isinstance(groups, (list, tuple)) else [groups]
    _need_cuda(*cols)
    b = cols[0].numel()
    dev = cols[0].device
    if len(cols) == 1 and cols[0].dtype is torch.int64 and cols[0].is_contiguous():
        ok = None if row_ok is None else row_ok.detach().reshape(-1).to(torch.uint8).contiguous()
        return cols[0].detach().view(1, -1), ok                  # integer ids are canonical as they are
    keys = torch.empty((len(cols), b), dtype=torch.int64, device=dev)
    ok = None if row_ok is None else row_ok.detach().reshape(-1).to(torch.uint8).contiguous().clone()
    for k, g in enumerate(cols):
        g = g.detach().reshape(-1)
        if g.numel() != b:
            raise ValueError("all group key tensors must hold the same number of elements")
        if g.dtype in (torch.float32, torch.float64, torch.float16, torch.bfloat16):
            if g.dtype in (torch.float16, torch.bfloat16):
                g = g.to(torch.float32)
            g = g.contiguous()
            and_into = 1
            if ok is None:
                ok = torch.empty(b, dtype=torch.uint8, device=dev)
                and_into = 0
            fn = lib().rn_canon_keys_f32 if g.dtype == torch.float32 else lib().rn_canon_keys_f64
            with _on_device(dev):
                check(fn(g.data_ptr(), b, keys[k].data_ptr(), ok.data_ptr(), and_into | (2 if inf_is_id else 0),
                         _stream(dev)), "rn_canon_keys")
        else:
            keys[k].copy_(g.to(torch.int64))
    return keys, ok


class _PairCall:
    """Per-thread reusable argument struct of rn_pairwise_fwd_bwd (building a 25-field ctypes struct per call costs
    more host time than the three kernel launches)."""
    __slots__ = ("args", "ref", "fn")

    def __init__(self):
        self.args = PairwiseArgs()
        self.ref = C.byref(self.args)
        self.fn = lib().rn_pairwise_fwd_bwd


import threading

_LABEL_FUNCS = {"step": _lib.RN_LABEL_STEP, "diff": _lib.RN_LABEL_DIFF, "gain2": _lib.RN_LABEL_GAIN2, "lut": _lib.RN_LABEL_LUT,
                "lambda": _lib.RN_LABEL_LAMBDA}
_tls = threading.local()             # (ctypes releases the GIL during the call: the argument struct is per thread)
_pair_scratch_bytes: dict = {}


def pairwise_fwd_bwd(logits, labels, keys, row_ok=None, rw_pos=None, rw_neg=None, label_func="step",
                     factor=1.0, power=0.0, only_wrong=False, reduce_mean=True, part=(0, 1),
                     want_row_pairs=False, deterministic=False, focal=None, pair_loss="logistic", margin=1.0,
                     weight_lut=None):
    """rn_pairwise_fwd_bwd.  keys: int64 [K,B] (canonical).  Returns dict of device tensors.
    label_func: "step" | "diff" | "gain2" | "lut" (weight_lut: float32 [8, 8] on the device, W by label level = label + 1 of
    integer labels -1 .. 6) | "lambda" (LambdaRank |delta NDCG| weights from the rows' score ranks); pair_loss: "logistic" (bpr_loss_func) | "hinge" (max(0, margin - x))."""
    _need_cuda(logits, labels, keys, row_ok, rw_pos, rw_neg)
    s, y = _f32(logits), _f32(labels)
    b = s.numel()
    if keys.dim() != 2 or not keys.is_contiguous():
        keys = keys.reshape(-1, b).contiguous()
    kk = keys.shape[0]
    dev = s.device
    rwp, rwn = _f32(rw_pos), _f32(rw_neg)
    ok = None if row_ok is None else row_ok.reshape(-1).to(torch.uint8).contiguous()
    for name, t in (("labels", y), ("rw_pos", rwp), ("rw_neg", rwn), ("row_ok", ok)):
        if t is not None and t.numel() != b:        # (a shorter column would be read out of bounds on the device)
            raise ValueError(f"{name} holds {t.numel()} elements, logits {b}")
    out = torch.empty(4, dtype=torch.float32, device=dev)          # loss, n_pair_f32, n_pair (int64 in [2:4])
    dlogits = torch.empty(b, dtype=torch.float32, device=dev)
    row_pairs = torch.empty(b, dtype=torch.int64, device=dev) if want_row_pairs else None
    nbytes = _pair_scratch_bytes.get((b, kk))
    if nbytes is None:
        nbytes = _pair_scratch_bytes[(b, kk)] = lib().rn_pairwise_scratch_bytes(b, kk)
    st = torch.cuda.current_stream(dev).cuda_stream
    scratch = _scratch(nbytes, dev, st, ("pair", kk))
    po = out.data_ptr()
    pc = getattr(_tls, "pair_call", None)
    if pc is None:
        pc = _tls.pair_call = _PairCall()
    a = pc.args
    a.B = b; a.K = kk; a.label_func = _LABEL_FUNCS[label_func]
    a.pair_loss = _lib.RN_LOSS_HINGE if pair_loss == "hinge" else _lib.RN_LOSS_LOGISTIC
    a.margin = margin
    if label_func == "lut":
        if weight_lut is None or weight_lut.numel() != 64 or weight_lut.device != dev:
            raise ValueError("label_func 'lut' needs weight_lut: 64 float32 on the logits' device")
        weight_lut = _f32(weight_lut)
        a.weight_lut = weight_lut.data_ptr()
    else:
        a.weight_lut = None
    a.keys = keys.data_ptr(); a.logits = s.data_ptr(); a.labels = y.data_ptr()
    a.row_ok = _ptr(ok); a.rw_pos = _ptr(rwp); a.rw_neg = _ptr(rwn)
    a.factor = factor; a.power = power; a.only_wrong = 1 if only_wrong else 0; a.reduce_mean = 1 if reduce_mean else 0
    a.part_rank, a.part_count = part
    a.loss = po; a.n_pair_f32 = po + 4; a.n_pair = po + 8
    a.dlogits = dlogits.data_ptr(); a.row_pairs = _ptr(row_pairs)
    a.block_rows = 0; a.block_stride = 0; a.out_chunk = 0
    a.scratch_persistent = 1; a.scratch_rows = 0; a.deterministic = 1 if deterministic else 0
    if focal is None:
        a.focal_weight = 0.0
    else:       # (weight, alpha or 0, gamma or 0, stop_weight_gradient): the fused focal term, see rn_pairwise_args
        a.focal_weight, a.focal_alpha, a.focal_gamma = focal[0], focal[1], focal[2]
        a.focal_stop_weight_gradient = 1 if focal[3] else 0
    with _on_device(dev):
        rc = pc.fn(pc.ref, scratch.data_ptr(), nbytes, st)
    if rc:
        check(rc, "rn_pairwise_fwd_bwd")
    return _PairOut(out, dlogits, row_pairs, scratch)


class _PairOut(dict):
    """Result of pairwise_fwd_bwd: a dict whose scalar views (loss, n_pair_f32, n_pair) are made on first use (each
    tensor slice costs a few microseconds of host time)."""

    def __init__(self, out, dlogits, row_pairs, scratch):
        super().__init__(dlogits=dlogits, row_pairs=row_pairs, _scratch=scratch, _out=out)

    def __missing__(self, key):
        out = dict.__getitem__(self, "_out")
        if key == "loss":
            v = out[0]
        elif key == "n_pair_f32":
            v = out[1]
        elif key == "n_pair":
            v = out[2:4].view(torch.int64)[0]
        else:
            raise KeyError(key)
        self[key] = v
        return v

    def get(self, key, default=None):
        try:
            return self[key]
        except KeyError:
            return default


def packed_block_layout(b_loc: int, kk: int, has_w: bool, has_ok: bool) -> dict:
    """Byte layout of one rank's packed row block (global mode): [keys K x int64][logits f32][labels f32][w f32]
    [row_ok u8], every column 16-byte aligned, block size a multiple of 16."""
    assert b_loc % 16 == 0
    off, o = {}, 0
    off["keys"] = o; o += 8 * kk * b_loc
    off["logits"] = o; o += 4 * b_loc
    off["labels"] = o; o += 4 * b_loc
    if has_w:
        off["w"] = o; o += 4 * b_loc
    if has_ok:
        off["ok"] = o; o += b_loc
    off["stride"] = (o + 15) // 16 * 16
    return off


def pack_row_block(keys, logits, labels, rw_pos, row_ok, stride: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """rn_pack_row_block: this rank's columns as one packed uint8 block (packed_block_layout), one launch.  `out`: a
    uint8 buffer of `stride` bytes to pack into (e.g. a peer-mapped symmetric buffer)."""
    _need_cuda(keys, logits, labels, rw_pos, row_ok)
    s, y, rwp = _f32(logits), _f32(labels), _f32(rw_pos)
    b_loc = s.numel()
    keys = keys.reshape(-1, b_loc)
    if keys.dtype is not torch.int64 or not keys.is_contiguous():
        keys = keys.to(torch.int64).contiguous()
    ok = None if row_ok is None else row_ok.reshape(-1).to(torch.uint8).contiguous()
    block = torch.empty(stride, dtype=torch.uint8, device=s.device) if out is None else out
    assert block.numel() == stride and block.dtype is torch.uint8
    with _on_device(s.device):
        check(lib().rn_pack_row_block(keys.data_ptr(), keys.shape[0], s.data_ptr(), y.data_ptr(), _ptr(rwp), _ptr(ok),
                                      b_loc, block.data_ptr(), stride, _stream(s.device)), "rn_pack_row_block")
    return block


def reduce_peer_chunks(peer_out_ptrs: Sequence[int], rank: int, chunk: int, dev: torch.device) -> torch.Tensor:
    """rn_reduce_peer_chunks: this rank's chunk of the chunked gradient buffers summed over all ranks, read from the
    peers' mapped memory (the reduce-scatter of the global mode without a collective call)."""
    world = len(peer_out_ptrs)
    mine = torch.empty(chunk, dtype=torch.float32, device=dev)
    arr = (C.c_void_p * world)(*peer_out_ptrs)
    with _on_device(dev):
        check(lib().rn_reduce_peer_chunks(arr, world, rank, chunk, mine.data_ptr(), _stream(dev)), "rn_reduce_peer_chunks")
    return mine


def pairwise_fwd_bwd_blocked(gbuf: torch.Tensor, world: int, b_loc: int, kk: int, has_w: bool, has_ok: bool,
                             label_func="step", factor=1.0, power=0.0, reduce_mean=True, part=(0, 1),
                             peer_blocks: Optional[Sequence[int]] = None, out: Optional[torch.Tensor] = None,
                             persistent: bool = False):
    """rn_pairwise_fwd_bwd on `world` packed row blocks (packed_block_layout) as ONE all-gather leaves them (or, with
    `peer_blocks` = the device pointers of every rank's block in peer-mapped memory, as the call's first kernel
    gathers them into `gbuf` itself), with the
    outputs laid out for ONE reduce-scatter: returns dict(out=float32[world * chunk] with chunk = b_loc + 4:
    d loss / d logits of global row r * b_loc + i at out[r * chunk + i], the partial loss at out[r * chunk + b_loc];
    n_pair (global, exact), chunk)."""
    _need_cuda(gbuf)
    lay = packed_block_layout(b_loc, kk, has_w, has_ok)
    assert gbuf.dtype is torch.uint8 and gbuf.numel() == world * lay["stride"] and gbuf.is_contiguous()
    dev = gbuf.device
    b = world * b_loc
    chunk = b_loc + 4
    scal = torch.empty(4, dtype=torch.float32, device=dev)
    if out is None:
        out = torch.empty(world * chunk, dtype=torch.float32, device=dev)
    assert out.numel() == world * chunk and out.dtype is torch.float32
    nbytes = lib().rn_pairwise_scratch_bytes(b, kk)
    st = torch.cuda.current_stream(dev).cuda_stream
    scratch = _scratch(nbytes, dev, st, ("blocked", kk))
    base, po = gbuf.data_ptr(), scal.data_ptr()
    a = PairwiseArgs(
        B=b, K=kk, label_func=_lib.RN_LABEL_DIFF if label_func == "diff" else _lib.RN_LABEL_STEP,
        keys=base + lay["keys"], logits=base + lay["logits"], labels=base + lay["labels"],
        row_ok=(base + lay["ok"]) if has_ok else None, rw_pos=(base + lay["w"]) if has_w else None, rw_neg=None,
        factor=float(factor), power=float(power), only_wrong=0, reduce_mean=int(bool(reduce_mean)),
        part_rank=int(part[0]), part_count=int(part[1]),
        loss=po, n_pair_f32=po + 4, n_pair=po + 8, dlogits=out.data_ptr(), row_pairs=None,
        block_rows=b_loc, block_stride=lay["stride"], out_chunk=chunk, scratch_persistent=1 if persistent else 0)
    if peer_blocks is not None:          # k_init gathers the ranks' blocks into gbuf over NVLink (no all-gather call)
        assert len(peer_blocks) == world
        for r in range(world):
            a.peer_blocks[r] = peer_blocks[r]
        a.gather_dst = base
    with _on_device(dev):
        check(lib().rn_pairwise_fwd_bwd(C.byref(a), scratch.data_ptr(), nbytes, C.c_void_p(st)), "rn_pairwise_fwd_bwd")
    return dict(out=out, n_pair=scal[2:4].view(torch.int64)[0], chunk=chunk, _scratch=scratch)


def device_error(scratch: torch.Tensor) -> int:
    err = C.c_int32(0)
    check(lib().rn_last_device_error(scratch.data_ptr(), C.byref(err), _stream()), "rn_last_device_error")
    return err.value


def last_segmentation_path(scratch: torch.Tensor) -> int:
    """Which segmentation the last finished pairwise call on `scratch` ran: 1 = counting (sort-free), 2 = radix sort, 3 = the one-CTA kernel of small batches (no segmentation pass)."""
    ts = (C.c_uint64 * 36)()
    check(lib().rn_debug_timestamps(scratch.data_ptr(), ts, 36, _stream()), "rn_debug_timestamps")
    return int(ts[35])


def _pair_args(s, y, keys, ok, rwp, rwn, label_func, only_wrong, factor=1.0, power=0.0, reduce_mean=True):
    b = s.numel()
    return PairwiseArgs(
        B=b, K=keys.shape[0], label_func=_lib.RN_LABEL_DIFF if label_func == "diff" else _lib.RN_LABEL_STEP,
        keys=keys.data_ptr(), logits=s.data_ptr(), labels=y.data_ptr(),
        row_ok=_ptr(ok), rw_pos=_ptr(rwp), rw_neg=_ptr(rwn), factor=float(factor), power=float(power),
        only_wrong=int(bool(only_wrong)), reduce_mean=int(bool(reduce_mean)), part_rank=0, part_count=1,
        loss=None, n_pair_f32=None, n_pair=None, dlogits=None, row_pairs=None)


def pair_indices(logits, labels, keys, row_ok=None, rw_pos=None, rw_neg=None, label_func="step",
                 only_wrong=False, label_cond=True, want_weights=False):
    """rn_pair_indices_count + _fill: (pos_idx int32[P], neg_idx int32[P], w f32[P] or None) in the reference's
    row-major pair order (PW:217).  Synchronises once (P is data dependent).  label_cond=False lists every
    same-group ordered pair i != j."""
    _need_cuda(logits, labels, keys, row_ok, rw_pos, rw_neg)
    s, y = _f32(logits), _f32(labels)
    b = s.numel()
    keys = keys.reshape(-1, b).contiguous()
    dev = s.device
    rwp, rwn = _f32(rw_pos), _f32(rw_neg)
    ok = None if row_ok is None else row_ok.reshape(-1).to(torch.uint8).contiguous()
    nbytes = lib().rn_pair_indices_scratch_bytes(b, keys.shape[0])
    scratch = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    a = _pair_args(s, y, keys, ok, rwp, rwn, label_func, only_wrong)
    n = C.c_int64(0)
    with torch.cuda.device(dev):
        check(lib().rn_pair_indices_count(C.byref(a), int(bool(label_cond)), scratch.data_ptr(), nbytes,
                                          C.byref(n), _stream()), "rn_pair_indices_count")
        p = n.value
        pos = torch.empty(p, dtype=torch.int32, device=dev)
        neg = torch.empty(p, dtype=torch.int32, device=dev)
        w = torch.empty(p, dtype=torch.float32, device=dev) if want_weights else None
        check(lib().rn_pair_indices_fill(C.byref(a), int(bool(label_cond)), scratch.data_ptr(), nbytes,
                                         pos.data_ptr(), neg.data_ptr(), _ptr(w), p, _stream()),
              "rn_pair_indices_fill")
    return pos, neg, w


def occurrence_power_weight(ids: torch.Tensor, power: float) -> torch.Tensor:
    """rn_occurrence_power_weight on canonical int64 ids."""
    _need_cuda(ids)
    ids = ids.reshape(-1).to(torch.int64).contiguous()
    n = ids.numel()
    out = torch.empty(n, dtype=torch.float32, device=ids.device)
    if n == 0:
        return out
    nbytes = lib().rn_occurrence_scratch_bytes(n)
    scratch = torch.empty(nbytes, dtype=torch.uint8, device=ids.device)
    with torch.cuda.device(ids.device):
        check(lib().rn_occurrence_power_weight(ids.data_ptr(), n, float(power), out.data_ptr(),
                                               scratch.data_ptr(), nbytes, _stream()), "rn_occurrence_power_weight")
    return out


class _LwCall:
    """Per-thread reusable argument struct of rn_listwise_fwd_bwd (see _PairCall)."""
    __slots__ = ("args", "ref", "fn")

    def __init__(self):
        self.args = ListwiseArgs()
        self.ref = C.byref(self.args)
        self.fn = lib().rn_listwise_fwd_bwd


class _LwOut(dict):
    """Result of listwise_fwd_bwd: scalar views (loss, n_valid, n_group) are made on first use."""

    def __missing__(self, key):
        out = dict.__getitem__(self, "_out")
        if key == "loss":
            v = out[0]
        elif key == "n_valid":
            v = out[1:2].view(torch.int32)[0]
        elif key == "n_group":
            v = out[2:3].view(torch.int32)[0]
        else:
            raise KeyError(key)
        self[key] = v
        return v

    def get(self, key, default=None):
        try:
            return self[key]
        except KeyError:
            return default


_lw_scratch_bytes: dict = {}


def listwise_fwd_bwd(keys, labels, logits, row_ok=None, list_w=None, pos_neg_th=0.5, do_reduce=True,
                     want_list_loss=False, sorted_form=False, temperature=1.0):
    """rn_listwise_fwd_bwd.  keys: canonical int64 [B].  Returns dict of device tensors (+ the scratch arena
    and args needed by listwise_dense).  sorted_form=True asks for the sorted (radix) form of the call, whose arena
    listwise_dense can read; otherwise a reduced loss without per-list weights / outputs runs as ONE sort-free kernel
    on a persistent arena."""
    _need_cuda(keys, labels, logits, row_ok, list_w)
    s, y = _f32(logits), _f32(labels)
    b = s.numel()
    if keys.dtype is not torch.int64 or not keys.is_contiguous() or keys.dim() != 1:
        keys = keys.reshape(-1).to(torch.int64).contiguous()
    dev = s.device
    ok = None if row_ok is None else row_ok.reshape(-1).to(torch.uint8).contiguous()
    lw = _f32(list_w)
    for name, t in (("labels", y), ("keys", keys), ("row_ok", ok)):
        if t is not None and t.numel() != b:        # (a shorter column would be read out of bounds on the device)
            raise ValueError(f"{name} holds {t.numel()} elements, logits {b}")
    out = torch.empty(4, dtype=torch.float32, device=dev)           # loss f32, n_valid i32, n_group i32
    dlogits = torch.empty(b, dtype=torch.float32, device=dev)
    list_loss = torch.empty(b, dtype=torch.float32, device=dev) if (want_list_loss or not do_reduce) else None
    nbytes = _lw_scratch_bytes.get(b)
    if nbytes is None:
        nbytes = _lw_scratch_bytes[b] = lib().rn_listwise_scratch_bytes(b)
    counting = do_reduce and lw is None and list_loss is None and not sorted_form
    st = torch.cuda.current_stream(dev).cuda_stream
    scratch = _scratch(nbytes, dev, st, "lw") if counting else torch.empty(nbytes, dtype=torch.uint8, device=dev)
    pc = getattr(_tls, "lw_call", None)
    if pc is None:
        pc = _tls.lw_call = _LwCall()
    a = pc.args
    po = out.data_ptr()
    a.B = b; a.keys = keys.data_ptr(); a.row_ok = _ptr(ok); a.labels = y.data_ptr(); a.logits = s.data_ptr()
    a.list_w = _ptr(lw); a.pos_neg_th = pos_neg_th; a.do_reduce = 1 if do_reduce else 0
    a.inv_temperature = 1.0 / float(temperature)
    a.loss = po; a.n_valid = po + 4; a.n_group = po + 8
    a.list_loss = _ptr(list_loss); a.dlogits = dlogits.data_ptr(); a.scratch_persistent = 1 if counting else 0
    with _on_device(dev):
        rc = pc.fn(pc.ref, scratch.data_ptr(), nbytes, st)
    if rc:
        check(rc, "rn_listwise_fwd_bwd")
    res = _LwOut(dlogits=dlogits, list_loss=list_loss, _scratch=scratch, _out=out, _keep=(keys, ok, lw, y, s),
                 _sorted=not counting)
    if not counting:
        if not do_reduce:
            out[0:1].zero_()                                         # (the kernel only writes a reduced loss)
        # listwise_dense re-reads the argument struct: give the sorted form its own copy
        res["_args"] = ListwiseArgs.from_buffer_copy(a)
    return res


def listwise_dense(fwd: dict, n_valid: int, do_mask_logits=True, value_of_masked_logit=-1e9):
    """rn_listwise_dense: the (V,B) dense_mask / dense_labels / dense_logits of to_listwise_sample (LW:142-145)."""
    if not fwd.get("_sorted", True):
        raise RuntimeError("listwise_dense needs the sorted form of the call (listwise_fwd_bwd(..., sorted_form=True))")
    a = fwd["_args"]
    b = int(a.B)
    dev = fwd["dlogits"].device
    dm = torch.empty((n_valid, b), dtype=torch.bool, device=dev)
    dl = torch.empty((n_valid, b), dtype=torch.float32, device=dev)
    dz = torch.empty((n_valid, b), dtype=torch.float32, device=dev)
    scratch = fwd["_scratch"]
    with torch.cuda.device(dev):
        check(lib().rn_listwise_dense(C.byref(a), scratch.data_ptr(), scratch.numel(), n_valid, dm.data_ptr(),
                                      dl.data_ptr(), dz.data_ptr(), int(bool(do_mask_logits)),
                                      float(value_of_masked_logit), _stream()), "rn_listwise_dense")
    return dm, dl, dz
