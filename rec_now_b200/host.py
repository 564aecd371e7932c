"""Host-buffer front end of the pairwise loss (``rn_host_pairwise_*`` of include/recnow_b200.h).

For callers whose batch lives in HOST memory (what a data loader or a CPU-placed TF2 graph hands to
``pairwise_loss``, pairwise_loss_from_batch.py:228-279): the library copies the columns to the device, runs
segmentation + the fused loss / gradient kernels and copies ``loss, n_pair, d loss / d logits`` back, pipelined over
``depth`` device slots so that the copies of one batch overlap the kernels of its neighbours.  Arrays are NumPy arrays
or CPU torch tensors; pinned memory keeps the copies asynchronous.  There is no CPU fallback: without a CUDA device
``create`` fails.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


def _ptr(x, dtype, n=None):
    """Host address of a contiguous NumPy array / CPU torch tensor of the given dtype."""
    if x is None:
        return None
    if hasattr(x, "data_ptr"):                       # torch tensor
        if x.device.type != "cpu":
            raise ValueError("HostPairwise takes HOST buffers (use ops.pairwise_fwd_bwd for device tensors)")
        if not x.is_contiguous() or np.dtype(str(x.dtype).replace("torch.", "")) != np.dtype(dtype):
            raise ValueError(f"expected a contiguous {np.dtype(dtype).name} buffer")
        if n is not None and x.numel() != n:
            raise ValueError(f"expected {n} elements, got {x.numel()}")
        return x.data_ptr()
    a = x
    if not isinstance(a, np.ndarray) or a.dtype != np.dtype(dtype) or not a.flags.c_contiguous:
        raise ValueError(f"expected a C-contiguous {np.dtype(dtype).name} ndarray")
    if n is not None and a.size != n:
        raise ValueError(f"expected {n} elements, got {a.size}")
    return a.ctypes.data


class BoundBatch:
    """Host buffers bound to a :class:`HostPairwise` (see :meth:`HostPairwise.bind`)."""

    def __init__(self, owner: "HostPairwise", args, buffers):
        self._owner, self._args, self._buffers = owner, args, buffers
        self._ref = C.byref(args)
        self._ticket = C.c_int32(-1)
        self._tref = C.byref(self._ticket)
        self._fn = _lib.lib().rn_host_pairwise_submit

    def submit(self) -> int:
        rc = self._fn(self._owner._h, self._ref, self._tref)
        if rc:
            _lib.check(rc, "rn_host_pairwise_submit")
        t = self._ticket.value
        self._owner._keep[t] = self
        return t


class HostPairwise:
    """``depth`` in-flight batches of at most ``B_max`` rows with ``K`` key columns."""

    def __init__(self, B_max: int, K: int = 1, depth: int = 2):
        self._h = C.c_void_p()
        _lib.check(_lib.lib().rn_host_pairwise_create(B_max, K, depth, C.byref(self._h)), "rn_host_pairwise_create")
        self.B_max, self.K, self.depth = B_max, K, depth
        self._keep = [None] * depth                  # the buffers of a submit stay referenced until its wait

    def submit(self, keys, logits, labels, *, loss, n_pair_f32, n_pair, dlogits, rw_pos=None, rw_neg=None, row_ok=None,
               row_pairs=None, label_func: str = "step", factor: float = 1.0, power: float = 0.0,
               only_wrong: bool = False, reduce_mean: bool = True, weight_lut=None) -> int:
        """Enqueue one batch; returns the ticket to pass to :meth:`wait`.  ``keys`` is int64 ``[K, B]``; the outputs
        ``loss`` (f32[1]), ``n_pair_f32`` (f32[1]), ``n_pair`` (i64[1]), ``dlogits`` (f32[B]) are filled by wait."""
        return self.bind(keys, logits, labels, loss=loss, n_pair_f32=n_pair_f32, n_pair=n_pair, dlogits=dlogits,
                         rw_pos=rw_pos, rw_neg=rw_neg, row_ok=row_ok, row_pairs=row_pairs, label_func=label_func,
                         factor=factor, power=power, only_wrong=only_wrong, reduce_mean=reduce_mean,
                         weight_lut=weight_lut).submit()

    def bind(self, keys, logits, labels, *, loss, n_pair_f32, n_pair, dlogits, rw_pos=None, rw_neg=None, row_ok=None,
             row_pairs=None, label_func: str = "step", factor: float = 1.0, power: float = 0.0,
             only_wrong: bool = False, reduce_mean: bool = True, weight_lut=None) -> "BoundBatch":
        """Validate a set of host buffers once and return a :class:`BoundBatch` whose ``submit()`` enqueues whatever
        they hold at that moment -- for loaders that refill the same staging buffers every step."""
        B = int(logits.numel() if hasattr(logits, "numel") else logits.size)
        a = _lib.PairwiseArgs()
        a.B, a.K = B, self.K
        a.label_func = {"step": _lib.RN_LABEL_STEP, "diff": _lib.RN_LABEL_DIFF, "gain2": _lib.RN_LABEL_GAIN2,
                        "lut": _lib.RN_LABEL_LUT, "lambda": _lib.RN_LABEL_LAMBDA}[label_func]
        a.weight_lut = _ptr(weight_lut, np.float32, 64)        # (label_func "lut": the 8 x 8 level table, a host buffer too)
        a.keys = _ptr(keys, np.int64, self.K * B)
        a.logits, a.labels = _ptr(logits, np.float32, B), _ptr(labels, np.float32, B)
        a.row_ok, a.rw_pos, a.rw_neg = _ptr(row_ok, np.uint8, B), _ptr(rw_pos, np.float32, B), _ptr(rw_neg, np.float32, B)
        a.factor, a.power = factor, power
        a.only_wrong, a.reduce_mean = int(only_wrong), int(reduce_mean)
        a.part_rank, a.part_count = 0, 1
        a.loss, a.n_pair_f32 = _ptr(loss, np.float32, 1), _ptr(n_pair_f32, np.float32, 1)
        a.n_pair, a.dlogits = _ptr(n_pair, np.int64, 1), _ptr(dlogits, np.float32, B)
        a.row_pairs = _ptr(row_pairs, np.int64, B)
        return BoundBatch(self, a, (keys, logits, labels, rw_pos, rw_neg, row_ok, loss, n_pair_f32, n_pair, dlogits, row_pairs,
                                    weight_lut))

    def wait(self, ticket: int) -> None:
        _lib.check(_lib.lib().rn_host_pairwise_wait(self._h, ticket), "rn_host_pairwise_wait")
        self._keep[ticket] = None

    def graph_steps(self) -> int:
        """Submits that went out as one launch of a slot's whole-step CUDA graph (same pinned buffers as the slot's previous
        submit: see rn_host_pairwise_graph_steps)."""
        return int(_lib.lib().rn_host_pairwise_graph_steps(self._h))

    def close(self) -> None:
        if self._h:
            _lib.lib().rn_host_pairwise_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
