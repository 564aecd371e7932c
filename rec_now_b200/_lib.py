"""ctypes binding of librecnow_b200.so (the C ABI declared in include/recnow_b200.h).

There is NO CPU fallback: if the shared library is missing this module raises, and every op raises if its
tensors are not on a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# (RN_LIB_PATH: developer override used to A/B compile-time variants of the kernels on the GPU box)
LIB_PATH = os.environ.get("RN_LIB_PATH") or os.path.join(HERE, "librecnow_b200.so")

RN_OK = 0
RN_LABEL_STEP, RN_LABEL_DIFF, RN_LABEL_GAIN2, RN_LABEL_LUT, RN_LABEL_LAMBDA = 0, 1, 2, 3, 4
RN_LOSS_LOGISTIC, RN_LOSS_HINGE = 0, 1
ERR_NAMES = {1: "RN_ERR_ARG", 2: "RN_ERR_ALIGN", 3: "RN_ERR_SCRATCH", 4: "RN_ERR_LAUNCH",
             5: "RN_ERR_UNSUPPORTED", 6: "RN_ERR_NO_DEVICE", 7: "RN_ERR_INTERNAL"}

# every symbol include/recnow_b200.h declares (tests check the .so exports all of them)
EXPORTS = [
    "rn_version", "rn_strerror", "rn_canon_keys_f32", "rn_canon_keys_f64",
    "rn_pairwise_scratch_bytes", "rn_pairwise_scratch_init", "rn_pairwise_fwd_bwd",
    "rn_pair_indices_scratch_bytes", "rn_pair_indices_count", "rn_pair_indices_fill",
    "rn_occurrence_scratch_bytes", "rn_occurrence_power_weight",
    "rn_gauc_scratch_bytes", "rn_gauc",
    "rn_listwise_scratch_bytes", "rn_listwise_fwd_bwd", "rn_listwise_dense",
    "rn_bench_mufu", "rn_profile_enable", "rn_profile_enable_ex", "rn_profile_collect", "rn_profile_disable", "rn_last_device_error", "rn_debug_timestamps", "rn_pairwise_launch_count", "rn_listwise_launch_count",
    "rn_debug_graph_launches", "rn_debug_arena_offset", "rn_pack_row_block", "rn_reduce_peer_chunks",
    "rn_global_buffer_bytes", "rn_global_gather_bytes", "rn_global_pairwise_fwd_bwd",
    "rn_host_pairwise_create", "rn_host_pairwise_submit", "rn_host_pairwise_wait", "rn_host_pairwise_destroy",
    "rn_host_pairwise_graph_steps",
    "rn_segment_pool_fwd", "rn_segment_pool_bwd",
]


class RnError(RuntimeError):
    def __init__(self, code: int, where: str):
        self.code = code
        msg = lib().rn_strerror(code).decode() if _lib is not None else "?"
        super().__init__(f"{where}: {ERR_NAMES.get(code, code)} ({msg})")


class PairwiseArgs(C.Structure):
    _fields_ = [
        ("B", C.c_int64), ("K", C.c_int32), ("label_func", C.c_int32),
        ("keys", C.c_void_p), ("logits", C.c_void_p), ("labels", C.c_void_p),
        ("row_ok", C.c_void_p), ("rw_pos", C.c_void_p), ("rw_neg", C.c_void_p),
        ("factor", C.c_float), ("power", C.c_float),
        ("only_wrong", C.c_int32), ("reduce_mean", C.c_int32),
        ("part_rank", C.c_int32), ("part_count", C.c_int32),
        ("loss", C.c_void_p), ("n_pair_f32", C.c_void_p), ("n_pair", C.c_void_p),
        ("dlogits", C.c_void_p), ("row_pairs", C.c_void_p),
        ("block_rows", C.c_int64), ("block_stride", C.c_int64), ("out_chunk", C.c_int64),
        ("peer_blocks", C.c_void_p * 8), ("gather_dst", C.c_void_p),
        ("scratch_persistent", C.c_int32), ("deterministic", C.c_int32), ("scratch_rows", C.c_int64),
        ("focal_weight", C.c_float), ("focal_alpha", C.c_float), ("focal_gamma", C.c_float),
        ("focal_stop_weight_gradient", C.c_int32),
        ("pair_loss", C.c_int32), ("margin", C.c_float),
        ("weight_lut", C.c_void_p),
    ]


class GlobalArgs(C.Structure):
    _fields_ = [("local", PairwiseArgs), ("world", C.c_int32), ("rank", C.c_int32),
                ("peer_buf", C.c_void_p * 8), ("gather_buf", C.c_void_p), ("step", C.c_int64)]


class GaucArgs(C.Structure):
    _fields_ = [("B", C.c_int64), ("K", C.c_int32), ("scratch_persistent", C.c_int32),
                ("keys", C.c_void_p), ("scores", C.c_void_p), ("labels", C.c_void_p), ("row_ok", C.c_void_p),
                ("gauc", C.c_void_p), ("auc_mean", C.c_void_p), ("n_valid_groups", C.c_void_p),
                ("n_pair", C.c_void_p), ("concordant2", C.c_void_p)]


class PoolArgs(C.Structure):
    _fields_ = [("B", C.c_int64), ("C", C.c_int64), ("T", C.c_int32), ("D", C.c_int32), ("mean", C.c_int32),
                ("reserved0", C.c_int32), ("slots", C.c_void_p), ("ids", C.c_void_p), ("weights", C.c_void_p),
                ("target_slots", C.c_void_p), ("table", C.c_void_p), ("V", C.c_int64)]


class ListwiseArgs(C.Structure):
    _fields_ = [
        ("B", C.c_int64), ("keys", C.c_void_p), ("row_ok", C.c_void_p),
        ("labels", C.c_void_p), ("logits", C.c_void_p), ("list_w", C.c_void_p),
        ("pos_neg_th", C.c_float), ("do_reduce", C.c_int32),
        ("loss", C.c_void_p), ("list_loss", C.c_void_p), ("n_valid", C.c_void_p),
        ("n_group", C.c_void_p), ("dlogits", C.c_void_p),
        ("scratch_persistent", C.c_int32), ("inv_temperature", C.c_float),
    ]


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing. Build it with `python -m rec_now_b200.build` (needs nvcc); "
            "rec_now_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, i64, i32, f32, sz = C.c_void_p, C.c_int64, C.c_int32, C.c_float, C.c_size_t
    L.rn_version.restype = C.c_int
    L.rn_strerror.restype = C.c_char_p
    L.rn_strerror.argtypes = [C.c_int]
    for name in ("rn_canon_keys_f32", "rn_canon_keys_f64"):
        getattr(L, name).argtypes = [vp, i64, vp, vp, C.c_int, vp]
    L.rn_pairwise_scratch_bytes.restype = sz
    L.rn_pairwise_scratch_bytes.argtypes = [i64, i32]
    L.rn_pairwise_scratch_init.argtypes = [vp, sz, vp]
    L.rn_pairwise_fwd_bwd.argtypes = [C.POINTER(PairwiseArgs), vp, sz, vp]
    L.rn_pairwise_fwd_bwd.restype = C.c_int
    L.rn_pair_indices_scratch_bytes.restype = sz
    L.rn_pair_indices_scratch_bytes.argtypes = [i64, i32]
    L.rn_pair_indices_count.argtypes = [C.POINTER(PairwiseArgs), i32, vp, sz, C.POINTER(i64), vp]
    L.rn_pair_indices_fill.argtypes = [C.POINTER(PairwiseArgs), i32, vp, sz, vp, vp, vp, i64, vp]
    L.rn_occurrence_scratch_bytes.restype = sz
    L.rn_occurrence_scratch_bytes.argtypes = [i64]
    L.rn_occurrence_power_weight.argtypes = [vp, i64, f32, vp, vp, sz, vp]
    L.rn_gauc_scratch_bytes.restype = sz
    L.rn_gauc_scratch_bytes.argtypes = [i64, i32]
    L.rn_gauc.argtypes = [C.POINTER(GaucArgs), vp, sz, vp]
    L.rn_listwise_scratch_bytes.restype = sz
    L.rn_listwise_scratch_bytes.argtypes = [i64]
    L.rn_listwise_fwd_bwd.argtypes = [C.POINTER(ListwiseArgs), vp, sz, vp]
    L.rn_listwise_dense.argtypes = [C.POINTER(ListwiseArgs), vp, sz, i64, vp, vp, vp, i32, f32, vp]
    L.rn_bench_mufu.argtypes = [i32, vp, C.POINTER(i64), vp]
    L.rn_profile_enable.argtypes = [i32]
    L.rn_profile_enable_ex.argtypes = [i32, i32]
    L.rn_profile_collect.argtypes = [C.POINTER(f32), i32, C.POINTER(i32)]
    L.rn_last_device_error.argtypes = [vp, C.POINTER(i32), vp]
    L.rn_debug_timestamps.argtypes = [vp, C.POINTER(C.c_uint64), i32, vp]
    L.rn_pairwise_launch_count.argtypes = [i64, i32]
    L.rn_listwise_launch_count.argtypes = [i64]
    L.rn_pack_row_block.argtypes = [vp, i32, vp, vp, vp, vp, i64, vp, i64, vp]
    L.rn_reduce_peer_chunks.argtypes = [C.POINTER(C.c_void_p), i32, i32, i64, vp, vp]
    L.rn_global_buffer_bytes.restype = sz
    L.rn_global_buffer_bytes.argtypes = [i64, i32, i32, i32, i32]
    L.rn_global_gather_bytes.restype = sz
    L.rn_global_gather_bytes.argtypes = [i64, i32, i32, i32, i32]
    L.rn_global_pairwise_fwd_bwd.argtypes = [C.POINTER(GlobalArgs), vp, sz, vp]
    L.rn_host_pairwise_create.argtypes = [i64, i32, i32, C.POINTER(vp)]
    L.rn_host_pairwise_submit.argtypes = [vp, C.POINTER(PairwiseArgs), C.POINTER(i32)]
    L.rn_host_pairwise_wait.argtypes = [vp, i32]
    L.rn_host_pairwise_destroy.argtypes = [vp]
    if hasattr(L, "rn_host_pairwise_graph_steps"):
        L.rn_host_pairwise_graph_steps.restype = i64
        L.rn_host_pairwise_graph_steps.argtypes = [vp]
    if hasattr(L, "rn_segment_pool_fwd"):       # (absent only from older builds loaded through RN_LIB_PATH for A/B timing)
        L.rn_segment_pool_fwd.argtypes = [C.POINTER(PoolArgs), vp, vp, vp]
        L.rn_segment_pool_bwd.argtypes = [C.POINTER(PoolArgs), vp, vp, vp, vp]
    L.rn_debug_arena_offset.restype = i64
    L.rn_debug_arena_offset.argtypes = [i64, i32, i32]
    L.rn_debug_graph_launches.restype = i64
    L.rn_debug_graph_launches.argtypes = []
    _lib = L
    return L


def check(code: int, where: str) -> None:
    if code != RN_OK:
        raise RnError(code, where)
