"""ctypes binding of ``include/xfmr_b200.h`` (libxfmr_b200.so).

There is deliberately NO fallback: if the shared library is missing or a call fails the
error is raised to the caller.  The library is built in-tree by
``__graft_entry__.build()`` / ``make -C transformer-recommenders_b200/csrc``.
"""

from __future__ import annotations

import ctypes as C
import pathlib

LIB_PATH = pathlib.Path(__file__).resolve().parent / "lib" / "libxfmr_b200.so"

XR_F32, XR_BF16 = 0, 1
XR_NUM_LOSSES = 7
XR_STATS_SLOTS = 16
TARGET_FIRST, TARGET_DIAGONAL, TARGET_EXPLICIT, TARGET_LAST = 0, 1, 2, 3

LOSS_KIND = {
    "AlignmentLoss": 0,
    "AlignmentContrastiveLoss": 1,
    "ContrastiveLoss": 2,
    "InfoNCELoss": 3,
    "NCELoss": 4,
    "PairwiseHingeLoss": 5,
    "PairwiseLogisticLoss": 6,
}


class XrLossConfig(C.Structure):
    """``xr_loss_config`` — LossConfig of xfmr_rec/losses.py:11-30."""

    _fields_ = [
        ("mask_false_negatives", C.c_int32),
        ("num_hard_negatives", C.c_int32),
        ("scale", C.c_float),
        ("margin", C.c_float),
        ("logits_bf16", C.c_int32),
    ]


_p, _i64, _int, _f, _sz, _u32 = C.c_void_p, C.c_int64, C.c_int, C.c_float, C.c_size_t, C.c_uint32
_cfgp = C.POINTER(XrLossConfig)

# name -> (restype, argtypes); must list every symbol declared in include/xfmr_b200.h
PROTOTYPES = {
    "xr_last_error": (C.c_char_p, []),
    "xr_abi_version": (_int, []),
    "xr_device_info": (_int, [C.POINTER(_int)] * 4),
    "xr_reserve_sms": (_int, [_int]),
    "xr_gather_rows": (_int, [_p, _i64, _i64, _int, _p, _p, _i64, _p, _int, _p, _p]),
    "xr_scatter_rows": (_int, [_p, _i64, _i64, _p, _p, _i64, _p]),
    "xr_row_nonzero": (_int, [_p, _i64, _i64, _int, _p, _p]),
    "xr_compact_workspace_bytes": (_sz, [_i64]),
    "xr_compact_positions": (_int, [_p, _p, _p, _i64, _i64, _p, _p, _p, _p, _p, _p, _p, _p]),
    "xr_scatter_scaled": (_int, [_p, _p, _p, _i64, _i64, _p, _int, _p]),
    "xr_normalize_rows": (_int, [_p, _i64, _i64, _int, _f, _p, _int, _p, _p]),
    "xr_logits_pool": (_int, [_p, _p, _p, _i64, _i64, _i64, _int, _p, _i64, _p]),
    "xr_logits_dense": (_int, [_p, _p, _i64, _i64, _i64, _int, _p, _p, _f, _p, _i64, _p]),
    "xr_logits_sampled": (_int, [_p, _p, _i64, _p, _i64, _i64, _i64, _int, _p, _p, _p, _i64, _p]),
    "xr_rowloss_workspace_bytes": (_sz, [_i64, _i64, _int]),
    "xr_rowloss": (_int, [_p, _i64, _i64, _i64, _int, _p, _cfgp, _u32, _int, _f, _p, _p, _p, _p, _p, _p]),
    "xr_dq_pool": (_int, [_p, _i64, _p, _p, _p, _i64, _i64, _i64, _int, _int, _p, _p, _p]),
    "xr_dq_dense": (_int, [_p, _i64, _p, _p, _i64, _i64, _i64, _int, _int, _p, _p, _p, _p]),
    "xr_dcand_dense": (_int, [_p, _p, _i64, _p, _p, _i64, _i64, _i64, _int, _int, _p, _p, _p, _p]),
    "xr_dq_sampled": (_int, [_p, _i64, _p, _p, _i64, _p, _i64, _i64, _i64, _int, _p, _p, _p, _p]),
    "xr_seq_sample_batch": (_int, [_p, _p, _p, _p, _p, _p, _p, _p, _i64, _i64, _int, _int, C.c_uint64,
                                   C.c_uint64, _p, _p, _p, _p, _p]),
    "xr_sampled_step_workspace_bytes": (_sz, [_i64]),
    "xr_sampled_step": (_int, [_p, _p, _i64, _p, _i64, _i64, _i64, _int, _p, _p, _cfgp, _int, _f, _p, _p,
                               _p, _p, _p]),
    "xr_fused_available": (_int, []),
    "xr_fused_wait_stats": (_int, [_int, C.POINTER(C.c_uint64)]),
    "xr_fused_timeline": (_int, [C.POINTER(C.c_int64)]),
    "xr_fused_profile": (_int, [_int]),
    "xr_fused_profile_read": (_int, [C.POINTER(C.c_float), _int]),
    "xr_fused_pool_workspace_bytes": (_sz, [_i64, _i64, _i64]),
    "xr_pool_step_workspace_bytes": (_sz, [_i64, _i64]),
    "xr_pool_step": (_int, [_p, _p, _p, _i64, _p, _int, _p, _p, _i64, _i64, _int, _cfgp, C.c_float, _p,
                            _int, _p, _p, _p, _p, _sz, _p]),
    "xr_pool_step_ingest": (_int, [_p, _p, _p, _i64, _p, _int, _p, _p, _i64, _i64, _p, _p, _p, _sz, _p]),
    "xr_pool_step_compute": (_int, [_i64, _i64, _int, _cfgp, C.c_float, _p, _int, _p, _p, _sz, _p]),
    "xr_pool_step_monitor_workspace_bytes": (_sz, [_i64, _i64]),
    "xr_pool_step_monitor": (_int, [_i64, _i64, _cfgp, _p, _p, _p, _p, _sz, _p]),
    "xr_fused_pool_loss_mon_workspace_bytes": (_sz, [_i64, _i64, _i64]),
    "xr_fused_pool_loss_mon": (_int, [_p, _p, _p, _i64, _i64, _i64, _int, _cfgp, C.c_float, _p, _p, _p, _p, _p, _p, _sz, _p]),
    "xr_pool_step_compute_mon": (_int, [_i64, _i64, _int, _cfgp, C.c_float, _p, _int, _p, _p, _p, _p, _p, _sz, _p]),
    "xr_fused_pool_loss": (_int, [_p, _p, _p, _i64, _i64, _i64, _int, _cfgp, _p, _f, _p, _p, _p, _p, _sz, _p]),
    "xr_fused_pool_all_workspace_bytes": (_sz, [_i64, _i64, _i64]),
    "xr_fused_pool_all": (_int, [_p, _p, _p, _i64, _i64, _i64, _int, _cfgp, _p, _p, _p, _sz, _p]),
    "xr_topk_workspace_bytes": (_sz, [_i64, _i64, _i64]),
    "xr_topk": (_int, [_p, _i64, _i64, _i64, _i64, _i64, _p, _p, _p, _sz, _p]),
    "xr_topk_merge_workspace_bytes": (_sz, [_i64, _i64]),
    "xr_topk_merge": (_int, [_p, _p, _i64, _i64, _i64, _p, _p, _p, _p]),
    "xr_topk_merge_peers": (_int, [_p, _p, _int, _i64, _i64, _p, _p, _p, _p]),
    "xr_scores": (_int, [_p, _i64, _p, _i64, _i64, _int, _p, _p, _p, _i64, _p]),
    "xr_mask_excluded": (_int, [_p, _i64, _i64, _i64, _i64, _p, _p, _p]),
    "xr_score_groupmax": (_int, [_p, _i64, _p, _i64, _i64, _i64, _p, _i64, _p]),
    "xr_score_filter_layout": (_int, [_i64, _i64, _i64, C.POINTER(_i64), C.POINTER(_i64)]),
    "xr_score_filter": (_int, [_p, _i64, _p, _i64, _i64, _p, _i64, _p, _p, _p, _i64, _i64, _p, _p, _p, _i64, _p]),
    "xr_filter_finalize": (_int, [_i64, _i64, _p, _p, _p, _i64, _i64, _p, _p, _p, _i64, _p, _i64, _i64, _i64, _i64,
                                  _p, _p, _i64, _p, _p, _p, _p]),
    "xr_kth_largest": (_int, [_p, _i64, _i64, _i64, _i64, _p, _p]),
    "xr_mask_excluded_ids": (_int, [_p, _p, _i64, _i64, _i64, _i64, _i64, _p, _p, _p]),
    "xr_groups_to_rows": (_int, [_p, _i64, _i64, _i64, _i64, _p, _p, _p]),
    "xr_score_groupmax_ld": (_i64, [_i64, _i64, _i64]),
    "xr_score_topk_workspace_bytes": (_sz, [_i64, _i64, _i64, _i64]),
    "xr_score_topk": (_int, [_p, _i64, _p, _i64, _i64, _i64, _i64, _p, _p, _i64, _p, _p, _p, _p, _sz, _p]),
    "xr_retrieval_metrics": (_int, [_p, _i64, _i64, _p, _p, _i64, _p, _p, _p]),
    "xr_enc_ln_workspace_bytes": (_sz, [_i64]),
    "xr_enc_embed_ln_fwd": (_int, [_p, _i64, _p, _p, _p, _p, _p, _i64, _i64, _i64, _f, _p, _p, _p, _p, _p, _p, _f, _int, _p]),
    "xr_enc_embed_ln_bwd": (_int, [_p, _i64, _p, _p, _p, _p, _p, _p, _p, _i64, _i64, _i64, _p, _p, _p, _p, _p, _p, _f, _int, _p]),
    "xr_enc_add_ln_fwd": (_int, [_p, _int, _p, _p, _p, _p, _i64, _i64, _f, _p, _p, _p, _p, _f, _int, _p]),
    "xr_enc_add_ln_bwd": (_int, [_p, _int, _p, _p, _p, _p, _p, _p, _i64, _i64, _p, _p, _p, _p, _p, _p, _p, _f, _int, _p]),
    "xr_enc_colsum_workspace_bytes": (_sz, [_i64]),
    "xr_enc_colsum": (_int, [_p, _int, _i64, _i64, _p, _p, _p]),
    "xr_enc_gelu": (_int, [_p, _p, _i64, _int, _p, _p]),
    "xr_enc_attention": (_int, [_p, _p, _p, _p, _p, _i64, _i64, _i64, _i64, _int, _p, _p, _f, _int, _p]),
}

_lib = None


class NativeError(RuntimeError):
    pass


def lib() -> C.CDLL:
    """Load libxfmr_b200.so (once) and attach prototypes.  Raises if it is not built."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise NativeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU / PyTorch fallback for the xfmr_b200 kernels)"
            )
        handle = C.CDLL(str(LIB_PATH))
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(handle, name)  # AttributeError if the .so lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().xr_last_error().decode(errors="replace")
        raise NativeError(f"{what} failed (code {rc}): {msg}")


def call(name: str, *args):
    fn = getattr(lib(), name)
    check(fn(*args), name)
