"""ctypes binding of the C ABI in include/pgbart_b200.h (libpgbart_b200.so).

The library is built in-tree by ``__graft_entry__.build()``.  There is NO CPU
fallback: if the shared object is missing or the CUDA device is absent, the
product path raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpgbart_b200.so")

BK_ABI_VERSION = 3
BK_MAX_NODES = 255
BK_LIK_NORMAL = 0
BK_LIK_BERNOULLI_LOGIT = 1
BK_LIK_NORMAL_HETERO = 2
BK_LIK_CATEGORICAL = 3
BK_MAX_OUTPUTS = 7
BK_RULE_CONTINUOUS = 0
BK_RULE_ONEHOT = 1
BK_RULE_SUBSET = 2
BK_SUBSET_MAX_CATS = 24     # category codes 0..23 (bk_spec.h)
BK_MAX_SUBSET_COLS = 8


class BkSettings(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32),
        ("n_rows", C.c_int32),
        ("n_cols", C.c_int32),
        ("n_trees", C.c_int32),
        ("n_particles", C.c_int32),
        ("n_chains", C.c_int32),
        ("likelihood", C.c_int32),
        ("qshift", C.c_int32),
        ("batch_tune", C.c_int32),
        ("batch_post", C.c_int32),
        ("seed", C.c_uint32),
        ("chain_base", C.c_uint32),
        ("init_sum", C.c_float),
        ("init_leaf", C.c_float),
        ("leaf_sd_init", C.c_float),
        ("device", C.c_int32),
        ("trace_capacity", C.c_int32),
        ("n_groups", C.c_int32),
        ("n_outputs", C.c_int32),
        ("reserved0", C.c_int32),
        ("p_leaf", C.POINTER(C.c_double)),
        ("split_prior", C.POINTER(C.c_double)),
        ("split_rules", C.POINTER(C.c_int32)),
    ]


class BkStepStats(C.Structure):
    _fields_ = [
        ("tree_updates", C.c_int32),
        ("rounds", C.c_int32),
        ("grow_events", C.c_int32),
        ("grow_root", C.c_int32),
        ("count_passes", C.c_int32),
        ("phases", C.c_int32),
        ("trace_len", C.c_int32),
        ("error_flags", C.c_int32),
        ("leaf_sd", C.c_float),
        ("iter", C.c_int32),
        ("us_control", C.c_int32),
        ("us_data", C.c_int32),
        ("us_sync", C.c_int32),
        ("us_total", C.c_int32),
        ("reserved", C.c_int32 * 2),
    ]


class BkTraceRec(C.Structure):
    _fields_ = [
        ("kind", C.c_int32),
        ("tree", C.c_int32),
        ("round", C.c_int32),
        ("particle", C.c_int32),
        ("node", C.c_int32),
        ("var", C.c_int32),
        ("n_left", C.c_int32),
        ("n_right", C.c_int32),
        ("split", C.c_float),
        ("val_left", C.c_float),
        ("val_right", C.c_float),
        ("ancestor", C.c_int32),
        ("log_w", C.c_double),
        ("aux", C.c_double),
    ]


class BkNode(C.Structure):
    _fields_ = [
        ("var", C.c_int32),
        ("split", C.c_float),
        ("left", C.c_int32),
        ("value", C.c_float),
        ("n", C.c_int32),
        ("depth", C.c_int32),
    ]


import numpy as np  # noqa: E402

TRACE_DTYPE = np.dtype(
    [
        ("kind", "<i4"), ("tree", "<i4"), ("round", "<i4"), ("particle", "<i4"),
        ("node", "<i4"), ("var", "<i4"), ("n_left", "<i4"), ("n_right", "<i4"),
        ("split", "<f4"), ("val_left", "<f4"), ("val_right", "<f4"), ("ancestor", "<i4"),
        ("log_w", "<f8"), ("aux", "<f8"),
    ]
)
NODE_DTYPE = np.dtype(
    [("var", "<i4"), ("split", "<f4"), ("left", "<i4"), ("value", "<f4"), ("n", "<i4"), ("depth", "<i4")]
)
assert TRACE_DTYPE.itemsize == C.sizeof(BkTraceRec) == 64
assert NODE_DTYPE.itemsize == C.sizeof(BkNode) == 24

# every symbol include/pgbart_b200.h declares
EXPORTS = (
    "bk_abi_version", "bk_last_error", "bk_padded_rows", "bk_query_bytes", "bk_create", "bk_destroy",
    "bk_step", "bk_step_launch", "bk_step_wait", "bk_run_launch", "bk_run_wait", "bk_set_draw_peers", "bk_stream", "bk_set_response", "bk_set_host_output", "bk_sum_trees_host", "bk_export_trees", "bk_read_trace", "bk_export_forest", "bk_export_leaf_ids",
    "bk_set_history", "bk_history_batch", "bk_history_values", "bk_history_batch_at", "bk_history_values_at", "bk_export_leaf_values", "bk_predict_history", "bk_pearson_r2",
)

_lib = None


def load():
    """Load libpgbart_b200.so (raises RuntimeError if it has not been built)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(the PGBART step has no CPU fallback)"
        )
    lib = C.CDLL(LIB_PATH)
    lib.bk_abi_version.restype = C.c_int
    lib.bk_last_error.restype = C.c_char_p
    lib.bk_padded_rows.argtypes = [C.c_int]
    lib.bk_padded_rows.restype = C.c_int
    lib.bk_query_bytes.argtypes = [C.POINTER(BkSettings), C.POINTER(C.c_size_t)]
    lib.bk_create.argtypes = [C.POINTER(BkSettings), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]
    lib.bk_destroy.argtypes = [C.c_void_p]
    lib.bk_destroy.restype = None
    lib.bk_step.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.bk_step_launch.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    lib.bk_step_wait.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.bk_run_launch.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    lib.bk_run_wait.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.bk_stream.argtypes = [C.c_void_p]
    lib.bk_stream.restype = C.c_void_p
    lib.bk_set_draw_peers.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    lib.bk_set_response.argtypes = [C.c_void_p, C.c_void_p]
    lib.bk_set_host_output.argtypes = [C.c_void_p, C.c_int]
    lib.bk_sum_trees_host.argtypes = [C.c_void_p]
    lib.bk_sum_trees_host.restype = C.POINTER(C.c_float)
    lib.bk_export_trees.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    lib.bk_read_trace.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    lib.bk_export_forest.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    lib.bk_export_leaf_ids.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    lib.bk_set_history.argtypes = [C.c_void_p, C.c_int]
    lib.bk_history_batch.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.c_void_p, C.c_void_p, C.POINTER(C.c_int64)]
    lib.bk_history_values.argtypes = [C.c_void_p, C.c_void_p]
    lib.bk_history_batch_at.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int32), C.c_void_p, C.c_void_p, C.POINTER(C.c_int64)]
    lib.bk_history_values_at.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    lib.bk_export_leaf_values.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    lib.bk_predict_history.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                       C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                       C.c_void_p, C.c_void_p]
    lib.bk_pearson_r2.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
    if lib.bk_abi_version() != BK_ABI_VERSION:
        raise RuntimeError("libpgbart_b200.so ABI version mismatch; rebuild")
    _lib = lib
    return lib


def check(rc: int, what: str):
    if rc != 0:
        msg = load().bk_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed ({rc}): {msg}")
