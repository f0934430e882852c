"""ctypes binding of csrc/libnlam_b200.so (C ABI: include/nlam_b200.h).

There is deliberately NO fallback: if the shared library is missing or a call
fails, a RuntimeError is raised.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libnlam_b200.so")

MAX_SRC = 3
TILE_ROWS = 64
FP32, BF16 = 0, 1

c_float_p = ctypes.c_void_p  # device pointers travel as integers
c_int_p = ctypes.c_void_p


class Src(ctypes.Structure):
    _fields_ = [
        ("ptr", c_float_p),
        ("idx", c_int_p),
        ("batch_stride", ctypes.c_int64),
        ("ld", ctypes.c_int32),
        ("width", ctypes.c_int32),
        ("shadow", ctypes.c_void_p),
        ("shadow_batch_stride", ctypes.c_int64),
        ("shadow_rows", ctypes.c_int64),
    ]


class MlpWeights(ctypes.Structure):
    _fields_ = [(n, c_float_p) for n in ("w1", "b1", "w2", "b2", "ln_g", "ln_b")]


class Agg(ctypes.Structure):
    _fields_ = [
        ("seg_ptr", c_int_p),
        ("tile_seg", c_int_p),
        ("scale", c_float_p),
        ("out", c_float_p),
        ("n_seg", ctypes.c_int32),
        ("out_bf16", ctypes.c_void_p),
    ]


class RowMlp(ctypes.Structure):
    _fields_ = [
        ("n_src", ctypes.c_int32),
        ("src", Src * MAX_SRC),
        ("batch", ctypes.c_int32),
        ("rows", ctypes.c_int32),
        ("d_hidden", ctypes.c_int32),
        ("d_out", ctypes.c_int32),
        ("w", MlpWeights),
        ("n_chunks", ctypes.c_int32),
        ("tile_ptr", c_int_p),
        ("tile_chunk", c_int_p),
        ("chunk_ptr", c_int_p),
        ("n_tiles", ctypes.c_int32),
        ("residual_src", ctypes.c_int32),
        ("out", c_float_p),
        ("out_res", c_float_p),
        ("out_idx", c_int_p),
        ("agg", Agg),
        ("precision", ctypes.c_int32),
        ("out_bf16", ctypes.c_void_p),
        ("out_res_bf16", ctypes.c_void_p),
    ]


class RowMlpBwd(ctypes.Structure):
    _fields_ = [
        ("fwd", RowMlp),
        ("g0", c_float_p),
        ("g1", c_float_p),
        ("g1_idx", c_int_p),
        ("g1_scale", c_float_p),
        ("g1_batch_stride", ctypes.c_int64),
        ("d_src", c_float_p * MAX_SRC),
        ("g0_idx", c_int_p),
        ("d_src_idx", c_int_p * MAX_SRC),
        ("reduce_src", ctypes.c_int32),
        ("reduce_accumulate", ctypes.c_int32),
        ("d_params", c_float_p),
        ("params_accumulate", ctypes.c_int32),
        ("workspace", c_float_p),
        ("workspace_floats", ctypes.c_size_t),
        ("stage_mask", ctypes.c_int32),
        ("g0_sum_count", ctypes.c_int32),
        ("g0_sum_stride", ctypes.c_int64),
        ("g0_bf16", ctypes.c_void_p),
        ("g1_bf16", ctypes.c_void_p),
        ("d_src_bf16", ctypes.c_void_p * MAX_SRC),
        ("sp_src", ctypes.c_int32),
        ("n_sp", ctypes.c_int32),
        ("sp_tile_ptr", c_int_p),
        ("sp_row_ptr", c_int_p),
        ("sp_rows", c_int_p),
        ("src0_batch_sum", ctypes.c_int32),
    ]


class SegSum(ctypes.Structure):
    _fields_ = [
        ("src", c_float_p),
        ("src_batch_stride", ctypes.c_int64),
        ("ptr", c_int_p),
        ("idx", c_int_p),
        ("scale", c_float_p),
        ("out", c_float_p),
        ("batch", ctypes.c_int32),
        ("n_out", ctypes.c_int32),
        ("width", ctypes.c_int32),
        ("accumulate", ctypes.c_int32),
    ]


class StateStep(ctypes.Structure):
    _fields_ = [(n, c_float_p) for n in ("net_out", "prev", "truth", "diff_std", "diff_mean",
                                         "inv_std", "interior", "new_state", "loss_partial",
                                         "loss_sum")] + [
        ("rows", ctypes.c_int64), ("nodes", ctypes.c_int32), ("features", ctypes.c_int32),
        ("prev_batch_stride", ctypes.c_int64), ("truth_batch_stride", ctypes.c_int64)]


class StateStepBwd(ctypes.Structure):
    _fields_ = [("fwd", StateStep), ("d_new", c_float_p), ("d_loss", c_float_p),
                ("d_net_out", c_float_p), ("d_prev", c_float_p)]


FEED_MAX_BATCH = 64


class FeedBatch(ctypes.Structure):
    _fields_ = [
        ("state", c_float_p), ("forcing", c_float_p), ("times", ctypes.c_void_p),
        ("sample_idx", ctypes.c_int64 * FEED_MAX_BATCH),
        ("batch", ctypes.c_int32), ("n_grid", ctypes.c_int32), ("d_state", ctypes.c_int32),
        ("d_forcing", ctypes.c_int32), ("ar_steps", ctypes.c_int32), ("past", ctypes.c_int32),
        ("future", ctypes.c_int32), ("ring_cap", ctypes.c_int32),
        ("t_lo", ctypes.c_int64), ("t_hi", ctypes.c_int64),
        ("init_states", c_float_p), ("target_states", c_float_p), ("forcing_out", c_float_p),
        ("target_times", ctypes.c_void_p),
    ]


# every symbol include/nlam_b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "nlam_last_error": (ctypes.c_char_p, []),
    "nlam_version": (ctypes.c_int, []),
    "nlam_launch_count": (ctypes.c_int64, []),
    "nlam_set_option": (ctypes.c_int, [ctypes.c_char_p, ctypes.c_int]),
    "nlam_csr_build": (
        ctypes.c_int,
        [c_int_p, ctypes.c_int64, ctypes.c_int32, c_int_p, c_int_p, c_float_p, c_int_p,
         ctypes.c_void_p],
    ),
    "nlam_rowmlp_fwd": (ctypes.c_int, [ctypes.POINTER(RowMlp), ctypes.c_void_p]),
    "nlam_rowmlp_bwd_workspace": (ctypes.c_size_t, [ctypes.POINTER(RowMlp)]),
    "nlam_rowmlp_param_floats": (ctypes.c_size_t, [ctypes.POINTER(RowMlp)]),
    "nlam_rowmlp_path": (ctypes.c_int, [ctypes.POINTER(RowMlp)]),
    "nlam_rowmlp_bwd_stages": (ctypes.c_int, [ctypes.POINTER(RowMlpBwd)]),
    "nlam_rowmlp_bwd_run": (ctypes.c_int, [ctypes.POINTER(RowMlpBwd), ctypes.c_void_p]),
    "nlam_rowmlp_bwd_flush": (ctypes.c_int, [ctypes.c_void_p]),
    "nlam_rowmlp_bwd_pending": (ctypes.c_int, []),
    "nlam_rowmlp_bwd_discard": (ctypes.c_int, [ctypes.c_void_p]),
    "nlam_segsum_run": (ctypes.c_int, [ctypes.POINTER(SegSum), ctypes.c_void_p]),
    "nlam_feed_standardize": (ctypes.c_int, [c_float_p, c_float_p, c_float_p, c_float_p,
                                             ctypes.c_int64, ctypes.c_int32, ctypes.c_void_p]),
    "nlam_feed_batch_run": (ctypes.c_int, [ctypes.POINTER(FeedBatch), ctypes.c_void_p]),
    "nlam_state_step_partials": (ctypes.c_int64, [ctypes.c_int64]),
    "nlam_state_step_fwd": (ctypes.c_int, [ctypes.POINTER(StateStep), ctypes.c_void_p]),
    "nlam_state_step_bwd_run": (ctypes.c_int, [ctypes.POINTER(StateStepBwd), ctypes.c_void_p]),
}

_lib = None


def load():
    """Load (once) and return the shared library; raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is not built -- run `python -c 'import __graft_entry__ as g; "
            "g.build()'` (there is no CPU or PyTorch fallback for this path)"
        )
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if a declared symbol is missing
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def check(status, what):
    if status != 0:
        msg = load().nlam_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed: {msg}")
