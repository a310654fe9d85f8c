"""ctypes binding of librb200.so (C ABI declared in include/rb200.h).

The shared library is the product: there is no PyTorch/CPU fallback.  If it is missing, or if a
tensor is not on a CUDA device, calls raise immediately.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import torch

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "librb200.so"

c_f = C.c_void_p  # all device pointers travel as void*


class RB200Error(RuntimeError):
    pass


class TowerJob(C.Structure):
    _fields_ = [("table", c_f), ("ids", c_f), ("extra", c_f), ("W1", c_f), ("b1", c_f), ("W2", c_f), ("b2", c_f),
                ("out", c_f), ("hid", c_f), ("denom", c_f), ("keep_mask", c_f), ("n_rows", C.c_int64),
                ("B", C.c_int), ("extra_dim", C.c_int), ("extra_by_id", C.c_int), ("img", c_f)]


class TowerBwdJob(C.Structure):
    _fields_ = [("table", c_f), ("ids", c_f), ("extra", c_f), ("n_rows", C.c_int64), ("B", C.c_int),
                ("extra_dim", C.c_int), ("extra_by_id", C.c_int), ("W1", c_f), ("W2", c_f), ("dY", c_f), ("y", c_f),
                ("denom", c_f), ("hid", c_f), ("dpre", c_f), ("dact", c_f), ("dRows", c_f), ("img", c_f)]


class SumsqSeg(C.Structure):
    _fields_ = [("x", c_f), ("n", C.c_int64), ("count", c_f), ("row_len", C.c_int)]


class OptState(C.Structure):   # host mirror of rb200_opt_state (128 bytes)
    _fields_ = [("lr", C.c_double), ("beta1", C.c_double), ("beta2", C.c_double), ("sumsq", C.c_double),
                ("step", C.c_int64), ("eps", C.c_float), ("weight_decay", C.c_float), ("max_norm", C.c_float),
                ("one_minus_beta1", C.c_float), ("one_minus_beta2", C.c_float), ("beta2_f", C.c_float),
                ("step_size", C.c_float), ("bias_corr2_sqrt", C.c_float), ("clip_coef", C.c_float),
                ("total_norm", C.c_float), ("loss", C.c_float), ("ticket", C.c_uint32), ("pad", C.c_float * 10)]


assert C.sizeof(OptState) == 128


class StepParams(C.Structure):
    _fields_ = [("D", C.c_int), ("H", C.c_int), ("extra_dim", C.c_int), ("B", C.c_int),
                ("n_user_rows", C.c_int64), ("n_item_rows", C.c_int64),
                ("user_table", c_f), ("user_table_m", c_f), ("user_table_v", c_f),
                ("item_table", c_f), ("item_table_m", c_f), ("item_table_v", c_f),
                ("user_mlp", c_f), ("user_mlp_m", c_f), ("user_mlp_v", c_f),
                ("item_mlp", c_f), ("item_mlp_m", c_f), ("item_mlp_v", c_f),
                ("user_row_slot", c_f), ("item_row_slot", c_f), ("opt", c_f),
                ("user_ids", c_f), ("pos_ids", c_f), ("neg_ids", c_f), ("pos_extra", c_f), ("neg_extra", c_f),
                ("extra_by_id", C.c_int), ("loss_kind", C.c_int), ("inbatch_mode", C.c_int), ("tower_mode", C.c_int),
                ("adam_mode", C.c_int),
                ("dropout_p", C.c_float), ("seed", C.c_uint64),
                ("keep_mask_user", c_f), ("keep_mask_pos", c_f), ("keep_mask_neg", c_f),
                ("padding_idx", C.c_int64), ("loss", c_f), ("err_flag", c_f),
                ("workspace", c_f), ("workspace_bytes", C.c_size_t), ("stage_events_host", c_f),
                ("grad_scale", C.c_float), ("dp_grads", c_f), ("next_batch", c_f)]


class Sampler(C.Structure):      # host mirror of rb200_sampler
    _fields_ = [("pos_users", c_f), ("pos_items", c_f), ("n_pos", C.c_int64), ("rated_offsets", c_f), ("rated_items", c_f),
                ("rated_bitmap", c_f), ("bitmap_words", C.c_int64), ("catalog", c_f), ("n_cat", C.c_int64),
                ("seed", C.c_uint64), ("batches_per_epoch", C.c_int64), ("rank", C.c_int64), ("world", C.c_int64)]


class StepViews(C.Structure):
    _fields_ = [("user_mlp_grad", c_f), ("item_mlp_grad", c_f), ("user_uniq_ids", c_f), ("item_uniq_ids", c_f),
                ("user_uniq_grads", c_f), ("item_uniq_grads", c_f), ("user_n_uniq", c_f), ("item_n_uniq", c_f),
                ("user_emb", c_f), ("pos_emb", c_f), ("neg_emb", c_f)]


I, I64, F, U64, SZ, P = C.c_int, C.c_int64, C.c_float, C.c_uint64, C.c_size_t, C.c_void_p

# name -> (restype, argtypes); must list every function include/rb200.h declares
SIGNATURES = {
    "rb200_version": (I, []),
    "rb200_last_error": (C.c_char_p, []),
    "rb200_sm_count": (I, []),
    "rb200_sizeof": (SZ, [I]),
    "rb200_launch_count": (U64, []),
    "rb200_stamp": (I, [P, I, P]),
    "rb200_tower_fwd": (I, [C.POINTER(TowerJob), I, I, I, F, U64, U64, P, I, P, P, SZ, P]),
    "rb200_tower_fwd_workspace_bytes": (SZ, [I, I, I, I, I]),
    "rb200_tower_img_bytes": (SZ, [I, I, I]),
    "rb200_tower_prep": (I, [P, P, I, I, I, P, P]),
    "rb200_tower_bwd_workspace_bytes": (SZ, [I, I, I]),
    "rb200_tower_bwd": (I, [C.POINTER(TowerBwdJob), I, I, I, F, I, P, I, P, SZ, P]),
    "rb200_bpr_pair": (I, [P, P, P, I, I, P, P, P, P, F, P, SZ, P]),
    "rb200_bpr_pair_workspace_bytes": (SZ, [I]),
    "rb200_bpr_inbatch_workspace_bytes": (SZ, [I, I]),
    "rb200_bpr_inbatch": (I, [P, P, I, I, I, P, P, P, F, P, SZ, P]),
    "rb200_scatter_workspace_bytes": (SZ, [I, I64]),
    "rb200_scatter_rows": (I, [P, P, I, I, I64, I64, P, P, P, P, P, P, SZ, P]),
    "rb200_scatter_plan": (I, [P, I, I64, I64, P, P, P, P, SZ, P]),
    "rb200_scatter_apply": (I, [P, I, I, I64, P, P, P, P, P, SZ, P]),
    "rb200_scatter_reset_slots": (I, [P, P, I, P, P]),
    "rb200_scatter_set_slots": (I, [P, P, I, P, P]),
    "rb200_gather_rows": (I, [P, P, I64, I, I64, P, P]),
    "rb200_gather_rows_sharded": (I, [P, P, I, P, I64, P, I64, I64, I64, I, P, P, P]),
    "rb200_push_rows_sharded": (I, [P, P, I, I, I64, P, P, I64, I, P, P]),
    "rb200_push_row_lists_sharded": (I, [P, I, I, I64, P, P]),
    "rb200_allreduce_oneshot": (I, [P, I, I64, P, P]),
    "rb200_allreduce_twoshot": (I, [P, I, I, I64, P]),
    "rb200_allreduce_multimem": (I, [P, I, I, I64, P]),
    "rb200_sharded_scalars_publish": (I, [P, P, F, P, P]),
    "rb200_sharded_scalars_reduce": (I, [P, I, P, P]),
    "rb200_sharded_scalars_finish": (I, [P, I, P, I64, P, P]),
    "rb200_adam_rows_dense2": (I, [P, P, P, I, P, P, P, I, P, P, P, P, I64, P, P, P, P, I64, P, P]),
    "rb200_sample_batch": (I, [P, P, I64, P, P, P, I64, I, U64, I64, I64, I64, I64, P, P, P, P]),
    "rb200_sample_batch_dev": (I, [P, I, P, P, P, P, P]),
    "rb200_route_plan_workspace_bytes": (SZ, [I64, I]),
    "rb200_route_plan": (I, [P, I64, P, I64, I, P, P, P, P, P, P, SZ, P]),
    "rb200_route_plan_padded_workspace_bytes": (SZ, [I64, I]),
    "rb200_route_plan_padded": (I, [P, I64, P, I64, I, P, I64, P, P, P, P, SZ, P]),
    "rb200_opt_begin_step": (I, [P, P]),
    "rb200_sumsq_accumulate": (I, [P, C.POINTER(SumsqSeg), I, P, SZ, P]),
    "rb200_sumsq_workspace_bytes": (SZ, []),
    "rb200_grad_norm_clip": (I, [P, P]),
    "rb200_adam_dense": (I, [P, P, P, P, I64, P, P]),
    "rb200_adam_table_dense": (I, [P, P, P, I64, I, P, P, P, P]),
    "rb200_adam_rows": (I, [P, P, P, I, P, P, P, I, P, P]),
    "rb200_bpr_step_workspace_bytes": (SZ, [I, I, I, I, I64, I64, I]),
    "rb200_bpr_step": (I, [C.POINTER(StepParams), P]),
    "rb200_bpr_apply": (I, [C.POINTER(StepParams), P]),
    "rb200_bpr_dp_grad_floats": (SZ, [I, I, I, I64, I64]),
    "rb200_bpr_step_views": (I, [C.POINTER(StepParams), C.POINTER(StepViews)]),
    "rb200_gemm_nt": (I, [P, I, P, I, I, I, P, I64, P, P]),
    "rb200_normalize_rows": (I, [P, I64, I, F, P, P]),
    "rb200_ivf_assign": (I, [P, I64, I, P, I, P, P, P]),
    "rb200_kmeans_update_workspace_bytes": (SZ, [I64, I]),
    "rb200_kmeans_update": (I, [P, I64, I, P, I, P, P, P, SZ, P]),
    "rb200_ivf_build_workspace_bytes": (SZ, [I64, I]),
    "rb200_ivf_build": (I, [P, I64, I, P, I, P, P, P, P, SZ, P]),
    "rb200_ivf_plan_workspace_bytes": (SZ, [I, I, I]),
    "rb200_ivf_search_plan": (I, [P, I, I, P, I, I, P, P, SZ, C.POINTER(I64), C.POINTER(I64), P]),
    "rb200_ivf_search_workspace_bytes": (SZ, [I64]),
    "rb200_ivf_search_status": (I, [P, SZ, I, I, I, P]),
    "rb200_ivf_search_run": (I, [P, I, I, I, I, P, P, P, I64, I64, P, P, I, I, P, SZ, I64, I64, P, P, P, SZ, P]),
    "rb200_flat_search_workspace_bytes": (SZ, [I, I64, I]),
    "rb200_flat_search": (I, [P, I, P, I64, I, I, I64, P, P, P, SZ, P]),
    "rb200_topk_merge_workspace_bytes": (SZ, [I, I, I]),
    "rb200_topk_merge": (I, [P, P, I, I, I, P, P, P, SZ, P]),
}

_lib = None


def load() -> C.CDLL:
    """dlopen librb200.so (built in-tree by ``__graft_entry__.build()`` / ``make -C csrc``)."""
    global _lib
    if _lib is not None:
        return _lib
    path = Path(os.environ.get("RB200_LIB", LIB_PATH))
    if not path.exists():
        raise RB200Error(
            f"{path} not found: the CUDA extension is not built. Run `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C recommendit_b200/csrc`. recommendit_b200 has no CPU fallback.")
    lib = C.CDLL(str(path))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)   # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    for which, cls in enumerate((TowerJob, TowerBwdJob, OptState, StepParams, StepViews, SumsqSeg, Sampler)):
        if lib.rb200_sizeof(which) != C.sizeof(cls):
            raise RB200Error(f"ABI mismatch: sizeof({cls.__name__}) is {C.sizeof(cls)} here, {lib.rb200_sizeof(which)} in {path}")
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().rb200_last_error().decode("utf-8", "replace")
        raise RB200Error(f"{what or 'rb200'} failed (code {rc}): {msg}")


def require_cuda(*tensors) -> torch.device:
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RB200Error(
                "recommendit_b200 computes only on CUDA devices (sm_100a kernels, no CPU fallback); "
                f"got a tensor on {t.device}. Move the model and its inputs to 'cuda'.")
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise RB200Error(f"tensors on different devices: {dev} and {t.device}")
    return dev


def ptr(t) -> int | None:
    return None if t is None else t.data_ptr()


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def workspace(nbytes: int, device) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
