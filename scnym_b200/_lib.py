"""ctypes binding of libscnym_b200.so (the C ABI declared in include/scnym_b200.h).

There is no CPU fallback: if the shared library is missing, `lib()` raises.  The
library is built in-tree by `scnym_b200/csrc/build.py` (see `__graft_entry__.build`).
"""
import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libscnym_b200.so")

P, I, L, F, U32 = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_float, ctypes.c_uint32

# name -> argtypes (order == include/scnym_b200.h)
SIGNATURES = {
    "scnb_csr_row_offsets": [P, P, L, I, P, P, P],
    "scnb_gather_i32": [P, P, I, I, P, P],
    "scnb_csr_gather_rows": [P, P, P, P, L, I, P, P, P, I, I, I, P, P, U32, F, P, P],
    "scnb_csr_gather_rows2": [P, P, P, P, I, I, U32, P, P, P, P, I, I, U32, P, P, P, P, P, F, P, I, P],
    "scnb_csr_gather_rows2_nocache": [P, P, P, P, I, I, U32, P, P, P, P, I, I, U32, P, P, P, P, P, F, P, I, P],
    "scnb_csr_fetch_rows2": [P, P, P, P, I, P, P, P, P, I, P, P, P, P, I, P],
    "scnb_csr_gather_rows_aug": [P, P, P, P, L, I, P, P, P, I, I, F, F, P, P, P, P, U32, P, I, P],
    "scnb_step_prep": [P, P, P, P, I, P, I, I, P, P, P, P, P, P, P, I, P, P, P, P, P, P, I, I, P],
    "scnb_fetch_step": [P, P, P, P, I, I, I, P, P, P, P, P, P],
    "scnb_host_gather_rows": [P, P, P, P, I, P, P, P, L, I],
    "scnb_host_gather_rows2": [P, P, P, P, I, P, P, P, L, P, P, P, P, I, P, P, P, L, I],
    "scnb_csr_linear_fwd": [P, P, P, I, P, P, I, P, P],
    "scnb_csr_linear_fwd_i64": [P, P, P, I, P, P, I, P, P],
    "scnb_csr_linear_bwd_dw": [P, P, P, I, P, I, P, P],
    "scnb_hs_classif_head": [P, P, P, P, P, P, P, F, I, I, I, P, P, P, I, I, I, P, P, P, P, P, P, F, P, P, P, P, P, P],
    "scnb_hs_dan_head": [P, P, P, P, P, P, P, F, I, I, I, I, P, P, P, P, P, I, P, P, P, I, P, F, P, P, P, P, P, P, P, P, P],
    "scnb_hs_bwd_in_unmix_planes": [P, P, I, I, I, P, P, P, P, P, P, P, P, P, P, I, P, P, P, P],
    "scnb_w1_planes": [P, I, I, P, P, P],
    "scnb_csr_linear_fwd_tc": [P, P, P, I, I, I, P, P, P, P, I, P],
    "scnb_csr_linear_fwd_tc_i64": [P, P, P, I, I, I, P, P, P, P, I, P],
    "scnb_csr_linear_fwd_tc_geometry": [I, I, I, P, P],
    "scnb_row_fwd_splits": [P, P, I, I, I, P, P],
    "scnb_csr_linear_fwd_tc_splits": [P, P, P, I, I, I, P, P, P, P, I, P, P],
    "scnb_w1tc_gene_block": [I],
    "scnb_row_gene_splits2": [P, P, I, I, I, P, P, P],
    "scnb_row_gene_splits": [P, P, I, I, P, P],
    "scnb_csr_w1_bwd_optim_tc": [I, P, P, P, I, I, I, P, P, P, P, P, P, P, P, P, P, P],
    "scnb_linear_planes": [P, I, I, P, P, P],
    "scnb_bn_eval_planes": [P, I, I, P, P, P, P, I, P, P, P, P, P],
    "scnb_linear_fwd_tc_planes": [P, P, I, I, I, P, P, P, P, P],
    "scnb_linear_fwd_tc_planes_bn_planes": [P, P, I, I, I, P, P, P, P, P, P, P, I, P, P, P],
    "scnb_linear_fwd_tc_planes_bn_head": [P, P, I, I, I, P, P, P, P, P, P, P, I, P, P, I, P, I, P, P],
    "scnb_csr_linear_fwd_tc_bn_planes_i64": [P, P, P, I, I, I, P, P, P, P, P, P, P, I, P, P, P],
    "scnb_head_logits": [P, P, P, I, I, I, P, I, P],
    "scnb_csc_count": [P, P, I, P, P],
    "scnb_csr_to_csc": [P, P, P, I, I, P, P, P, P, P],
    "scnb_csr_w1_bwd_optim": [I, P, P, P, I, I, I, P, P, P, P, P, P, P],
    "scnb_sgemm": [I, I, I, I, I, P, I, P, I, P, I, F, F, P, I, P, P],
    "scnb_gemm_tc_min_flops": [L],
    "scnb_gemm_tc_min_flops_get": [P],
    "scnb_gemm_tc_workspace": [P, P, L],
    "scnb_gemm_tc_workspace_bytes": [I, I, I, P],
    "scnb_gemm_tc_planes_count": [P],
    "scnb_gemm_tc_bind_planes": [P, I, I, P, P],
    "scnb_gemm_tc_bound_count": [P],
    "scnb_colsum": [P, I, I, I, P, I, P],
    "scnb_relu_bwd": [P, P, P, L, P],
    "scnb_bn_eval_bwd": [P, P, I, I, I, P, P, P, P],
    "scnb_bn_act_fwd": [P, P, P, I, I, P, I, I, P, P, P, P, P, P, P, U32, F, I, P, P, P],
    "scnb_bn_act_bwd": [P, P, P, I, I, P, I, P, P, P, P, U32, F, I, P, P, P, P, P, P],
    "scnb_bn_act_fwd_peer": [P, P, P, I, I, P, I, P, P, P, P, P, U32, F, I, P, I, I, I, I, P, L, P, P],
    "scnb_bn_act_bwd_peer": [P, P, P, I, I, P, I, P, P, P, P, U32, F, I, P, P, P, P, I, I, I, I, P, L, P],
    "scnb_bn_peer_bytes": [I, I, I, I],
    "scnb_peer_alloc": [L, P, P],
    "scnb_host_alloc_mapped": [L, P, P],
    "scnb_host_free_mapped": [P],
    "scnb_publish_f32": [P, I, P, P],
    "scnb_host_register_mapped": [P, L, P],
    "scnb_host_unregister": [P],
    "scnb_peer_open": [P, P],
    "scnb_peer_close": [P],
    "scnb_peer_free": [P],
    "scnb_bn_running_update": [P, P, P, P, P, P, P, P, I, I, F, P, P],
    "scnb_mix_rows": [P, I, P, P, P, P, I, P, P],
    "scnb_hs_mix": [P, I, P, P, P, I, P, P, P, P, U32, F, P],
    "scnb_hs_act": [P, P, I, I, I, P, P, P, P, P, P, P, U32, F, P],
    "scnb_hs_fwd": [P, P, P, P, P, P, P, F, P, P, P, P, P, P, U32, F, I, I, I, I, P, P, P, P],
    "scnb_hs_bwd_act": [P, P, P, I, I, I, P, P, F, P, P, P],
    "scnb_hs_bwd": [P, P, P, P, P, P, P, P, P, P, P, P, F, P, P, I, I, I, I, P, P, P],
    "scnb_hs_dw": [P, I, P, I, I, P, P, P],
    "scnb_hs_bwd_in": [P, P, I, I, I, P, P, P, P, P, P, P, P, P, P],
    "scnb_hs_peer_bytes": [I, I, I, I],
    "scnb_hs_debug_read": [P],
    "scnb_hh_debug_read": [P],
    "scnb_hs_xchg": [I, P, I, I, I, P, P, I, I, I, I, I, P, L, P, P, P, P, P],
    "scnb_unmix_rows": [P, I, P, P, I, P, P],
    "scnb_unmix_rows_planes": [P, I, P, P, I, P, P, P, P],
    "scnb_pseudolabel": [P, I, I, I, F, I, F, P, P, P, P, P, I, P],
    "scnb_mixmatch_plan": [P, P, P, I, I, I, I, I, P, P, P, P, P, P, I, P, P, P, P],
    "scnb_soft_ce_fwd_bwd": [P, I, I, P, P, P, P, P, I, P, P, I, F, P, P, P, I, P, P],
    "scnb_softmax_mse_fwd_bwd": [P, I, I, P, P, P, P, P, I, I, P, F, P, P, P],
    "scnb_classif_mixmatch_fused": [P, P, P, I, I, I, I, P, P, P, P, P, P, F, P, P, P, P, P],
    "scnb_classif_supervised_fused": [P, P, P, I, I, I, P, P, P, P, P, P, P],
    "scnb_dan_tail_fused": [P, P, P, I, I, I, P, P, P, I, P, P, P, P, P, P, P],
    "scnb_teacher_head_fused": [P, P, P, I, I, I, I, F, I, F, P, P, P, P, P, P, I, P],
    "scnb_teacher_head_bn_fused": [P, P, P, P, P, P, P, I, I, I, I, F, I, F, P, P, P, P, P, P, I, P],
    "scnb_onehot_rows": [P, P, I, I, I, P, P],
    "scnb_step_begin": [P, P],
    "scnb_mm_step_inc": [P, P],
    "scnb_zero2": [P, L, P, L, P],
    "scnb_optim_step": [I, P, P, P, P, L, P, P, P],
    "scnb_dp_reduce_optim_bcast": [I, P, I, I, L, L, L, L, P, P, P, P, P, P],
    "scnb_dp_reduce_optim_bcast2": [I, P, I, I, L, L, L, L, L, L, P, P, P, P, P, P],
    "scnb_dp_slice_floats": [L, I, P],
    "scnb_csr_w1_bwd_push_tc": [P, P, P, I, I, I, P, P, P, I, I, L, L, P, P],
    "scnb_ema_update": [P, P, L, F, P, P],
    "scnb_record_step": [P, P, I, P, I, P, I, P, P],
    "scnb_predict_head": [P, I, I, I, P, P, P, P],
    "scnb_csr_remap_workspace_len": [L],
    "scnb_csr_remap_count": [P, P, P, I, L, P, P, P],
    "scnb_csr_remap_fill": [P, P, P, P, I, I, L, P, P, P, P],
}

OPTIMIZER_CODES = {"none": -1, "adadelta": 0, "adam": 1, "adamw": 2, "sgd": 3}

_lib = None
_lock = threading.Lock()
launch_count = 0  # number of C-ABI compute calls issued (bench.py reports kernels per step from this)


class ScnbError(RuntimeError):
    pass


def lib():
    """Load (once) and return the ctypes handle.  Raises if the library is absent."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise ScnbError(
                f"{LIB_PATH} not found: build it with `python -m scnym_b200.csrc.build` "
                "(scnym_b200 has no CPU or eager fallback)")
        h = ctypes.CDLL(LIB_PATH)
        h.scnb_last_error.restype = ctypes.c_char_p
        h.scnb_last_error.argtypes = []
        h.scnb_version.restype = I
        h.scnb_device_arch.restype = I
        for name, argtypes in SIGNATURES.items():
            fn = getattr(h, name)
            fn.argtypes = argtypes
            fn.restype = I
        _lib = h
    return _lib


def call(name, *args):
    """Invoke a C-ABI entry point; raise ScnbError on a non-zero return."""
    global launch_count
    h = lib()
    rc = getattr(h, name)(*args)
    launch_count += 1
    if rc != 0:
        raise ScnbError(f"{name} failed ({rc}): {h.scnb_last_error().decode()}")
    return rc


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL)."""
    if t is None:
        return None
    return t.data_ptr()


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise ScnbError("scnym_b200 kernels need CUDA tensors; there is no CPU fallback")


def current_stream():
    import torch
    return torch.cuda.current_stream().cuda_stream
