"""ctypes binding of libdmf_b200.so (the C ABI declared in include/dmf_b200.h).

There is NO fallback: if the shared library is missing the import fails loudly, and every entry
point raises on a non-B200 device (``dmf_device_check``).
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdmf_b200.so")


class DmfError(RuntimeError):
    pass


if not os.path.isfile(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} not found: build the sm_100a kernels first "
        "(python -c 'import __graft_entry__ as g; g.build()'  or  make -C disentagled_multimodal_fusion_b200/csrc). "
        "This package has no CPU / PyTorch fallback.")

lib = C.CDLL(LIB_PATH)

c_f = C.c_float
c_i = C.c_int
c_ll = C.c_longlong
c_p = C.c_void_p
c_sz = C.c_size_t
c_ull = C.c_ulonglong


class GemmDesc(C.Structure):
    _fields_ = [("A", c_p), ("a_rs", c_ll), ("a_cs", c_ll),
                ("B", c_p), ("b_rs", c_ll), ("b_cs", c_ll),
                ("C", c_p), ("ldc", c_ll),
                ("bias", c_p),
                ("aux", c_p), ("ldaux", c_ll),
                ("rowsum_a", c_p),
                ("M", c_i), ("N", c_i), ("K", c_i),
                ("accumulate", c_i)]


class TcGemmDesc(C.Structure):
    _fields_ = [("A", c_p), ("lda", c_ll),
                ("B", c_p), ("ldb", c_ll),
                ("out_f32", c_p), ("ldo_f32", c_ll),
                ("out_bf16", c_p), ("ldo_bf16", c_ll),
                ("bias", c_p),
                ("mask_bf16", c_p), ("ldmask", c_ll),
                ("M", c_i), ("N", c_i), ("K", c_i),
                ("out_bf16_t", c_p), ("ldo_t", c_ll),
                ("split_k", c_i), ("mn_major", c_i)]


class EdlParams(C.Structure):
    _fields_ = [("B", c_i), ("V", c_i), ("C", c_i), ("agg", c_i),
                ("coef", c_f), ("dc_weight", c_f), ("inv_B_global", c_f)]


class HeadGemmDesc(C.Structure):
    _fields_ = [("A", c_p), ("lda", c_ll), ("W", c_p), ("ldw", c_ll), ("bias", c_p),
                ("pre_f32", c_p), ("ld_pre_f32", c_ll), ("pre_bf16", c_p), ("ld_pre_bf16", c_ll),
                ("out_f32", c_p), ("ld_out_f32", c_ll), ("out_bf16", c_p), ("ld_out_bf16", c_ll),
                ("inv_norm", c_p), ("noise_w", c_p), ("noise_v", c_p),
                ("eps", c_f), ("M", c_i), ("N", c_i), ("K", c_i)]


EPI_NONE, EPI_BIAS, EPI_BIAS_RELU, EPI_RELU_MASK, EPI_BIAS_EVIDENCE = range(5)
AGG = {"cml": 0, "avg": 1, "joint": 2, "disentangled": 3, "dbf": 4}

_SIGS = {
    "dmf_version": ([], c_i),
    "dmf_last_error": ([C.c_char_p, c_sz], c_i),
    "dmf_device_check": ([], c_i),
    "dmf_launch_count": ([], c_ll),
    "dmf_grouped_gemm_f32": ([C.POINTER(GemmDesc), c_i, c_i, c_p], c_i),
    "dmf_grouped_gemm_bf16_tc": ([C.POINTER(TcGemmDesc), c_i, c_i, c_p], c_i),
    "dmf_head_gemm_bf16": ([C.POINTER(HeadGemmDesc), c_i, c_i, c_p], c_i),
    "dmf_colsum_f32": ([c_p, c_ll, c_i, c_i, c_p, c_i, c_p], c_i),
    "dmf_cast_f32_to_bf16": ([c_p, c_ll, c_p, c_ll, c_i, c_i, c_p], c_i),
    "dmf_cast_transpose_f32_to_bf16": ([c_p, c_ll, c_p, c_ll, c_i, c_i, c_p], c_i),
    "dmf_transpose_bf16": ([c_p, c_ll, c_p, c_ll, c_i, c_i, c_p], c_i),
    "dmf_cast_dual_bf16": ([c_p, c_ll, c_p, c_ll, c_p, c_ll, c_p, c_i, c_i, c_p], c_i),
    "dmf_colsum_bf16": ([c_p, c_ll, c_i, c_i, c_p, c_p], c_i),
    "dmf_rowlse": ([c_p, c_ll, c_i, c_p, c_ll, c_i, c_i, c_f, c_p, c_p, c_ll, c_p, c_p, c_sz, c_i, c_p], c_i),
    "dmf_infonce_rowcol_sums": ([c_p, c_ll, c_i, c_p, c_ll, c_i, c_i, c_f, c_f, c_i, c_i, c_p, c_p, c_ll, c_p, c_p], c_i),
    "dmf_rowlse_workspace_bytes": ([c_i, c_i], c_sz),
    "dmf_infonce_rowcol_sums_store": ([c_p, c_ll, c_i, c_p, c_ll, c_i, c_i, c_f, c_f, c_i, c_i, c_p, c_p, c_ll, c_p, c_p, c_p], c_i),
    "dmf_infonce_e_bytes": ([c_i, c_i], c_sz),
    "dmf_infonce_bwd_stored_work_floats": ([c_i, c_i], c_sz),
    "dmf_infonce_bwd_stored": ([c_p, c_i, c_i, c_p, c_p, c_f, c_p, c_ll, c_i, c_i, c_f, c_p, c_ll, c_p, c_ll, c_i, c_p, c_p], c_i),
    "dmf_infonce_finalize": ([c_p, c_p, c_p, c_p, c_p, c_p, c_i, c_f, c_f, c_i, c_p, c_p, c_p], c_i),
    "dmf_infonce_bwd": ([c_p, c_ll, c_i, c_p, c_p, c_ll, c_p, c_ll, c_i, c_p, c_i, c_f, c_f, c_p, c_ll, c_p, c_ll,
                         c_i, c_i, c_p], c_i),
    "dmf_infonce_bwd_needs_transposed": ([c_i], c_i),
    "dmf_row_normalize_fwd": ([c_p, c_ll, c_i, c_i, c_f, c_p, c_ll, c_p, c_ll, c_p, c_p], c_i),
    "dmf_row_normalize_bwd": ([c_p, c_ll, c_p, c_p, c_ll, c_i, c_i, c_p, c_ll, c_i, c_p], c_i),
    "dmf_sumsq_f32": ([c_p, c_ll, c_p, c_p], c_i),
    "dmf_vmf_fwd": ([c_p, c_ll, c_p, c_p, c_i, c_i, c_p, c_ll, c_p, c_ll, c_p], c_i),
    "dmf_vmf_bwd": ([c_p, c_ll, c_p, c_p, c_p, c_ll, c_i, c_i, c_p, c_ll, c_i, c_p], c_i),
    "dmf_vmf_draw": ([c_p, c_p, c_i, c_i, c_f, c_ull, c_ull, c_p], c_i),
    "dmf_augment": ([c_p, c_ll, c_p, c_ll, c_i, c_i, c_f, c_i, c_ull, c_ull, c_p, c_p], c_i),
    "dmf_vmf_draw_ctr": ([c_p, c_p, c_i, c_i, c_f, c_ull, c_ull, c_p, c_p], c_i),
    "dmf_augment_ctr": ([c_p, c_ll, c_p, c_ll, c_i, c_i, c_f, c_i, c_ull, c_ull, c_p, c_p, c_p], c_i),
    "dmf_counter_add": ([c_p, c_ull, c_p], c_i),
    "dmf_dmvae_head_fwd": ([C.POINTER(c_p), c_p, c_i, c_i, c_i, c_f, C.POINTER(c_p), c_p, c_p], c_i),
    "dmf_dmvae_head_bwd": ([C.POINTER(c_p), c_p, C.POINTER(c_p), c_i, c_i, c_i, c_f, c_p, C.POINTER(c_p), c_p], c_i),
    "dmf_dmvae_poe_mean": ([C.POINTER(c_p), c_i, c_i, c_i, c_f, c_p, c_p], c_i),
    "dmf_dmvae_mse_fwd_bwd": ([c_p, c_ll, c_p, c_ll, c_i, c_i, c_i, c_i, c_f, c_f, c_p, c_p, c_p, c_ll, c_p], c_i),
    "dmf_edl_fused": ([c_p, c_p, C.POINTER(EdlParams), c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p], c_i),
    "dmf_eval_reduce": ([c_p, c_p, c_p, c_i, c_i, c_i, c_p, c_p, c_p, c_p, c_p], c_i),
    "dmf_evidence_fwd": ([c_p, c_p, c_ll, c_p], c_i),
    "dmf_evidence_bwd": ([c_p, c_p, c_p, c_p, c_ll, c_p], c_i),
    "dmf_adam_step": ([c_p, c_p, c_p, c_p, c_ll, c_f, c_f, c_f, c_f, c_f, c_i, c_i, c_f, c_p, c_p], c_i),
    "dmf_adam_step_dev": ([c_p, c_p, c_p, c_p, c_ll, c_p, c_f, c_f, c_f, c_f, c_i, c_f, c_p, c_p], c_i),
    "dmf_fill_f32": ([c_p, c_ll, c_f, c_p], c_i),
}

EXPORTS = tuple(_SIGS)

for _name, (_args, _res) in _SIGS.items():
    _fn = getattr(lib, _name)          # AttributeError here == header/library mismatch
    _fn.argtypes = _args
    _fn.restype = _res


def last_error() -> str:
    buf = C.create_string_buffer(512)
    lib.dmf_last_error(buf, 512)
    return buf.value.decode(errors="replace")


def check(rc: int) -> None:
    if rc != 0:
        raise DmfError(f"libdmf_b200 error {rc}: {last_error()}")


_device_ok = False


def require_device() -> None:
    """Fail loudly unless a B200 (sm_100) is the current device."""
    global _device_ok
    if _device_ok:
        return
    if not torch.cuda.is_available():
        raise DmfError("disentagled_multimodal_fusion_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    check(lib.dmf_device_check())
    _device_ok = True


def stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def ptr(t) -> int:
    return 0 if t is None else t.data_ptr()


def launch_count() -> int:
    return int(lib.dmf_launch_count())


def ptr_array(tensors):
    arr = (c_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = ptr(t)
    return arr
