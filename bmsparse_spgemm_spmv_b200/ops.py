"""The two operators of the reference, same names and argument order, over the C ABI.

  bmSparse_SpMV(A, v, u, batched)                    src/bmSparse_SPMV.cu:191-230
  bmSparse_mult(A, B, C, mode, VERBOSE, tc_version)  src/bmSparse_SPGEMM.cu:827-1223
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib as L
from .matrix import bmSpMatrix, _dt, _stream_ptr


def bmSparse_SpMV(A: bmSpMatrix, v: torch.Tensor, u: torch.Tensor | None = None, batched: bool = False, stream=None):
    """u = A v.  v: CUDA fp32 or fp16 [num_cols]; u: CUDA fp32 [num_rows] (allocated when None).
    `batched` selected the reference's second kernel (broken as shipped); accepted, ignored."""
    view = A._view()
    if not v.is_cuda or v.numel() != view.num_cols:
        raise ValueError("v must be a CUDA tensor with num_cols elements")
    v = v.contiguous()
    if u is None:
        u = torch.empty(view.num_rows, dtype=torch.float32, device=v.device)
    if not u.is_cuda or u.dtype != torch.float32 or u.numel() != view.num_rows or not u.is_contiguous():
        raise ValueError("u must be a contiguous CUDA fp32 tensor with num_rows elements")
    L.check(L.lib().bmsp_spmv(A._h, C.c_void_p(v.data_ptr()), _dt(v.dtype), C.c_void_p(u.data_ptr()), _stream_ptr(stream)))
    return u


def bmSparse_SpMV_host(A: bmSpMatrix, v: torch.Tensor, u: torch.Tensor | None = None, stream=None):
    """u = A v with v and u in HOST memory (pin them for full PCIe speed): what the reference's driver does by hand
    around the operator (cudaMemcpy v, bmSparse_SpMV, cudaMemcpy u -- SPMV.cu:276-309), as one pipelined call.
    u is complete after the stream (default: the current one) is synchronised."""
    view = A._view()
    if v.is_cuda or v.numel() != view.num_cols or not v.is_contiguous():
        raise ValueError("v must be a contiguous host tensor with num_cols elements")
    if u is None:
        u = torch.empty(view.num_rows, dtype=torch.float32).pin_memory()
    if u.is_cuda or u.dtype != torch.float32 or u.numel() != view.num_rows or not u.is_contiguous():
        raise ValueError("u must be a contiguous host fp32 tensor with num_rows elements")
    L.check(L.lib().bmsp_spmv_host(A._h, C.c_void_p(v.data_ptr()), _dt(v.dtype), C.c_void_p(u.data_ptr()), _stream_ptr(stream)))
    return u


def bmSparse_mult(A: bmSpMatrix, B: bmSpMatrix, C_out: bmSpMatrix | None = None, mode=0, VERBOSE: bool = False,
                  tc_version: int = 5, numeric_path: int = -1, brow_range=None, stream=None):
    """C = A * B with B in transposed-operand form (built with transpose=True, SPGEMM.cu:1262).
    Returns (C, info).  `mode` / `tc_version` are accepted for drop-in compatibility and ignored."""
    opts = L.SpgemmOpts(int(mode), int(tc_version), int(bool(VERBOSE)), int(numeric_path),
                        brow_range[0] if brow_range is not None else 0, brow_range[1] if brow_range is not None else 0,
                        int(brow_range is not None))
    info = L.SpgemmInfo()
    out = C_out if C_out is not None else bmSpMatrix()
    if out._h:
        L.lib().bmsp_destroy(out._h)
        out._h = C.c_void_p()
    L.check(L.lib().bmsp_spgemm(A._h, B._h, C.byref(opts), _stream_ptr(stream), C.byref(out._h), C.byref(info)))
    return out, info
