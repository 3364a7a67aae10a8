"""Python face of ``mtts_gemm`` (``csrc/gemm_sm100.cu``): the tcgen05 / TMEM / TMA contraction every dense op of
the bf16 teacher-forced path runs on (``mamba_decoder.py:29,39-43,61,72-77,88,118`` and their backward).

    gemm(a, b)  ->  C[..., m, n] = sum_k a[..., m, k] * b[..., n, k]

Operands are *views*: the kernel reads any bf16 tensor whose last two dimensions have one unit stride --
``(.., m, k)`` with ``stride(-1) == 1`` is K-major, with ``stride(-2) == 1`` it is MN-major (e.g.
``y.transpose(1, 2)`` of a channel-major activation, ``w.t()`` of a weight) -- and up to two leading batch
dimensions with arbitrary (or zero = broadcast) strides, e.g. the ``(B, H, T, dh)`` view of a ``(B, T, E)``
projection.  Nothing is copied or transposed.  CUDA only; no fallback.
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import ptr

_EPI = {"store": _lib.EPI_STORE, "gelu": _lib.EPI_GELU, "gelu_bwd": _lib.EPI_GELU_BWD,
        "softmax": _lib.EPI_SOFTMAX, "dsoftmax": _lib.EPI_DSOFTMAX, "mul_aux": _lib.EPI_MUL_AUX}


def _as4(t):
    if t.dim() < 2 or t.dim() > 4:
        raise RuntimeError("gemm operands are (..., rows, k) views with at most two batch dimensions")
    if t.dim() == 3:          # a single batch dimension is the OUTER one
        t = t.unsqueeze(1)
    while t.dim() < 4:
        t = t.unsqueeze(0)
    return t


def _operand(t, name):
    """-> (major, ld, bo_stride, bi_stride) of a 4-D view; batch dimensions of extent 1 count as broadcast."""
    if t.dtype != torch.bfloat16:
        raise RuntimeError(f"gemm operand {name} must be bf16, got {t.dtype}")
    rows, k = t.shape[-2], t.shape[-1]
    s_r, s_k = t.stride(-2), t.stride(-1)
    if s_k == 1 or k == 1:
        major, ld = 0, (s_r if rows > 1 else max(s_r, k))
    elif s_r == 1 or rows == 1:
        major, ld = 1, (s_k if k > 1 else max(s_k, rows))
    else:
        raise RuntimeError(f"gemm operand {name}: one of the last two dimensions must have unit stride "
                           f"(shape {tuple(t.shape)}, strides {t.stride()})")
    bo = t.stride(0) if t.shape[0] > 1 else 0
    bi = t.stride(1) if t.shape[1] > 1 else 0
    if ld % 8 or bo % 8 or bi % 8 or t.data_ptr() % 16:
        raise RuntimeError(f"gemm operand {name}: leading dimension / batch strides must be multiples of 8 "
                           f"elements and the base 16-byte aligned (ld {ld}, batch strides {bo}, {bi})")
    return major, ld, bo, bi


def gemm(a, b, out=None, out_dtype=None, bias_n=None, bias_m=None, epilogue="store", aux=None, mask=None,
         scale=1.0, accumulate=False, reduce_batch=False, split_k=1, single_cta=False, aux_gelu_grad=False,
         _debug=0):
    """C = epilogue(a @ b^T) on the tensor cores.

    a (.., m, k), b (.., n, k): bf16 views (see module docstring); leading dimensions broadcast against each other.
    out: optional (.., m, n) tensor with unit stride along n (bf16 or fp32), else allocated in ``out_dtype``.
    bias_n (n) / bias_m (m): fp32, added before the activation.
    epilogue: "store" | "gelu" (aux = optional bf16 output receiving the pre-activation, or gelu'(pre) with
      ``aux_gelu_grad``) | "gelu_bwd" (aux = the forward's pre-activation) | "mul_aux" (C = acc * aux) | "softmax" (row softmax of scale * acc under the key ``mask`` (bo, n) uint8,
      n <= 256) | "dsoftmax" (aux = P; scale * P o (acc - rowsum(P o acc))).
    accumulate: C += result.  reduce_batch: the contraction also runs over the outermost batch dimension (weight
      gradients), C is (m, n); split_k = -1 lets the library split it across the SMs (fp32 out).
    single_cta: keep to 128-row tiles of one CTA (cta_group::1) instead of 256-row tiles of a CTA pair."""
    _lib.require_cuda(a, b, out, bias_n, bias_m, aux, mask)
    a4, b4 = _as4(a), _as4(b)
    m, k = a4.shape[-2:]
    n, k2 = b4.shape[-2:]
    if k != k2:
        raise RuntimeError(f"gemm: contraction sizes differ ({k} vs {k2})")
    nbo, nbi = max(a4.shape[0], b4.shape[0]), max(a4.shape[1], b4.shape[1])
    for t in (a4, b4):
        if t.shape[0] not in (1, nbo) or t.shape[1] not in (1, nbi):
            raise RuntimeError("gemm: batch dimensions do not broadcast")
    a_major, lda, a_bo, a_bi = _operand(a4, "a")
    b_major, ldb, b_bo, b_bi = _operand(b4, "b")
    k_batches = 1
    if reduce_batch:
        if nbi != 1:
            raise RuntimeError("gemm: reduce_batch contracts over a single batch dimension")
        k_batches, out_bo = nbo, 1
    else:
        out_bo = nbo
    if out is None:
        dt = out_dtype or torch.bfloat16
        shape = (m, n) if reduce_batch else tuple(torch.broadcast_shapes(a.shape[:-2], b.shape[:-2])) + (m, n)
        out = torch.empty(shape, dtype=dt, device=a.device)
    o4 = _as4(out)
    if o4.shape[-2:] != (m, n) or (o4.stride(-1) != 1 and n > 1) or o4.shape[0] != out_bo or o4.shape[1] != nbi:
        raise RuntimeError(f"gemm: out must be (.., {m}, {n}) with unit stride along n, got {tuple(out.shape)}")
    if out.dtype not in (torch.bfloat16, torch.float32):
        raise RuntimeError("gemm: out must be bf16 or fp32")
    aux4 = None
    if aux is not None:
        aux4 = _as4(aux)
        if aux4.shape != o4.shape or aux4.dtype != torch.bfloat16 or (aux4.stride(-1) != 1 and n > 1):
            raise RuntimeError("gemm: aux must be a bf16 tensor shaped like out")
    if bias_n is not None and (bias_n.dtype != torch.float32 or bias_n.shape != (n,) or not bias_n.is_contiguous()):
        raise RuntimeError("gemm: bias_n must be contiguous fp32 (n)")
    if bias_m is not None and (bias_m.dtype != torch.float32 or bias_m.shape != (m,) or not bias_m.is_contiguous()):
        raise RuntimeError("gemm: bias_m must be contiguous fp32 (m)")
    m8 = None
    if mask is not None:
        m8 = mask if mask.dtype == torch.uint8 else mask.to(torch.uint8)
        if m8.shape != (nbo, n) or not m8.is_contiguous():
            m8 = m8.reshape(nbo, n).contiguous()
    if m == 0 or n == 0 or out.numel() == 0:
        return out
    if k == 0:
        raise RuntimeError("gemm: empty contraction")
    bs = lambda t, d: t.stride(d) if t.shape[d] > 1 else 0
    p = _lib.GemmParams(
        m=m, n=n, k=k, batch_outer=out_bo, batch_inner=nbi, k_batches=k_batches, a_major=a_major, b_major=b_major,
        out_dtype=_lib.io_dtype(out), epilogue=_EPI[epilogue], accumulate=int(bool(accumulate)), split_k=int(split_k),
        a=ptr(a4), lda=lda, a_bo_stride=a_bo, a_bi_stride=a_bi,
        b=ptr(b4), ldb=ldb, b_bo_stride=b_bo, b_bi_stride=b_bi,
        out=ptr(o4), ldc=o4.stride(-2) if m > 1 else max(o4.stride(-2), n), c_bo_stride=bs(o4, 0), c_bi_stride=bs(o4, 1),
        bias_n=ptr(bias_n), bias_m=ptr(bias_m),
        aux=ptr(aux4), ld_aux=0 if aux4 is None else (aux4.stride(-2) if m > 1 else max(aux4.stride(-2), n)),
        aux_bo_stride=0 if aux4 is None else bs(aux4, 0), aux_bi_stride=0 if aux4 is None else bs(aux4, 1),
        mask=ptr(m8), mask_bo_stride=0 if m8 is None else n, scale=float(scale),
        flags=(_lib.GEMM_SINGLE_CTA if single_cta else 0) | (_lib.GEMM_AUX_GELU_GRAD if aux_gelu_grad else 0)
        | (_debug << 8))
    _lib.call("mtts_gemm", p)
    return out
