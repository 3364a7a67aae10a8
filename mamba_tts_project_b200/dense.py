"""Dense contractions of the bf16 teacher-forced path as autograd nodes over ``mtts_gemm`` (tcgen05 / TMEM /
TMA, ``csrc/gemm_sm100.cu``): projections (``mamba_decoder.py:29,61,118``), the FiLM'd FFN (``:39-43,88``) and
the cross-attention to ``[ref || text]`` (``:32-36,72-77``) -- forward, data gradients and weight gradients.

* weights: fp32 masters; a bf16 *shadow* per parameter is made once per parameter version (i.e. once per
  optimizer step, not once per use) -- except while a CUDA graph is being captured, where the cast must be
  part of the graph;
* weight gradients come out of the GEMM in fp32 (split-k over the SMs, vector REDs): no ``bmm + sum(0)``, no cast;
* FFN: bias + GELU ride in the first GEMM's epilogue (the pre-activation is its second output), GELU' in the
  epilogue of the data-gradient GEMM;
* attention: QK^T with scale + key mask + softmax in the epilogue (a 256-column tile holds a whole row of
  ``[ref || text]`` scores), PV, and in the backward dP = dO V^T with the softmax backward in the epilogue.
  K/V are projected once per layer and forward.  Longer memories (> 256 keys) take the library SDPA path.

fp32 (the 1e-4 parity / debug precision) stays on the library GEMMs: the tensor-core path is bf16.
"""
from __future__ import annotations

import math
import os
import weakref

import torch

from . import _lib
from ._lib import ptr
from .gemm import gemm

_BACKEND = "tc"          # "tc": mtts_gemm for bf16;  "library": torch / cuBLASLt everywhere (A/B measurements)
MAX_FUSED_KEYS = 256


def set_backend(name: str) -> None:
    global _BACKEND
    if name not in ("tc", "library"):
        raise ValueError("backend must be 'tc' or 'library'")
    _BACKEND = name


def tc_enabled(dtype) -> bool:
    return _BACKEND == "tc" and dtype == torch.bfloat16


_shadow = {}          # id(parameter) -> (weakref to it, version, data_ptr, bf16 copy); entries die with the parameter


def bf16_weight(w: torch.Tensor) -> torch.Tensor:
    """bf16 shadow of an fp32 master weight, refreshed when the parameter's version changes."""
    if w.dtype == torch.bfloat16:
        return w.detach()
    if torch.cuda.is_current_stream_capturing() or not w.is_leaf:
        return w.detach().to(torch.bfloat16)           # the cast belongs to the captured step / a temporary
    key = id(w)
    hit = _shadow.get(key)
    if hit is not None and hit[0]() is w and hit[1] == w._version and hit[2] == w.data_ptr():
        return hit[3]
    s = w.detach().to(torch.bfloat16)
    if hit is None or hit[0]() is not w:
        weakref.finalize(w, _shadow.pop, key, None)    # no strong reference: the model can be freed
    _shadow[key] = (weakref.ref(w), w._version, w.data_ptr(), s)
    return s


def _colsum(x2):
    out = torch.zeros(x2.shape[1], dtype=torch.float32, device=x2.device)
    if x2.numel():
        _lib.call("mtts_colsum", _lib.BiasGeluParams(rows=x2.shape[0], cols=x2.shape[1], io_dtype=_lib.io_dtype(x2),
                                                     reserved=0, ld=x2.stride(0), x=x2.data_ptr(), bias=None,
                                                     dout=None, out=None, colsum=out.data_ptr()))
    return out


def _f32(t):
    return None if t is None else t.detach().float().contiguous()


def _rows(t):
    t2 = t.reshape(-1, t.shape[-1])
    return t2 if t2.is_contiguous() else t2.contiguous()


class _LinearTC(torch.autograd.Function):
    """out = x @ W^T (+ bias): x (.., K) bf16, W (N, K) fp32 master."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        x2 = _rows(x if x.dtype == torch.bfloat16 else x.to(torch.bfloat16))
        wb = bf16_weight(weight)
        out = gemm(x2, wb, bias_n=_f32(bias))
        ctx.save_for_backward(x2, wb)
        ctx.meta = (x.shape, x.dtype, None if bias is None else bias.dtype)
        return out.view(*x.shape[:-1], weight.shape[0])

    @staticmethod
    def backward(ctx, dy):
        x2, wb = ctx.saved_tensors
        shape, t_x, t_b = ctx.meta
        dy2 = _rows(dy if dy.dtype == torch.bfloat16 else dy.to(torch.bfloat16))
        need_x, need_w, need_b = ctx.needs_input_grad
        dx = dw = db = None
        if need_x:
            dx = gemm(dy2, wb.t()).view(shape).to(t_x)
        if need_w:
            dw = gemm(dy2.t(), x2.t(), out_dtype=torch.float32, split_k=-1)
        if need_b and t_b is not None:
            db = _colsum(dy2).to(t_b)
        return dx, dw, db


def linear(x, weight, bias=None):
    return _LinearTC.apply(x, weight, bias)


class _FfnTC(torch.autograd.Function):
    """f = W2 gelu(W1 h + b1)  (``ff[0]``, ``nn.GELU()``, ``ff[2]`` of ``mamba_decoder.py:39-43``; ``ff[2].bias`` is
    added by the caller's next fused residual + LayerNorm launch).  One library call per direction
    (``mtts_film_ffn_fwd`` / ``_bwd``): the forward leaves gelu'(pre) -- from the same tanh as gelu(pre) -- so the
    backward's data-gradient GEMM multiplies by it in its epilogue."""

    @staticmethod
    def forward(ctx, h, w1, b1, w2):
        h2 = _rows(h if h.dtype == torch.bfloat16 else h.to(torch.bfloat16))
        w1b, w2b = bf16_weight(w1), bf16_weight(w2)
        T, D, Fd = h2.shape[0], w1.shape[1], w1.shape[0]
        bf, dev = torch.bfloat16, h.device
        act = torch.empty((T, Fd), dtype=bf, device=dev)
        gprime = torch.empty((T, Fd), dtype=bf, device=dev)
        f = torch.empty((T, D), dtype=bf, device=dev)
        b32 = _f32(b1)
        p = _lib.FilmFfnParams(tokens=T, d_model=D, d_ff=Fd, h=ptr(h2), w1=ptr(w1b), b1=ptr(b32), w2=ptr(w2b),
                               act=ptr(act), gprime=ptr(gprime), f=ptr(f))
        _lib.require_cuda(h2, w1b, w2b, b32)
        _lib.call("mtts_film_ffn_fwd", p, launches=2)
        ctx.save_for_backward(h2, gprime, act, w1b, w2b)
        ctx.meta = (h.shape, h.dtype, None if b1 is None else b1.dtype)
        return f.view(*h.shape[:-1], D)

    @staticmethod
    def backward(ctx, df):
        h2, gprime, act, w1b, w2b = ctx.saved_tensors
        shape, t_h, t_b = ctx.meta
        df2 = _rows(df if df.dtype == torch.bfloat16 else df.to(torch.bfloat16))
        T, D, Fd = h2.shape[0], h2.shape[1], act.shape[1]
        dev, f32 = df.device, torch.float32
        dpre = torch.empty_like(act)
        dw1 = torch.empty((Fd, D), dtype=f32, device=dev)
        dw2 = torch.empty((D, Fd), dtype=f32, device=dev)
        db1 = None if t_b is None else torch.zeros(Fd, dtype=f32, device=dev)
        need_h = ctx.needs_input_grad[0]
        dh = torch.empty((T, D), dtype=torch.bfloat16, device=dev) if need_h else None
        p = _lib.FilmFfnParams(tokens=T, d_model=D, d_ff=Fd, h=ptr(h2), w1=ptr(w1b), w2=ptr(w2b), act=ptr(act), gprime=ptr(gprime), df=ptr(df2), dpre=ptr(dpre),
                               dw1=ptr(dw1), db1=ptr(db1), dw2=ptr(dw2), dh=ptr(dh))
        _lib.call("mtts_film_ffn_bwd", p, launches=4 + int(need_h))
        return (dh.view(shape).to(t_h) if need_h else None), dw1, (None if db1 is None else db1.to(t_b)), dw2


def ffn(h, w1, b1, w2):
    return _FfnTC.apply(h, w1, b1, w2)


FUSED_ATTENTION_BACKWARD = os.environ.get("MTTS_ATTN_FUSED_BWD", "1") != "0"     # A/B switch (measurements)


class _CrossAttentionTC(torch.autograd.Function):
    """``nn.MultiheadAttention(batch_first=True)`` forward without its out-projection bias
    (``mamba_decoder.py:72-77``; packed ``in_proj_weight`` (3E, E) = [Wq; Wk; Wv]), T_kv <= 256, as one library call
    per direction (``mtts_cross_attn_fwd`` / ``_bwd``).  q is scaled in the softmax epilogue (fp32), which for
    power-of-two head sizes is bit-identical to torch's scaling of q before QK^T.  With 64-wide heads the forward
    keeps only the per-row log-sum-exp (not the (B, H, T, T_kv) probabilities) and the backward recomputes P inside
    ``mtts_attn_core_bwd`` (scores, probabilities and their gradients never leave the SM)."""

    @staticmethod
    def forward(ctx, query, memory, w_in, b_in, w_out, mask, heads):
        B, T, E = query.shape
        Tk = memory.shape[1]
        bf, dev = torch.bfloat16, query.device
        q_in = _rows(query if query.dtype == bf else query.to(bf))
        m_in = _rows(memory if memory.dtype == bf else memory.to(bf))
        wb, wob = bf16_weight(w_in), bf16_weight(w_out)
        b32 = _f32(b_in)
        tkp = Tk + (-Tk) % 8
        fused = FUSED_ATTENTION_BACKWARD and E == heads * 64
        q = torch.empty((B, T, E), dtype=bf, device=dev)
        kv = torch.empty((B, Tk, 2 * E), dtype=bf, device=dev)
        P = None if fused else torch.empty((B, heads, T, tkp), dtype=bf, device=dev)
        o = torch.empty((B, T, E), dtype=bf, device=dev)
        out = torch.empty((B, T, E), dtype=bf, device=dev)
        lse2 = torch.empty((B, heads, T), dtype=torch.float32, device=dev) if fused else None
        _lib.require_cuda(q_in, m_in, wb, wob, b32, mask)
        p = _lib.CrossAttnParams(batch=B, t_q=T, t_kv=Tk, d_model=E, heads=heads, query=ptr(q_in), memory=ptr(m_in),
                                 w_in=ptr(wb), b_in=ptr(b32), w_out=ptr(wob), mask=ptr(mask), q=ptr(q), kv=ptr(kv),
                                 p=ptr(P), o=ptr(o), out=ptr(out), lse2=ptr(lse2))
        _lib.call("mtts_cross_attn_fwd", p, launches=4 if fused else 5)
        if fused:
            ctx.save_for_backward(q_in, m_in, wb, wob, q, kv, lse2, o, mask)     # P dies here
        else:
            ctx.save_for_backward(q_in, m_in, wb, wob, q, kv, P, o, mask)
        ctx.meta = (query.dtype, memory.dtype, b_in.dtype, heads, fused)
        return out

    @staticmethod
    def backward(ctx, dout):
        q_in, m_in, wb, wob, q, kv, P_or_lse, o, mask = ctx.saved_tensors
        t_q, t_m, t_b, H, fused = ctx.meta
        B, T, E = q.shape
        Tk = kv.shape[1]
        bf, f32, dev = torch.bfloat16, torch.float32, q.device
        d2 = _rows(dout if dout.dtype == bf else dout.to(bf))
        d_o, dq, dkv = torch.empty_like(o), torch.empty_like(q), torch.empty_like(kv)
        dS = None if fused else torch.empty_like(P_or_lse)
        dw_in = torch.empty((3 * E, E), dtype=f32, device=dev)
        dw_out = torch.empty((E, E), dtype=f32, device=dev)
        db_in = torch.zeros(3 * E, dtype=f32, device=dev)
        need_q, need_m = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        dquery = torch.empty((B, T, E), dtype=bf, device=dev) if need_q else None
        dmem = torch.empty((B, Tk, E), dtype=bf, device=dev) if need_m else None
        p = _lib.CrossAttnParams(batch=B, t_q=T, t_kv=Tk, d_model=E, heads=H, query=ptr(q_in), memory=ptr(m_in),
                                 w_in=ptr(wb), b_in=ptr(db_in), w_out=ptr(wob), mask=ptr(mask), q=ptr(q), kv=ptr(kv),
                                 p=None if fused else ptr(P_or_lse), o=ptr(o), lse2=ptr(P_or_lse) if fused else None,
                                 dout=ptr(d2), d_o=ptr(d_o), ds=ptr(dS), dq=ptr(dq), dkv=ptr(dkv),
                                 dw_in=ptr(dw_in), db_in=ptr(db_in), dw_out=ptr(dw_out), dquery=ptr(dquery),
                                 dmemory=ptr(dmem))
        _lib.call("mtts_cross_attn_bwd", p, launches=(7 if fused else 10) + int(need_q) + int(need_m))
        return (None if dquery is None else dquery.to(t_q), None if dmem is None else dmem.to(t_m), dw_in,
                db_in.to(t_b), dw_out, None, None)


def cross_attention(query, memory, w_in, b_in, w_out, mask, heads):
    """mask: (B, T_kv) bool / uint8, True = attend, or None."""
    m8 = None
    if mask is not None:
        m8 = (mask if mask.dtype == torch.uint8 else mask.to(torch.uint8)).contiguous()
    return _CrossAttentionTC.apply(query, memory, w_in, b_in, w_out, m8, heads)
