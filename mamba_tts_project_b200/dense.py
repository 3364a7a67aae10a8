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

import torch

from . import _lib
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


_shadow = {}


def bf16_weight(w: torch.Tensor) -> torch.Tensor:
    """bf16 shadow of an fp32 master weight, refreshed when the parameter's version changes."""
    if w.dtype == torch.bfloat16:
        return w.detach()
    if torch.cuda.is_current_stream_capturing():
        return w.detach().to(torch.bfloat16)           # the cast belongs to the captured step
    key = id(w)
    hit = _shadow.get(key)
    if hit is not None and hit[0] is w and hit[1] == w._version and hit[2] == w.data_ptr():
        return hit[3]
    s = w.detach().to(torch.bfloat16)
    _shadow[key] = (w, w._version, w.data_ptr(), s)
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
    added by the caller's next fused residual + LayerNorm launch)."""

    @staticmethod
    def forward(ctx, h, w1, b1, w2):
        h2 = _rows(h if h.dtype == torch.bfloat16 else h.to(torch.bfloat16))
        w1b, w2b = bf16_weight(w1), bf16_weight(w2)
        pre = torch.empty((h2.shape[0], w1.shape[0]), dtype=torch.bfloat16, device=h.device)
        act = gemm(h2, w1b, bias_n=_f32(b1), epilogue="gelu", aux=pre)
        f = gemm(act, w2b)
        ctx.save_for_backward(h2, pre, act, w1b, w2b)
        ctx.meta = (h.shape, h.dtype, None if b1 is None else b1.dtype)
        return f.view(*h.shape[:-1], w2.shape[0])

    @staticmethod
    def backward(ctx, df):
        h2, pre, act, w1b, w2b = ctx.saved_tensors
        shape, t_h, t_b = ctx.meta
        df2 = _rows(df if df.dtype == torch.bfloat16 else df.to(torch.bfloat16))
        dpre = gemm(df2, w2b.t(), epilogue="gelu_bwd", aux=pre)        # (tokens, d_ff)
        dw2 = gemm(df2.t(), act.t(), out_dtype=torch.float32, split_k=-1)
        dw1 = gemm(dpre.t(), h2.t(), out_dtype=torch.float32, split_k=-1)
        db1 = None if t_b is None else _colsum(dpre).to(t_b)
        dh = gemm(dpre, w1b.t()).view(shape).to(t_h) if ctx.needs_input_grad[0] else None
        return dh, dw1, db1, dw2


def ffn(h, w1, b1, w2):
    return _FfnTC.apply(h, w1, b1, w2)


class _CrossAttentionTC(torch.autograd.Function):
    """``nn.MultiheadAttention(batch_first=True)`` forward without its out-projection bias
    (``mamba_decoder.py:72-77``; packed ``in_proj_weight`` (3E, E) = [Wq; Wk; Wv]), T_kv <= 256.
    q is scaled in the softmax epilogue (fp32), which for power-of-two head sizes is bit-identical to torch's
    scaling of q before QK^T."""

    @staticmethod
    def forward(ctx, query, memory, w_in, b_in, w_out, mask, heads):
        B, T, E = query.shape
        Tk = memory.shape[1]
        H, dh = heads, E // heads
        bf = torch.bfloat16
        q_in = _rows(query if query.dtype == bf else query.to(bf))
        m_in = _rows(memory if memory.dtype == bf else memory.to(bf))
        wb, wob = bf16_weight(w_in), bf16_weight(w_out)
        b32 = _f32(b_in)
        q = gemm(q_in, wb[:E], bias_n=b32[:E]).view(B, T, E)
        kv = gemm(m_in, wb[E:], bias_n=b32[E:]).view(B, Tk, 2 * E)
        qv = q.view(B, T, H, dh).transpose(1, 2)                       # (B, H, T, dh)
        kvw = kv[..., :E].view(B, Tk, H, dh).transpose(1, 2)           # (B, H, Tk, dh)
        vvw = kv[..., E:].view(B, Tk, H, dh).transpose(1, 2)
        tkp = Tk + (-Tk) % 8
        P = torch.empty((B, H, T, tkp), dtype=bf, device=query.device)
        scale = 1.0 / math.sqrt(dh)
        gemm(qv, kvw, out=P[..., :Tk], epilogue="softmax", mask=mask, scale=scale)
        o = torch.empty((B, T, E), dtype=bf, device=query.device)
        gemm(P[..., :Tk], vvw.transpose(-1, -2), out=o.view(B, T, H, dh).transpose(1, 2))
        out = gemm(o.view(B * T, E), wob).view(B, T, E)
        ctx.save_for_backward(q_in, m_in, wb, wob, q, kv, P, o)
        ctx.meta = (query.dtype, memory.dtype, b_in.dtype, H, scale, Tk)
        return out

    @staticmethod
    def backward(ctx, dout):
        q_in, m_in, wb, wob, q, kv, P, o = ctx.saved_tensors
        t_q, t_m, t_b, H, scale, Tk = ctx.meta
        B, T, E = q.shape
        dh = E // H
        bf = torch.bfloat16
        f32 = torch.float32
        d2 = _rows(dout if dout.dtype == bf else dout.to(bf))
        do = gemm(d2, wob.t()).view(B, T, E)
        dwo = gemm(d2.t(), o.view(B * T, E).t(), out_dtype=f32, split_k=-1)
        dov = do.view(B, T, H, dh).transpose(1, 2)                     # (B, H, T, dh)
        qv = q.view(B, T, H, dh).transpose(1, 2)
        kvw = kv[..., :E].view(B, Tk, H, dh).transpose(1, 2)
        vvw = kv[..., E:].view(B, Tk, H, dh).transpose(1, 2)
        Pv = P[..., :Tk]
        dkv = torch.empty_like(kv)
        dkw = dkv[..., :E].view(B, Tk, H, dh).transpose(1, 2)
        dvw = dkv[..., E:].view(B, Tk, H, dh).transpose(1, 2)
        gemm(Pv.transpose(-1, -2), dov.transpose(-1, -2), out=dvw)                      # dV = P^T dO
        dS = torch.empty_like(P)
        gemm(dov, vvw, out=dS[..., :Tk], epilogue="dsoftmax", aux=Pv, scale=scale)      # dS from dP = dO V^T
        dq = torch.empty_like(q)
        gemm(dS[..., :Tk], kvw.transpose(-1, -2), out=dq.view(B, T, H, dh).transpose(1, 2))   # dQ = dS K
        gemm(dS[..., :Tk].transpose(-1, -2), qv.transpose(-1, -2), out=dkw)             # dK = dS^T Q
        dq2, dkv2 = dq.view(B * T, E), dkv.view(B * Tk, 2 * E)
        dw_in = torch.empty((3 * E, E), dtype=f32, device=q.device)
        gemm(dq2.t(), q_in.t(), out=dw_in[:E], split_k=-1)
        gemm(dkv2.t(), m_in.t(), out=dw_in[E:], split_k=-1)
        db_in = torch.cat([_colsum(dq2), _colsum(dkv2)]).to(t_b)
        dquery = gemm(dq2, wb[:E].t()).view(B, T, E).to(t_q) if ctx.needs_input_grad[0] else None
        dmem = gemm(dkv2, wb[E:].t()).view(B, Tk, E).to(t_m) if ctx.needs_input_grad[1] else None
        return dquery, dmem, dw_in, db_in, dwo, None, None


def cross_attention(query, memory, w_in, b_in, w_out, mask, heads):
    """mask: (B, T_kv) bool / uint8, True = attend, or None."""
    m8 = None
    if mask is not None:
        m8 = (mask if mask.dtype == torch.uint8 else mask.to(torch.uint8)).contiguous()
    return _CrossAttentionTC.apply(query, memory, w_in, b_in, w_out, m8, heads)
