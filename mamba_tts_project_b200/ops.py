"""Operator surface of the hot path -- same names, argument meaning and error behaviour as the
third-party ops the reference reaches through ``Mamba(d_model)`` (``mamba_decoder.py:4,29,61,63``):

    selective_scan_fn        <- mamba_ssm.ops.selective_scan_interface.selective_scan_fn
    selective_state_update   <- mamba_ssm.ops.triton.selective_state_update.selective_state_update
    causal_conv1d_fn         <- causal_conv1d.causal_conv1d_fn
    causal_conv1d_update     <- causal_conv1d.causal_conv1d_update
    mamba_inner_fn           <- mamba_ssm.ops.selective_scan_interface.mamba_inner_fn

Every op is a thin ``torch.autograd.Function`` (or plain function) that fills a POD struct and calls
the sm_100a library through the C ABI (``_lib.call``).  CUDA tensors only; nothing here computes on
the host, and there is no fallback.  Variants upstream offers that the decoder never uses (complex A,
constant / grouped 4-D B and C, ``seq_idx``) raise ``NotImplementedError``.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import _lib
from ._lib import ptr

__all__ = ["selective_scan_fn", "selective_state_update", "causal_conv1d_fn",
           "causal_conv1d_update", "mamba_inner_fn", "mamba_decode_step", "cross_attn_decode",
           "add_layernorm", "skinny_linear", "gemm_bf16", "bias_gelu", "colsum", "linear",
           "cross_attn_block_decode", "cross_attn_block_decode_supported", "decode_embed", "decode_greedy", "length_regulate"]


def _unit_last_stride(t):
    return t if t is None or t.stride(-1) == 1 else t.contiguous()


def _f32c(t):
    return None if t is None else t.detach().to(torch.float32).contiguous()


# ------------------------------------------------------------------------------------------------
# causal_conv1d
# ------------------------------------------------------------------------------------------------
def _conv_args(x, weight, activation):
    if activation not in (None, "silu", "swish"):
        raise NotImplementedError("activation must be None, silu, or swish")
    if x.dim() != 3:
        raise RuntimeError("x must be (batch, dim, seqlen)")
    dim, width = weight.shape
    if x.shape[1] != dim:
        raise RuntimeError("weight must be (dim, width) matching x's dim")
    if not 2 <= width <= _lib.MAX_CONV_WIDTH:
        raise RuntimeError("causal_conv1d only supports width between 2 and 4")


class _CausalConv1dFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, initial_states, activation):
        _lib.require_cuda(x, weight, bias, initial_states)
        _conv_args(x, weight, activation)
        x = _unit_last_stride(x)
        w32, b32 = _f32c(weight), _f32c(bias)
        init = None
        if initial_states is not None:
            init = _unit_last_stride(initial_states.detach().to(x.dtype))
            if init.shape != (x.shape[0], x.shape[1], weight.shape[1] - 1):
                raise RuntimeError("initial_states must be (batch, dim, width - 1)")
        out = torch.empty_like(x, memory_format=torch.contiguous_format)
        B, Dm, L = x.shape
        ctx.save_for_backward(x, w32, b32, init)
        ctx.activation = activation
        ctx.wdtype = weight.dtype
        ctx.bdtype = None if bias is None else bias.dtype
        if x.numel() == 0:  # empty input: nothing to launch (data_ptr() would be NULL)
            return out
        p = _lib.Conv1dFwdParams(
            batch=B, dim=Dm, seqlen=L, width=weight.shape[1], io_dtype=_lib.io_dtype(x),
            silu=int(activation is not None), x=ptr(x), x_batch_stride=x.stride(0),
            x_dim_stride=x.stride(1), weight=ptr(w32), bias=ptr(b32), initial_states=ptr(init),
            init_batch_stride=0 if init is None else init.stride(0),
            init_dim_stride=0 if init is None else init.stride(1),
            out=ptr(out), out_batch_stride=out.stride(0), out_dim_stride=out.stride(1))
        _lib.call("mtts_causal_conv1d_fwd", p)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, w32, b32, init = ctx.saved_tensors
        dout = _unit_last_stride(dout.to(x.dtype))
        B, Dm, L = x.shape
        dx = torch.empty_like(x, memory_format=torch.contiguous_format)
        dw = torch.zeros_like(w32)
        db = torch.zeros(Dm, dtype=torch.float32, device=x.device)
        if x.numel() == 0:
            return (dx, dw.to(ctx.wdtype), None if ctx.bdtype is None else db.to(ctx.bdtype),
                    None, None)
        p = _lib.Conv1dBwdParams(
            batch=B, dim=Dm, seqlen=L, width=w32.shape[1], io_dtype=_lib.io_dtype(x),
            silu=int(ctx.activation is not None), x=ptr(x), x_batch_stride=x.stride(0),
            x_dim_stride=x.stride(1), weight=ptr(w32), bias=ptr(b32), initial_states=ptr(init),
            init_batch_stride=0 if init is None else init.stride(0),
            init_dim_stride=0 if init is None else init.stride(1),
            dout=ptr(dout), dout_batch_stride=dout.stride(0), dout_dim_stride=dout.stride(1),
            dx=ptr(dx), dx_batch_stride=dx.stride(0), dx_dim_stride=dx.stride(1),
            dweight=ptr(dw), dbias=ptr(db))
        _lib.call("mtts_causal_conv1d_bwd", p)
        return (dx, dw.to(ctx.wdtype), None if ctx.bdtype is None else db.to(ctx.bdtype),
                None, None)


def causal_conv1d_fn(x, weight, bias=None, seq_idx=None, initial_states=None,
                     return_final_states=False, final_states_out=None, activation=None):
    """x: (batch, dim, seqlen); weight: (dim, width); bias: (dim,); initial_states:
    (batch, dim, width - 1).  Returns out (batch, dim, seqlen) [, final_states (batch, dim, width-1)]."""
    if seq_idx is not None:
        raise NotImplementedError("seq_idx (packed variable-length batches) is not on this path")
    out = _CausalConv1dFn.apply(x, weight, bias, initial_states, activation)
    if not return_final_states:
        return out
    width = weight.shape[1]
    with torch.no_grad():
        hist = x if initial_states is None else torch.cat([initial_states.to(x.dtype), x], dim=-1)
        final = F.pad(hist, (width - 1 - hist.shape[-1], 0))[..., -(width - 1):] \
            if hist.shape[-1] < width - 1 else hist[..., -(width - 1):]
        if final_states_out is not None:
            final_states_out.copy_(final)
            final = final_states_out
    return out, final


def causal_conv1d_update(x, conv_state, weight, bias=None, activation=None):
    """x: (batch, dim); conv_state: (batch, dim, width) rolled left IN PLACE, x written last
    (``Mamba.step`` convention).  Returns (batch, dim)."""
    if activation not in (None, "silu", "swish"):
        raise NotImplementedError("activation must be None, silu, or swish")
    _lib.require_cuda(x, conv_state, weight, bias)
    B, Dm = x.shape
    width = weight.shape[1]
    if conv_state.shape != (B, Dm, width) or not conv_state.is_contiguous():
        raise RuntimeError("conv_state must be contiguous (batch, dim, width)")
    if conv_state.dtype != x.dtype:
        raise RuntimeError("conv_state and x must have the same dtype")
    x = _unit_last_stride(x)
    out = torch.empty(B, Dm, dtype=x.dtype, device=x.device)
    w32, b32 = _f32c(weight), _f32c(bias)
    p = _lib.Conv1dUpdateParams(
        batch=B, dim=Dm, width=width, io_dtype=_lib.io_dtype(x), silu=int(activation is not None),
        x=ptr(x), x_batch_stride=x.stride(0), conv_state=ptr(conv_state), weight=ptr(w32),
        bias=ptr(b32), out=ptr(out), out_batch_stride=out.stride(0))
    _lib.call("mtts_causal_conv1d_update", p)
    return out


# ------------------------------------------------------------------------------------------------
# selective scan
# ------------------------------------------------------------------------------------------------
def scan_num_chunks(seqlen):
    return (seqlen + _lib.SCAN_CHUNK - 1) // _lib.SCAN_CHUNK


class _SelectiveScanFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, u, delta, A, B, C, D, z, delta_bias, delta_softplus, return_last_state,
                initial_state):
        _lib.require_cuda(u, delta, A, B, C, D, z, delta_bias, initial_state)
        if A.is_complex():
            raise NotImplementedError("complex A is not on this path")
        if B.dim() != 3 or C.dim() != 3:
            raise NotImplementedError("only input-dependent (batch, dstate, seqlen) B and C")
        batch, dim, L = u.shape
        N = A.shape[1]
        if A.shape[0] != dim or delta.shape != u.shape or B.shape != (batch, N, L) or \
                C.shape != (batch, N, L):
            raise RuntimeError("shape mismatch: u/delta (b,d,l), A (d,n), B/C (b,n,l)")
        if N > _lib.MAX_DSTATE:
            raise RuntimeError(f"selective_scan only supports state dimension <= {_lib.MAX_DSTATE}")
        io = u.dtype
        ctx.dtypes = tuple(None if t is None else t.dtype for t in (delta, A, B, C, D, z, delta_bias))
        # from the ORIGINAL arguments: the .to()/.contiguous() copies below do not require grad
        needs_grad = any(ctx.needs_input_grad[:8])
        u = _unit_last_stride(u)
        delta = _unit_last_stride(delta.to(io))
        Bm = _unit_last_stride(B.to(io))
        Cm = _unit_last_stride(C.to(io))
        zz = None if z is None else _unit_last_stride(z.to(io))
        A32, D32, db32 = _f32c(A), _f32c(D), _f32c(delta_bias)
        h0 = _f32c(initial_state)
        if h0 is not None and h0.shape != (batch, dim, N):
            raise RuntimeError("initial_state must be (batch, dim, dstate)")

        out = torch.empty((batch, dim, L), dtype=io, device=u.device)
        last = torch.empty((batch, dim, N), dtype=torch.float32, device=u.device) \
            if return_last_state else None
        chk = torch.empty((batch, dim, max(scan_num_chunks(L), 1), N), dtype=torch.float32,
                          device=u.device) if needs_grad else None
        # y before the gate, for the backward's dz (upstream saves the same tensor next to out_z)
        ypre = torch.empty_like(out) if needs_grad and zz is not None else None
        ctx.delta_softplus = bool(delta_softplus)
        ctx.save_for_backward(u, delta, A32, Bm, Cm, D32, zz, db32, chk, ypre)
        if u.numel() == 0:  # empty input: the state passes through, nothing to launch
            if return_last_state:
                last = h0.clone() if h0 is not None else last.zero_()
                ctx.mark_non_differentiable(last)
                return out, last
            return out
        p = _lib.ScanFwdParams(
            batch=batch, dim=dim, seqlen=L, dstate=N, io_dtype=_lib.io_dtype(u),
            delta_softplus=int(bool(delta_softplus)),
            u=ptr(u), u_batch_stride=u.stride(0), u_dim_stride=u.stride(1),
            delta=ptr(delta), delta_batch_stride=delta.stride(0), delta_dim_stride=delta.stride(1),
            A=ptr(A32),
            B=ptr(Bm), B_batch_stride=Bm.stride(0), B_state_stride=Bm.stride(1),
            C=ptr(Cm), C_batch_stride=Cm.stride(0), C_state_stride=Cm.stride(1),
            D=ptr(D32), delta_bias=ptr(db32),
            z=ptr(zz), z_batch_stride=0 if zz is None else zz.stride(0),
            z_dim_stride=0 if zz is None else zz.stride(1),
            initial_state=ptr(h0),
            out=ptr(out), out_batch_stride=out.stride(0), out_dim_stride=out.stride(1),
            last_state=ptr(last), checkpoints=ptr(chk),
            y_pre=ptr(ypre), y_batch_stride=0 if ypre is None else ypre.stride(0),
            y_dim_stride=0 if ypre is None else ypre.stride(1))
        _lib.call("mtts_selective_scan_fwd", p)
        if return_last_state:
            ctx.mark_non_differentiable(last)
            return out, last
        return out

    @staticmethod
    def backward(ctx, dout, *unused):
        u, delta, A32, Bm, Cm, D32, zz, db32, chk, ypre = ctx.saved_tensors
        batch, dim, L = u.shape
        N = A32.shape[1]
        dev = u.device
        dout = _unit_last_stride(dout.to(u.dtype))
        du = torch.empty((batch, dim, L), dtype=u.dtype, device=dev)
        ddelta = torch.empty((batch, dim, L), dtype=u.dtype, device=dev)
        dz = torch.empty((batch, dim, L), dtype=u.dtype, device=dev) if zz is not None else None
        dA = torch.zeros((dim, N), dtype=torch.float32, device=dev)
        dB = torch.zeros((batch, N, L), dtype=torch.float32, device=dev)
        dC = torch.zeros((batch, N, L), dtype=torch.float32, device=dev)
        dD = torch.zeros(dim, dtype=torch.float32, device=dev) if D32 is not None else None
        ddb = torch.zeros(dim, dtype=torch.float32, device=dev) if db32 is not None else None
        t_delta, t_A, t_B, t_C, t_D, t_z, t_db = ctx.dtypes
        if u.numel() == 0:
            return (du, ddelta.to(t_delta), dA.to(t_A), dB.to(t_B), dC.to(t_C),
                    None if dD is None else dD.to(t_D), None if dz is None else dz.to(t_z),
                    None if ddb is None else ddb.to(t_db), None, None, None)
        p = _lib.ScanBwdParams(
            batch=batch, dim=dim, seqlen=L, dstate=N, io_dtype=_lib.io_dtype(u),
            delta_softplus=int(ctx.delta_softplus),
            u=ptr(u), u_batch_stride=u.stride(0), u_dim_stride=u.stride(1),
            delta=ptr(delta), delta_batch_stride=delta.stride(0), delta_dim_stride=delta.stride(1),
            A=ptr(A32),
            B=ptr(Bm), B_batch_stride=Bm.stride(0), B_state_stride=Bm.stride(1),
            C=ptr(Cm), C_batch_stride=Cm.stride(0), C_state_stride=Cm.stride(1),
            D=ptr(D32), delta_bias=ptr(db32),
            z=ptr(zz), z_batch_stride=0 if zz is None else zz.stride(0),
            z_dim_stride=0 if zz is None else zz.stride(1),
            dout=ptr(dout), dout_batch_stride=dout.stride(0), dout_dim_stride=dout.stride(1),
            checkpoints=ptr(chk),
            du=ptr(du), du_batch_stride=du.stride(0), du_dim_stride=du.stride(1),
            ddelta=ptr(ddelta), ddelta_batch_stride=ddelta.stride(0),
            ddelta_dim_stride=ddelta.stride(1),
            dz=ptr(dz), dz_batch_stride=0 if dz is None else dz.stride(0),
            dz_dim_stride=0 if dz is None else dz.stride(1),
            dA=ptr(dA), dB=ptr(dB), dC=ptr(dC), dD=ptr(dD), ddelta_bias=ptr(ddb),
            y_pre=ptr(ypre), y_batch_stride=0 if ypre is None else ypre.stride(0),
            y_dim_stride=0 if ypre is None else ypre.stride(1))
        _lib.call("mtts_selective_scan_bwd", p)
        return (du, ddelta.to(t_delta), dA.to(t_A), dB.to(t_B), dC.to(t_C),
                None if dD is None else dD.to(t_D), None if dz is None else dz.to(t_z),
                None if ddb is None else ddb.to(t_db), None, None, None)


def selective_scan_fn(u, delta, A, B, C, D=None, z=None, delta_bias=None, delta_softplus=False,
                      return_last_state=False, initial_state=None):
    """Upstream signature (+ ``initial_state``, needed by the reference's stateful ``Mamba`` contract).

    u, delta, z: (batch, dim, seqlen); A: (dim, dstate); B, C: (batch, dstate, seqlen);
    D, delta_bias: (dim,).  Returns out [, last_state (batch, dim, dstate) fp32]."""
    return _SelectiveScanFn.apply(u, delta, A, B, C, D, z, delta_bias, delta_softplus,
                                  return_last_state, initial_state)


def selective_state_update(state, x, dt, A, B, C, D=None, z=None, dt_bias=None, dt_softplus=False):
    """state (batch, dim, dstate) fp32, updated IN PLACE; x, dt, z (batch, dim); A (dim, dstate);
    B, C (batch, dstate); D, dt_bias (dim,).  Returns out (batch, dim)."""
    _lib.require_cuda(state, x, dt, A, B, C, D, z, dt_bias)
    if x.dim() != 2 or B.dim() != 2:
        raise NotImplementedError("only the single-head (batch, dim) layout is on this path")
    batch, dim = x.shape
    N = A.shape[1]
    if state.shape != (batch, dim, N) or state.dtype != torch.float32 or not state.is_contiguous():
        raise RuntimeError("state must be contiguous fp32 (batch, dim, dstate)")
    io = x.dtype
    x, dt = _unit_last_stride(x), _unit_last_stride(dt.to(io))
    Bm, Cm = _unit_last_stride(B.to(io)), _unit_last_stride(C.to(io))
    zz = None if z is None else _unit_last_stride(z.to(io))
    A32, D32, b32 = _f32c(A), _f32c(D), _f32c(dt_bias)
    out = torch.empty(batch, dim, dtype=io, device=x.device)
    p = _lib.StateUpdateParams(
        batch=batch, dim=dim, dstate=N, io_dtype=_lib.io_dtype(x), dt_softplus=int(bool(dt_softplus)),
        state=ptr(state), x=ptr(x), x_batch_stride=x.stride(0), dt=ptr(dt),
        dt_batch_stride=dt.stride(0), A=ptr(A32), B=ptr(Bm), B_batch_stride=Bm.stride(0),
        C=ptr(Cm), C_batch_stride=Cm.stride(0), D=ptr(D32), z=ptr(zz),
        z_batch_stride=0 if zz is None else zz.stride(0), dt_bias=ptr(b32), out=ptr(out),
        out_batch_stride=out.stride(0))
    _lib.call("mtts_selective_state_update", p)
    return out


def mamba_inner_fn(xz, conv1d_weight, conv1d_bias, x_proj_weight, delta_proj_weight,
                   out_proj_weight, out_proj_bias, A, B=None, C=None, D=None, delta_bias=None,
                   B_proj_bias=None, C_proj_bias=None, delta_softplus=True):
    """Upstream's fused block body: conv -> x_proj -> dt_proj -> scan -> out_proj.
    xz: (batch, 2*dim, seqlen).  The projections are library GEMMs; conv and scan are ours."""
    if B is not None or C is not None or B_proj_bias is not None or C_proj_bias is not None:
        raise NotImplementedError("only input-dependent B and C")
    x, z = xz.chunk(2, dim=1)
    w2d = conv1d_weight.squeeze(1) if conv1d_weight.dim() == 3 else conv1d_weight
    x = causal_conv1d_fn(x, w2d, conv1d_bias, activation="silu")
    R = delta_proj_weight.shape[1]
    N = A.shape[1]
    x_dbl = torch.matmul(x_proj_weight.to(x.dtype), x)              # (batch, R + 2N, L)
    delta = torch.matmul(delta_proj_weight.to(x.dtype), x_dbl[:, :R])
    y = selective_scan_fn(x, delta, A, x_dbl[:, R:R + N], x_dbl[:, R + N:], D, z=z,
                          delta_bias=delta_bias, delta_softplus=delta_softplus)
    return F.linear(y.transpose(1, 2), out_proj_weight.to(y.dtype),
                    None if out_proj_bias is None else out_proj_bias.to(y.dtype))


class _MambaBlockFn(torch.autograd.Function):
    """The whole Mamba block body as ONE autograd node (upstream's ``MambaInnerFn``):
    in_proj -> conv+SiLU -> x_proj -> dt_proj -> scan -> out_proj on a (B, T, D) input.

    Why not the composition of the individual Functions: their backward hands autograd dx and dz as two
    tensors (it then ``cat``s them for ``xz.chunk``), du and the x_proj input gradient as two tensors (it
    then ``add_``s them), and every weight gets its own cast.  Here conv and scan backward write dx | dz
    into the two halves of one (B, 2 Di, T) buffer, du is the accumulator of the x_proj GEMM, and nothing
    is copied.  All activations stay channel-major (B, C, T)."""

    @staticmethod
    def forward(ctx, h, in_w, conv_w, conv_b, xproj_w, dtproj_w, dt_bias, A_log, D, out_w, dtype):
        _lib.require_cuda(h, in_w, conv_w, xproj_w, dtproj_w, A_log, out_w)
        Bsz, T, Dm = h.shape
        Di, W = conv_w.shape
        R = dtproj_w.shape[1]
        N = A_log.shape[1]
        dev = h.device
        hc = h if h.dtype == dtype else h.to(dtype)
        if not hc.is_contiguous():
            hc = hc.contiguous()
        from . import dense
        # bf16: every projection on mtts_gemm (tcgen05), reading the channel-major activations in place
        tc = dense.tc_enabled(dtype) and T % 8 == 0 and Dm % 8 == 0 and Di % 8 == 0 and R % 8 == 0 and N % 4 == 0
        if tc:
            Wi, Wx, Wdt, Wo = (dense.bf16_weight(w) for w in (in_w, xproj_w, dtproj_w, out_w))
        else:
            Wi, Wx, Wdt, Wo = (w.detach().to(dtype) for w in (in_w, xproj_w, dtproj_w, out_w))
        bc = lambda w: w.unsqueeze(0).expand(Bsz, -1, -1)
        cw32, cb32 = _f32c(conv_w), _f32c(conv_b)
        A = -torch.exp(A_log.detach().float())
        D32, db32 = _f32c(D), _f32c(dt_bias)
        io = _lib.io_dtype(hc)
        with torch.autocast("cuda", enabled=False):
            if tc:
                xz = dense.gemm(bc(Wi), hc)                                              # (B, 2Di, T)
            else:
                xz = torch.bmm(bc(Wi), hc.transpose(1, 2))
            x, z = xz[:, :Di], xz[:, Di:]
            xc = torch.empty((Bsz, Di, T), dtype=dtype, device=dev)
            _lib.call("mtts_causal_conv1d_fwd", _lib.Conv1dFwdParams(
                batch=Bsz, dim=Di, seqlen=T, width=W, io_dtype=io, silu=1, x=ptr(x),
                x_batch_stride=x.stride(0), x_dim_stride=x.stride(1), weight=ptr(cw32), bias=ptr(cb32),
                initial_states=None, init_batch_stride=0, init_dim_stride=0, out=ptr(xc),
                out_batch_stride=xc.stride(0), out_dim_stride=xc.stride(1)))
            if tc:
                x_dbl = dense.gemm(bc(Wx), xc.transpose(1, 2))                           # (B, R+2N, T)
                delta = dense.gemm(bc(Wdt), x_dbl[:, :R].transpose(1, 2))                # (B, Di, T)
            else:
                x_dbl = torch.bmm(bc(Wx), xc)
                delta = torch.bmm(bc(Wdt), x_dbl[:, :R])
            Bm, Cm = x_dbl[:, R:R + N], x_dbl[:, R + N:]
            y = torch.empty((Bsz, Di, T), dtype=dtype, device=dev)
            last = torch.empty((Bsz, Di, N), dtype=torch.float32, device=dev)
            chk = torch.empty((Bsz, Di, max(scan_num_chunks(T), 1), N), dtype=torch.float32, device=dev)
            ypre = torch.empty((Bsz, Di, T), dtype=dtype, device=dev)   # y before the gate, for dz
            _lib.call("mtts_selective_scan_fwd", _lib.ScanFwdParams(
                batch=Bsz, dim=Di, seqlen=T, dstate=N, io_dtype=io, delta_softplus=1,
                u=ptr(xc), u_batch_stride=xc.stride(0), u_dim_stride=xc.stride(1),
                delta=ptr(delta), delta_batch_stride=delta.stride(0), delta_dim_stride=delta.stride(1),
                A=ptr(A), B=ptr(Bm), B_batch_stride=Bm.stride(0), B_state_stride=Bm.stride(1),
                C=ptr(Cm), C_batch_stride=Cm.stride(0), C_state_stride=Cm.stride(1),
                D=ptr(D32), delta_bias=ptr(db32), z=ptr(z), z_batch_stride=z.stride(0),
                z_dim_stride=z.stride(1), initial_state=None, out=ptr(y),
                out_batch_stride=y.stride(0), out_dim_stride=y.stride(1), last_state=ptr(last),
                checkpoints=ptr(chk), y_pre=ptr(ypre), y_batch_stride=ypre.stride(0),
                y_dim_stride=ypre.stride(1)))
            if tc:
                out = dense.gemm(y.transpose(1, 2), bc(Wo))                              # (B, T, D)
            else:
                out = torch.bmm(y.transpose(1, 2), bc(Wo.t()))
            new_conv = F.pad(x[..., -W:], (max(0, W - T), 0)) if T < W else x[..., -W:].clone()
        ctx.tc = tc
        ctx.save_for_backward(hc, xz, xc, x_dbl, delta, y, chk, A, D32, db32, cw32, cb32, Wi, Wx, Wdt, Wo,
                              ypre)
        ctx.meta = (h.dtype, tuple(t.dtype for t in (in_w, conv_w, conv_b, xproj_w, dtproj_w, dt_bias,
                                                      A_log, D, out_w)))
        ctx.mark_non_differentiable(new_conv, last)
        return out, new_conv, last

    @staticmethod
    def backward(ctx, dout, _dconv, _dlast):
        (hc, xz, xc, x_dbl, delta, y, chk, A, D32, db32, cw32, cb32, Wi, Wx, Wdt, Wo,
         ypre) = ctx.saved_tensors
        t_h, (t_in, t_cw, t_cb, t_xw, t_dtw, t_dtb, t_A, t_D, t_ow) = ctx.meta
        Bsz, T, Dm = hc.shape
        Di, W = cw32.shape
        R = Wdt.shape[1]
        N = A.shape[1]
        dev, dtype = hc.device, hc.dtype
        io = _lib.io_dtype(hc)
        dout = dout.to(dtype)
        if not dout.is_contiguous():
            dout = dout.contiguous()
        f32 = torch.float32
        tc = ctx.tc
        if tc:
            from . import dense
            g = dense.gemm
        bc = lambda w: w.unsqueeze(0).expand(Bsz, -1, -1)
        with torch.autocast("cuda", enabled=False):
            doutT = dout.transpose(1, 2)                                                  # (B, D, T)
            if tc:
                dy = g(bc(Wo.t()), dout)                                                  # (B, Di, T)
                dWo = g(doutT, y, out_dtype=f32, reduce_batch=True, split_k=-1)           # (D, Di)
            else:
                dy = torch.bmm(bc(Wo.t()), doutT)
                dWo = torch.bmm(doutT, y.transpose(1, 2)).sum(0)
            dxz = torch.empty_like(xz)
            dxh, dzh = dxz[:, :Di], dxz[:, Di:]
            x, z = xz[:, :Di], xz[:, Di:]
            Bm, Cm = x_dbl[:, R:R + N], x_dbl[:, R + N:]
            du = torch.empty((Bsz, Di, T), dtype=dtype, device=dev)
            ddelta = torch.empty((Bsz, Di, T), dtype=dtype, device=dev)
            # every accumulated-into fp32 gradient of the block lives in ONE zero-filled buffer (one fill launch)
            sizes = (Di * N, 2 * Bsz * N * T, Di, Di, Di * W, Di)
            pool = torch.zeros(sum(sizes), dtype=f32, device=dev)
            dA, dBC, dD, ddb, dcw, dcb = (t.view(sh) for t, sh in zip(
                pool.split(sizes), ((Di, N), (2, Bsz, N, T), (Di,), (Di,), (Di, W), (Di,))))
            _lib.call("mtts_selective_scan_bwd", _lib.ScanBwdParams(
                batch=Bsz, dim=Di, seqlen=T, dstate=N, io_dtype=io, delta_softplus=1,
                u=ptr(xc), u_batch_stride=xc.stride(0), u_dim_stride=xc.stride(1),
                delta=ptr(delta), delta_batch_stride=delta.stride(0), delta_dim_stride=delta.stride(1),
                A=ptr(A), B=ptr(Bm), B_batch_stride=Bm.stride(0), B_state_stride=Bm.stride(1),
                C=ptr(Cm), C_batch_stride=Cm.stride(0), C_state_stride=Cm.stride(1),
                D=ptr(D32), delta_bias=ptr(db32), z=ptr(z), z_batch_stride=z.stride(0),
                z_dim_stride=z.stride(1), dout=ptr(dy), dout_batch_stride=dy.stride(0),
                dout_dim_stride=dy.stride(1), checkpoints=ptr(chk),
                du=ptr(du), du_batch_stride=du.stride(0), du_dim_stride=du.stride(1),
                ddelta=ptr(ddelta), ddelta_batch_stride=ddelta.stride(0),
                ddelta_dim_stride=ddelta.stride(1),
                dz=ptr(dzh), dz_batch_stride=dzh.stride(0), dz_dim_stride=dzh.stride(1),
                dA=ptr(dA), dB=ptr(dBC[0]), dC=ptr(dBC[1]), dD=ptr(dD), ddelta_bias=ptr(ddb),
                y_pre=ptr(ypre), y_batch_stride=ypre.stride(0), y_dim_stride=ypre.stride(1)))
            # d x_dbl = [W_dt^T ddelta | dB | dC]  (B, R + 2N, T), small
            dx_dbl = torch.empty((Bsz, R + 2 * N, T), dtype=dtype, device=dev)
            if tc:
                g(bc(Wdt.t()), ddelta.transpose(1, 2), out=dx_dbl[:, :R])
            else:
                dx_dbl[:, :R].copy_(torch.bmm(bc(Wdt.t()), ddelta))
            dx_dbl[:, R:].view(Bsz, 2, N, T).copy_(dBC.transpose(0, 1))           # dB | dC in one cast-copy
            if tc:
                dWdt = g(ddelta, x_dbl[:, :R], out_dtype=f32, reduce_batch=True, split_k=-1)      # (Di, R)
                # d xc = du + W_x^T d x_dbl : the GEMM accumulates onto du
                dxc = g(bc(Wx.t()), dx_dbl.transpose(1, 2), out=du, accumulate=True)
                dWx = g(dx_dbl, xc, out_dtype=f32, reduce_batch=True, split_k=-1)                 # (R+2N, Di)
            else:
                dWdt = torch.bmm(ddelta, x_dbl[:, :R].transpose(1, 2)).sum(0)
                dxc = du.baddbmm_(bc(Wx.t()), dx_dbl)
                dWx = torch.bmm(dx_dbl, xc.transpose(1, 2)).sum(0)
            _lib.call("mtts_causal_conv1d_bwd", _lib.Conv1dBwdParams(
                batch=Bsz, dim=Di, seqlen=T, width=W, io_dtype=io, silu=1, x=ptr(x),
                x_batch_stride=x.stride(0), x_dim_stride=x.stride(1), weight=ptr(cw32), bias=ptr(cb32),
                initial_states=None, init_batch_stride=0, init_dim_stride=0,
                dout=ptr(dxc), dout_batch_stride=dxc.stride(0), dout_dim_stride=dxc.stride(1),
                dx=ptr(dxh), dx_batch_stride=dxh.stride(0), dx_dim_stride=dxh.stride(1),
                dweight=ptr(dcw), dbias=ptr(dcb)))
            if tc:
                dh = g(dxz.transpose(1, 2), bc(Wi.t()))                                           # (B, T, D)
                dWi = g(dxz, hc.transpose(1, 2), out_dtype=f32, reduce_batch=True, split_k=-1)    # (2Di, D)
            else:
                dh = torch.bmm(dxz.transpose(1, 2), bc(Wi))
                dWi = torch.bmm(dxz, hc).sum(0)
            dA_log = dA * A                                                                # A = -exp(A_log)
        return (dh.to(t_h), dWi.to(t_in), dcw.to(t_cw), None if t_cb is None else dcb.to(t_cb),
                dWx.to(t_xw), dWdt.to(t_dtw), ddb.to(t_dtb), dA_log.to(t_A), dD.to(t_D), dWo.to(t_ow),
                None)


def mamba_block_fn(h, in_proj_weight, conv_weight, conv_bias, x_proj_weight, dt_proj_weight, dt_bias,
                   A_log, D, out_proj_weight, dtype=None):
    """h (batch, T, d_model) -> (out (batch, T, d_model), conv_state (batch, d_inner, width),
    ssm_state (batch, d_inner, dstate) fp32): the bias-free Mamba block from the zero state, as one
    autograd node.  conv_weight (d_inner, width)."""
    if dtype is None:
        dtype = torch.get_autocast_dtype("cuda") if torch.is_autocast_enabled("cuda") else h.dtype
    return _MambaBlockFn.apply(h, in_proj_weight, conv_weight, conv_bias, x_proj_weight, dt_proj_weight,
                               dt_bias, A_log, D, out_proj_weight, dtype)


# ------------------------------------------------------------------------------------------------
# decode-step kernels (inference only)
# ------------------------------------------------------------------------------------------------
def mamba_decode_step(xz, conv_state, ssm_state, conv_weight, conv_bias, x_proj_w, dt_proj_w,
                      dt_bias, A, D, out=None, prefetch=None):
    """The inner part of ``Mamba.step`` in one launch; both states updated IN PLACE.
    ``prefetch`` = (a, b): two contiguous (batch, ...) tensors of equal size that the kernel pulls into L2
    after its own loads (the decoder passes the layer's cached K / V for the attention that follows).
    xz (batch, 2*dim); conv_state (batch, dim, width) same dtype; ssm_state (batch, dim, dstate)
    fp32; conv_weight (dim, width) / conv_bias / dt_bias / A / D fp32; x_proj_w (R + 2N, dim) and
    dt_proj_w (dim, R) in xz's dtype.  Returns y (batch, dim)."""
    _lib.require_cuda(xz, conv_state, ssm_state)
    batch, two_dim = xz.shape
    dim = two_dim // 2
    N = ssm_state.shape[2]
    R = dt_proj_w.shape[1]
    width = conv_state.shape[2]
    io = xz.dtype
    for name, t, shape, dt in (("conv_state", conv_state, (batch, dim, width), io),
                               ("ssm_state", ssm_state, (batch, dim, N), torch.float32),
                               ("x_proj_w", x_proj_w, (R + 2 * N, dim), io),
                               ("dt_proj_w", dt_proj_w, (dim, R), io),
                               ("conv_weight", conv_weight, (dim, width), torch.float32),
                               ("dt_bias", dt_bias, (dim,), torch.float32),
                               ("A", A, (dim, N), torch.float32), ("D", D, (dim,), torch.float32)):
        if tuple(t.shape) != shape or t.dtype != dt or not t.is_contiguous():
            raise RuntimeError(f"{name} must be contiguous {shape} {dt}, got {tuple(t.shape)} {t.dtype}")
    xz = _unit_last_stride(xz)
    y = out if out is not None else torch.empty(batch, dim, dtype=io, device=xz.device)
    p = _lib.DecodeStepParams(
        batch=batch, dim=dim, dstate=N, dt_rank=R, width=width, io_dtype=_lib.io_dtype(xz),
        xz=ptr(xz), xz_batch_stride=xz.stride(0), conv_state=ptr(conv_state),
        ssm_state=ptr(ssm_state), conv_weight=ptr(conv_weight), conv_bias=ptr(conv_bias),
        x_proj_w=ptr(x_proj_w), dt_proj_w=ptr(dt_proj_w), dt_bias=ptr(dt_bias), A=ptr(A), D=ptr(D),
        y=ptr(y), y_batch_stride=y.stride(0))
    if prefetch is not None:
        pa, pb = prefetch
        _lib.require_cuda(pa, pb)
        if not (pa.is_contiguous() and pb.is_contiguous()) or pa.shape[0] != batch or pb.shape[0] != batch \
                or pa.numel() * pa.element_size() != pb.numel() * pb.element_size():
            raise RuntimeError("prefetch tensors must be contiguous (batch, ...) of equal byte size")
        p.prefetch_a, p.prefetch_b = ptr(pa), ptr(pb)
        p.prefetch_bytes = pa.numel() * pa.element_size() // batch
    _lib.call("mtts_mamba_decode_step", p)
    return y


def cross_attn_decode(q, k, v, heads, mask=None, out=None):
    """q (batch, d_model) projected, unscaled; k, v (batch, t_kv, d_model) contiguous;
    mask (batch, t_kv) bool/uint8, True = attend.  Returns (batch, d_model)."""
    _lib.require_cuda(q, k, v, mask)
    batch, dm = q.shape
    if k.shape != v.shape or k.shape[0] != batch or k.shape[2] != dm or dm % heads:
        raise RuntimeError("k, v must be (batch, t_kv, d_model); d_model divisible by heads")
    if not (q.is_contiguous() and k.is_contiguous() and v.is_contiguous()):
        raise RuntimeError("q, k, v must be contiguous")
    if k.dtype != q.dtype or v.dtype != q.dtype:
        raise RuntimeError("q, k, v must share a dtype")
    m8 = None
    if mask is not None:
        m8 = mask.to(torch.uint8).contiguous() if mask.dtype != torch.uint8 else mask.contiguous()
    o = out if out is not None else torch.empty_like(q)
    p = _lib.CrossAttnDecodeParams(batch=batch, heads=heads, head_dim=dm // heads, t_kv=k.shape[1],
                                   io_dtype=_lib.io_dtype(q), q=ptr(q), k=ptr(k), v=ptr(v),
                                   mask=ptr(m8), out=ptr(o))
    _lib.call("mtts_cross_attn_decode", p)
    return o


def cross_attn_block_decode_supported(dtype, d_model, heads, t_kv):
    """Shapes ``cross_attn_block_decode`` has a kernel for (else use the separate ops)."""
    return dtype in (torch.bfloat16, torch.float32) and heads == 8 and d_model == 512 and t_kv <= 256


def cross_attn_block_decode(x, delta, lnq, wq, bq, k, v, heads, wo=None, bo=None, lno=None, mask=None,
                            gamma=None, beta=None, out=None):
    """One launch for the cross-attention branch of a decode step (``mtts_cross_attn_block_decode``):
    x (batch, d) fp32 residual stream; delta (batch, d) or None;
    lnq / lno = (weight, bias, eps) of the LayerNorms before the attention / before the FFN; wq, bq, wo,
    bo in the io dtype; k, v (batch, t_kv, d); gamma, beta (batch, d) fp32 FiLM or None.
    Returns ``(x_new, out)``.  Whole branch (wo given): x_new = x + delta + o, written IN PLACE into x (the
    8 head CTAs of a batch element are one cluster and synchronise before the write), out =
    FiLM(LN(x_new; lno)) (batch, d) in the io dtype.  Front half only (wo=None): x_new = x + delta in a NEW
    tensor -- the head CTAs of a batch element are independent there and all of them read the whole row of x,
    so the row must not be overwritten by the launch -- and out = the attention output before the out
    projection."""
    _lib.require_cuda(x, delta, wq, bq, k, v, wo, bo, mask, gamma, beta)
    batch, dm = x.shape
    if x.dtype != torch.float32 or not x.is_contiguous():
        raise RuntimeError("the residual stream x must be contiguous fp32")
    dt = k.dtype
    front_only = wo is None  # x <- x + delta, returns the attention output (before the out projection)
    for t in (wq, bq, v) + (() if front_only else (wo, bo)) + (() if delta is None else (delta,)):
        if t.dtype != dt or not t.is_contiguous():
            raise RuntimeError("delta, wq, bq, wo, bo, k, v must be contiguous and share a dtype")
    if k.shape != v.shape or k.shape[0] != batch or k.shape[2] != dm or not k.is_contiguous():
        raise RuntimeError("k, v must be contiguous (batch, t_kv, d_model)")
    if wq.shape != (dm, dm) or bq.shape != (dm,) or (
            not front_only and (wo.shape != (dm, dm) or bo.shape != (dm,))):
        raise RuntimeError("wq, wo must be (d, d); bq, bo (d)")
    if (gamma is None) != (beta is None):
        raise RuntimeError("gamma and beta come together")
    m8 = None
    if mask is not None:
        m8 = mask.to(torch.uint8).contiguous() if mask.dtype != torch.uint8 else mask.contiguous()
    o = out if out is not None else torch.empty(batch, dm, device=x.device, dtype=dt)
    x_new = torch.empty_like(x) if front_only else x
    p = _lib.CrossAttnBlockParams(
        batch=batch, heads=heads, head_dim=dm // heads, t_kv=k.shape[1], io_dtype=_lib.io_dtype(k),
        eps_q=lnq[2], eps_o=0.0 if front_only else lno[2], x=ptr(x), delta=ptr(delta), x_out=ptr(x_new),
        lnq_weight=ptr(lnq[0]), lnq_bias=ptr(lnq[1]), wq=ptr(wq), bq=ptr(bq), k=ptr(k), v=ptr(v),
        mask=ptr(m8), wo=ptr(wo), bo=ptr(bo), lno_weight=None if front_only else ptr(lno[0]),
        lno_bias=None if front_only else ptr(lno[1]),
        film_gamma=ptr(gamma), film_beta=ptr(beta), out=ptr(o))
    _lib.call("mtts_cross_attn_block_decode", p)
    return x_new, o


def decode_embed(tok, pos, tok_embed, pos_embed, x, step=None):
    """x[b] = tok_embed[tok[b]] + pos_embed[pos] into the static fp32 buffer x (batch, d); then step += 1.
    tok (batch) int64, pos / step one-element int64 device tensors (``mtts_decode_embed``)."""
    _lib.require_cuda(tok, pos, tok_embed, pos_embed, x, step)
    if tok.dtype != torch.long or pos.dtype != torch.long or (step is not None and step.dtype != torch.long):
        raise RuntimeError("tok, pos, step must be int64")
    if tok_embed.dtype != torch.float32 or pos_embed.dtype != torch.float32 or x.dtype != torch.float32:
        raise RuntimeError("embedding tables and x must be fp32")
    if not (tok.is_contiguous() and tok_embed.is_contiguous() and pos_embed.is_contiguous() and x.is_contiguous()):
        raise RuntimeError("operands must be contiguous")
    batch, dim = x.shape
    if tok.shape != (batch,) or tok_embed.shape[1] != dim or pos_embed.shape[1] != dim:
        raise RuntimeError("shape mismatch")
    p = _lib.DecodeEmbedParams(batch=batch, dim=dim, tok=ptr(tok), pos=ptr(pos), tok_embed=ptr(tok_embed),
                               pos_embed=ptr(pos_embed), x=ptr(x), step=ptr(step))
    _lib.call("mtts_decode_embed", p)
    return x


def decode_greedy(logits, tok, out=None, step=None, pos=None, eos_id=None, pad_id=0, lengths=None):
    """tok[b] = argmax logits[b] (lowest index on ties), out[b, step] = tok[b]; then pos += 1
    (``mtts_decode_greedy``).  logits (batch, vocab) fp32 / bf16; tok (batch), out (batch, n) int64.
    With ``eos_id``: rows that produced it earlier get ``pad_id``; ``lengths`` (batch) int64, -1 while a
    row is running, receives the token count up to and including the eos."""
    _lib.require_cuda(logits, tok, out, step, pos)
    if not logits.is_contiguous() or logits.dim() != 2 or tok.shape != (logits.shape[0],):
        raise RuntimeError("logits must be contiguous (batch, vocab), tok (batch)")
    if tok.dtype != torch.long or (out is not None and (out.dtype != torch.long or out.stride(1) != 1)):
        raise RuntimeError("tok / out must be int64, out unit-stride along steps")
    if out is not None and step is None:
        raise RuntimeError("out needs the step counter")
    if eos_id is not None and (lengths is None or lengths.dtype != torch.long or lengths.shape != tok.shape):
        raise RuntimeError("eos_id needs lengths (batch) int64")
    _lib.require_cuda(lengths)
    p = _lib.DecodeGreedyParams(batch=logits.shape[0], vocab=logits.shape[1], io_dtype=_lib.io_dtype(logits),
                                reserved=0, logits=ptr(logits), tok=ptr(tok), out=ptr(out),
                                out_stride=0 if out is None else out.stride(0), step=ptr(step), pos=ptr(pos),
                                eos_id=-1 if eos_id is None else int(eos_id), pad_id=int(pad_id),
                                lengths=ptr(lengths))
    _lib.call("mtts_decode_greedy", p)
    return tok


class _LengthRegulateFn(torch.autograd.Function):
    """``mtts_length_regulate_fwd / _bwd``: frame-level expansion of phoneme rows by their durations."""

    @staticmethod
    def forward(ctx, hidden, durations, max_len):
        _lib.require_cuda(hidden, durations)
        if hidden.dim() != 3 or durations.shape != hidden.shape[:2]:
            raise RuntimeError("hidden must be (B, T_text, D) and durations (B, T_text)")
        B, T, D = hidden.shape
        h = hidden.contiguous()
        dur = durations.detach().float().contiguous()
        expanded = torch.empty(B, max_len, D, dtype=h.dtype, device=h.device)
        lengths = torch.empty(B, dtype=torch.long, device=h.device)
        if B > 0 and max_len == 0:  # nothing to expand, but the lengths are still defined
            lengths.copy_(torch.clamp(torch.round(dur), min=0).long().sum(1))
        p = _lib.LengthRegulateFwdParams(batch=B, t_text=T, dim=D, max_len=max_len, io_dtype=_lib.io_dtype(h),
                                         reserved=0, hidden=ptr(h), durations=ptr(dur), expanded=ptr(expanded),
                                         output_lengths=ptr(lengths), frame_index=None)
        _lib.call("mtts_length_regulate_fwd", p)
        ctx.save_for_backward(dur)
        ctx.meta = (B, T, D, max_len, hidden.dtype)
        ctx.mark_non_differentiable(lengths)
        return expanded, lengths

    @staticmethod
    def backward(ctx, dexp, _dlen):
        (dur,) = ctx.saved_tensors
        B, T, D, max_len, dt = ctx.meta
        g = dexp.to(dt).contiguous()
        dh = torch.empty(B, T, D, dtype=dt, device=g.device)
        p = _lib.LengthRegulateBwdParams(batch=B, t_text=T, dim=D, max_len=max_len, io_dtype=_lib.io_dtype(g),
                                         reserved=0, durations=ptr(dur), dexpanded=ptr(g), dhidden=ptr(dh))
        _lib.call("mtts_length_regulate_bwd", p)
        return dh, None, None


def length_regulate(hidden, durations, max_len=None):
    """``LengthRegulator.forward`` (``style_cross_attention.py:155-198``): (expanded (B, max_len, D),
    output_lengths (B,) int64).  ``max_len=None`` takes the longest row of the batch, which -- exactly as in
    the reference (``:178-179``) -- costs one host sync; pass it to stay asynchronous."""
    if max_len is None:
        _lib.require_cuda(hidden, durations)
        dur = torch.clamp(torch.round(durations.detach().float()), min=0).long()
        max_len = int(dur.sum(dim=1).max().item()) if dur.shape[0] > 0 else 0
    return _LengthRegulateFn.apply(hidden, durations, int(max_len))


class _AddLayerNormFn(torch.autograd.Function):
    """(x_out, out) = add_layernorm(x, delta, w, b, gamma, beta): see ``mtts_add_layernorm_fwd``."""

    @staticmethod
    def forward(ctx, x, delta, weight, bias, gamma, beta, eps, out_dtype, inplace, delta_bias):
        _lib.require_cuda(x, delta, weight, bias, gamma, beta, delta_bias)
        if delta_bias is not None and delta is None:
            raise RuntimeError("delta_bias needs delta")
        if x.dtype != torch.float32:
            raise RuntimeError("the residual stream x must be fp32")
        shape = x.shape
        dim = shape[-1]
        x2 = x.reshape(-1, dim)
        if not x2.is_contiguous():
            x2 = x2.contiguous()
        rows = x2.shape[0]
        rpb = max(rows, 1)  # no FiLM: the whole tensor is one "batch element" for the column sums
        if gamma is not None:
            if gamma.dim() != 2 or gamma.shape[1] != dim or rows % gamma.shape[0]:
                raise RuntimeError("gamma/beta must be (batch, dim) with batch dividing the rows")
            rpb = rows // gamma.shape[0]
        d2 = None
        if delta is not None:
            d2 = delta.reshape(-1, dim)
            if d2.dtype != out_dtype:
                d2 = d2.to(out_dtype)
            if not d2.is_contiguous():
                d2 = d2.contiguous()
        w32, b32, g32, be32 = _f32c(weight), _f32c(bias), _f32c(gamma), _f32c(beta)
        db32 = _f32c(delta_bias)
        need = any(ctx.needs_input_grad)
        out = torch.empty((rows, dim), dtype=out_dtype, device=x.device)
        if d2 is None:
            x_out = x2
        else:
            x_out = x2 if inplace else torch.empty_like(x2)
        mean = torch.empty(rows, dtype=torch.float32, device=x.device) if need else None
        rstd = torch.empty(rows, dtype=torch.float32, device=x.device) if need else None
        if rows:
            p = _lib.AddLayerNormFwdParams(
                rows=rows, dim=dim, rows_per_batch=rpb, io_dtype=_lib.io_dtype(out), eps=eps,
                x=ptr(x2), delta=ptr(d2), delta_bias=ptr(db32),
                x_out=None if d2 is None else ptr(x_out),
                ln_weight=ptr(w32), ln_bias=ptr(b32), film_gamma=ptr(g32), film_beta=ptr(be32),
                out=ptr(out), mean=ptr(mean), rstd=ptr(rstd))
            _lib.call("mtts_add_layernorm_fwd", p)
        ctx.save_for_backward(x_out, mean, rstd, w32, b32, g32)
        ctx.meta = (rows, dim, rpb, shape, delta is not None,
                    None if delta is None else delta.dtype, weight.dtype, bias.dtype,
                    None if gamma is None else gamma.dtype,
                    None if delta_bias is None else delta_bias.dtype)
        if inplace and d2 is not None:
            if x2.data_ptr() != x.data_ptr():
                raise RuntimeError("inplace add_layernorm needs a contiguous x")
            ctx.mark_dirty(x)
            return x, out.view(shape)
        return x_out.view(shape), out.view(shape)

    @staticmethod
    def backward(ctx, dx_out, dout):
        x_out, mean, rstd, w32, b32, g32 = ctx.saved_tensors
        rows, dim, rpb, shape, has_delta, t_delta, t_w, t_b, t_g, t_db = ctx.meta
        dev = x_out.device
        batch = rows // rpb
        io = dout.dtype if dout is not None else torch.float32
        if dout is None:
            dout = torch.zeros((rows, dim), dtype=io, device=dev)
        dout2 = dout.reshape(-1, dim)
        if not dout2.is_contiguous():
            dout2 = dout2.contiguous()
        dxo = None
        if dx_out is not None:
            dxo = dx_out.reshape(-1, dim).to(torch.float32)
            if not dxo.is_contiguous():
                dxo = dxo.contiguous()
        dx = torch.empty((rows, dim), dtype=torch.float32, device=dev)
        ddelta = torch.empty((rows, dim), dtype=t_delta, device=dev) \
            if has_delta and t_delta == io else None
        colsum = torch.zeros((batch, 3, dim), dtype=torch.float32, device=dev)
        if rows:
            p = _lib.AddLayerNormBwdParams(
                rows=rows, dim=dim, rows_per_batch=rpb, io_dtype=_lib.io_dtype(dout2),
                x_out=ptr(x_out), mean=ptr(mean), rstd=ptr(rstd), ln_weight=ptr(w32),
                film_gamma=ptr(g32), dout=ptr(dout2), dx_out=ptr(dxo), dx=ptr(dx),
                ddelta=ptr(ddelta), colsum=ptr(colsum))
            _lib.call("mtts_add_layernorm_bwd", p)
        s1, s2 = colsum[:, 0], colsum[:, 1]
        if g32 is None:
            # no FiLM: one "batch element", the column sums ARE the gradients (no reduction kernels)
            dw, db, dgamma, dbeta = s1[0], s2[0], None, None
        else:
            # (batch, dim)-sized finishing (dw, db, dgamma, dbeta, d delta_bias) in one launch
            fin = torch.empty((3 + 2 * batch, dim), dtype=torch.float32, device=dev)
            dw, db, ddb_f = fin[0], fin[1], fin[2]
            dgamma, dbeta = fin[3:3 + batch], fin[3 + batch:]
            _lib.call("mtts_add_layernorm_bwd_finish", _lib.AddLayerNormFinishParams(
                batch=batch, dim=dim, colsum=ptr(colsum), film_gamma=ptr(g32), ln_weight=ptr(w32), ln_bias=ptr(b32),
                dweight=ptr(dw), dbias=ptr(db), dgamma=ptr(dgamma), dbeta=ptr(dbeta),
                ddelta_bias=ptr(ddb_f) if t_db is not None else None))
            dgamma, dbeta = dgamma.to(t_g), dbeta.to(t_g)
        if has_delta and ddelta is None:
            ddelta = dx.to(t_delta)
        ddb = None
        if t_db is not None:
            if g32 is not None:
                ddb = ddb_f.to(t_db)
            else:
                ddb = (colsum[0, 2] if batch == 1 else colsum[:, 2].sum(0)).to(t_db)
        return (dx.view(shape), None if not has_delta else ddelta.view(shape), dw.to(t_w),
                db.to(t_b), dgamma, dbeta, None, None, None, ddb)


def add_layernorm(x, delta, weight, bias, eps=1e-5, gamma=None, beta=None, out_dtype=None,
                  inplace=False, delta_bias=None):
    """Fused ``x_out = x + delta ; out = LN(x_out) [* gamma_b + beta_b]`` (``mamba_decoder.py:59-89``).

    x: (..., dim) fp32 residual stream; delta: same shape, activation dtype, or None; weight/bias:
    LayerNorm affine; gamma/beta: (batch, dim) FiLM terms or None.  Returns (x_out fp32, out in
    ``out_dtype``).  ``inplace`` writes x_out over x (inference only).  ``delta_bias`` (dim,): bias of the
    Linear that produced delta, added here (x_out = x + delta + delta_bias) so that neither the add nor
    its gradient (a column sum) needs a kernel of its own."""
    if out_dtype is None:
        out_dtype = delta.dtype if delta is not None else x.dtype
    return _AddLayerNormFn.apply(x, delta, weight, bias, gamma, beta, eps, out_dtype, inplace,
                                 delta_bias)


# ------------------------------------------------------------------------------------------------
# FFN / projection glue: bias + GELU, and bias gradients as column sums
# ------------------------------------------------------------------------------------------------
def _glue_params(x2, bias32=None, dout2=None, out2=None, colsum=None):
    return _lib.BiasGeluParams(rows=x2.shape[0], cols=x2.shape[1], io_dtype=_lib.io_dtype(x2), reserved=0,
                               ld=x2.stride(0), x=ptr(x2), bias=ptr(bias32), dout=ptr(dout2),
                               out=ptr(out2), colsum=ptr(colsum))


def _rows2d(t):
    t2 = t.reshape(-1, t.shape[-1])
    return t2 if t2.is_contiguous() else t2.contiguous()


def colsum(x):
    """Sum over every dimension but the last -> fp32 (dim,): the bias gradient of a Linear whose output
    gradient is ``x``.  One streaming launch (``mtts_colsum``)."""
    _lib.require_cuda(x)
    x2 = _rows2d(x)
    out = torch.zeros(x2.shape[1], dtype=torch.float32, device=x.device)
    if x2.numel():
        _lib.call("mtts_colsum", _glue_params(x2, colsum=out))
    return out


class _BiasGeluFn(torch.autograd.Function):
    """out = gelu(x + bias) (exact erf, ``nn.GELU()`` of ``mamba_decoder.py:41``); the backward writes
    dx = dout * gelu'(x + bias) and its column sum (= d bias) in one pass."""

    @staticmethod
    def forward(ctx, x, bias):
        _lib.require_cuda(x, bias)
        x2 = _rows2d(x)
        b32 = _f32c(bias)
        out = torch.empty_like(x2)
        if x2.numel():
            _lib.call("mtts_bias_gelu_fwd", _glue_params(x2, b32, out2=out))
        ctx.save_for_backward(x2, b32)
        ctx.meta = (x.shape, None if bias is None else bias.dtype)
        return out.view(x.shape)

    @staticmethod
    def backward(ctx, dout):
        x2, b32 = ctx.saved_tensors
        shape, t_b = ctx.meta
        d2 = _rows2d(dout.to(x2.dtype))
        dx = torch.empty_like(x2)
        cs = torch.zeros(x2.shape[1], dtype=torch.float32, device=x2.device)
        if x2.numel():
            _lib.call("mtts_bias_gelu_bwd", _glue_params(x2, b32, d2, dx, cs))
        return dx.view(shape), None if t_b is None else cs.to(t_b)


def bias_gelu(x, bias=None):
    """``gelu(x + bias)`` over the last dimension (fp32 or bf16 ``x``; 16-byte aligned rows)."""
    return _BiasGeluFn.apply(x, bias)


class _LinearFn(torch.autograd.Function):
    """``F.linear(x, weight, bias)`` in ``dtype`` with the bias gradient taken by ``mtts_colsum`` (aten's
    column reduction of a (32768, 512) bf16 gradient runs at 0.7 TB/s)."""

    @staticmethod
    def forward(ctx, x, weight, bias, dtype):
        xc = x if x.dtype == dtype else x.to(dtype)
        wc = weight if weight.dtype == dtype else weight.to(dtype)
        ctx.save_for_backward(xc, wc)
        ctx.meta = (x.dtype, weight.dtype, None if bias is None else bias.dtype)
        with torch.autocast("cuda", enabled=False):
            return F.linear(xc, wc, None if bias is None else bias.to(dtype))

    @staticmethod
    def backward(ctx, dy):
        xc, wc = ctx.saved_tensors
        t_x, t_w, t_b = ctx.meta
        dy = dy.to(wc.dtype)
        dy2, x2 = _rows2d(dy), _rows2d(xc)
        need_x, need_w, need_b = ctx.needs_input_grad[:3]
        dx = dw = db = None
        with torch.autocast("cuda", enabled=False):
            if need_x:   # e.g. text_hidden under the K/V projection needs none
                dx = (dy2 @ wc).view(xc.shape).to(t_x)
            if need_w:
                dw = (dy2.t() @ x2).to(t_w)
        if need_b and t_b is not None:
            db = colsum(dy2).to(t_b)
        return dx, dw, db, None


def linear(x, weight, bias=None, dtype=None):
    """``F.linear`` computing in ``dtype`` (default: the autocast dtype, else x's)."""
    if dtype is None:
        dtype = torch.get_autocast_dtype("cuda") if torch.is_autocast_enabled("cuda") else x.dtype
    return _LinearFn.apply(x, weight, bias, dtype)


def skinny_linear(w, bias=None, a=None, x=None, delta=None, x_out=None, ln_weight=None, ln_bias=None,
                  eps=1e-5, gamma=None, beta=None, gelu=False):
    """decode_step projection for <= 64 rows (bf16): ``out = act(A @ w.T + bias)`` with ``A = a`` or
    ``A = FiLM(LN(x + delta))`` (then ``x_out`` receives ``x + delta``; it must not alias ``x``).
    w (n, k), bias (n), a / delta (m, k) bf16; x, x_out (m, k) fp32; gamma / beta (m, k) fp32."""
    _lib.require_cuda(w, bias, a, x, delta, x_out, gamma, beta)
    ln = a is None
    src = x if ln else a
    m, k = src.shape
    n = w.shape[0]
    if w.shape[1] != k or w.dtype != torch.bfloat16 or not w.is_contiguous():
        raise RuntimeError("w must be contiguous bf16 (n, k)")
    for t in (a, x, delta, x_out, gamma, beta, bias):
        if t is not None and not t.is_contiguous():
            raise RuntimeError("skinny_linear operands must be contiguous")
    if ln and (x.dtype != torch.float32 or (delta is not None and delta.dtype != torch.bfloat16)):
        raise RuntimeError("x must be fp32 and delta bf16")
    if not ln and a.dtype != torch.bfloat16:
        raise RuntimeError("a must be bf16")
    out = torch.empty((m, n), dtype=torch.bfloat16, device=w.device)
    p = _lib.SkinnyLinearParams(
        m=m, n=n, k=k, io_dtype=_lib.BF16, ln_mode=int(ln), gelu=int(bool(gelu)), eps=eps,
        a=ptr(a), x=ptr(x), delta=ptr(delta), x_out=ptr(x_out), ln_weight=ptr(ln_weight),
        ln_bias=ptr(ln_bias), film_gamma=ptr(gamma), film_beta=ptr(beta), w=ptr(w), bias=ptr(bias),
        out=ptr(out))
    _lib.call("mtts_skinny_linear", p)
    return out


def gemm_bf16(a, w, bias=None, gelu=False, return_pre=False):
    """tcgen05 tensor-core GEMM with fused epilogue: ``act(a @ w.T + bias)``.
    a (..., k) bf16 with unit stride along k; w (n, k) bf16; bias (n) fp32.  Returns out (..., n) bf16
    [, pre-activation (..., n) bf16 when ``return_pre``]."""
    _lib.require_cuda(a, w, bias)
    if a.dtype != torch.bfloat16 or w.dtype != torch.bfloat16:
        raise RuntimeError("gemm_bf16 takes bf16 operands")
    k = a.shape[-1]
    a2 = a.reshape(-1, k)
    if a2.stride(-1) != 1:
        a2 = a2.contiguous()
    if w.stride(-1) != 1 or w.shape[1] != k:
        raise RuntimeError("w must be (n, k) with unit stride along k")
    m, n = a2.shape[0], w.shape[0]
    b32 = _f32c(bias)
    out = torch.empty((m, n), dtype=torch.bfloat16, device=a.device)
    pre = torch.empty_like(out) if return_pre else None
    p = _lib.GemmBf16Params(m=m, n=n, k=k, gelu=int(bool(gelu)), a=ptr(a2), lda=a2.stride(0), w=ptr(w),
                            ldw=w.stride(0), bias=ptr(b32), out=ptr(out), ldo=n, pre_out=ptr(pre))
    _lib.call("mtts_gemm_bf16", p)
    out = out.view(*a.shape[:-1], n)
    return (out, pre.view(*a.shape[:-1], n)) if return_pre else out


# ------------------------------------------------------------------------------------------------
# the caller's side of the training step (train.py:31-42,115-131,152-159,230-235)
# ------------------------------------------------------------------------------------------------
class _EmbedSumFn(torch.autograd.Function):
    """x = token_embed[tokens] + pos_embed[pos_ids] (+ quant_embed[quant_ids]) in one launch, fp32."""

    @staticmethod
    def forward(ctx, tokens, pos_ids, quant_ids, tok_w, pos_w, quant_w):
        _lib.require_cuda(tokens, pos_ids, quant_ids, tok_w, pos_w, quant_w)
        B, L = tokens.shape
        D = tok_w.shape[1]
        if tokens.dtype != torch.long or pos_ids.dtype != torch.long or pos_ids.shape != (L,):
            raise RuntimeError("tokens (B, L) and pos_ids (L) must be int64")
        if (quant_ids is None) != (quant_w is None):
            raise RuntimeError("quant_ids and quant_embed come together")
        for w in (tok_w, pos_w, quant_w):
            if w is not None and (w.dtype != torch.float32 or not w.is_contiguous() or w.shape[1] != D):
                raise RuntimeError("embedding tables must be contiguous fp32 (rows, dim)")
        tokens, pos_ids = tokens.contiguous(), pos_ids.contiguous()
        quant_ids = None if quant_ids is None else quant_ids.to(torch.long).contiguous()
        x = torch.empty((B, L, D), dtype=torch.float32, device=tokens.device)
        p = _lib.EmbedSumParams(batch=B, seqlen=L, dim=D, reserved=0, tokens=ptr(tokens), pos_ids=ptr(pos_ids),
                                quant_ids=ptr(quant_ids), token_embed=ptr(tok_w), pos_embed=ptr(pos_w),
                                quant_embed=ptr(quant_w), x=ptr(x))
        _lib.call("mtts_embed_sum_fwd", p)
        ctx.save_for_backward(tokens, pos_ids, quant_ids)
        ctx.shapes = (tok_w.shape, pos_w.shape, None if quant_w is None else quant_w.shape)
        return x

    @staticmethod
    def backward(ctx, dx):
        tokens, pos_ids, quant_ids = ctx.saved_tensors
        st, sp, sq = ctx.shapes
        B, L = tokens.shape
        dx = dx.float().contiguous()
        need = ctx.needs_input_grad
        z = lambda shape, on: torch.zeros(shape, dtype=torch.float32, device=dx.device) if on and shape else None
        dt, dp, dq = z(st, need[3]), z(sp, need[4]), z(sq, need[5])
        p = _lib.EmbedSumParams(batch=B, seqlen=L, dim=st[1], reserved=0, tokens=ptr(tokens), pos_ids=ptr(pos_ids),
                                quant_ids=ptr(quant_ids) if dq is not None else None, token_embed=ptr(dt),
                                pos_embed=ptr(dp), quant_embed=ptr(dq), x=ptr(dx))
        _lib.call("mtts_embed_sum_bwd", p)
        return None, None, None, dt, dp, dq


def embed_sum(tokens, pos_ids, quant_ids, token_embed, pos_embed, quant_embed=None):
    """(B, L) ids -> (B, L, D) fp32: ``tok + pos + quant`` of ``mamba_decoder.py:167-171`` /
    ``train.py:115-131``.  pos_ids / quant_ids are per position (L,), shared by the batch."""
    return _EmbedSumFn.apply(tokens, pos_ids, quant_ids, token_embed, pos_embed, quant_embed)


class _CeLossFn(torch.autograd.Function):
    """Sum of token cross entropies over targets != ignore_index and, in the same pass, its gradient."""

    @staticmethod
    def forward(ctx, logits, targets, n_valid, ignore_index):
        _lib.require_cuda(logits, targets, n_valid)
        V = logits.shape[-1]
        l2 = logits.reshape(-1, V)
        if l2.stride(-1) != 1:
            l2 = l2.contiguous()
        t = targets.reshape(-1).contiguous()
        nv = n_valid.detach().float().reshape(1).contiguous()
        loss = torch.zeros((), dtype=torch.float32, device=logits.device)
        need = ctx.needs_input_grad[0]
        dl = torch.empty_like(l2, memory_format=torch.contiguous_format) if need else None
        p = _lib.CeLossParams(rows=l2.shape[0], vocab=V, io_dtype=_lib.io_dtype(l2), ld=l2.stride(0),
                              ignore_index=int(ignore_index), logits=ptr(l2), targets=ptr(t), n_valid=ptr(nv),
                              grad_scale=1.0, reserved=0, loss_sum=ptr(loss), row_loss=None, dlogits=ptr(dl))
        _lib.call("mtts_ce_loss", p)
        ctx.save_for_backward(dl)
        ctx.shape = logits.shape
        return loss / nv[0].clamp(min=1.0)

    @staticmethod
    def backward(ctx, g):
        (dl,) = ctx.saved_tensors
        # in place: dl is this node's private buffer (a second backward through the same node is not supported)
        return dl.mul_(g.to(dl.dtype)).view(ctx.shape), None, None, None


def ce_loss(logits, targets, ignore_index=0, n_valid=None):
    """``F.cross_entropy(logits.view(-1, V), targets.view(-1), ignore_index=...)`` (``train.py:31-42``) on the
    logits' own dtype (no fp32 copy), gradient computed in the forward pass.  ``n_valid``: the divisor (default:
    this call's count of non-ignored targets; pass the global count under data parallelism / micro-batching)."""
    if n_valid is None:
        n_valid = (targets != ignore_index).sum()
    return _CeLossFn.apply(logits, targets, n_valid, ignore_index)


class FusedClipAdam:
    """``clip_grad_norm_(params, max_norm)`` + ``torch.optim.Adam(params, lr).step()`` (``train.py:152-159,233-234``)
    as two launches over a device-resident table of the parameters (``mtts_grad_sumsq`` + ``mtts_adam_step``):
    the clip coefficient is applied inside the update, gradients are never rescaled in memory.  Plain Adam
    (no weight decay, no amsgrad), fp32 parameters with contiguous fp32 gradients."""

    def __init__(self, params, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, max_norm=1.0):
        self.params = [p for p in params if p.requires_grad]
        self.lr, self.betas, self.eps, self.max_norm = lr, betas, eps, max_norm
        self.step_count = 0
        dev = self.params[0].device
        self.exp_avg = [torch.zeros_like(p, memory_format=torch.contiguous_format) for p in self.params]
        self.exp_avg_sq = [torch.zeros_like(p, memory_format=torch.contiguous_format) for p in self.params]
        self.sumsq = torch.zeros(1, dtype=torch.float32, device=dev)
        self._table_key = None
        self._dev = dev

    def zero_grad(self, set_to_none=True):
        for p in self.params:
            if set_to_none:
                p.grad = None
            elif p.grad is not None:
                p.grad.zero_()

    def _tables(self):
        import ctypes as C
        act = [(p, m, v) for p, m, v in zip(self.params, self.exp_avg, self.exp_avg_sq) if p.grad is not None]
        key = tuple((p.data_ptr(), p.grad.data_ptr()) for p, _, _ in act)
        if key != self._table_key:
            chunk = _lib.load().mtts_adam_chunk_elems()
            arr = (_lib.AdamTensor * max(len(act), 1))()
            chunks = []
            for i, (p, m, v) in enumerate(act):
                if p.dtype != torch.float32 or p.grad.dtype != torch.float32 or not p.is_contiguous() \
                        or not p.grad.is_contiguous():
                    raise RuntimeError("FusedClipAdam needs contiguous fp32 parameters and gradients")
                arr[i] = _lib.AdamTensor(param=p.data_ptr(), grad=p.grad.data_ptr(), exp_avg=m.data_ptr(),
                                         exp_avg_sq=v.data_ptr(), numel=p.numel())
                chunks += [(i, c) for c in range((p.numel() + chunk - 1) // chunk)]
            raw = bytes(memoryview(arr).cast("B"))[: C.sizeof(_lib.AdamTensor) * len(act)]
            self._tensors = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(self._dev)
            self._chunks = torch.tensor(chunks, dtype=torch.int32).reshape(-1, 2).to(self._dev)
            self._table_key = key
        return self._tensors, self._chunks

    @torch.no_grad()
    def step(self):
        tensors, chunks = self._tables()
        self.step_count += 1
        b1, b2 = self.betas
        p = _lib.AdamParams(tensors=tensors.data_ptr(), chunks=chunks.data_ptr(), num_chunks=chunks.shape[0],
                            max_norm=float(self.max_norm or 0.0), grad_sumsq=self.sumsq.data_ptr(),
                            step_size=self.lr / (1.0 - b1 ** self.step_count), beta1=b1, beta2=b2, eps=self.eps,
                            bias_correction2_sqrt=(1.0 - b2 ** self.step_count) ** 0.5, reserved=0)
        if self.max_norm:
            _lib.call("mtts_grad_sumsq", p)
        _lib.call("mtts_adam_step", p)

    def grad_norm(self):
        """Total gradient norm seen by the last step (before clipping)."""
        return self.sumsq.sqrt()[0]
