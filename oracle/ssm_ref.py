"""CPU restatement of the SSM operator arithmetic (TEST INFRASTRUCTURE, see oracle/__init__.py).

Follows the *published* reference functions of the third-party packages the
reference imports at ``/root/reference/mamba_decoder.py:4`` (``mamba_ssm``; un-pinned):

* ``selective_scan_ref``           <- mamba_ssm/ops/selective_scan_interface.py::selective_scan_ref
* ``selective_state_update_ref``   <- mamba_ssm/ops/triton/selective_state_update.py::selective_state_update_ref
* ``causal_conv1d_ref`` / ``causal_conv1d_update_ref``
                                   <- causal_conv1d/causal_conv1d_interface.py

Only the Mamba-1 case the decoder uses is restated: real ``A (Di, N)``,
input-dependent ``B, C (batch, N, T)`` (one group), optional ``D``, ``z``,
``delta_bias``, ``delta_softplus``.  Two extensions over upstream's reference, both
required by the ``(out, state)`` contract of ``mamba_decoder.py:9-15``:

* ``initial_state`` -- start the recurrence from a given ``h_{-1}`` (prompt-then-decode).
* ``dim_block``     -- evaluate channels in slices so the ``(batch, Di, T, N)`` fp32
  intermediates of the upstream formulation stay bounded (channels are independent,
  so slicing does not change a single bit of the result).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def selective_scan_ref(u, delta, A, B, C, D=None, z=None, delta_bias=None,
                       delta_softplus=False, return_last_state=False,
                       initial_state=None, dim_block=None):
    """u, delta, z: (batch, Di, T); A: (Di, N); B, C: (batch, N, T); D, delta_bias: (Di,).

    h_l = exp(delta_l * A) * h_{l-1} + delta_l * B_l * u_l ;  y_l = <C_l, h_l> + D u_l ;
    out_l = y_l * silu(z_l).  All arithmetic in fp32, result cast back to u.dtype.
    """
    if B.dim() != 3 or C.dim() != 3 or A.is_complex():
        raise NotImplementedError("oracle restates only real A with 3-D input-dependent B/C")
    dtype_in = u.dtype
    batch, dim, T = u.shape
    N = A.shape[1]
    A = A.float()
    Bf, Cf = B.float(), C.float()
    blk = dim if dim_block is None else int(dim_block)
    outs, lasts = [], []
    for d0 in range(0, dim, blk):
        sl = slice(d0, min(dim, d0 + blk))
        uf = u[:, sl].float()
        dl = delta[:, sl].float()
        if delta_bias is not None:
            dl = dl + delta_bias[sl].float()[None, :, None]
        if delta_softplus:
            dl = F.softplus(dl)
        x = (A.new_zeros((batch, uf.shape[1], N)) if initial_state is None
             else initial_state[:, sl].float().clone())
        deltaA = torch.exp(torch.einsum("bdl,dn->bdln", dl, A[sl]))
        deltaB_u = torch.einsum("bdl,bnl,bdl->bdln", dl, Bf, uf)
        # (unbind, not deltaA[:, :, i]: same values, but the backward of T slices is ONE stack instead of T
        # zero-filled (batch, Di, T, N) tensors -- the C2-shape fixtures take minutes instead of hours)
        dA_t, dBu_t, C_t = deltaA.unbind(2), deltaB_u.unbind(2), Cf.unbind(2)
        ys = []
        for i in range(T):
            x = dA_t[i] * x + dBu_t[i]
            ys.append(torch.einsum("bdn,bn->bd", x, C_t[i]))
        y = torch.stack(ys, dim=2) if T > 0 else uf.new_zeros((batch, uf.shape[1], 0))
        if D is not None:
            y = y + uf * D[sl].float()[None, :, None]
        if z is not None:
            y = y * F.silu(z[:, sl].float())
        outs.append(y)
        lasts.append(x)
    out = torch.cat(outs, dim=1).to(dtype_in)
    if return_last_state:
        return out, torch.cat(lasts, dim=1)
    return out


def selective_state_update_ref(state, x, dt, A, B, C, D=None, z=None, dt_bias=None,
                               dt_softplus=False):
    """One-token selective scan.  state (batch, Di, N) is updated IN PLACE.

    x, dt, z: (batch, Di); A: (Di, N); B, C: (batch, N); D, dt_bias: (Di,).  Returns (batch, Di).
    """
    dtf = dt.float()
    if dt_bias is not None:
        dtf = dtf + dt_bias.float()
    if dt_softplus:
        dtf = F.softplus(dtf)
    dA = torch.exp(dtf[:, :, None] * A.float()[None])
    dBx = dtf[:, :, None] * B.float()[:, None, :] * x.float()[:, :, None]
    state.copy_((state.float() * dA + dBx).to(state.dtype))
    out = torch.einsum("bdn,bn->bd", state.float(), C.float())
    if D is not None:
        out = out + x.float() * D.float()
    if z is not None:
        out = out * F.silu(z.float())
    return out.to(x.dtype)


def causal_conv1d_ref(x, weight, bias=None, initial_states=None, return_final_states=False,
                      activation=None):
    """Depthwise causal conv.  x (batch, Di, T); weight (Di, W); bias (Di,);
    initial_states (batch, Di, W-1) = the W-1 inputs that precede x[..., 0]."""
    if activation not in (None, "silu", "swish"):
        raise NotImplementedError("activation must be None, silu or swish")
    dtype_in = x.dtype
    xf = x.to(weight.dtype)
    T = xf.shape[-1]
    dim, width = weight.shape
    if initial_states is None:
        full = xf
        out = F.conv1d(full, weight.unsqueeze(1), bias, padding=width - 1, groups=dim)
    else:
        full = torch.cat([initial_states.to(weight.dtype), xf], dim=-1)
        out = F.conv1d(full, weight.unsqueeze(1), bias, padding=0, groups=dim)
    out = out[..., :T]
    out = (out if activation is None else F.silu(out)).to(dtype_in)
    if not return_final_states:
        return out
    final = F.pad(full, (width - 1 - full.shape[-1], 0))[..., -(width - 1):].to(dtype_in)
    return out, final


def causal_conv1d_update_ref(x, conv_state, weight, bias=None, activation=None):
    """Single-token conv.  x (batch, Di); conv_state (batch, Di, W) rolled left IN PLACE
    with x written to the last column (the ``Mamba.step`` convention)."""
    if activation not in (None, "silu", "swish"):
        raise NotImplementedError("activation must be None, silu or swish")
    dtype_in = x.dtype
    conv_state.copy_(torch.roll(conv_state, shifts=-1, dims=-1))
    conv_state[:, :, -1] = x.to(conv_state.dtype)
    out = torch.sum(conv_state.to(weight.dtype) * weight[None], dim=-1)
    if bias is not None:
        out = out + bias
    return (out if activation is None else F.silu(out)).to(dtype_in)
