"""``Mamba`` block -- drop-in for ``from mamba_ssm import Mamba`` as the reference uses it.

The reference builds ``Mamba(d_model)`` (``mamba_decoder.py:29``) and calls it as
``out, new_state = self.mamba(h)`` / ``self.mamba(h, mamba_state)`` (``:61,63``, contract documented
at ``:9-15``).  Parameter names, shapes and initialisation are upstream's
(``mamba_ssm.modules.mamba_simple.Mamba``), so state_dicts interchange.

    forward(h, state=None) -> (out, (conv_state, ssm_state))
        state is None            full sequence from the zero state
        state given, T == 1      ``Mamba.step``: one fused launch, states updated IN PLACE
        state given, T  > 1      continue a sequence (prompt, then decode)

    conv_state (batch, d_inner, d_conv)  activation dtype: last d_conv pre-conv inputs, zero padded
    ssm_state  (batch, d_inner, d_state) fp32

Conv, scan and the fused step run in the sm_100a library; the four projections are GEMMs.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops


def compute_dtype(t):
    """Activation / GEMM-operand dtype: the autocast dtype when autocast is on, else t's dtype."""
    if torch.is_autocast_enabled("cuda"):
        return torch.get_autocast_dtype("cuda")
    return t.dtype


class _ChannelMajorLinear(torch.autograd.Function):
    """y[b] = W @ x[b]:  W (O, I), x (B, I, T) in any cuBLAS-addressable striding -> y (B, O, T)
    contiguous.  One strided-batched GEMM (W broadcast with batch stride 0): the projection lands
    directly in the channel-major layout conv/scan want, with no transposing copy in either
    direction (``torch.matmul`` would fold the batch and hand back a transposed view)."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=None)
    def forward(ctx, weight, x, dtype):
        Wc = weight.to(dtype)
        xc = x if x.dtype == dtype else x.to(dtype)
        ctx.save_for_backward(Wc, xc)
        ctx.wdtype = weight.dtype
        ctx.xdtype = x.dtype
        with torch.autocast("cuda", enabled=False):
            return torch.bmm(Wc.unsqueeze(0).expand(xc.shape[0], -1, -1), xc)

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, dy):
        Wc, xc = ctx.saved_tensors
        B = xc.shape[0]
        dy = dy.to(Wc.dtype)
        with torch.autocast("cuda", enabled=False):
            dx = torch.bmm(Wc.t().unsqueeze(0).expand(B, -1, -1), dy)
            dW = torch.bmm(dy, xc.transpose(1, 2)).sum(0)
        return dW.to(ctx.wdtype), dx.to(ctx.xdtype), None


class _TokenMajorLinear(torch.autograd.Function):
    """out[b] = y[b]^T @ W^T:  y (B, I, T) channel-major, W (O, I) -> out (B, T, O) contiguous;
    the backward hands dy back channel-major.  Again one strided-batched GEMM each, no copies."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=None)
    def forward(ctx, weight, y, dtype):
        Wc = weight.to(dtype)
        yc = y if y.dtype == dtype else y.to(dtype)
        ctx.save_for_backward(Wc, yc)
        ctx.wdtype = weight.dtype
        ctx.ydtype = y.dtype
        with torch.autocast("cuda", enabled=False):
            return torch.bmm(yc.transpose(1, 2), Wc.t().unsqueeze(0).expand(yc.shape[0], -1, -1))

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, dout):
        Wc, yc = ctx.saved_tensors
        B = yc.shape[0]
        dout = dout.to(Wc.dtype)
        with torch.autocast("cuda", enabled=False):
            dy = torch.bmm(Wc.t().unsqueeze(0).expand(B, -1, -1), dout.transpose(1, 2))
            dW = torch.bmm(dout.transpose(1, 2), yc.transpose(1, 2)).sum(0)
        return dW.to(ctx.wdtype), dy.to(ctx.ydtype), None


class Mamba(nn.Module):
    def __init__(self, d_model, d_state=16, d_conv=4, expand=2, dt_rank="auto", dt_min=0.001,
                 dt_max=0.1, dt_init="random", dt_scale=1.0, dt_init_floor=1e-4, conv_bias=True,
                 bias=False, use_fast_path=True, layer_idx=None, device=None, dtype=None):
        factory = {"device": device, "dtype": dtype}
        super().__init__()
        self.d_model, self.d_state, self.d_conv, self.expand = d_model, d_state, d_conv, expand
        self.d_inner = int(expand * d_model)
        self.dt_rank = math.ceil(d_model / 16) if dt_rank == "auto" else dt_rank
        self.use_fast_path = use_fast_path
        self.layer_idx = layer_idx
        Di, R, N = self.d_inner, self.dt_rank, d_state

        self.in_proj = nn.Linear(d_model, 2 * Di, bias=bias, **factory)
        self.conv1d = nn.Conv1d(Di, Di, kernel_size=d_conv, groups=Di, padding=d_conv - 1,
                                bias=conv_bias, **factory)
        self.activation = "silu"
        self.x_proj = nn.Linear(Di, R + 2 * N, bias=False, **factory)
        self.dt_proj = nn.Linear(R, Di, bias=True, **factory)

        std = R ** -0.5 * dt_scale
        if dt_init == "constant":
            nn.init.constant_(self.dt_proj.weight, std)
        elif dt_init == "random":
            nn.init.uniform_(self.dt_proj.weight, -std, std)
        else:
            raise NotImplementedError
        dt = torch.exp(torch.rand(Di, **factory) * (math.log(dt_max) - math.log(dt_min))
                       + math.log(dt_min)).clamp(min=dt_init_floor)
        with torch.no_grad():
            self.dt_proj.bias.copy_(dt + torch.log(-torch.expm1(-dt)))  # softplus^-1
        self.dt_proj.bias._no_reinit = True

        A = torch.arange(1, N + 1, dtype=torch.float32, device=device).repeat(Di, 1).contiguous()
        self.A_log = nn.Parameter(torch.log(A))
        self.A_log._no_weight_decay = True
        self.D = nn.Parameter(torch.ones(Di, device=device))
        self.D._no_weight_decay = True
        self.out_proj = nn.Linear(Di, d_model, bias=bias, **factory)
        self._step_cache = None

    # -- state helpers ---------------------------------------------------------------------------
    def allocate_inference_cache(self, batch_size, max_seqlen=None, dtype=None):
        dev = self.out_proj.weight.device
        cdt = self.conv1d.weight.dtype if dtype is None else dtype
        return (torch.zeros(batch_size, self.d_inner, self.d_conv, device=dev, dtype=cdt),
                torch.zeros(batch_size, self.d_inner, self.d_state, device=dev, dtype=torch.float32))

    def _step_weights(self, dtype):
        """Weights of the step path in the layouts/dtypes the fused kernel wants, rebuilt only when
        a parameter changed (version counters)."""
        params = (self.in_proj.weight, self.conv1d.weight, self.conv1d.bias, self.x_proj.weight,
                  self.dt_proj.weight, self.dt_proj.bias, self.A_log, self.D, self.out_proj.weight)
        key = (dtype, tuple(None if p is None else (p.data_ptr(), p._version) for p in params))
        if self._step_cache is None or self._step_cache[0] != key:
            with torch.no_grad():
                w = {
                    "in_proj": self.in_proj.weight.to(dtype).contiguous(),
                    "in_bias": None if self.in_proj.bias is None else self.in_proj.bias.to(dtype),
                    "conv_w": self.conv1d.weight.squeeze(1).float().contiguous(),
                    "conv_b": None if self.conv1d.bias is None else self.conv1d.bias.float().contiguous(),
                    "x_proj": self.x_proj.weight.to(dtype).contiguous(),
                    "dt_proj": self.dt_proj.weight.to(dtype).contiguous(),
                    "dt_bias": self.dt_proj.bias.float().contiguous(),
                    "A": (-torch.exp(self.A_log.float())).contiguous(),
                    "D": self.D.float().contiguous(),
                    "out_proj": self.out_proj.weight.to(dtype).contiguous(),
                    "out_bias": None if self.out_proj.bias is None else self.out_proj.bias.to(dtype),
                }
            self._step_cache = (key, w)
        return self._step_cache[1]

    # -- forward -----------------------------------------------------------------------------------
    def forward(self, hidden_states, state=None):
        """hidden_states (batch, T, d_model) -> (out (batch, T, d_model), (conv_state, ssm_state))."""
        if not hidden_states.is_cuda:
            raise RuntimeError("Mamba (mamba_tts_project_b200) is CUDA-only: there is no CPU path")
        if state is not None and hidden_states.shape[1] == 1:
            conv_state, ssm_state = state
            out = self.step(hidden_states, conv_state, ssm_state)[0]
            return out, (conv_state, ssm_state)
        return self._forward_sequence(hidden_states, state)

    def _forward_sequence(self, h, state):
        Bsz, T, _ = h.shape
        R, N, W = self.dt_rank, self.d_state, self.d_conv
        cdt = compute_dtype(h)
        if state is None and self.in_proj.bias is None and self.out_proj.bias is None \
                and self.conv1d.bias is not None and T > 0 and Bsz > 0 \
                and cdt in (torch.float32, torch.bfloat16):
            # the training hot path: the whole block body as one autograd node
            out, new_conv, last = ops.mamba_block_fn(
                h, self.in_proj.weight, self.conv1d.weight.squeeze(1), self.conv1d.bias,
                self.x_proj.weight, self.dt_proj.weight, self.dt_proj.bias, self.A_log, self.D,
                self.out_proj.weight, cdt)
            return out, (new_conv, last)
        # channel-major projection: (2Di, D) @ (B, D, T) -> (B, 2Di, T), no transposing copy
        xz = _ChannelMajorLinear.apply(self.in_proj.weight, h.transpose(1, 2), cdt)
        if self.in_proj.bias is not None:
            xz = xz + self.in_proj.bias.to(xz.dtype)[:, None]
        x, z = xz.chunk(2, dim=1)

        prev_conv, h0 = (None, None) if state is None else state
        with torch.no_grad():
            xd = x.detach()
            if prev_conv is None:
                new_conv = F.pad(xd[..., -W:], (max(0, W - T), 0)) if T < W else xd[..., -W:].clone()
            else:
                new_conv = torch.cat([prev_conv.to(xd.dtype), xd], dim=-1)[..., -W:].contiguous()

        w2d = self.conv1d.weight.squeeze(1)
        init = None if prev_conv is None else prev_conv[..., 1:]
        xc = ops.causal_conv1d_fn(x, w2d, self.conv1d.bias, initial_states=init,
                                  activation=self.activation)
        x_dbl = _ChannelMajorLinear.apply(self.x_proj.weight, xc, cdt)           # (B, R + 2N, T)
        delta = _ChannelMajorLinear.apply(self.dt_proj.weight, x_dbl[:, :R], cdt)  # (B, Di, T)
        A = -torch.exp(self.A_log.float())
        y, last = ops.selective_scan_fn(xc, delta, A, x_dbl[:, R:R + N], x_dbl[:, R + N:],
                                        self.D.float(), z=z, delta_bias=self.dt_proj.bias.float(),
                                        delta_softplus=True, return_last_state=True,
                                        initial_state=h0)
        out = _TokenMajorLinear.apply(self.out_proj.weight, y, cdt)               # (B, T, D)
        if self.out_proj.bias is not None:
            out = out + self.out_proj.bias.to(out.dtype)
        return out, (new_conv, last)

    def step(self, hidden_states, conv_state, ssm_state):
        """Upstream ``Mamba.step`` signature: hidden_states (batch, 1, d_model); both states are
        updated in place.  Returns (out (batch, 1, d_model), conv_state, ssm_state)."""
        assert hidden_states.shape[1] == 1, "Only support decoding with 1 token at a time for now"
        dtype = hidden_states.dtype
        w = self._step_weights(dtype)
        if conv_state.dtype != dtype:
            raise RuntimeError("conv_state dtype must match the activations")
        xz = F.linear(hidden_states[:, 0], w["in_proj"], w["in_bias"])
        y = ops.mamba_decode_step(xz, conv_state, ssm_state, w["conv_w"], w["conv_b"], w["x_proj"],
                                  w["dt_proj"], w["dt_bias"], w["A"], w["D"])
        out = F.linear(y, w["out_proj"], w["out_bias"])
        return out.unsqueeze(1), conv_state, ssm_state
