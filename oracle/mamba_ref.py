"""CPU restatement of the Mamba-1 block (TEST INFRASTRUCTURE, see oracle/__init__.py).

Follows the non-fused branch of ``mamba_ssm.modules.mamba_simple.Mamba.forward`` and the
non-Triton branch of ``Mamba.step`` (published algorithm, package un-pinned by the reference;
constructed at ``/root/reference/mamba_decoder.py:29`` as ``Mamba(d_model)`` and called at
``:61`` / ``:63``).  Parameter names and shapes are upstream's, so state_dicts interchange.

Contract implemented (the one the reference decoder documents at ``mamba_decoder.py:9-15``
and relies on at ``:61,63`` -- upstream's ``forward`` returns a single tensor):

    out, (conv_state, ssm_state) = block(h)            # full sequence from zero state
    out, (conv_state, ssm_state) = block(h, state)     # continue: T == 1 is ``Mamba.step``

``conv_state`` (batch, Di, W) = last W pre-conv inputs, left zero padded; ``ssm_state``
(batch, Di, N) fp32 = h after the last token.  The oracle never mutates the state it is given.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from .ssm_ref import causal_conv1d_ref, selective_scan_ref


class MambaRef(nn.Module):
    def __init__(self, d_model, d_state=16, d_conv=4, expand=2, dt_rank="auto",
                 dt_min=0.001, dt_max=0.1, dt_init="random", dt_scale=1.0,
                 dt_init_floor=1e-4, conv_bias=True, bias=False, scan_dim_block=256):
        super().__init__()
        self.d_model, self.d_state, self.d_conv, self.expand = d_model, d_state, d_conv, expand
        self.d_inner = int(expand * d_model)
        self.dt_rank = math.ceil(d_model / 16) if dt_rank == "auto" else dt_rank
        self.scan_dim_block = scan_dim_block
        Di, R, N = self.d_inner, self.dt_rank, d_state

        self.in_proj = nn.Linear(d_model, 2 * Di, bias=bias)
        self.conv1d = nn.Conv1d(Di, Di, kernel_size=d_conv, groups=Di, padding=d_conv - 1,
                                bias=conv_bias)
        self.x_proj = nn.Linear(Di, R + 2 * N, bias=False)
        self.dt_proj = nn.Linear(R, Di, bias=True)

        # dt_proj init: weight U(+-R^-0.5) ("random") or constant; bias = softplus^-1(dt),
        # dt log-uniform in [dt_min, dt_max] clamped at dt_init_floor.
        std = R ** -0.5 * dt_scale
        if dt_init == "constant":
            nn.init.constant_(self.dt_proj.weight, std)
        elif dt_init == "random":
            nn.init.uniform_(self.dt_proj.weight, -std, std)
        else:
            raise NotImplementedError(dt_init)
        dt = torch.exp(torch.rand(Di) * (math.log(dt_max) - math.log(dt_min))
                       + math.log(dt_min)).clamp(min=dt_init_floor)
        with torch.no_grad():
            self.dt_proj.bias.copy_(dt + torch.log(-torch.expm1(-dt)))

        # S4D-real init: A[d, n] = -(n + 1)
        self.A_log = nn.Parameter(torch.log(torch.arange(1, N + 1, dtype=torch.float32))
                                  .repeat(Di, 1).contiguous())
        self.D = nn.Parameter(torch.ones(Di))
        self.out_proj = nn.Linear(Di, d_model, bias=bias)

    def allocate_state(self, batch, dtype=torch.float32):
        return (torch.zeros(batch, self.d_inner, self.d_conv, dtype=dtype),
                torch.zeros(batch, self.d_inner, self.d_state, dtype=torch.float32))

    def forward(self, hidden_states, state=None):
        """hidden_states (batch, T, d_model) -> (out (batch, T, d_model), (conv_state, ssm_state))."""
        Bsz, T, _ = hidden_states.shape
        R, N, W = self.dt_rank, self.d_state, self.d_conv
        A = -torch.exp(self.A_log.float())

        xz = self.in_proj(hidden_states).transpose(1, 2)            # (B, 2Di, T)
        x, z = xz.chunk(2, dim=1)

        if state is None:
            prev_conv = x.new_zeros(Bsz, self.d_inner, W)
            h0 = None
        else:
            prev_conv, h0 = state
        window = torch.cat([prev_conv.to(x.dtype), x], dim=-1)        # (B, Di, W + T)
        new_conv_state = window[..., -W:].clone()

        w2d = self.conv1d.weight.squeeze(1)                          # (Di, W)
        x = causal_conv1d_ref(x, w2d, self.conv1d.bias, initial_states=prev_conv[..., 1:],
                              activation="silu")

        x_dbl = self.x_proj(x.transpose(1, 2))                        # (B, T, R + 2N)
        dt, Bm, Cm = torch.split(x_dbl, [R, N, N], dim=-1)
        dt = (dt @ self.dt_proj.weight.t()).transpose(1, 2)           # bias added inside the scan
        Bm = Bm.transpose(1, 2).contiguous()
        Cm = Cm.transpose(1, 2).contiguous()

        y, last = selective_scan_ref(x, dt, A, Bm, Cm, self.D.float(), z=z,
                                     delta_bias=self.dt_proj.bias.float(), delta_softplus=True,
                                     return_last_state=True, initial_state=h0,
                                     dim_block=self.scan_dim_block)
        out = self.out_proj(y.transpose(1, 2))
        return out, (new_conv_state, last)

    def step(self, hidden_states, conv_state, ssm_state):
        """``Mamba.step`` restated verbatim (roll / dot / softplus / exp / einsum); states are
        updated in place like upstream.  hidden_states (batch, 1, d_model)."""
        assert hidden_states.shape[1] == 1
        R, N = self.dt_rank, self.d_state
        xz = self.in_proj(hidden_states.squeeze(1))
        x, z = xz.chunk(2, dim=-1)
        conv_state.copy_(torch.roll(conv_state, shifts=-1, dims=-1))
        conv_state[:, :, -1] = x
        x = torch.sum(conv_state * self.conv1d.weight.squeeze(1), dim=-1)
        if self.conv1d.bias is not None:
            x = x + self.conv1d.bias
        x = F.silu(x)
        x_db = self.x_proj(x)
        dt, Bm, Cm = torch.split(x_db, [R, N, N], dim=-1)
        dt = F.linear(dt, self.dt_proj.weight)
        A = -torch.exp(self.A_log.float())
        dt = F.softplus(dt + self.dt_proj.bias)
        dA = torch.exp(torch.einsum("bd,dn->bdn", dt, A))
        dB = torch.einsum("bd,bn->bdn", dt, Bm)
        ssm_state.copy_(ssm_state * dA + x[:, :, None] * dB)
        y = torch.einsum("bdn,bn->bd", ssm_state, Cm)
        y = y + self.D * x
        y = y * F.silu(z)
        return self.out_proj(y).unsqueeze(1), conv_state, ssm_state
