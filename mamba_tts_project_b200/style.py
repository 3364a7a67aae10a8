"""Caller-side components next to the decoder: ``/root/reference/style_cross_attention.py``.

Same classes, constructor arguments, module tree / state_dict keys and return values as the reference, so its
checkpoints load unchanged and ``train.py:58,206`` switches by changing one import:

``LengthRegulator`` (``:144-213``)  same ``forward(hidden, durations, max_len=None) -> (expanded, output_lengths)``
    and ``forward_with_target`` -- one CUDA launch (``mtts_length_regulate_fwd``) instead of a Python loop with one
    host sync per phoneme, differentiable with respect to ``hidden``.
``StyleProjection`` (``:16-66``), ``StyleTextCrossAttention`` (``:69-141``), ``StyleDecoderCrossAttention``
    (``:215-286``), ``StyleConditioningPipeline`` (``:289-354``).

What differs underneath.  Both cross-attention blocks attend to ONE key / value token (the projected style
vector), so the softmax is the constant 1 and ``nn.MultiheadAttention``'s output does not depend on the query:
``attn_out[b, t] = W_o (W_v V_b + b_v) + b_o`` for every ``t``.  The blocks therefore compute that vector once per
batch element (the query / key projections and the T x 1 score matrix of the reference are never formed; their
weights receive the same zero gradient the reference gives them) and the rest of the block -- residual add +
LayerNorm, FFN, residual add + LayerNorm -- runs on the decoder's kernels: ``mtts_add_layernorm_*`` (the attention row
broadcast into its delta operand; the second Linear's bias rides in as ``delta_bias``), ``mtts_gemm`` (tcgen05) with
bias + GELU in the epilogue under bf16 autocast, library GEMMs + ``mtts_bias_gelu`` in fp32.  Attention-weight dropout
(training mode, ``dropout > 0``) is the only thing that makes the output depend on ``t``; it is applied to the single
weight per (batch, head, position) exactly where ``nn.MultiheadAttention`` applies it.  CUDA only, like the rest of the
package.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import dense, ops
from .decoder import CrossAttention
from .mamba import compute_dtype


class LengthRegulator(nn.Module):
    def __init__(self):
        super().__init__()

    def forward(self, hidden, durations, max_len=None):
        return ops.length_regulate(hidden, durations, max_len=max_len)

    def forward_with_target(self, hidden, target_durations):
        return self.forward(hidden, target_durations)


class StyleProjection(nn.Module):
    """``style_emb (B, d_style) -> K, V (B, 1, d_model)``: Linear + LayerNorm + Dropout each (``:16-66``)."""

    def __init__(self, d_style, d_model, dropout=0.1):
        super().__init__()
        self.d_style, self.d_model = d_style, d_model
        self.key_proj = nn.Sequential(nn.Linear(d_style, d_model), nn.LayerNorm(d_model), nn.Dropout(dropout))
        self.value_proj = nn.Sequential(nn.Linear(d_style, d_model), nn.LayerNorm(d_model), nn.Dropout(dropout))

    def forward(self, style_emb):
        # (B, d_style): a few rows -- library GEMV + LayerNorm, not worth a kernel of its own
        return self.key_proj(style_emb).unsqueeze(1), self.value_proj(style_emb).unsqueeze(1)


class _StyleCrossAttentionBlock(nn.Module):
    """Shared body of the two blocks (the reference spells the same module twice, ``:69-141`` and ``:215-286``)."""

    def __init__(self, d_model, num_heads=8, dropout=0.1):
        super().__init__()
        self.d_model, self.num_heads = d_model, num_heads
        self.cross_attn = CrossAttention(d_model, num_heads)      # nn.MultiheadAttention's parameter names
        self.attn_dropout = dropout                               # (its dropout acts on the attention weights)
        self.norm = nn.LayerNorm(d_model)
        self.dropout = nn.Dropout(dropout)
        self.ffn = nn.Sequential(nn.Linear(d_model, d_model * 4), nn.GELU(), nn.Dropout(dropout),
                                 nn.Linear(d_model * 4, d_model), nn.Dropout(dropout))
        self.ffn_norm = nn.LayerNorm(d_model)

    def _attention_rows(self, B, T, style_V):
        """attn_out: (B, 1, E) when it is the same for every position, (B, T, E) under attention dropout."""
        ca = self.cross_attn
        E, H = ca.embed_dim, ca.num_heads
        v = F.linear(style_V.reshape(B, E), ca.in_proj_weight[2 * E:], ca.in_proj_bias[2 * E:])   # (B, E)
        if self.training and self.attn_dropout > 0:
            keep = F.dropout(torch.ones(B, H, T, 1, device=v.device, dtype=v.dtype), self.attn_dropout, True)
            o = (keep * v.view(B, H, 1, E // H)).transpose(1, 2).reshape(B, T, E)
            return F.linear(o, ca.out_proj.weight, ca.out_proj.bias)
        return F.linear(v, ca.out_proj.weight, ca.out_proj.bias).unsqueeze(1)

    def _block(self, hidden, style_K, style_V):
        B, T, E = hidden.shape
        if style_K.shape[1] != 1 or style_V.shape[1] != 1:
            raise NotImplementedError("the style blocks attend to a single style token (style_cross_attention.py:52-66)")
        cdt = compute_dtype(hidden)
        attn = self.dropout(self._attention_rows(B, T, style_V))
        x = hidden.float()
        delta = attn.expand(B, T, E).to(cdt).contiguous() if attn.shape[1] == 1 else attn.to(cdt)
        # hidden = norm(hidden + attn_out): post-LN, the normalised rows ARE the next residual stream
        _, h = ops.add_layernorm(x, delta, self.norm.weight, self.norm.bias, self.norm.eps, out_dtype=torch.float32)
        hc = h if cdt == torch.float32 else h.to(cdt)
        w1, b1, w2, b2 = self.ffn[0].weight, self.ffn[0].bias, self.ffn[3].weight, self.ffn[3].bias
        drop = self.training and self.ffn[2].p > 0
        if dense.tc_enabled(cdt) and E % 8 == 0 and not drop:
            f = dense.ffn(hc, w1, b1, w2)                          # tcgen05, bias + GELU in the epilogue
        else:
            a = ops.bias_gelu(F.linear(hc, w1.to(cdt)), b1)
            f = F.linear(self.ffn[2](a), w2.to(cdt))
        if self.training and self.ffn[4].p > 0:                    # dropout after the second Linear (bias included)
            f = self.ffn[4](f + b2.to(f.dtype))
            b2 = None
        _, out = ops.add_layernorm(h, f, self.ffn_norm.weight, self.ffn_norm.bias, self.ffn_norm.eps,
                                   out_dtype=torch.float32 if hidden.dtype == torch.float32 else cdt, delta_bias=b2)
        return out


class StyleTextCrossAttention(_StyleCrossAttentionBlock):
    """Cross-attention #1, text x style, before the duration predictor (``:69-141``)."""

    def forward(self, text_hidden, style_K, style_V, text_mask=None):
        return self._block(text_hidden, style_K, style_V)


class StyleDecoderCrossAttention(_StyleCrossAttentionBlock):
    """Cross-attention #2, upsampled frames x style, before the codec generator (``:215-286``)."""

    def forward(self, upsampled_hidden, style_K, style_V, frame_mask=None):
        return self._block(upsampled_hidden, style_K, style_V)


class StyleConditioningPipeline(nn.Module):
    """``style_cross_attention.py:289-354``: style projection -> block #1 -> length regulator -> block #2."""

    def __init__(self, d_style=256, d_model=512, num_heads=8, dropout=0.1):
        super().__init__()
        self.style_proj = StyleProjection(d_style, d_model, dropout)
        self.cross_attn_1 = StyleTextCrossAttention(d_model, num_heads, dropout)
        self.cross_attn_2 = StyleDecoderCrossAttention(d_model, num_heads, dropout)
        self.length_regulator = LengthRegulator()

    def forward(self, text_hidden, style_emb, durations, text_mask=None, max_frame_len=None):
        style_K, style_V = self.style_proj(style_emb)
        styled_text = self.cross_attn_1(text_hidden, style_K, style_V, text_mask)
        upsampled, output_lengths = self.length_regulator(styled_text, durations, max_len=max_frame_len)
        styled_frames = self.cross_attn_2(upsampled, style_K, style_V)
        return styled_frames, output_lengths, style_K, style_V
