"""B200 decoder: host-side mirror of ``/root/reference/mamba_decoder.py``.

Same classes, constructor arguments, call signatures, module tree / state_dict keys and error
behaviour as the reference's ``MambaTTSDecoderLayer`` (``:25-91``) and ``MambaTTSDecoder``
(``:94-256``), so a reference checkpoint loads unchanged and a caller (``train.py:62-67,220-227``)
switches by changing one import.  What differs is underneath:

* ``forward``      teacher-forced path: Mamba conv + scan on the sm_100a library (autograd through
                   the recompute backward), attention without the discarded weights (D8).
* ``decode_step``  same signature and return value; K/V of [ref || text], the FiLM (gamma, beta)
                   and the dtype-converted weights are cached across steps instead of being
                   recomputed every step (D9) -- numerically identical.
* ``generate``     the missing caller of ``decode_step``: greedy / sampled loop,
                   one CUDA graph per step, states updated in place.

Decisions on the reference's defects are listed in DESIGN.md.
"""
from __future__ import annotations

import math
import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import dense, ops
from .mamba import Mamba, compute_dtype


def _join_memory(text_hidden, text_mask, ref_hidden, ref_mask):
    """[ref || text] memory and validity mask (``mamba_decoder.py:148-165,226-241``); True = attend."""
    if ref_hidden is None:
        return text_hidden, text_mask
    B = text_hidden.shape[0]
    assert ref_hidden.dim() == 3 and ref_hidden.shape[0] == B, "ref_hidden must be (B, T_ref, d_model)"
    if ref_mask is None:
        ref_mask = torch.ones(B, ref_hidden.shape[1], dtype=torch.bool, device=ref_hidden.device)
    else:
        assert ref_mask.dim() == 2 and ref_mask.shape[0] == B, "ref_mask must be (B, T_ref) bool"
    memory = torch.cat([ref_hidden, text_hidden], dim=1)
    mask = ref_mask if text_mask is None else torch.cat([ref_mask, text_mask], dim=1)
    return memory, mask


class CrossAttention(nn.Module):
    """``nn.MultiheadAttention(embed_dim, num_heads, batch_first=True)`` with the same parameter
    names (``in_proj_weight``, ``in_proj_bias``, ``out_proj.{weight,bias}``) and arithmetic
    (``mamba_decoder.py:32-36,72-77``); never materialises the averaged weights (D8)."""

    def __init__(self, embed_dim, num_heads):
        super().__init__()
        assert embed_dim % num_heads == 0, "embed_dim must be divisible by num_heads"
        self.embed_dim, self.num_heads, self.head_dim = embed_dim, num_heads, embed_dim // num_heads
        self.in_proj_weight = nn.Parameter(torch.empty(3 * embed_dim, embed_dim))
        self.in_proj_bias = nn.Parameter(torch.zeros(3 * embed_dim))
        self.out_proj = nn.Linear(embed_dim, embed_dim)
        nn.init.xavier_uniform_(self.in_proj_weight)
        nn.init.constant_(self.out_proj.bias, 0.0)

    def project_kv(self, memory, dtype=None):
        E = self.embed_dim
        w, b = self.in_proj_weight, self.in_proj_bias
        if dtype is not None:
            w, b, memory = w.to(dtype), b.to(dtype), memory.to(dtype)
        kv = ops.linear(memory, w[E:], b[E:]) if dtype is None else F.linear(memory, w[E:], b[E:])
        # (one GEMM for both projections; the bias gradient is a streaming column sum)
        return kv[..., :E], kv[..., E:]

    def forward(self, query, memory, key_padding_mask=None, add_out_bias=True):
        """query (B, T, E); memory (B, T_kv, E); key_padding_mask (B, T_kv) True = IGNORE.
        ``add_out_bias=False`` leaves out_proj.bias to the caller (it is folded into the next fused
        residual-add + LayerNorm launch together with its gradient)."""
        B, T, E = query.shape
        H, dh = self.num_heads, self.head_dim
        Tk = memory.shape[1]
        if dense.tc_enabled(compute_dtype(query)) and E % 8 == 0 and dh % 8 == 0 and 0 < Tk <= dense.MAX_FUSED_KEYS \
                and T > 0:
            # bf16: projections, QK^T + masked softmax, PV and the whole backward on mtts_gemm (tcgen05)
            out = dense.cross_attention(query, memory, self.in_proj_weight, self.in_proj_bias,
                                        self.out_proj.weight,
                                        None if key_padding_mask is None else ~key_padding_mask, H)
            return out + self.out_proj.bias.to(out.dtype) if add_out_bias else out
        q = ops.linear(query, self.in_proj_weight[:E], self.in_proj_bias[:E])
        k, v = self.project_kv(memory)
        q = q.view(B, T, H, dh).transpose(1, 2)
        k = k.view(B, -1, H, dh).transpose(1, 2)
        v = v.view(B, -1, H, dh).transpose(1, 2)
        bias = None
        if key_padding_mask is not None:
            bias = torch.zeros(B, 1, 1, k.shape[2], dtype=q.dtype, device=q.device)
            bias.masked_fill_(key_padding_mask[:, None, None, :], float("-inf"))
        o = F.scaled_dot_product_attention(q, k, v, attn_mask=bias)
        o = o.transpose(1, 2).reshape(B, T, E)
        return F.linear(o, self.out_proj.weight, self.out_proj.bias if add_out_bias else None)


class MambaTTSDecoderLayer(nn.Module):
    def __init__(self, d_model, n_heads, d_ff, d_style, d_state=16, d_conv=4, expand=2):
        super().__init__()
        self.norm_mamba = nn.LayerNorm(d_model)
        self.mamba = Mamba(d_model, d_state=d_state, d_conv=d_conv, expand=expand)
        self.norm_cross = nn.LayerNorm(d_model)
        self.cross_attn = CrossAttention(d_model, n_heads)
        self.norm_ff = nn.LayerNorm(d_model)
        self.ff = nn.Sequential(nn.Linear(d_model, d_ff), nn.GELU(), nn.Linear(d_ff, d_model))
        self.style_mlp = nn.Sequential(nn.Linear(d_style, 2 * d_model), nn.Tanh())

    def forward_fused(self, x, delta, text_hidden, z_style, text_mask=None, mamba_state=None,
                      delta_bias=None, film=None):
        """The layer with every ``x = x + branch`` folded into the LayerNorm that follows it.

        x: fp32 residual stream (B, T, D); delta (+ delta_bias): branch output still to be added to x
        (or None).  Returns (x, delta_out, delta_bias_out, new_state): the layer's result is
        ``x + delta_out + delta_bias_out``."""
        cdt = compute_dtype(x)
        x, h = ops.add_layernorm(x, delta, self.norm_mamba.weight, self.norm_mamba.bias,
                                 self.norm_mamba.eps, out_dtype=cdt, delta_bias=delta_bias)
        h_mamba, new_state = self.mamba(h) if mamba_state is None else self.mamba(h, mamba_state)

        x, h = ops.add_layernorm(x, h_mamba, self.norm_cross.weight, self.norm_cross.bias,
                                 self.norm_cross.eps, out_dtype=cdt)
        key_padding_mask = None if text_mask is None else ~text_mask
        attn_out = self.cross_attn(h, text_hidden, key_padding_mask=key_padding_mask,
                                   add_out_bias=False)

        # film: (gamma, beta) already computed for all layers at once (MambaTTSDecoder.film_terms)
        gamma, beta = torch.chunk(self.style_mlp(z_style), 2, dim=-1) if film is None else film
        x, h = ops.add_layernorm(x, attn_out, self.norm_ff.weight, self.norm_ff.bias,
                                 self.norm_ff.eps, gamma=gamma, beta=beta, out_dtype=cdt,
                                 delta_bias=self.cross_attn.out_proj.bias)
        # ff[0].bias is added inside the GELU kernel (and its gradient is that kernel's column sum);
        # ff[2].bias rides with delta into the next LayerNorm
        d_ff = self.ff[0].out_features
        if dense.tc_enabled(cdt) and h.shape[-1] % 8 == 0 and d_ff % 8 == 0:
            # both GEMMs on mtts_gemm; bias + GELU in the first one's epilogue, GELU' in the backward's
            f = dense.ffn(h, self.ff[0].weight, self.ff[0].bias, self.ff[2].weight)
        else:
            f = F.linear(ops.bias_gelu(F.linear(h, self.ff[0].weight), self.ff[0].bias), self.ff[2].weight)
        return x, f, self.ff[2].bias, new_state

    def forward(self, x, text_hidden, z_style, text_mask=None, mamba_state=None):
        """Reference signature (``mamba_decoder.py:50-91``): returns (x, new_state)."""
        xs, delta, dbias, new_state = self.forward_fused(x.float(), None, text_hidden, z_style,
                                                         text_mask, mamba_state)
        return (xs + delta.float() + dbias.float()).to(x.dtype), new_state


class _LayerStepWeights:
    """Per-layer weights in the decode dtype + the generation-constant tensors (K, V, gamma, beta)."""

    def __init__(self, layer: MambaTTSDecoderLayer, dtype, memory, z_style):
        E = layer.cross_attn.embed_dim
        ca = layer.cross_attn
        f32 = torch.float32
        self.mamba = layer.mamba._step_weights(dtype)
        self.ln1 = (layer.norm_mamba.weight.detach().to(f32).contiguous(),
                    layer.norm_mamba.bias.detach().to(f32).contiguous(), layer.norm_mamba.eps)
        self.ln2 = (layer.norm_cross.weight.detach().to(f32).contiguous(),
                    layer.norm_cross.bias.detach().to(f32).contiguous(), layer.norm_cross.eps)
        self.ln3 = (layer.norm_ff.weight.detach().to(f32).contiguous(),
                    layer.norm_ff.bias.detach().to(f32).contiguous(), layer.norm_ff.eps)
        self.wq = ca.in_proj_weight.detach()[:E].to(dtype).contiguous()
        self.bq = ca.in_proj_bias.detach()[:E].to(dtype).contiguous()
        self.wo = ca.out_proj.weight.detach().to(dtype).contiguous()
        self.bo = ca.out_proj.bias.detach().to(dtype).contiguous()
        self.w1 = layer.ff[0].weight.detach().to(dtype).contiguous()
        self.b1 = layer.ff[0].bias.detach().to(dtype).contiguous()
        self.w2 = layer.ff[2].weight.detach().to(dtype).contiguous()
        self.b2 = layer.ff[2].bias.detach().to(dtype).contiguous()
        self.heads = ca.num_heads
        k, v = ca.project_kv(memory, dtype)
        self.k, self.v = k.contiguous(), v.contiguous()
        gb = layer.style_mlp(z_style.to(layer.style_mlp[0].weight.dtype)).to(dtype).float()
        gamma, beta = gb.chunk(2, dim=-1)
        self.gamma, self.beta = gamma.contiguous(), beta.contiguous()


class GenerationContext:
    """Everything that is constant over one generation (``decode_step``'s D9 recomputations)."""

    def __init__(self, decoder: "MambaTTSDecoder", text_hidden, z_style, text_mask=None,
                 ref_hidden=None, ref_mask=None, dtype=None, fused_projections=False):
        dtype = dtype if dtype is not None else text_hidden.dtype
        self.dtype = dtype
        # opt-in: LN + projection (+GELU) as one skinny_linear launch (8 launches per layer instead
        # of 14).  Measured on B200 at B = 64: 0.89 ms/step vs 0.74 ms/step for LayerNorm kernel +
        # library GEMM, so it is off by default until the kernel's latency chain is shortened.
        self.fused_projections = fused_projections
        with torch.no_grad():
            memory, mask = _join_memory(text_hidden, text_mask, ref_hidden, ref_mask)
            self.batch = memory.shape[0]
            self.mask = None if mask is None else mask.to(torch.uint8).contiguous()
            self.layers = [_LayerStepWeights(l, dtype, memory, z_style) for l in decoder.layers]
            f32 = torch.float32
            self.ln_out = (decoder.norm_out.weight.detach().to(f32).contiguous(),
                           decoder.norm_out.bias.detach().to(f32).contiguous(), decoder.norm_out.eps)
            self.head_w = decoder.head.weight.detach().to(dtype).contiguous()
            self.head_b = decoder.head.bias.detach().to(dtype).contiguous()
            # skinny_linear needs every plain-mode K to be <= 512 or a multiple of 512
            self.fused_ok = all(k <= 512 or k % 512 == 0 for k in
                                (decoder.layers[0].mamba.d_inner, decoder.layers[0].ff[0].out_features)) \
                and decoder.layers[0].mamba.d_inner % 16 == 0 \
                and decoder.layers[0].ff[0].out_features % 16 == 0
            # the front of the cross-attention branch (residual add + LN + q projection + attention) as one
            # launch where the kernel exists (8 heads x 64, t_kv <= 256).  MTTS_FUSED_ATTENTION=0 keeps the
            # separate launches, =2 runs the whole branch incl. out projection + LN + FiLM as one cluster
            # launch (the parity tests run all three).  Measured at B 64 inside the replayed graph:
            # 0: 749 us/step, 1: 674 us/step, 2: 704 us/step -- the per-(batch, head) out projection pulls
            # another 32 MB of weight rows through L2 and two cluster barriers, more than the two launches cost.
            self.fused_attention = int(os.environ.get("MTTS_FUSED_ATTENTION", "1")) \
                if ops.cross_attn_block_decode_supported(dtype, memory.shape[2], self.layers[0].heads,
                                                         memory.shape[1]) else 0
            # the step kernel pulls the layer's K / V into L2 for the attention that follows (MTTS_PREFETCH_KV=0: off)
            self.prefetch_kv = os.environ.get("MTTS_PREFETCH_KV", "1") != "0"
            self.tok = decoder.token_embed.weight.detach().float().contiguous()
            self.pos = decoder.pos_embed.weight.detach().float().contiguous()


def flatten_codes(codes):
    """(B, Q, T) codec ids -> the flattened (B, Q*T) order the decoder is trained on: all frames of quantizer 0,
    then quantizer 1, ... (``train.py:181-182``: ``codec.permute(0, 2, 1).reshape(B, -1)``)."""
    if codes.dim() != 3:
        raise ValueError("codes must be (B, Q, T)")
    return codes.reshape(codes.shape[0], -1)


def unflatten_codes(tokens, num_quantizers):
    """Inverse of ``flatten_codes``: (B, Q*T) generated ids -> (B, Q, T).  A length that is not a multiple of Q
    (generation stopped inside the last quantizer) is an error rather than a silent truncation."""
    if tokens.dim() != 2 or tokens.shape[1] % num_quantizers:
        raise ValueError(f"tokens must be (B, Q*T) with Q = {num_quantizers}; got {tuple(tokens.shape)}")
    return tokens.reshape(tokens.shape[0], num_quantizers, -1)


class MambaTTSDecoder(nn.Module):
    def __init__(self, vocab_size_audio, d_model=512, n_layers=8, n_heads=8, d_ff=2048, d_style=256,
                 max_len=8192, num_quantizers=1, d_state=16, d_conv=4, expand=2):
        super().__init__()
        self.vocab_size_audio = vocab_size_audio
        self.token_embed = nn.Embedding(vocab_size_audio, d_model)
        self.pos_embed = nn.Embedding(max_len, d_model)
        self.quant_embed = nn.Embedding(num_quantizers, d_model)
        self.layers = nn.ModuleList([
            MambaTTSDecoderLayer(d_model, n_heads, d_ff, d_style, d_state, d_conv, expand)
            for _ in range(n_layers)])
        self.norm_out = nn.LayerNorm(d_model)
        self.head = nn.Linear(d_model, vocab_size_audio)
        self._gen_key = None
        self._gen_ctx = None
        self.last_generate_events = None
        self.last_generate_lengths = None

    # ---- teacher-forced path (mamba_decoder.py:120-186) -----------------------------------------
    def forward(self, audio_tokens, text_hidden, z_style, text_mask=None, ref_hidden=None,
                ref_mask=None):
        if audio_tokens.dim() == 3:
            B, Q, T = audio_tokens.shape
            audio_tokens = audio_tokens.reshape(B, Q * T)
            quant_row = torch.arange(Q, device=audio_tokens.device).repeat_interleave(T)
            pos_ids = torch.arange(T, device=audio_tokens.device).repeat(Q)  # train.py:123 (D4)
        elif audio_tokens.dim() == 2:
            B, T = audio_tokens.shape
            quant_row = torch.zeros(T, dtype=torch.long, device=audio_tokens.device)
            pos_ids = torch.arange(T, device=audio_tokens.device)
        else:
            raise ValueError("audio_tokens must be (B, T) or (B, Q, T)")
        if text_mask is not None:
            assert text_mask.dim() == 2 and text_mask.shape[0] == B, (
                "text_mask must be shape (B, T_text) with dtype=bool")
        memory, mask = _join_memory(text_hidden, text_mask, ref_hidden, ref_mask)

        if self.token_embed.weight.dtype == torch.float32 and self.token_embed.weight.shape[1] % 4 == 0:
            # one launch: gather + (tok + pos) + quant, backward = scatter-add with the batch pre-summed for pos / quant
            x = ops.embed_sum(audio_tokens, pos_ids, quant_row, self.token_embed.weight, self.pos_embed.weight,
                              self.quant_embed.weight)
        else:
            x = (self.token_embed(audio_tokens) + self.pos_embed(pos_ids)[None]
                 + self.quant_embed(quant_row)[None]).float()
        delta, dbias = None, None
        films = self.film_terms(z_style)
        for layer, film in zip(self.layers, films):
            x, delta, dbias, _ = layer.forward_fused(x, delta, memory, z_style, text_mask=mask,
                                                     delta_bias=dbias, film=film)
        cdt = compute_dtype(x)
        _, h = ops.add_layernorm(x, delta, self.norm_out.weight, self.norm_out.bias,
                                 self.norm_out.eps, out_dtype=cdt, delta_bias=dbias)
        if dense.tc_enabled(cdt) and h.shape[-1] % 8 == 0 and self.vocab_size_audio % 8 == 0:
            return dense.linear(h, self.head.weight, self.head.bias)
        return self.head(h)

    def film_terms(self, z_style):
        """(gamma, beta) of every layer's ``style_mlp`` (``mamba_decoder.py:45-48,81-83``: Linear + Tanh on the same
        ``z_style``) from ONE contraction over the stacked weights instead of one small GEMM + tanh per layer (and
        three more per layer in the backward); -> list of ((B, D), (B, D)) fp32, contiguous."""
        L, B = len(self.layers), z_style.shape[0]
        D2 = self.layers[0].style_mlp[0].out_features
        w = torch.cat([l.style_mlp[0].weight for l in self.layers])            # (L * 2D, d_style)
        b = torch.cat([l.style_mlp[0].bias for l in self.layers])
        if dense.tc_enabled(compute_dtype(z_style)) and z_style.shape[1] % 8 == 0 and D2 % 8 == 0:
            g = dense.linear(z_style, w, b)                                     # tcgen05 (M = B rows)
        else:
            g = F.linear(z_style, w, b)
        g = torch.tanh(g.float()).view(B, L, 2, D2 // 2).permute(1, 2, 0, 3).contiguous()   # (L, 2, B, D)
        return [(g[i, 0], g[i, 1]) for i in range(L)]

    # ---- incremental path (mamba_decoder.py:188-256) --------------------------------------------
    def prepare_generation(self, text_hidden, z_style, text_mask=None, ref_hidden=None,
                           ref_mask=None, dtype=None, fused_projections=False):
        return GenerationContext(self, text_hidden, z_style, text_mask, ref_hidden, ref_mask, dtype,
                                 fused_projections)

    def allocate_states(self, batch, dtype):
        return [l.mamba.allocate_inference_cache(batch, dtype=dtype) for l in self.layers]

    def _context_for(self, text_hidden, z_style, text_mask, ref_hidden, ref_mask):
        """The cached ``GenerationContext`` of ``decode_step``: valid only for the very same input TENSORS
        (identity, not address: the entry keeps strong references to them, so the allocator cannot hand
        their storage to the next utterance while the entry lives) at the same ``_version``, and for
        unchanged parameters."""
        inputs = (text_hidden, z_style, text_mask, ref_hidden, ref_mask)
        versions = tuple(None if t is None else t._version for t in inputs)
        pkey = tuple((p.data_ptr(), p._version) for p in self.parameters())
        hit = (self._gen_key is not None
               and all(a is b for a, b in zip(self._gen_key[0], inputs))
               and self._gen_key[1] == versions and self._gen_key[2] == pkey)
        if not hit:
            self._gen_ctx = self.prepare_generation(text_hidden, z_style, text_mask, ref_hidden, ref_mask)
            self._gen_key = (inputs, versions, pkey)
        return self._gen_ctx

    def _step_core(self, ctx: GenerationContext, x, states):
        """x (B, d_model) fp32 residual stream of the new token -> logits (B, V); states in place."""
        if ctx.fused_projections and ctx.dtype == torch.bfloat16 and x.shape[0] <= 64 \
                and x.shape[1] in (128, 256, 512, 1024) and ctx.fused_ok:
            return self._step_core_fused(ctx, x, states)
        delta = None  # pending branch output, folded into the next LayerNorm launch
        dt = ctx.dtype
        for lw, (conv_state, ssm_state) in zip(ctx.layers, states):
            x, h = ops.add_layernorm(x, delta, lw.ln1[0], lw.ln1[1], lw.ln1[2], out_dtype=dt,
                                     inplace=True)
            xz = F.linear(h, lw.mamba["in_proj"], lw.mamba["in_bias"])
            y = ops.mamba_decode_step(xz, conv_state, ssm_state, lw.mamba["conv_w"],
                                      lw.mamba["conv_b"], lw.mamba["x_proj"], lw.mamba["dt_proj"],
                                      lw.mamba["dt_bias"], lw.mamba["A"], lw.mamba["D"],
                                      prefetch=(lw.k, lw.v) if ctx.prefetch_kv else None)
            m = F.linear(y, lw.mamba["out_proj"], lw.mamba["out_bias"])
            if ctx.fused_attention == 2:
                # residual add + LN + q projection + attention + out projection + residual add + LN + FiLM:
                # one cluster launch instead of five
                x, h = ops.cross_attn_block_decode(x, m, lw.ln2, lw.wq, lw.bq, lw.k, lw.v, lw.heads, lw.wo,
                                                   lw.bo, lw.ln3, mask=ctx.mask, gamma=lw.gamma, beta=lw.beta)
            elif ctx.fused_attention == 1:
                # residual add + LN + q projection + attention in one launch of independent (batch, head) CTAs
                x, a = ops.cross_attn_block_decode(x, m, lw.ln2, lw.wq, lw.bq, lw.k, lw.v, lw.heads,
                                                   mask=ctx.mask)
                o = F.linear(a, lw.wo, lw.bo)
                x, h = ops.add_layernorm(x, o, lw.ln3[0], lw.ln3[1], lw.ln3[2], gamma=lw.gamma,
                                         beta=lw.beta, out_dtype=dt, inplace=True)
            else:
                x, h = ops.add_layernorm(x, m, lw.ln2[0], lw.ln2[1], lw.ln2[2], out_dtype=dt,
                                         inplace=True)
                q = F.linear(h, lw.wq, lw.bq)
                a = ops.cross_attn_decode(q, lw.k, lw.v, lw.heads, mask=ctx.mask)
                o = F.linear(a, lw.wo, lw.bo)
                x, h = ops.add_layernorm(x, o, lw.ln3[0], lw.ln3[1], lw.ln3[2], gamma=lw.gamma,
                                         beta=lw.beta, out_dtype=dt, inplace=True)
            f = F.gelu(F.linear(h, lw.w1, lw.b1))
            delta = F.linear(f, lw.w2, lw.b2)
        _, h = ops.add_layernorm(x, delta, ctx.ln_out[0], ctx.ln_out[1], ctx.ln_out[2], out_dtype=dt,
                                 inplace=True)
        return F.linear(h, ctx.head_w, ctx.head_b)

    def _step_core_fused(self, ctx: GenerationContext, x, states):
        """bf16, batch <= 64: every (residual add + LayerNorm [+ FiLM] + projection [+ GELU]) is ONE
        skinny_linear launch -- 8 launches per layer instead of 14.  The fp32 residual stream
        ping-pongs between two buffers (x_out must not alias x inside a launch)."""
        cur, other = x, torch.empty_like(x)
        delta = None

        def ln_linear(ln, w, bias, gamma=None, beta=None, gelu=False):
            nonlocal cur, other, delta
            out = ops.skinny_linear(w, bias, x=cur, delta=delta, x_out=None if delta is None else other,
                                    ln_weight=ln[0], ln_bias=ln[1], eps=ln[2], gamma=gamma, beta=beta,
                                    gelu=gelu)
            if delta is not None:
                cur, other = other, cur
            return out

        for lw, (conv_state, ssm_state) in zip(ctx.layers, states):
            xz = ln_linear(lw.ln1, lw.mamba["in_proj"], lw.mamba["in_bias"])
            y = ops.mamba_decode_step(xz, conv_state, ssm_state, lw.mamba["conv_w"],
                                      lw.mamba["conv_b"], lw.mamba["x_proj"], lw.mamba["dt_proj"],
                                      lw.mamba["dt_bias"], lw.mamba["A"], lw.mamba["D"])
            delta = ops.skinny_linear(lw.mamba["out_proj"], lw.mamba["out_bias"], a=y)
            q = ln_linear(lw.ln2, lw.wq, lw.bq)
            a = ops.cross_attn_decode(q, lw.k, lw.v, lw.heads, mask=ctx.mask)
            delta = ops.skinny_linear(lw.wo, lw.bo, a=a)
            f = ln_linear(lw.ln3, lw.w1, lw.b1, gamma=lw.gamma, beta=lw.beta, gelu=True)
            delta = ops.skinny_linear(lw.w2, lw.b2, a=f)
        return ln_linear(ctx.ln_out, ctx.head_w, ctx.head_b)

    @torch.no_grad()
    def decode_step(self, last_token, text_hidden, z_style, mamba_states, step_index: int,
                    text_mask=None, ref_hidden=None, ref_mask=None):
        """Reference signature.  Returns (logits (B, 1, V), new_states); ``mamba_states`` entries may
        be None on the first step; given states are updated IN PLACE and returned."""
        if not last_token.is_cuda:
            raise RuntimeError("MambaTTSDecoder (mamba_tts_project_b200) is CUDA-only")
        ctx = self._context_for(text_hidden, z_style, text_mask, ref_hidden, ref_mask)
        B = last_token.shape[0]
        states = []
        for i, layer in enumerate(self.layers):
            st = None if mamba_states is None else mamba_states[i]
            states.append(st if st is not None else
                          layer.mamba.allocate_inference_cache(B, dtype=ctx.dtype))
        x = ctx.tok[last_token[:, 0]] + ctx.pos[step_index]   # no quant_embed here (reference D5)
        logits = self._step_core(ctx, x.float().contiguous(), states)
        return logits.unsqueeze(1), states

    @torch.no_grad()
    def generate(self, first_token, n_steps, text_hidden, z_style, text_mask=None, ref_hidden=None,
                 ref_mask=None, start_index=0, temperature=0.0, use_cuda_graph=True, dtype=None,
                 generator=None, eos_id=None, pad_id=0, check_every=64):
        """Autoregressive loop around ``decode_step`` (greedy when temperature == 0).

        first_token (B, 1) int64.  Returns tokens (B, n_steps) int64 (the generated ids).  With
        ``use_cuda_graph`` one decode step is captured once and replayed: the token, the position
        counter and the per-layer states live in static device buffers, nothing syncs the host.

        ``eos_id``: a row that produces it is finished -- the eos is kept, every later position of the row is
        ``pad_id`` (and pad is what the finished row feeds back, rows are independent); the loop stops early
        once all rows are done, looked at every ``check_every`` steps (the only host sync).  The per-row token
        counts (including the eos; ``n_steps`` for rows that never finished) are left in
        ``self.last_generate_lengths``.  The ids are in the flattened (B, Q*T) codec order of ``train.py:181-182``;
        ``unflatten_codes`` turns them into (B, Q, T)."""
        ctx = self.prepare_generation(text_hidden, z_style, text_mask, ref_hidden, ref_mask, dtype)
        B = first_token.shape[0]
        dev = first_token.device
        states = self.allocate_states(B, ctx.dtype)
        tok = first_token[:, 0].clone()
        pos = torch.full((1,), start_index, dtype=torch.long, device=dev)
        out = torch.empty(B, n_steps, dtype=torch.long, device=dev)
        col = torch.zeros(1, dtype=torch.long, device=dev)
        greedy = temperature == 0.0
        if not greedy and use_cuda_graph:
            use_cuda_graph = False  # sampling draws host-side RNG state per step

        if int(tok.min()) < 0 or int(tok.max()) >= ctx.tok.shape[0] or start_index + n_steps > ctx.pos.shape[0]:
            raise IndexError("first_token / positions outside the embedding tables")
        xbuf = torch.empty(B, ctx.tok.shape[1], dtype=torch.float32, device=dev)
        if greedy:
            col.fill_(-1)  # decode_embed bumps it at the start of every step
        lengths = torch.full((B,), -1, dtype=torch.long, device=dev) if eos_id is not None else None
        if eos_id is not None:
            out.fill_(pad_id)  # columns never reached after an early stop

        def one_step():
            if greedy:  # token plumbing as two library launches, counters stay on the device
                ops.decode_embed(tok, pos, ctx.tok, ctx.pos, xbuf, step=col)
                logits = self._step_core(ctx, xbuf, states)
                ops.decode_greedy(logits, tok, out=out, step=col, pos=pos, eos_id=eos_id, pad_id=pad_id,
                                  lengths=lengths)
                return
            x = (ctx.tok.index_select(0, tok) + ctx.pos.index_select(0, pos)).float()
            logits = self._step_core(ctx, x, states)
            probs = torch.softmax(logits.float() / temperature, dim=-1)
            nxt = torch.multinomial(probs, 1, generator=generator)[:, 0]
            if eos_id is not None:
                done = lengths >= 0
                lengths.copy_(torch.where(~done & (nxt == eos_id), col + 1, lengths))
                nxt = torch.where(done, torch.full_like(nxt, pad_id), nxt)
            tok.copy_(nxt)
            out.index_copy_(1, col, nxt[:, None])
            pos.add_(1)
            col.add_(1)

        def all_done(i):  # one host sync every check_every steps, only when an eos is in play
            return eos_id is not None and (i + 1) % check_every == 0 and bool((lengths >= 0).all())

        def finish(n_done):
            if lengths is not None:
                self.last_generate_lengths = torch.where(lengths >= 0, lengths, torch.full_like(lengths, n_done))
            else:
                self.last_generate_lengths = None
            return out

        ev0 = torch.cuda.Event(enable_timing=True)
        ev1 = torch.cuda.Event(enable_timing=True)
        if not use_cuda_graph:
            ev0.record()
            n_done = n_steps
            for i in range(n_steps):
                one_step()
                if all_done(i):
                    n_done = i + 1
                    break
            ev1.record()
            self.last_generate_events = (ev0, ev1, n_done)
            return finish(n_done)

        # warm up on a side stream (cuBLAS workspaces, lazy module loads), then restore the state
        snap = [(c.clone(), s.clone()) for c, s in states]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            one_step()
        torch.cuda.current_stream().wait_stream(side)
        for (c, s), (c0, s0) in zip(states, snap):
            c.copy_(c0)
            s.copy_(s0)
        tok.copy_(first_token[:, 0])
        pos.fill_(start_index)
        col.fill_(-1 if greedy else 0)
        if lengths is not None:
            lengths.fill_(-1)
            out.fill_(pad_id)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            one_step()
        # capture does not execute: state is still the initial one
        ev0.record()
        n_done = n_steps
        for i in range(n_steps):
            graph.replay()
            if all_done(i):
                n_done = i + 1
                break
        ev1.record()
        self.last_generate_events = (ev0, ev1, n_done)  # steady-state loop only (bench.py)
        return finish(n_done)
