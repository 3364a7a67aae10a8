"""CPU restatement of the reference decoder (TEST INFRASTRUCTURE, see oracle/__init__.py).

Behaviour of ``/root/reference/mamba_decoder.py``:

* layer  (``:25-91``)  x += Mamba(LN(x)); x += MHA(LN(x), kv, kv, key_padding_mask=~mask);
                        x += FFN(gamma * LN(x) + beta), (gamma, beta) = chunk(tanh(Linear(z_style)))
* stack  (``:94-186``) tok + pos + quant embeddings, [ref || text] memory, layers, LN, head
* step   (``:188-256``) tok + pos only (no quant embedding -- SURVEY.md D5, reproduced as written)

Decisions on the reference's defects, fixed here once:

* D1  the Mamba block returns ``(out, state)`` -- the documented contract (``:9-15``), not the
      accidental tensor unpack.
* D3  masks are ``True = attend`` (what ``:70`` implements with ``~text_mask``).
* D4  3-D ``(B, Q, T)`` tokens: the file as written fails for Q > 1 (``pos`` is built for T, the
      sequence is Q*T).  The caller's own convention, ``train.py:123`` ``arange(T).repeat(Q)``,
      is used so the multi-quantizer path is defined; Q == 1 is unchanged.
* D8  attention weights are not returned (the reference discards them, ``:72,78``).

Module/parameter names mirror the reference so a state_dict moves between the reference, this
oracle and the CUDA decoder unchanged.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .mamba_ref import MambaRef


def _join_memory(text_hidden, text_mask, ref_hidden, ref_mask):
    """[ref || text] memory and its validity mask (``mamba_decoder.py:148-165`` / ``:226-241``)."""
    if ref_hidden is None:
        return text_hidden, text_mask
    B = text_hidden.shape[0]
    if ref_hidden.dim() != 3 or ref_hidden.shape[0] != B:
        raise AssertionError("ref_hidden must be (B, T_ref, d_model)")
    if ref_mask is None:
        ref_mask = torch.ones(B, ref_hidden.shape[1], dtype=torch.bool, device=ref_hidden.device)
    elif ref_mask.dim() != 2 or ref_mask.shape[0] != B:
        raise AssertionError("ref_mask must be (B, T_ref) bool")
    memory = torch.cat([ref_hidden, text_hidden], dim=1)
    mask = ref_mask if text_mask is None else torch.cat([ref_mask, text_mask], dim=1)
    return memory, mask


class MambaTTSDecoderLayerRef(nn.Module):
    def __init__(self, d_model, n_heads, d_ff, d_style, d_state=16, d_conv=4, expand=2):
        super().__init__()
        self.norm_mamba = nn.LayerNorm(d_model)
        self.mamba = MambaRef(d_model, d_state=d_state, d_conv=d_conv, expand=expand)
        self.norm_cross = nn.LayerNorm(d_model)
        self.cross_attn = nn.MultiheadAttention(embed_dim=d_model, num_heads=n_heads,
                                                batch_first=True)
        self.norm_ff = nn.LayerNorm(d_model)
        self.ff = nn.Sequential(nn.Linear(d_model, d_ff), nn.GELU(), nn.Linear(d_ff, d_model))
        self.style_mlp = nn.Sequential(nn.Linear(d_style, 2 * d_model), nn.Tanh())

    def forward(self, x, text_hidden, z_style, text_mask=None, mamba_state=None):
        mixed, new_state = self.mamba(self.norm_mamba(x), mamba_state)
        x = x + mixed

        pad = None if text_mask is None else ~text_mask
        attended, _ = self.cross_attn(query=self.norm_cross(x), key=text_hidden,
                                      value=text_hidden, key_padding_mask=pad)
        x = x + attended

        gamma, beta = self.style_mlp(z_style).chunk(2, dim=-1)
        x = x + self.ff(gamma.unsqueeze(1) * self.norm_ff(x) + beta.unsqueeze(1))
        return x, new_state


class MambaTTSDecoderRef(nn.Module):
    def __init__(self, vocab_size_audio, d_model=512, n_layers=8, n_heads=8, d_ff=2048,
                 d_style=256, max_len=8192, num_quantizers=1, d_state=16, d_conv=4, expand=2):
        super().__init__()
        self.vocab_size_audio = vocab_size_audio
        self.token_embed = nn.Embedding(vocab_size_audio, d_model)
        self.pos_embed = nn.Embedding(max_len, d_model)
        self.quant_embed = nn.Embedding(num_quantizers, d_model)
        self.layers = nn.ModuleList([
            MambaTTSDecoderLayerRef(d_model, n_heads, d_ff, d_style, d_state, d_conv, expand)
            for _ in range(n_layers)])
        self.norm_out = nn.LayerNorm(d_model)
        self.head = nn.Linear(d_model, vocab_size_audio)

    def forward(self, audio_tokens, text_hidden, z_style, text_mask=None, ref_hidden=None,
                ref_mask=None, return_states=False):
        if audio_tokens.dim() == 3:
            B, Q, T = audio_tokens.shape
            audio_tokens = audio_tokens.reshape(B, Q * T)
            quant_ids = torch.arange(Q, device=audio_tokens.device).repeat_interleave(T)
            quant_ids = quant_ids.unsqueeze(0).expand(B, -1)
            pos_ids = torch.arange(T, device=audio_tokens.device).repeat(Q)
        elif audio_tokens.dim() == 2:
            B, T = audio_tokens.shape
            quant_ids = torch.zeros_like(audio_tokens)
            pos_ids = torch.arange(T, device=audio_tokens.device)
        else:
            raise ValueError("audio_tokens must be (B, T) or (B, Q, T)")
        if text_mask is not None and (text_mask.dim() != 2 or text_mask.shape[0] != B):
            raise AssertionError("text_mask must be shape (B, T_text) with dtype=bool")
        memory, mask = _join_memory(text_hidden, text_mask, ref_hidden, ref_mask)

        x = (self.token_embed(audio_tokens) + self.pos_embed(pos_ids)[None]
             + self.quant_embed(quant_ids))
        states = []
        for layer in self.layers:
            x, st = layer(x, memory, z_style, text_mask=mask, mamba_state=None)
            states.append(st)
        logits = self.head(self.norm_out(x))
        return (logits, states) if return_states else logits

    def decode_step(self, last_token, text_hidden, z_style, mamba_states, step_index,
                    text_mask=None, ref_hidden=None, ref_mask=None):
        memory, mask = _join_memory(text_hidden, text_mask, ref_hidden, ref_mask)
        pos = self.pos_embed(torch.tensor([step_index], device=last_token.device))
        x = self.token_embed(last_token) + pos[None]
        new_states = []
        for i, layer in enumerate(self.layers):
            prev = None if mamba_states is None else mamba_states[i]
            x, st = layer(x, memory, z_style, text_mask=mask, mamba_state=prev)
            new_states.append(st)
        return self.head(self.norm_out(x)), new_states
