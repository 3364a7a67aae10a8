"""The caller contract either side of the decoder's forward/backward:
what ``/root/reference/train.py`` does around ``decoder(...)``, restated for the B200 path.

    embed_codec_tokens   <- train.py:115-131   ref_hidden from the decoder's own embeddings
    codec_ce_loss        <- train.py:31-42     cross entropy with ignore_index = pad_id
    TrainStep            <- train.py:152-159, 220-235  fwd + loss + bwd (+ DP all-reduce) + clip + Adam

``TrainStep`` supports micro-batching (gradient accumulation) so that a global batch that does not fit
one GPU (BASELINE config C5: 24 x d1024, B 64, T 4096) keeps identical numerics at every GPU count:
the loss of each micro-batch is the SUM of token losses divided by the GLOBAL number of valid tokens.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def embed_codec_tokens(tokens_3d, decoder):
    """tokens_3d (B, Q, T) codec ids -> (ref_hidden (B, Q*T, d_model), mask (B, Q*T) True = pad)."""
    B, Q, T_ref = tokens_3d.shape
    flat = tokens_3d.reshape(B, Q * T_ref)
    quant_ids = torch.arange(Q, device=flat.device).repeat_interleave(T_ref)
    pos_ids = torch.arange(T_ref, device=flat.device).repeat(Q)
    ref_hidden = (decoder.token_embed(flat) + decoder.pos_embed(pos_ids)[None]
                  + decoder.quant_embed(quant_ids)[None])
    return ref_hidden, (tokens_3d == 0).reshape(B, Q * T_ref)


def codec_ce_loss(logits, targets, pad_id=0):
    """logits (B, T, V), targets (B, T) -> mean cross entropy over targets != pad_id."""
    B, T, V = logits.shape
    return F.cross_entropy(logits.reshape(B * T, V).float(), targets.reshape(B * T),
                           ignore_index=pad_id)


class TrainStep:
    """One optimisation step of the decoder: forward, CE loss, backward, gradient all-reduce,
    ``clip_grad_norm_(1.0)``, Adam -- the loop body of ``train.py:220-235`` for the decoder."""

    def __init__(self, decoder, lr=1e-4, max_norm=1.0, pad_id=0, reducer=None, world_size=1,
                 amp_dtype=torch.bfloat16, micro_batch=None, fused_adam=True):
        self.decoder, self.max_norm, self.pad_id = decoder, max_norm, pad_id
        self.reducer, self.world_size = reducer, world_size
        self.amp_dtype, self.micro_batch = amp_dtype, micro_batch
        self.optim = torch.optim.Adam(decoder.parameters(), lr=lr, fused=fused_adam)

    def __call__(self, audio_tokens, text_hidden, z_style, targets=None, text_mask=None,
                 ref_hidden=None, ref_mask=None, ref_tokens=None):
        """``ref_tokens`` (B, Q, T_ref) codec ids of the voice prompt: ref_hidden is then built per
        micro-batch from the decoder's own embeddings (train.py:213-217) so they receive gradient;
        padding (id 0) is masked out."""
        targets = audio_tokens if targets is None else targets   # train.py:228 (unshifted, D7)
        B = audio_tokens.shape[0]
        mb = B if self.micro_batch is None else min(self.micro_batch, B)
        n_valid = (targets != self.pad_id).sum().clamp(min=1).float()
        if self.world_size > 1:
            torch.distributed.all_reduce(n_valid)
        self.optim.zero_grad(set_to_none=True)
        total = torch.zeros((), device=audio_tokens.device)
        starts = list(range(0, B, mb))
        for k, s in enumerate(starts):
            sl = slice(s, s + mb)
            opt = lambda t: None if t is None else t[sl]
            rh, rm = opt(ref_hidden), opt(ref_mask)
            if ref_tokens is not None:
                rh, pad = embed_codec_tokens(ref_tokens[sl], self.decoder)
                rm = ~pad                      # the decoder's masks are True = attend (D3)
            with torch.autocast("cuda", dtype=self.amp_dtype, enabled=self.amp_dtype is not None):
                logits = self.decoder(audio_tokens[sl], text_hidden[sl], z_style[sl], opt(text_mask),
                                      rh, rm)
            V = logits.shape[-1]
            loss_sum = F.cross_entropy(logits.reshape(-1, V).float(), targets[sl].reshape(-1),
                                       ignore_index=self.pad_id, reduction="sum")
            # averaged gradients x world_size / global token count == gradient of the global mean
            loss = loss_sum * (self.world_size / n_valid)
            if self.reducer is not None and k + 1 < len(starts):
                self.reducer.pause()          # only the last micro-batch triggers the all-reduce
            loss.backward()
            if self.reducer is not None and k + 1 < len(starts):
                self.reducer.resume()
            total = total + loss_sum.detach()
        if self.reducer is not None:
            self.reducer.finish()
        torch.nn.utils.clip_grad_norm_(self.decoder.parameters(), self.max_norm)
        self.optim.step()
        return total / n_valid   # this rank's share of the global mean loss
