"""The caller contract either side of the decoder's forward/backward:
what ``/root/reference/train.py`` does around ``decoder(...)``, restated for the B200 path.

    embed_codec_tokens   <- train.py:115-131   ref_hidden from the decoder's own embeddings
    codec_ce_loss        <- train.py:31-42     cross entropy with ignore_index = pad_id
    TrainStep            <- train.py:152-159, 220-235  fwd + loss + bwd (+ DP all-reduce) + clip + Adam
    GraphedForwardBackward   the same forward + loss + backward captured once in a CUDA graph and replayed

``TrainStep`` supports micro-batching (gradient accumulation) so that a global batch that does not fit
one GPU (BASELINE config C5: 24 x d1024, B 64, T 4096) keeps identical numerics at every GPU count:
the loss of each micro-batch is the SUM of token losses divided by the GLOBAL number of valid tokens.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import ops


def embed_codec_tokens(tokens_3d, decoder):
    """tokens_3d (B, Q, T) codec ids -> (ref_hidden (B, Q*T, d_model), mask (B, Q*T) True = pad)."""
    B, Q, T_ref = tokens_3d.shape
    flat = tokens_3d.reshape(B, Q * T_ref)
    quant_ids = torch.arange(Q, device=flat.device).repeat_interleave(T_ref)
    pos_ids = torch.arange(T_ref, device=flat.device).repeat(Q)
    if flat.is_cuda and decoder.token_embed.weight.dtype == torch.float32 and decoder.token_embed.weight.shape[1] % 4 == 0:
        ref_hidden = ops.embed_sum(flat, pos_ids, quant_ids, decoder.token_embed.weight, decoder.pos_embed.weight,
                                   decoder.quant_embed.weight)
    else:
        ref_hidden = (decoder.token_embed(flat) + decoder.pos_embed(pos_ids)[None]
                      + decoder.quant_embed(quant_ids)[None])
    return ref_hidden, (tokens_3d == 0).reshape(B, Q * T_ref)


def codec_ce_loss(logits, targets, pad_id=0):
    """logits (B, T, V), targets (B, T) -> mean cross entropy over targets != pad_id."""
    B, T, V = logits.shape
    if logits.is_cuda:      # one pass over the logits in their own dtype: loss and d loss / d logits together
        return ops.ce_loss(logits, targets, ignore_index=pad_id)
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
        # clip_grad_norm_ + Adam as two launches over all parameters (ops.FusedClipAdam); fused_adam=False keeps
        # torch.optim.Adam + torch's clip (the arithmetic the fused kernels are tested against)
        self.fused = bool(fused_adam) and next(decoder.parameters()).is_cuda
        self.optim = (ops.FusedClipAdam(decoder.parameters(), lr=lr, max_norm=max_norm) if self.fused
                      else torch.optim.Adam(decoder.parameters(), lr=lr))

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
            # averaged gradients x world_size / global token count == gradient of the global mean
            if logits.is_cuda:
                loss = ops.ce_loss(logits, targets[sl], ignore_index=self.pad_id, n_valid=n_valid / self.world_size)
                loss_sum = loss.detach() * (n_valid / self.world_size)
            else:
                loss_sum = F.cross_entropy(logits.reshape(-1, V).float(), targets[sl].reshape(-1),
                                           ignore_index=self.pad_id, reduction="sum")
                loss = loss_sum * (self.world_size / n_valid)
            if self.reducer is not None and k + 1 < len(starts):
                self.reducer.pause()          # only the last micro-batch triggers the all-reduce
            loss.backward()
            if self.reducer is not None and k + 1 < len(starts):
                self.reducer.resume()
            total = total + loss_sum.detach()
        if self.reducer is not None:
            self.reducer.finish()
        if not self.fused:
            torch.nn.utils.clip_grad_norm_(self.decoder.parameters(), self.max_norm)
        self.optim.step()
        return total / n_valid   # this rank's share of the global mean loss


class GraphedForwardBackward:
    """``decoder(...)`` + mean cross entropy + ``backward()`` of ``train.py:226-231`` for FIXED shapes, captured
    once in a CUDA graph and replayed: one graph launch per step instead of ~1600 kernel launches, so the
    host never sits between the GPU and its next kernel (eagerly, a step that ends in ``loss.item()`` cannot
    queue ahead and pays the launch latency of every small kernel burst).

        step = GraphedForwardBackward(decoder, tokens, text, z, targets)   # example tensors fix the shapes
        loss = step(tokens, text, z, targets)      # host (pinned) or device tensors; returns a 0-d tensor

    After a call every ``p.grad`` holds this step's gradient (in the graph's memory pool: consume it --
    optimizer, clipping, all-reduce -- before the next call).  ``pad_id`` (default 0, the codec padding id of
    ``codec_ce_loss``, ``train.py:31-42``) is the ``ignore_index`` of the loss; ``None`` averages over every token.  Data parallel: call
    ``GradAllReducer.finish()`` after it (the all-reduce then runs after the backward, not under it)."""

    def __init__(self, decoder, tokens, text_hidden, z_style, targets=None, amp_dtype=torch.bfloat16,
                 pad_id=0, warmup=3, reducer=None, overlap=None):
        """``reducer``: a ``dp.GradAllReducer`` over ``decoder`` -- its bucketed NCCL all-reduces are then captured
        INSIDE the graph (``finish()`` must not be called after a replay: it is part of it).  ``overlap=True``: each
        all-reduce is forked off the backward at the point where its bucket is complete, like the eager hooks do;
        ``overlap=False``: all of them follow the backward, still inside the graph.  Default: environment variable
        ``MTTS_DP_OVERLAP`` (1 / 0), else False -- measured on 2 and 8 B200s: the persistent one-CTA-per-SM GEMMs of
        the backward wait for the SMs a concurrent NCCL kernel holds, which costs more than the overlap hides.  If the capture of the collectives fails (backend that
        cannot be captured), the graph is re-captured without them and ``reduce_after_replay`` is set: the caller's
        ``reducer.finish()`` then follows each replay as before."""
        self.decoder, self.amp_dtype, self.pad_id = decoder, amp_dtype, pad_id
        self.reducer, self.reduce_after_replay = reducer, False
        if overlap is None:
            import os
            overlap = os.environ.get("MTTS_DP_OVERLAP", "0") == "1"
        self.overlap = bool(overlap)
        dev = next(decoder.parameters()).device
        targets = tokens if targets is None else targets
        self._in = [torch.empty(t.shape, dtype=t.dtype, device=dev)
                    for t in (tokens, text_hidden, z_style, targets)]
        for dst, src in zip(self._in, (tokens, text_hidden, z_style, targets)):
            dst.copy_(src, non_blocking=True)
        if reducer is not None:
            reducer.resume()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):          # warm-up: lazy initialisation, autotuning, workspaces, NCCL channels
            for _ in range(warmup):
                self._fwd_bwd()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        from . import _lib
        try:
            self._capture(_lib)
        except Exception:
            if reducer is None:
                raise
            torch.cuda.synchronize()
            reducer.reset()
            reducer.pause()
            self.reducer, self.reduce_after_replay = None, True
            self._capture(_lib)
            reducer.resume()
        # the graph writes its gradients into THESE tensors on every replay
        self._params = [p for p in decoder.parameters() if p.grad is not None]
        self._grads = [p.grad for p in self._params]

    def _capture(self, _lib):
        self.graph = torch.cuda.CUDAGraph()
        self.decoder.zero_grad(set_to_none=True)
        n0 = _lib.launch_count
        with torch.cuda.graph(self.graph):
            self.loss = self._fwd_bwd()
        self.library_launches = _lib.launch_count - n0     # kernels of the C-ABI library inside one replay

    def _fwd_bwd(self):
        tok, text, z, tgt = self._in
        self.decoder.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=self.amp_dtype, enabled=self.amp_dtype is not None):
            logits = self.decoder(tok, text, z)
        # (ignore_index = -1 never matches a codec id: every token counts)
        loss = ops.ce_loss(logits, tgt, ignore_index=-1 if self.pad_id is None else self.pad_id)
        if self.reducer is not None and not self.overlap:
            self.reducer.pause()           # no all-reduce from the gradient hooks: finish() launches them all
        loss.backward()
        if self.reducer is not None:
            self.reducer.resume()
            self.reducer.finish()          # launches what is pending, waits, re-points the gradients
        return loss.detach()

    def __call__(self, tokens, text_hidden, z_style, targets=None):
        targets = tokens if targets is None else targets
        for dst, src in zip(self._in, (tokens, text_hidden, z_style, targets)):
            dst.copy_(src, non_blocking=True)
        self.graph.replay()
        for p, g in zip(self._params, self._grads):   # a caller (e.g. the DP reducer) may have re-pointed .grad
            p.grad = g
        return self.loss
