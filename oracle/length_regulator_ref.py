"""CPU restatement of the reference's LengthRegulator.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows ``/root/reference/style_cross_attention.py:144-198`` (``LengthRegulator.forward``): durations are
rounded (half to even, ``torch.round``) and clamped at 0, every phoneme row ``hidden[b, t]`` is repeated
``durations[b, t]`` times along the frame axis, the result is cut at ``max_len`` (default: the longest row of
the batch) and zero-padded; ``output_lengths`` are the un-truncated sums.  The reference is a Python double loop
with ``.item()`` per phoneme (``:185-196``); this restatement is the same arithmetic without the loop, and it is
PINNED: ``oracle/make_golden_length_regulator.py`` imports the reference class itself (pure PyTorch, importable
here) and commits its outputs as ``tests/golden/ref_length_regulator_*.pt``.
"""
from __future__ import annotations

import torch


def length_regulator_ref(hidden: torch.Tensor, durations: torch.Tensor, max_len=None):
    """hidden (B, T, D), durations (B, T) float or int -> (expanded (B, max_len, D), output_lengths (B,) int64)."""
    B, T, D = hidden.shape
    dur = torch.clamp(torch.round(durations.float()), min=0).long()       # :172
    output_lengths = dur.sum(dim=1)                                        # :175
    if max_len is None:
        max_len = int(output_lengths.max().item()) if B > 0 else 0         # :178-179
    cum = dur.cumsum(dim=1)                                                # end frame (exclusive) of phoneme t
    frames = torch.arange(max_len, device=hidden.device)[None, :].expand(B, -1)
    # frame f belongs to the first phoneme whose cumulative end exceeds f   (:185-193)
    idx = torch.searchsorted(cum, frames.contiguous(), right=True)
    valid = frames < output_lengths[:, None]
    idx = idx.clamp(max=max(T - 1, 0))
    expanded = torch.gather(hidden, 1, idx[:, :, None].expand(-1, -1, D)) if T > 0 else \
        hidden.new_zeros(B, max_len, D)
    expanded = torch.where(valid[:, :, None], expanded, torch.zeros((), dtype=hidden.dtype, device=hidden.device))
    return expanded, output_lengths
