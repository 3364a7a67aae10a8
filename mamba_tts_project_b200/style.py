"""Caller-side components next to the decoder.

``LengthRegulator`` mirrors ``/root/reference/style_cross_attention.py:144-213`` -- same constructor, same
``forward(hidden, durations, max_len=None) -> (expanded, output_lengths)`` and ``forward_with_target`` -- on one
CUDA launch (``mtts_length_regulate_fwd``) instead of a Python loop with one host sync per phoneme, and it is
differentiable with respect to ``hidden``.  CUDA only, like the rest of the package.
"""
from __future__ import annotations

import torch.nn as nn

from . import ops


class LengthRegulator(nn.Module):
    def __init__(self):
        super().__init__()

    def forward(self, hidden, durations, max_len=None):
        return ops.length_regulate(hidden, durations, max_len=max_len)

    def forward_with_target(self, hidden, target_durations):
        return self.forward(hidden, target_durations)
