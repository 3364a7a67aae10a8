"""Name-keyed seeded weights for the decoder module tree (TEST INFRASTRUCTURE, see oracle/__init__.py).

The reference-pinned fixtures (``oracle/make_golden_reference_decoder.py``) must give the REFERENCE classes,
the oracle and the CUDA decoder bit-identical weights without shipping megabytes of state_dict: every tensor
is drawn from its own generator, seeded by ``crc32(name) ^ seed``, so the values depend on the parameter's
name and shape only -- not on construction order, RNG consumption of a constructor, or module class.
"""
from __future__ import annotations

import math
import zlib

import torch


def seeded_tensor(name: str, shape, seed: int) -> torch.Tensor:
    g = torch.Generator().manual_seed((zlib.crc32(name.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
    return torch.randn(tuple(shape), generator=g, dtype=torch.float32)


def seeded_state_dict(reference_sd, seed: int, round_to=None):
    """A state_dict with the keys / shapes of ``reference_sd`` and well-conditioned seeded values.

    Scales: matrices ~ N(0, 1/fan_in); embeddings N(0, 0.5); LayerNorm weight 1 + 0.1 n, biases 0.1 n;
    ``A_log`` = log(1..N) + 0.3 n (S4D-real, perturbed); ``D`` = 1 + 0.5 n; ``dt_proj.bias`` = n - 3
    (softplus^-1 of dt around 0.05).  ``round_to``: every tensor with dim > 1 is rounded through that dtype
    (the bf16 parity convention: oracle and CUDA path see the same rounded weights)."""
    out = {}
    for name, ref in reference_sd.items():
        shape = tuple(ref.shape)
        n = seeded_tensor(name, shape, seed)
        leaf = name.split(".")[-1]
        if name.endswith("A_log"):
            N = shape[1]
            t = torch.log(torch.arange(1, N + 1, dtype=torch.float32))[None, :] + 0.3 * n
        elif leaf == "D":
            t = 1.0 + 0.5 * n
        elif name.endswith("dt_proj.bias"):
            t = n - 3.0
        elif "embed" in name:
            t = 0.5 * n
        elif name.startswith("norm") or ".norm" in name:
            t = 1.0 + 0.1 * n if leaf == "weight" else 0.1 * n
        elif n.dim() >= 2:
            fan_in = math.prod(shape[1:])
            t = n / math.sqrt(fan_in)
        else:
            t = 0.1 * n
        if round_to is not None and t.dim() > 1:
            t = t.to(round_to).float()
        out[name] = t.contiguous()
    return out
