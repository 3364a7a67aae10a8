"""Golden vectors for the LengthRegulator, produced by THE REFERENCE ITSELF (TEST INFRASTRUCTURE).

``/root/reference/style_cross_attention.py`` is pure PyTorch, so its ``LengthRegulator`` (``:144-213``) can be
imported in the build container (it cannot travel to the GPU box: only the vectors are committed).

    python -m oracle.make_golden_length_regulator          # needs /root/reference
"""
from __future__ import annotations

import importlib.util
import os

import torch

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
REF = "/root/reference/style_cross_attention.py"


def main():
    spec = importlib.util.spec_from_file_location("ref_style_cross_attention", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    lr = mod.LengthRegulator()
    cases = {
        # name: (B, T, D, max duration, max_len or None, seed)
        "small": (3, 11, 16, 4.4, None, 0),
        "truncated": (2, 9, 8, 6.0, 17, 1),      # max_len cuts the longest row
        "padded": (2, 5, 8, 2.0, 40, 2),         # max_len beyond every row: zero padding
        "zeros": (2, 7, 8, 0.9, None, 3),        # many durations round to 0 (and .5 ties round to even)
    }
    for name, (B, T, D, dmax, max_len, seed) in cases.items():
        g = torch.Generator().manual_seed(seed)
        hidden = torch.randn(B, T, D, generator=g)
        durations = torch.rand(B, T, generator=g) * dmax
        durations[0, 0] = 2.5   # round-half-to-even: 2
        durations[0, 1] = 3.5   # -> 4
        durations[-1, -1] = -1.0  # clamped to 0
        expanded, lengths = lr(hidden, durations, max_len=max_len)
        torch.save({"hidden": hidden, "durations": durations, "max_len": max_len,
                    "expanded": expanded, "output_lengths": lengths},
                   os.path.join(OUT, f"ref_length_regulator_{name}.pt"))
        print(name, tuple(expanded.shape), lengths.tolist())


if __name__ == "__main__":
    main()
