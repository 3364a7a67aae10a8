"""Golden vectors of ``StyleConditioningPipeline`` produced by the REFERENCE's own module (TEST INFRASTRUCTURE).

    python -m oracle.make_golden_style_pipeline        # from the repo root, in the build container

``/root/reference/style_cross_attention.py`` imports only torch, so it is imported unmodified and run in eval mode
(dropout is the only stochastic element) on seeded weights and inputs:

* ``ref_style_pipeline_small.pt``   d_style 16, d_model 64, 4 heads, B 3, T_text 9, durations 0..4 with zeros and a
  ``max_frame_len`` that truncates: every output and the gradient of every parameter and of ``text_hidden`` /
  ``style_emb`` under a seeded cotangent;
* ``ref_style_pipeline_default.pt`` the reference's own smoke-test shape (``style_cross_attention.py:357-382``:
  B 4, T_text 20, d_style 256, d_model 512, 8 heads, durations 1..4): outputs and gradient norms.
"""
from __future__ import annotations

import importlib.util
import os

import torch

from .make_golden_reference_decoder import OUT
from .seeded import seeded_state_dict, seeded_tensor

REFERENCE = "/root/reference/style_cross_attention.py"
SMALL = dict(cfg=dict(d_style=16, d_model=64, num_heads=4, dropout=0.1), B=3, T_text=9, seed=31, max_frame_len=17,
             dur_hi=5, dur_lo=0)
DEFAULT = dict(cfg=dict(d_style=256, d_model=512, num_heads=8, dropout=0.1), B=4, T_text=20, seed=32,
               max_frame_len=None, dur_hi=5, dur_lo=1)


def import_reference():
    spec = importlib.util.spec_from_file_location("reference_style_cross_attention", REFERENCE)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def pipeline_inputs(case):
    g = torch.Generator().manual_seed(case["seed"])
    B, T, cfg = case["B"], case["T_text"], case["cfg"]
    text = seeded_tensor("style.text_hidden", (B, T, cfg["d_model"]), case["seed"])
    emb = seeded_tensor("style.style_emb", (B, cfg["d_style"]), case["seed"])
    dur = torch.randint(case["dur_lo"], case["dur_hi"], (B, T), generator=g).float()
    dur = dur + 0.3 * (torch.rand(B, T, generator=g) - 0.5)          # predicted durations are not integers
    return text, emb, dur


def cotangent(case, shape):
    return seeded_tensor("style.cotangent", shape, case["seed"])


def run(mod, case):
    pipe = mod.StyleConditioningPipeline(**case["cfg"]).eval()
    pipe.load_state_dict(seeded_state_dict(pipe.state_dict(), case["seed"]))
    text, emb, dur = pipeline_inputs(case)
    text.requires_grad_()
    emb.requires_grad_()
    frames, lengths, K, V = pipe(text, emb, dur, max_frame_len=case["max_frame_len"])
    (frames * cotangent(case, frames.shape)).sum().backward()
    grads = {k: (torch.zeros_like(p) if p.grad is None else p.grad.clone()) for k, p in pipe.named_parameters()}
    return dict(case=case, styled_frames=frames.detach(), output_lengths=lengths, style_K=K.detach(),
                style_V=V.detach(), grads=grads, d_text=text.grad.clone(), d_style_emb=emb.grad.clone(),
                source=REFERENCE + " (imported unmodified, eval mode)")


def main():
    mod = import_reference()
    res = run(mod, SMALL)
    torch.save(res, os.path.join(OUT, "ref_style_pipeline_small.pt"))
    print("small", tuple(res["styled_frames"].shape), res["output_lengths"].tolist())
    res = run(mod, DEFAULT)
    res["grad_norms"] = {k: g.norm() for k, g in res["grads"].items()}
    res["grads"] = {k: g for k, g in res["grads"].items() if g.numel() <= 4096}
    res["styled_frames_sub"] = res.pop("styled_frames")[:, ::3].clone()
    res["d_text"] = res["d_text"][:, ::4].clone()
    torch.save(res, os.path.join(OUT, "ref_style_pipeline_default.pt"))
    print("default", tuple(res["styled_frames_sub"].shape), res["output_lengths"].tolist())


if __name__ == "__main__":
    main()
