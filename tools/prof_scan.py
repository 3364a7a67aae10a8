"""Minimal launcher for ncu: a few selective-scan / conv launches at BASELINE C2 layer shape
(B 16, d_inner 1024, T 2048, N 16, bf16) or C4 (--c4).  python tools/prof_scan.py [--c4] [--n64]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mamba_tts_project_b200 import causal_conv1d_fn, selective_scan_fn  # noqa: E402

dev, dt = "cuda", torch.bfloat16
B, Di, T = (32, 2048, 4096) if "--c4" in sys.argv else (16, 1024, 2048)
N = 64 if "--n64" in sys.argv else 16
torch.manual_seed(0)
u = torch.randn(B, Di, T, device=dev, dtype=dt).requires_grad_()
delta = (0.5 * torch.rand(B, Di, T, device=dev)).to(dt).requires_grad_()
A = (-0.5 * torch.rand(Di, N, device=dev)).requires_grad_()
Bm = torch.randn(B, N, T, device=dev, dtype=dt).requires_grad_()
Cm = torch.randn(B, N, T, device=dev, dtype=dt).requires_grad_()
D = torch.randn(Di, device=dev).requires_grad_()
z = torch.randn(B, Di, T, device=dev, dtype=dt).requires_grad_()
bias = (0.5 * torch.rand(Di, device=dev)).requires_grad_()
dout = torch.randn(B, Di, T, device=dev, dtype=dt)
w = torch.randn(Di, 4, device=dev).requires_grad_()
cb = torch.randn(Di, device=dev).requires_grad_()
for _ in range(2):
    y = selective_scan_fn(u, delta, A, Bm, Cm, D, z=z, delta_bias=bias, delta_softplus=True)
    torch.autograd.grad(y, [u, delta, A, Bm, Cm, D, z, bias], dout)
    c = causal_conv1d_fn(u, w, cb, activation="silu")
    torch.autograd.grad(c, [u, w, cb], dout)
torch.cuda.synchronize()
print("ok")
