"""Selective scan forward / backward at the C2 layer shape (B 16 x d_inner 1024 x T 2048, N 16, bf16), L2 flushed
between launches: median CUDA-event time of the library calls.  MTTS_LIB selects another build of the library."""
import os, statistics, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mamba_tts_project_b200 import _lib, selective_scan_fn
B, Di, T, N = (int(v) for v in (sys.argv[1:5] if len(sys.argv) > 4 else (16, 1024, 2048, 16)))
dev, dt = "cuda", torch.bfloat16
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
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
hook = {"mtts_selective_scan_fwd": [], "mtts_selective_scan_bwd": []}
for it in range(3 + 10):
    if it == 3:
        _lib.event_hook = hook
    flush.zero_()
    y = selective_scan_fn(u, delta, A, Bm, Cm, D, z=z, delta_bias=bias, delta_softplus=True)
    flush.zero_()
    torch.autograd.grad(y, [u, delta, A, Bm, Cm, D, z, bias], dout)
torch.cuda.synchronize()
f = statistics.median(a.elapsed_time(b) for a, b in hook["mtts_selective_scan_fwd"])
b = statistics.median(a.elapsed_time(b) for a, b in hook["mtts_selective_scan_bwd"])
print(f"{os.environ.get('MTTS_LIB', 'default').split('/')[-1]:40s} fwd {f*1e3:8.1f} us   bwd {b*1e3:8.1f} us")
