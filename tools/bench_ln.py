"""Fused add+LayerNorm(+FiLM) forward / backward timing at the C2 shape (32768 x 512, bf16 branch)."""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mamba_tts_project_b200 import add_layernorm

dev = "cuda"
B, T, D = 16, 2048, 512
torch.manual_seed(0)
x = torch.randn(B, T, D, device=dev, requires_grad=True)
delta = torch.randn(B, T, D, device=dev).bfloat16().requires_grad_()
w = torch.randn(D, device=dev, requires_grad=True); b = torch.randn(D, device=dev, requires_grad=True)
gam = torch.randn(B, D, device=dev, requires_grad=True); bet = torch.randn(B, D, device=dev, requires_grad=True)
dbias = torch.randn(D, device=dev, requires_grad=True)
gx = torch.randn(B, T, D, device=dev); gh = torch.randn(B, T, D, device=dev).bfloat16()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

def timeit(fn, n=10):
    ts = []
    for _ in range(n):
        flush.zero_()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(e))
    return sorted(ts)[len(ts) // 2]

for film in (False, True):
    kw = dict(gamma=gam, beta=bet) if film else {}
    def fwd():
        return add_layernorm(x, delta, w, b, 1e-5, out_dtype=torch.bfloat16, delta_bias=dbias, **kw)
    xo, h = fwd()
    leaves = [x, delta, w, b, dbias] + ([gam, bet] if film else [])
    def bwd():
        torch.autograd.grad([xo, h], leaves, [gx, gh], retain_graph=True)
    for _ in range(3):
        with torch.no_grad(): fwd()
        bwd()
    with torch.no_grad():
        tf = timeit(fwd)
    tb = timeit(bwd)
    e = 2
    fb = B * T * D * (4 + e + 4 + e); bb = B * T * D * (4 + e + 4 + 4 + e)
    print(json.dumps({"film": film, "fwd_ms": round(tf, 4), "fwd_GBs": round(fb / tf / 1e6, 1),
                      "bwd_ms": round(tb, 4), "bwd_GBs": round(bb / tb / 1e6, 1)}))
