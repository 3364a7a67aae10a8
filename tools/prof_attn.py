"""The fused attention core at the C2 layer shape (B 16 x 8 heads x T 2048 x 256 keys) for ncu: one forward and one
backward of dense.cross_attention (mtts_cross_attn_fwd / _bwd -> attn_fwd_kernel / attn_bwd_kernel)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mamba_tts_project_b200 import dense
B, T, Tk, E, H = 16, 2048, 256, 512, 8
g = torch.Generator().manual_seed(0)
rnd = lambda *s, sc=1.0: (torch.randn(*s, generator=g) * sc).to(torch.bfloat16).cuda()
query, memory = rnd(B, T, E).requires_grad_(), rnd(B, Tk, E).requires_grad_()
w_in, w_out = rnd(3 * E, E, sc=E ** -0.5).float().requires_grad_(), rnd(E, E, sc=E ** -0.5).float().requires_grad_()
b_in = torch.zeros(3 * E, device="cuda", requires_grad=True)
dout = rnd(B, T, E)
for _ in range(int(os.environ.get("REPS", "2"))):
    out = dense.cross_attention(query, memory, w_in, b_in, w_out, None, H)
    torch.autograd.grad(out, [query, memory, w_in, b_in, w_out], dout)
torch.cuda.synchronize()
print("ok")
