"""Epilogue-bound mtts_gemm launches of the C2 step for ncu: dt_proj fwd (k 32), x_proj dgrad (k 64, accumulate),
attention scores + softmax, dS = dsoftmax(dO V^T), PV."""
import math, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mamba_tts_project_b200.gemm import gemm
dev, bf = "cuda", torch.bfloat16
torch.manual_seed(0)
B, T, D, Di, H, Tk, R, N = 16, 2048, 512, 1024, 8, 256, 32, 16
r = lambda *s, sc=1.0: (torch.randn(*s, device=dev) * sc).to(bf)
q, kk = r(B, T, D), r(B, Tk, D)
qv = q.view(B, T, H, D // H).transpose(1, 2); kv = kk.view(B, Tk, H, D // H).transpose(1, 2)
Wdt, xdbl = r(Di, R, sc=R ** -0.5), r(B, 64, T)
Wx = r(64, Di, sc=Di ** -0.5)
du = r(B, Di, T)
P = torch.softmax(torch.randn(B, H, T, Tk, device=dev), -1).to(bf)
for _ in range(int(os.environ.get("REPS", "2"))):
    gemm(Wdt.unsqueeze(0).expand(B, -1, -1), xdbl[:, :R].transpose(1, 2))                 # (B, Di, T), k 32
    gemm(Wx.t().unsqueeze(0).expand(B, -1, -1), xdbl.transpose(1, 2), out=du, accumulate=True)   # (B, Di, T), k 64
    gemm(qv, kv, epilogue="softmax", scale=1 / math.sqrt(D // H))
    gemm(qv, kv, epilogue="dsoftmax", aux=P, scale=1 / math.sqrt(D // H))
    gemm(P, kv.transpose(-1, -2))
torch.cuda.synchronize()
print("ok")
