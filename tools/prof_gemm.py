"""Three representative mtts_gemm launches for ncu: short-K large-N (ffn1 fwd), one-k-block row-softmax (attention
scores), tiny-K (dt_proj)."""
import math, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mamba_tts_project_b200.gemm import gemm
dev, bf = "cuda", torch.bfloat16
torch.manual_seed(0)
B, T, D, Di, F, H, Tk, R = 16, 2048, 512, 1024, 2048, 8, 256, 32
r = lambda *s, sc=1.0: (torch.randn(*s, device=dev) * sc).to(bf)
X, W1 = r(B * T, D), r(F, D, sc=D ** -0.5)
A1, W2 = r(B * T, F), r(D, F, sc=F ** -0.5)
q, kk = r(B, T, D), r(B, Tk, D)
qv = q.view(B, T, H, D // H).transpose(1, 2); kv = kk.view(B, Tk, H, D // H).transpose(1, 2)
Wdt, xdbl = r(Di, R, sc=R ** -0.5), r(B, 64, T)
for _ in range(int(os.environ.get("REPS", "2"))):
    gemm(X, W1)
    gemm(A1, W2)
    gemm(qv, kv, epilogue="softmax", scale=1 / math.sqrt(D // H))
    gemm(Wdt.unsqueeze(0).expand(B, -1, -1), xdbl[:, :R].transpose(1, 2))
torch.cuda.synchronize()
print("ok")
