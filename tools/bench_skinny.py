import os, sys, statistics, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mamba_tts_project_b200 import skinny_linear
dev="cuda"; bf=torch.bfloat16
def t(fn, n=20):
    for _ in range(5): fn()
    ts=[]
    for _ in range(n):
        a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b)*1e3)
    return statistics.median(ts)
M=64
for K,N in ((512,2048),(512,512),(1024,512),(2048,512)):
    w=(torch.randn(N,K,device=dev)*K**-0.5).to(bf); b=torch.randn(N,device=dev).to(bf)
    a=torch.randn(M,K,device=dev).to(bf)
    print(f"plain K{K} N{N}: {t(lambda: skinny_linear(w,b,a=a)):.1f} us   torch F.linear: {t(lambda: torch.nn.functional.linear(a,w,b)):.1f} us")
    if K<=1024:
        x=torch.randn(M,K,device=dev); dl=torch.randn(M,K,device=dev).to(bf); xo=torch.empty_like(x)
        lw,lb=torch.randn(K,device=dev),torch.randn(K,device=dev)
        g,be=torch.randn(M,K,device=dev),torch.randn(M,K,device=dev)
        print(f"  ln          : {t(lambda: skinny_linear(w,b,x=x,ln_weight=lw,ln_bias=lb)):.1f} us")
        print(f"  ln+delta    : {t(lambda: skinny_linear(w,b,x=x,delta=dl,x_out=xo,ln_weight=lw,ln_bias=lb)):.1f} us")
        print(f"  ln+film+gelu: {t(lambda: skinny_linear(w,b,x=x,delta=dl,x_out=xo,ln_weight=lw,ln_bias=lb,gamma=g,beta=be,gelu=True)):.1f} us")
