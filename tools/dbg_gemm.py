import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mamba_tts_project_b200 import _lib
from mamba_tts_project_b200.gemm import gemm
bf = torch.bfloat16
def run(name, a, b, epi="gelu", **kw):
    dbg = torch.zeros(4096, 8, device="cuda", dtype=bf)   # aux-shaped scratch: (m, n) bf16 is required; use a big view
    m, n = a.shape[-2], b.shape[-2]
    aux = torch.zeros(m, n, device="cuda", dtype=bf)
    for _ in range(3):
        gemm(a, b, epilogue=epi, aux=aux, _debug=16 | 1, **kw)
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(); gemm(a, b, epilogue=epi, aux=aux, _debug=16 | 1, **kw); ev1.record(); torch.cuda.synchronize()
    d = aux.view(torch.int64).flatten()[:10].tolist()
    t = aux.view(torch.int64).flatten()[16:16 + 296].view(148, 2).cpu()
    t0 = int(t[:, 0].min())
    ent, ex = (t[:, 0] - t0).tolist(), (t[:, 1] - t0).tolist()
    print("   CTA entry ns min/max", min(ent), max(ent), " exit min/max", min(ex), max(ex), " dur min/max", min(b - a for a, b in zip(ent, ex)), max(b - a for a, b in zip(ent, ex)))
    # graph timing of the same call
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(10): gemm(a, b, epilogue=epi, aux=aux, _debug=1, **kw)
    g.replay(); torch.cuda.synchronize()
    ev0.record(); g.replay(); ev1.record(); torch.cuda.synchronize()
    print("   graph ms per call", ev0.elapsed_time(ev1) / 10)
    print("   event ms", ev0.elapsed_time(ev1), "ns: entry->ready", d[6] - d[5], "ready->mma loop end", d[7] - d[6], "-> epilogue warp done", d[8] - d[7], "-> dealloc", d[9] - d[8])
    print(name, "mma total clk", d[0], "wait tempty", d[1], "wait full", d[2], "| producer total", d[3], "wait empty", d[4])
X, W1 = torch.randn(32768, 512, device="cuda").to(bf), torch.randn(2048, 512, device="cuda").to(bf)
A1, W2 = torch.randn(32768, 2048, device="cuda").to(bf), torch.randn(512, 2048, device="cuda").to(bf)
run("ffn1 K=512 pair gelu", X, W1)
run("ffn1 K=512 pair STORE", X, W1, epi="store")
run("ffn1 K=512 single", X, W1, single_cta=True)
run("ffn2 K=2048 pair", A1, W2)


def gt(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (3 * n)

out = torch.empty(32768, 2048, device="cuda", dtype=bf)
bias = torch.randn(2048, device="cuda")
print("store, out reused          ", gt(lambda: gemm(X, W1, out=out)))
print("store + bias               ", gt(lambda: gemm(X, W1, out=out, bias_n=bias)))
print("no stores (debug 1)        ", gt(lambda: gemm(X, W1, out=out, _debug=1)))
print("gelu + aux                 ", gt(lambda: gemm(X, W1, out=out, bias_n=bias, epilogue="gelu", aux=out)))
print("gelu no aux                ", gt(lambda: gemm(X, W1, out=out, bias_n=bias, epilogue="gelu")))
print("cublas                     ", gt(lambda: torch.matmul(X, W1.t(), out=out)))
print("store single-cta reused    ", gt(lambda: gemm(X, W1, out=out, single_cta=True)))
out2 = torch.empty(32768, 512, device="cuda", dtype=bf)
print("ffn2 K=2048 store reused   ", gt(lambda: gemm(A1, W2, out=out2)))
print("ffn2 cublas                ", gt(lambda: torch.matmul(A1, W2.t(), out=out2)))
