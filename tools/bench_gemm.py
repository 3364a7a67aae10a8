"""tcgen05 GEMM (mtts_gemm) vs the library GEMM on the contraction shapes of the C2 train step:
every layout the Mamba block, the attention and the FFN use, forward / dgrad / wgrad.  TFLOP/s by CUDA events.

    python tools/bench_gemm.py [out.json]
"""
import json, math, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mamba_tts_project_b200.gemm import gemm
dev = "cuda"; bf = torch.bfloat16
torch.manual_seed(0)


def t(fn, n=10):
    """ms per call with the calls replayed from a CUDA graph: device time, no host launch overhead."""
    for _ in range(3): fn()
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n): fn()
    g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3): g.replay()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / (3 * n)


def r(*shape, s=1.0):
    return (torch.randn(*shape, device=dev) * s).to(bf)


def main():
    B, T, D, Di, F, H, Tk, R, N = 16, 2048, 512, 1024, 2048, 8, 256, 32, 16
    dh = D // H
    h = r(B, T, D); Wi = r(2 * Di, D, s=D ** -0.5); xz = r(B, 2 * Di, T); y = r(B, Di, T); Wo = r(D, Di, s=Di ** -0.5)
    X = r(B * T, D); W1 = r(F, D, s=D ** -0.5); A1 = r(B * T, F); W2 = r(D, F, s=F ** -0.5)
    q = r(B, T, D); kk = r(B, Tk, D); P = r(B, H, T, Tk, s=0.1)
    Wx = r(R + 2 * N, Di, s=Di ** -0.5); xdbl = r(B, R + 2 * N, T); Wdt = r(Di, R, s=R ** -0.5)
    qv = q.view(B, T, H, dh).transpose(1, 2); kv = kk.view(B, Tk, H, dh).transpose(1, 2)
    bias = torch.randn(F, device=dev)
    ex = lambda w: w.unsqueeze(0).expand(B, -1, -1)
    cases = [
        ("ffn1 fwd (K,K)", 2 * B * T * D * F, lambda: gemm(X, W1), lambda: X @ W1.t()),
        ("ffn1 fwd (K,K) single-CTA tiles", 2 * B * T * D * F, lambda: gemm(X, W1, single_cta=True), lambda: X @ W1.t()),
        ("ffn2 fwd k2048 single-CTA tiles", 2 * B * T * D * F, lambda: gemm(A1, W2, single_cta=True), lambda: A1 @ W2.t()),
        ("ffn1 fwd + bias + gelu", 2 * B * T * D * F, lambda: gemm(X, W1, bias_n=bias, epilogue="gelu"),
         lambda: torch.nn.functional.gelu(torch.nn.functional.linear(X, W1, bias.to(bf)))),
        ("ffn2 fwd (K,K) k2048", 2 * B * T * D * F, lambda: gemm(A1, W2), lambda: A1 @ W2.t()),
        ("ffn1 dgrad (K,N)", 2 * B * T * D * F, lambda: gemm(A1, W1.t()), lambda: A1 @ W1),
        ("ffn1 wgrad (M,N) splitk", 2 * B * T * D * F,
         lambda: gemm(A1.t(), X.t(), out_dtype=torch.float32, split_k=-1), lambda: A1.t() @ X),
        ("in_proj fwd (K,K) batched", 2 * B * T * D * 2 * Di, lambda: gemm(ex(Wi), h),
         lambda: torch.bmm(ex(Wi), h.transpose(1, 2))),
        ("in_proj dgrad (M,N)", 2 * B * T * D * 2 * Di, lambda: gemm(xz.transpose(1, 2), ex(Wi.t())),
         lambda: torch.bmm(xz.transpose(1, 2), ex(Wi))),
        ("in_proj wgrad (K,N) kbatch", 2 * B * T * D * 2 * Di,
         lambda: gemm(xz, h.transpose(1, 2), out_dtype=torch.float32, reduce_batch=True, split_k=-1),
         lambda: torch.bmm(xz, h).sum(0)),
        ("out_proj fwd (M,K)", 2 * B * T * D * Di, lambda: gemm(y.transpose(1, 2), ex(Wo)),
         lambda: torch.bmm(y.transpose(1, 2), ex(Wo.t()))),
        ("x_proj fwd m64", 2 * B * T * Di * (R + 2 * N), lambda: gemm(ex(Wx), y.transpose(1, 2)),
         lambda: torch.bmm(ex(Wx), y)),
        ("dt_proj fwd k32", 2 * B * T * Di * R, lambda: gemm(ex(Wdt), xdbl[:, :R].transpose(1, 2)),
         lambda: torch.bmm(ex(Wdt), xdbl[:, :R])),
        ("attn scores + softmax", 2 * B * T * Tk * D,
         lambda: gemm(qv, kv, epilogue="softmax", scale=1 / math.sqrt(dh)),
         lambda: torch.softmax((qv @ kv.transpose(-1, -2)) / math.sqrt(dh), -1)),
        ("attn dS = dsoftmax(dO V^T)", 2 * B * T * Tk * D,
         lambda: gemm(qv, kv, epilogue="dsoftmax", aux=P, scale=1 / math.sqrt(dh)),
         lambda: P * ((qv @ kv.transpose(-1, -2)) - 0.5)),
        ("ffn dgrad + gelu' (K,N)", 2 * B * T * D * F, lambda: gemm(X, W2.t(), epilogue="gelu_bwd", aux=A1),
         lambda: torch.nn.functional.gelu(A1) * (X @ W2)),
        ("attn PV n64", 2 * B * T * Tk * D, lambda: gemm(P, kv.transpose(-1, -2)), lambda: P @ kv),
        ("attn dV (M,N) n64 k2048", 2 * B * T * Tk * D, lambda: gemm(P.transpose(-1, -2), qv.transpose(-1, -2)),
         lambda: P.transpose(-1, -2) @ qv),
    ]
    res = []
    for name, fl, ours, lib in cases:
        o, l = ours(), lib()
        err = (o.float() - l.float()).abs().max().item() / max(l.float().abs().max().item(), 1e-9)
        to, tl = t(ours), t(lib)
        rec = {"case": name, "ours_ms": round(to, 4), "ours_TFLOPs": round(fl / to / 1e9, 1),
               "library_ms": round(tl, 4), "library_TFLOPs": round(fl / tl / 1e9, 1), "rel_diff": float(f"{err:.2e}")}
        print(json.dumps(rec), flush=True)
        res.append(rec)
    if len(sys.argv) > 1:
        json.dump(res, open(sys.argv[1], "w"), indent=1)


if __name__ == "__main__":
    main()
