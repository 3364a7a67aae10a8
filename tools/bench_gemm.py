"""tcgen05 GEMM (mtts_gemm_bf16): correctness vs torch and TFLOP/s vs the library GEMM (+ GELU)."""
import json, os, statistics, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mamba_tts_project_b200 import gemm_bf16
dev = "cuda"; bf = torch.bfloat16
torch.manual_seed(0)

def check(m, n, k, gelu):
    a = torch.randn(m, k, device=dev).to(bf); w = (torch.randn(n, k, device=dev) * k ** -0.5).to(bf)
    b = torch.randn(n, device=dev)
    ref = torch.nn.functional.linear(a.float(), w.float(), b)
    if gelu: ref = torch.nn.functional.gelu(ref)
    out = gemm_bf16(a, w, b, gelu=gelu)
    torch.cuda.synchronize()
    err = (out.float() - ref).abs().max().item() / ref.abs().max().item()
    print(f"check m{m} n{n} k{k} gelu{int(gelu)}: rel err {err:.3e}", flush=True)
    return err

def t(fn, n=20):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n

if __name__ == "__main__":
    errs = [check(128, 128, 64, False), check(128, 128, 512, False), check(256, 384, 512, True),
            check(1000, 520, 264, True), check(32768, 2048, 512, True)]
    assert max(errs) < 2e-2, errs
    for (m, n, k, gelu) in ((32768, 2048, 512, True), (32768, 512, 2048, False), (32768, 2048, 512, False)):
        a = torch.randn(m, k, device=dev).to(bf); w = (torch.randn(n, k, device=dev) * k ** -0.5).to(bf)
        b = torch.randn(n, device=dev); bb = b.to(bf)
        ours = t(lambda: gemm_bf16(a, w, b, gelu=gelu))
        lib = t(lambda: torch.nn.functional.gelu(torch.nn.functional.linear(a, w, bb)) if gelu
                else torch.nn.functional.linear(a, w, bb))
        fl = 2.0 * m * n * k
        print(json.dumps({"m": m, "n": n, "k": k, "gelu": gelu, "ours_ms": round(ours, 4),
                          "ours_TFLOPs": round(fl / ours / 1e9, 1), "library_ms": round(lib, 4),
                          "library_TFLOPs": round(fl / lib / 1e9, 1)}), flush=True)
