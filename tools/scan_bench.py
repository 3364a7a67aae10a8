"""Isolated selective-scan timing (BASELINE config C4): HBM GB/s against the measured peak.
Run on the GPU box:  python tools/scan_bench.py [--bwd] [--dtype bf16|fp32]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mamba_tts_project_b200 import selective_scan_fn  # noqa: E402


def scan_bytes(B, Di, T, N, e, bwd):
    fwd = e * (4 * B * Di * T + 2 * B * N * T) + 4 * (Di * N + 2 * Di)
    if not bwd:
        return fwd
    return (e * (4 * B * Di * T + 2 * B * N * T) + e * 3 * B * Di * T + 4 * 2 * B * N * T
            + 4 * (2 * Di * N + 4 * Di))


def time_cuda(fn, iters, flush):
    start = [torch.cuda.Event(enable_timing=True) for _ in range(iters)]
    end = [torch.cuda.Event(enable_timing=True) for _ in range(iters)]
    for i in range(iters):
        flush.zero_()  # > L2: evict
        start[i].record()
        fn()
        end[i].record()
    torch.cuda.synchronize()
    ts = sorted(s.elapsed_time(e) for s, e in zip(start, end))
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dtype", default="bf16")
    ap.add_argument("--tokens", type=int, default=131072)
    ap.add_argument("--dim", type=int, default=2048)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--only", default="", help="N,T: a single point of the sweep")
    args = ap.parse_args()
    dt = torch.bfloat16 if args.dtype == "bf16" else torch.float32
    e = 2 if dt == torch.bfloat16 else 4
    peak = 6545.6
    try:
        peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    dev = "cuda"
    rows = []
    only = tuple(int(v) for v in args.only.split(",")) if args.only else None
    for N in ((only[0],) if only else (16, 64)):
        for T in ((only[1],) if only else (4096, 16384, 65536)):
            B, Di = args.tokens // T, args.dim
            u = torch.randn(B, Di, T, device=dev, dtype=dt).requires_grad_()
            delta = (0.5 * torch.rand(B, Di, T, device=dev)).to(dt).requires_grad_()
            A = (-0.5 * torch.rand(Di, N, device=dev)).requires_grad_()
            Bm = torch.randn(B, N, T, device=dev, dtype=dt).requires_grad_()
            Cm = torch.randn(B, N, T, device=dev, dtype=dt).requires_grad_()
            D = torch.randn(Di, device=dev).requires_grad_()
            z = torch.randn(B, Di, T, device=dev, dtype=dt).requires_grad_()
            bias = (0.5 * torch.rand(Di, device=dev)).requires_grad_()
            dout = torch.randn(B, Di, T, device=dev, dtype=dt)

            def fwd_only():
                with torch.no_grad():
                    selective_scan_fn(u, delta, A, Bm, Cm, D, z=z, delta_bias=bias, delta_softplus=True)

            out = selective_scan_fn(u, delta, A, Bm, Cm, D, z=z, delta_bias=bias, delta_softplus=True)

            def bwd_only():
                torch.autograd.grad(out, [u, delta, A, Bm, Cm, D, z, bias], dout, retain_graph=True)

            for _ in range(3):
                fwd_only()
                bwd_only()
            tf = time_cuda(fwd_only, args.iters, flush)
            tb = time_cuda(bwd_only, args.iters, flush)
            bf, bb = scan_bytes(B, Di, T, N, e, False), scan_bytes(B, Di, T, N, e, True)
            U = B * Di * T * N
            row = dict(N=N, T=T, B=B, dtype=args.dtype, fwd_ms=round(tf, 4), bwd_ms=round(tb, 4),
                       fwd_GBs=round(bf / tf / 1e6, 1), bwd_GBs=round(bb / tb / 1e6, 1),
                       fwd_frac=round(bf / tf / 1e6 / peak, 4), bwd_frac=round(bb / tb / 1e6 / peak, 4),
                       fwd_Gupd_s=round(U / tf / 1e6, 1), bwd_Gupd_s=round(U / tb / 1e6, 1))
            print(json.dumps(row), flush=True)
            rows.append(row)
            del out
    return rows


if __name__ == "__main__":
    main()
