"""Fused decode-step kernel alone: algorithmic HBM bytes (states r+w, conv state r+w, xz in, y out)
over CUDA-event time, L2 flushed between launches."""
import json, os, statistics, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from mamba_tts_project_b200 import mamba_decode_step
dev = torch.device("cuda", 0)
peak, _ = bench.measured_peaks()
Di, N, R, W = 1024, 16, 32, 4
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for Bk in (64, 256, 1024, 8192):
    xz = torch.randn(Bk, 2 * Di, device=dev, dtype=torch.bfloat16)
    cs = torch.randn(Bk, Di, W, device=dev, dtype=torch.bfloat16)
    ss = torch.randn(Bk, Di, N, device=dev)
    cw, cb = torch.randn(Di, W, device=dev), torch.randn(Di, device=dev)
    xp = (torch.randn(R + 2 * N, Di, device=dev) * 0.05).bfloat16()
    dp = (torch.randn(Di, R, device=dev) * 0.1).bfloat16()
    dtb, A, D = torch.randn(Di, device=dev), -torch.rand(Di, N, device=dev), torch.randn(Di, device=dev)
    y = torch.empty(Bk, Di, device=dev, dtype=torch.bfloat16)
    ts = []
    for it in range(10):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); mamba_decode_step(xz, cs, ss, cw, cb, xp, dp, dtb, A, D, out=y); b.record()
        torch.cuda.synchronize()
        if it >= 3: ts.append(a.elapsed_time(b))
    ms = statistics.median(ts)
    nbytes = 2 * Bk * Di * (4 * N + W * 2) + 2 * Bk * (3 * Di)
    print(json.dumps({"B": Bk, "ms": round(ms, 4), "GBs": round(nbytes / ms / 1e6, 1), "frac_hbm": round(nbytes / ms / 1e6 / peak, 4)}))
