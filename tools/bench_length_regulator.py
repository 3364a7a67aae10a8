"""LengthRegulator timing: HBM GB/s of the one-launch expansion against the measured peak, with
the reference-style Python loop timed on a bounded sample beside it.  python tools/bench_length_regulator.py"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mamba_tts_project_b200.ops import length_regulate  # noqa: E402

dev = "cuda"
peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
for (B, T, D, mean_dur, dt) in ((16, 256, 512, 8.0, torch.bfloat16), (64, 256, 512, 16.0, torch.bfloat16),
                                (64, 256, 1024, 16.0, torch.float32)):
    torch.manual_seed(0)
    h = torch.randn(B, T, D, device=dev).to(dt)
    dur = torch.rand(B, T, device=dev) * 2 * mean_dur
    max_len = int(torch.clamp(torch.round(dur), min=0).sum(1).max())
    ts = []
    for _ in range(8):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out, lens = length_regulate(h, dur, max_len=max_len)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ms = sorted(ts)[len(ts) // 2]
    e = h.element_size()
    algo = e * (B * max_len * D + B * T * D) + 4 * B * T   # frames written + rows read once + durations
    row = dict(B=B, T_text=T, D=D, max_len=max_len, dtype=str(dt).split(".")[-1], ms=round(ms, 4),
               GBs=round(algo / ms / 1e6, 1), frac_hbm=round(algo / ms / 1e6 / peak, 4))
    if B == 16:  # the reference's loop (style_cross_attention.py:185-196) on the same tensors, once
        d_int = torch.clamp(torch.round(dur), min=0).long()
        t0 = time.time()
        exp = torch.zeros(B, max_len, D, device=dev, dtype=dt)
        for b_ in range(B):
            pos = 0
            for t in range(T):
                n = d_int[b_, t].item()
                if n > 0 and pos < max_len:
                    end = min(pos + n, max_len)
                    exp[b_, pos:end] = h[b_, t].unsqueeze(0).repeat(end - pos, 1)
                    pos = end
        torch.cuda.synchronize()
        row["reference_loop_ms"] = round((time.time() - t0) * 1e3, 1)
        assert torch.equal(exp, out)
    print(json.dumps(row), flush=True)
