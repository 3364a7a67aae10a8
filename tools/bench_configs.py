"""All five BASELINE.json configs in one run -> one JSON document (profiles/rN_configs.json).
   python tools/bench_configs.py [--skip-c5]"""
import json
import os
import statistics
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from mamba_tts_project_b200 import (MambaTTSDecoder, TrainStep, _lib, embed_codec_tokens,  # noqa: E402
                                    mamba_decode_step, selective_scan_fn)

dev = torch.device("cuda", 0)
peak, _ = bench.measured_peaks()
out = {"peak_hbm_gbs": peak}


def ev_time(fn, iters=5, flush=None):
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return statistics.median(ts)


# ---- C3: batched decode, B 64, 1500 steps ------------------------------------------------------
cfg = bench.C2
model = bench.build_decoder(cfg, dev).eval()
B = 64
text = torch.randn(B, cfg["t_text"], cfg["d_model"], device=dev)
z = torch.randn(B, cfg["d_style"], device=dev)
first = torch.ones(B, 1, dtype=torch.long, device=dev)
c3 = {}
for name, dt in (("bf16", torch.bfloat16), ("fp32", torch.float32)):
    model.generate(first, 8, text, z, dtype=dt)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    model.generate(first, 1500, text, z, dtype=dt)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    e0, e1, ns = model.last_generate_events
    ms = e0.elapsed_time(e1) / ns
    c3[name] = {"steps": 1500, "batch": B, "steady_ms_per_step": round(ms, 4),
                "steady_tokens_per_s": round(B / ms * 1e3, 1), "wall_tokens_per_s": round(B * 1500 / wall, 1)}
out["C3_decode"] = c3
del model

# ---- decode-step kernel alone: HBM bytes (states r+w, xz in, y out) vs time ---------------------
ks = {}
Di, N, R, W = 1024, 16, 32, 4
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for Bk in (64, 1024, 8192):
    xz = torch.randn(Bk, 2 * Di, device=dev, dtype=torch.bfloat16)
    cs = torch.randn(Bk, Di, W, device=dev, dtype=torch.bfloat16)
    ss = torch.randn(Bk, Di, N, device=dev)
    cw, cb = torch.randn(Di, W, device=dev), torch.randn(Di, device=dev)
    xp = (torch.randn(R + 2 * N, Di, device=dev) * 0.05).bfloat16()
    dp = (torch.randn(Di, R, device=dev) * 0.1).bfloat16()
    dtb, A, D = torch.randn(Di, device=dev), -torch.rand(Di, N, device=dev), torch.randn(Di, device=dev)
    y = torch.empty(Bk, Di, device=dev, dtype=torch.bfloat16)
    fn = lambda: mamba_decode_step(xz, cs, ss, cw, cb, xp, dp, dtb, A, D, out=y)
    for _ in range(3):
        fn()
    ms = ev_time(fn, 7, flush)
    nbytes = 2 * Bk * Di * (4 * N + W * 2) + 2 * Bk * (3 * Di)
    ks[f"B{Bk}"] = {"ms": round(ms, 4), "GBs": round(nbytes / ms / 1e6, 1), "frac_hbm": round(nbytes / ms / 1e6 / peak, 4)}
out["decode_step_kernel"] = ks

# ---- C4: isolated scan sweep ------------------------------------------------------------------
c4 = []
for dtn, dt, e in (("bf16", torch.bfloat16, 2), ("fp32", torch.float32, 4)):
    for N in (16, 64):
        for T in (4096, 16384, 65536):
            Bq, Dq = 131072 // T, 2048
            u = torch.randn(Bq, Dq, T, device=dev, dtype=dt).requires_grad_()
            dl = (0.5 * torch.rand(Bq, Dq, T, device=dev)).to(dt).requires_grad_()
            A = (-0.5 * torch.rand(Dq, N, device=dev)).requires_grad_()
            Bm = torch.randn(Bq, N, T, device=dev, dtype=dt).requires_grad_()
            Cm = torch.randn(Bq, N, T, device=dev, dtype=dt).requires_grad_()
            D = torch.randn(Dq, device=dev).requires_grad_()
            zz = torch.randn(Bq, Dq, T, device=dev, dtype=dt).requires_grad_()
            bias = (0.5 * torch.rand(Dq, device=dev)).requires_grad_()
            dout = torch.randn(Bq, Dq, T, device=dev, dtype=dt)
            hook = {"mtts_selective_scan_fwd": [], "mtts_selective_scan_bwd": []}
            for it in range(2 + 4):
                if it == 2:
                    _lib.event_hook = hook
                flush.zero_()
                yv = selective_scan_fn(u, dl, A, Bm, Cm, D, z=zz, delta_bias=bias, delta_softplus=True)
                flush.zero_()
                torch.autograd.grad(yv, [u, dl, A, Bm, Cm, D, zz, bias], dout)
            torch.cuda.synchronize()
            _lib.event_hook = {}
            row = {"dtype": dtn, "N": N, "T": T, "B": Bq}
            for nm, bwd in (("fwd", False), ("bwd", True)):
                ms = statistics.median(a.elapsed_time(b) for a, b in hook["mtts_selective_scan_" + nm])
                gbs = bench.scan_alg_bytes(Bq, Dq, T, N, e, bwd) / ms / 1e6
                row[nm] = {"ms": round(ms, 4), "GBs": round(gbs, 1), "frac_hbm": round(gbs / peak, 4)}
            c4.append(row)
            del u, dl, Bm, Cm, zz, dout, yv
            torch.cuda.empty_cache()
out["C4_scan"] = c4

# ---- C5: 24 x d1024 full training step, global B 64 as 8 micro-batches of 8 --------------------
if "--skip-c5" not in sys.argv:
    torch.manual_seed(0)
    dec = MambaTTSDecoder(1024, d_model=1024, n_layers=24, n_heads=16, d_ff=4096, d_style=256,
                          max_len=8192, num_quantizers=1).to(dev)
    nparams = sum(p.numel() for p in dec.parameters())
    Bc, Tc = 64, 4096
    tok = torch.randint(1, 1024, (Bc, Tc), device=dev)
    textc = torch.randn(Bc, 256, 1024, device=dev)
    zc = torch.randn(Bc, 256, device=dev)
    voice = torch.randint(1, 1024, (Bc, 1, 256), device=dev)
    tmask = torch.ones(Bc, 256, dtype=torch.bool, device=dev)
    step = TrainStep(dec, micro_batch=8)
    times = []
    for it in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        loss = step(tok, textc, zc, text_mask=tmask, ref_tokens=voice)   # ref_hidden via embeddings (a12)
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
    s = min(times[1:])
    out["C5_train_step_1gpu"] = {"params": nparams, "global_batch": Bc, "micro_batch": 8, "T": Tc,
                                 "sec_per_step": round(s, 3), "tokens_per_s": round(Bc * Tc / s, 1),
                                 "loss": round(loss.item(), 4),
                                 "peak_mem_GB": round(torch.cuda.max_memory_allocated() / 1e9, 1),
                                 "includes": "fwd + CE + bwd over 8 micro-batches, clip_grad_norm, fused Adam"}
print(json.dumps(out, indent=1))
