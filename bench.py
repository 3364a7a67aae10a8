#!/usr/bin/env python
"""Benchmark of the MambaTTSDecoder hot path on B200.

    python bench.py --gpus N --steps K --warmup W            # our arm
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference's CPU path

    python bench.py --workload c3|c5 [--decode-weak] ...     # the other multi-GPU configs of BASELINE.json

Default workload at every N: BASELINE.json configs[1] ("C2"): 12-layer d_model 512 MambaTTSDecoder, bf16
(fp32 master weights, bf16 activations/GEMMs, fp32 scan state), teacher-forced forward + backward,
B = 16 per GPU, T_audio 2048, T_text 256, cross-attn + FiLM.  N > 1 is batch-sharded data
parallelism (weak scaling: 16 samples per GPU) with a bucketed NCCL gradient all-reduce.
A "step" = forward + loss + backward (+ all-reduce) over one batch.

Two timed passes per run, both over the same K steps after W warm-up steps:
  1. eager (one launch per kernel, all-reduce overlapped with backward): every call into the C-ABI library is
     bracketed by CUDA events -> per-kernel durations for `roofline` (config.eager_ms_per_step);
  2. the reported one: forward + loss + backward (+ the gradient all-reduce, captured in the same graph) replayed from a
     CUDA graph (mamba_tts_project_b200.GraphedForwardBackward) -> `value`, `ms_per_step`, and,
     with pinned host inputs copied in and loss.item() read back every step, `e2e`.

One JSON line on stdout (rank 0).  Besides the contract keys it carries
  roofline      the dominant hand-written kernel (selective-scan backward) against the measured HBM peak; `others` = the
                other memory-bound kernels; `tensor_core_branches` = FLOP/s of the FFN / attention C-ABI calls (tcgen05
                GEMM family + fused attention) against the measured sustained dense-bf16 throughput
  cpu_baseline  the CPU oracle (port of the reference path) timed on this box's host cores
  e2e           same metric through the public API with host (pinned) inputs copied in every step
  extra         the other two headline quantities of BASELINE.json's metric: isolated scan GB/s
                (config C4) and decode_step tokens/s (config C3)
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

C2 = dict(vocab=1024, d_model=512, n_layers=12, n_heads=8, d_ff=2048, d_style=256, d_state=16,
          batch=16, t_audio=2048, t_text=256)
METRIC = "decoder_train_tokens_per_sec"
UNIT = "tokens/s"


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return p["hbm_gbs"], "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------
# clocks sampling during the timed region
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
            except Exception:
                continue
            for n, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
# workload
# ---------------------------------------------------------------------------------------------
def make_inputs(cfg, batch, device, pinned=False, seed=0):
    g = torch.Generator().manual_seed(seed)
    t = dict(
        tokens=torch.randint(0, cfg["vocab"], (batch, cfg["t_audio"]), generator=g),
        targets=torch.randint(0, cfg["vocab"], (batch, cfg["t_audio"]), generator=g),
        text=torch.randn(batch, cfg["t_text"], cfg["d_model"], generator=g),
        z=torch.randn(batch, cfg["d_style"], generator=g))
    if pinned:
        return {k: v.pin_memory() for k, v in t.items()}
    return {k: v.to(device) for k, v in t.items()}


def build_decoder(cfg, device):
    from mamba_tts_project_b200 import MambaTTSDecoder
    torch.manual_seed(0)
    return MambaTTSDecoder(cfg["vocab"], d_model=cfg["d_model"], n_layers=cfg["n_layers"],
                           n_heads=cfg["n_heads"], d_ff=cfg["d_ff"], d_style=cfg["d_style"],
                           max_len=8192, num_quantizers=1, d_state=cfg["d_state"]).to(device)


def train_step(model, inp, reducer=None):
    from mamba_tts_project_b200 import codec_ce_loss
    model.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        logits = model(inp["tokens"], inp["text"], inp["z"])
    loss = codec_ce_loss(logits, inp["targets"])      # train.py:31-42 (ignore_index = 0), fused CE kernel
    loss.backward()
    if reducer is not None:
        reducer.finish()
    # detached: a live autograd graph would keep this iteration's AccumulateGrad nodes (and their stream)
    # alive, which breaks a later CUDA-graph capture of the same model
    return loss.detach()


def scan_alg_bytes(B, Di, T, N, e, bwd):
    fwd = e * (4 * B * Di * T + 2 * B * N * T) + 4 * (Di * N + 2 * Di)
    if not bwd:
        return fwd
    return (e * (4 * B * Di * T + 2 * B * N * T) + e * 3 * B * Di * T + 4 * 2 * B * N * T
            + 4 * (2 * Di * N + 4 * Di))


def conv_alg_bytes(B, Di, T, e, bwd):
    return e * (4 if bwd else 2) * B * Di * T


# ---------------------------------------------------------------------------------------------
# extras: isolated scan (C4) and decode (C3)
# ---------------------------------------------------------------------------------------------
def extra_scan(peak):
    from mamba_tts_project_b200 import selective_scan_fn
    dev, dt = "cuda", torch.bfloat16
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    out = {}
    for N in (16, 64):
        B, Di, T = 32, 2048, 4096
        u = torch.randn(B, Di, T, device=dev, dtype=dt).requires_grad_()
        delta = (0.5 * torch.rand(B, Di, T, device=dev)).to(dt).requires_grad_()
        A = (-0.5 * torch.rand(Di, N, device=dev)).requires_grad_()
        Bm = torch.randn(B, N, T, device=dev, dtype=dt).requires_grad_()
        Cm = torch.randn(B, N, T, device=dev, dtype=dt).requires_grad_()
        D = torch.randn(Di, device=dev).requires_grad_()
        z = torch.randn(B, Di, T, device=dev, dtype=dt).requires_grad_()
        bias = (0.5 * torch.rand(Di, device=dev)).requires_grad_()
        dout = torch.randn(B, Di, T, device=dev, dtype=dt)
        from mamba_tts_project_b200 import _lib
        hook = {"mtts_selective_scan_fwd": [], "mtts_selective_scan_bwd": []}
        for it in range(3 + 5):
            if it == 3:
                _lib.event_hook = hook
            flush.zero_()
            y = selective_scan_fn(u, delta, A, Bm, Cm, D, z=z, delta_bias=bias, delta_softplus=True)
            flush.zero_()
            torch.autograd.grad(y, [u, delta, A, Bm, Cm, D, z, bias], dout)
        torch.cuda.synchronize()
        _lib.event_hook = {}
        for name, bwd in (("mtts_selective_scan_fwd", False), ("mtts_selective_scan_bwd", True)):
            ms = statistics.median(a.elapsed_time(b) for a, b in hook[name])
            gbs = scan_alg_bytes(B, Di, T, N, 2, bwd) / ms / 1e6
            out[f"scan_{'bwd' if bwd else 'fwd'}_N{N}"] = {
                "ms": round(ms, 4), "GBs": round(gbs, 1), "frac_hbm": round(gbs / peak, 4),
                "state_updates_per_s": round(B * Di * T * N / ms * 1e3, -6)}
        del u, delta, Bm, Cm, z, dout, y
    out["shape"] = "C4: d_inner 2048, B*T = 32*4096, bf16, L2 flushed between launches"
    return out


def extra_decode(cfg, steps=512):
    model = build_decoder(cfg, "cuda").eval()
    B = 64
    g = torch.Generator().manual_seed(1)
    text = torch.randn(B, cfg["t_text"], cfg["d_model"], generator=g).cuda()
    z = torch.randn(B, cfg["d_style"], generator=g).cuda()
    first = torch.ones(B, 1, dtype=torch.long, device="cuda")
    from mamba_tts_project_b200 import _lib
    res = {}
    for name, dt in (("bf16", torch.bfloat16),):
        model.generate(first, 8, text, z, dtype=dt)  # warm-up (captures its own graph)
        torch.cuda.synchronize()
        n0 = _lib.launch_count
        t0 = time.perf_counter()
        toks = model.generate(first, steps, text, z, dtype=dt)
        torch.cuda.synchronize()
        dtm = time.perf_counter() - t0
        e0, e1, ns = model.last_generate_events
        loop_ms = e0.elapsed_time(e1) / ns
        # algorithmic HBM bytes of one step: per layer the conv / SSM states read + written, the
        # cached K and V read once, every weight read once; plus the head
        e = 2 if dt == torch.bfloat16 else 4
        D, L = cfg["d_model"], cfg["n_layers"]
        Di, N, W = 2 * D, cfg["d_state"], 4
        t_kv = cfg["t_text"]
        layer_params = sum(p.numel() for p in model.layers[0].parameters())
        step_bytes = L * (2 * B * Di * N * 4 + 2 * B * Di * W * e + 2 * B * t_kv * D * e + e * layer_params) \
            + e * (cfg["vocab"] * D)
        peak, _ = measured_peaks()
        res[name] = {"tokens_per_s": round(B * steps / dtm, 1), "ms_per_step": round(dtm / steps * 1e3, 4),
                     "algorithmic_bytes_per_step": step_bytes,
                     "steady_state_GBs": round(step_bytes / loop_ms / 1e6, 1),
                     "steady_state_frac_hbm": round(step_bytes / loop_ms / 1e6 / peak, 4),
                     "steps": steps, "batch": B,
                     "includes": "whole generate(): K/V + FiLM precompute, warm-up step, graph capture, replay",
                     "steady_state_ms_per_step": round(loop_ms, 4),
                     "steady_state_tokens_per_s": round(B / loop_ms * 1e3, 1),
                     "library_launches_captured_per_step": (_lib.launch_count - n0) // 2}
    del model
    return res


# ---------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the oracle (port of the reference path) on the host cores
# ---------------------------------------------------------------------------------------------
def cpu_oracle_tokens_per_s(cfg, steps, warmup, t_sample, batch=1):
    from oracle.decoder_ref import MambaTTSDecoderRef
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    dec = MambaTTSDecoderRef(cfg["vocab"], d_model=cfg["d_model"], n_layers=cfg["n_layers"],
                             n_heads=cfg["n_heads"], d_ff=cfg["d_ff"], d_style=cfg["d_style"],
                             max_len=8192, num_quantizers=1, d_state=cfg["d_state"])
    g = torch.Generator().manual_seed(0)
    tok = torch.randint(0, cfg["vocab"], (batch, t_sample), generator=g)
    tgt = torch.randint(0, cfg["vocab"], (batch, t_sample), generator=g)
    text = torch.randn(batch, cfg["t_text"], cfg["d_model"], generator=g)
    z = torch.randn(batch, cfg["d_style"], generator=g)

    def step():
        dec.zero_grad(set_to_none=True)
        lg = dec(tok, text, z)
        F.cross_entropy(lg.view(-1, lg.shape[-1]), tgt.view(-1)).backward()

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    sample = (f"oracle (fp32 port of mamba_decoder.py + selective_scan_ref), same {cfg['n_layers']}-layer "
              f"d{cfg['d_model']} model, B={batch} T_audio={t_sample} T_text={cfg['t_text']}, fwd+bwd, "
              f"{steps} steps of {dt:.2f}s")
    return batch * t_sample / dt, dt, cores, sample


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, args.steps)
    warm = min(args.warmup, 1)
    tps, dt, cores, sample = cpu_oracle_tokens_per_s(C2, steps, warm, t_sample=128)
    line = {
        "impl": "reference", "metric": METRIC, "value": round(tps, 2), "unit": UNIT,
        "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": round(dt * 1e3, 2),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": "C2: 12-layer d_model 512 MambaTTSDecoder fwd+bwd, T_text 256 "
                               "(CPU reference path on a bounded sample: B=1, T_audio=128 per step, the same sample as cpu_baseline)"},
        "cpu_baseline": {"value": round(tps, 2), "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": sample},
        "e2e": {"value": round(tps, 2), "unit": UNIT, "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist
    from mamba_tts_project_b200 import _lib
    from mamba_tts_project_b200.dp import GradAllReducer, broadcast_parameters

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly one JSON line: park the real stdout and send everything else that
    # writes to fd 1 (e.g. NCCL's version banner) to stderr
    sys.stdout.flush()
    json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: there is no CPU path for the product")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("NCCL_MIN_NCHANNELS", "32")   # the collective runs alone after the backward: use the SMs
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()
    cfg = C2
    B, T = cfg["batch"], cfg["t_audio"]
    peak, peak_src = measured_peaks()

    model = build_decoder(cfg, dev).train()
    reducer = None
    if world > 1:
        broadcast_parameters(model)
        # the all-reduce follows the backward inside the graph: one bucket (fewer, larger NCCL launches), measured
        # 34.65 -> 34.06 ms per step at N = 2 together with NCCL_MIN_NCHANNELS = 32 (set in run_ours below)
        reducer = GradAllReducer(model, bucket_bytes=int(os.environ.get("MTTS_DP_BUCKET_MB", "512")) << 20)
    inp = make_inputs(cfg, B, dev, seed=rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        train_step(model, inp, reducer)
    barrier()

    # ---- device-timed region: inputs resident in HBM ----
    hook = {"mtts_selective_scan_fwd": [], "mtts_selective_scan_bwd": [],
            "mtts_causal_conv1d_fwd": [], "mtts_causal_conv1d_bwd": [],
            "mtts_film_ffn_fwd": [], "mtts_film_ffn_bwd": [], "mtts_cross_attn_fwd": [], "mtts_cross_attn_bwd": []}
    _lib.event_hook = hook
    sampler = ClockSampler(local)
    sampler.start()
    n0 = _lib.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        loss = train_step(model, inp, reducer)
    e1.record()
    barrier()
    clocks = sampler.stop()
    _lib.event_hook = {}
    eager_ms = e0.elapsed_time(e1) / args.steps
    tms = torch.tensor([eager_ms], device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    eager_ms = tms.item()
    ms = eager_ms   # denominator of the per-kernel shares below (same launches, eager order)

    # ---- per-kernel roofline from the events recorded inside the timed region ----
    Di, N, e = 2 * cfg["d_model"], cfg["d_state"], 2
    kern = {}
    for name, recs in hook.items():
        if recs:
            kern[name] = statistics.mean(a.elapsed_time(b) for a, b in recs)
    alg = {"mtts_selective_scan_fwd": scan_alg_bytes(B, Di, T, N, e, False),
           "mtts_selective_scan_bwd": scan_alg_bytes(B, Di, T, N, e, True),
           "mtts_causal_conv1d_fwd": conv_alg_bytes(B, Di, T, e, False),
           "mtts_causal_conv1d_bwd": conv_alg_bytes(B, Di, T, e, True)}
    # the tensor-core branches (several mtts_gemm launches per call): FLOP/s against the measured dense bf16 throughput
    tensor = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            tf_peak = json.load(f).get("bf16_tflops_sustained") or 1396.9
    except Exception:
        tf_peak = 1396.9
    Dm, Fd, Tk_, rows = cfg["d_model"], cfg["d_ff"], cfg["t_text"], B * T
    flops = {"mtts_film_ffn_fwd": 4 * rows * Dm * Fd, "mtts_film_ffn_bwd": 8 * rows * Dm * Fd,
             # projections (q, o: rows x D x D; k | v: B Tk x 2D x D) + the two score-sized contractions per head
             "mtts_cross_attn_fwd": 4 * rows * Dm * Dm + 4 * B * Tk_ * Dm * Dm + 4 * rows * Tk_ * Dm,
             "mtts_cross_attn_bwd": 8 * rows * Dm * Dm + 8 * B * Tk_ * Dm * Dm + 10 * rows * Tk_ * Dm}
    for name, fl in flops.items():
        if kern.get(name):
            tf = fl / kern[name] / 1e9
            tensor[name] = {"bound": "tensor", "ms_per_call": round(kern[name], 4), "TFLOPs": round(tf, 1),
                            "frac": round(tf / tf_peak, 4), "peak": tf_peak, "unit": "TFLOP/s",
                            "share_of_step": round(kern[name] * (len(hook[name]) / args.steps) / ms, 4)}
            del kern[name]
    dom = max(kern, key=lambda k: kern[k])
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic = json.load(f).get(dom, {}).get("dram_bytes_per_launch")
    except Exception:
        pass
    ach = alg[dom] / kern[dom] / 1e6
    roofline = {"kernel": dom, "bound": "hbm", "achieved": round(ach, 1), "peak": peak, "unit": "GB/s",
                "frac": round(ach / peak, 4), "traffic": traffic,
                "traffic_source": "profiles/traffic.json: dram__bytes_read.sum + dram__bytes_write.sum per launch from an "
                                  "ncu --set full capture of this kernel at this shape (not measurable inside the run)",
                "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg[dom], "ms_per_launch": round(kern[dom], 4),
                "launches_per_step": len(hook[dom]) // args.steps,
                "share_of_step": round(kern[dom] * (len(hook[dom]) / args.steps) / ms, 4),
                "co_bound": "shared-memory / shuffle data pipe (LSU wavefronts) and warp-level latency; MUFU.EX2 "
                            "(16 per clk per SM) caps the scan at 36% of HBM peak for bf16 N=16; see DESIGN.md",
                "state_updates_per_s": round(B * Di * T * N / kern[dom] * 1e3, -6),
                "others": {k: {"ms_per_launch": round(v, 4), "GBs": round(alg[k] / v / 1e6, 1),
                               "frac": round(alg[k] / v / 1e6 / peak, 4),
                               "share_of_step": round(v * (len(hook[k]) / args.steps) / ms, 4)}
                           for k, v in kern.items() if k != dom},
                "tensor_core_branches": tensor}

    # ---- the reported step: forward + loss + backward captured ONCE in a CUDA graph and replayed
    # (mamba_tts_project_b200.GraphedForwardBackward).  Same kernels as the eager pass above, one graph launch
    # per step: the ~1600 per-kernel launch gaps disappear.  Data parallel: the gradient all-reduce follows
    # the replay (GradAllReducer.finish()).
    from mamba_tts_project_b200 import GraphedForwardBackward
    gstep = GraphedForwardBackward(model, inp["tokens"], inp["text"], inp["z"], inp["targets"], reducer=reducer)

    def graphed_step(src):
        loss = gstep(src["tokens"], src["text"], src["z"], src["targets"])
        if gstep.reduce_after_replay:      # the collectives could not be captured: reduce after the replay
            reducer.finish()
        return loss

    for _ in range(max(args.warmup, 3)):
        graphed_step(inp)
    if os.environ.get("MTTS_BENCH_PROFILE"):
        # ncu --profile-from-start off --metrics gpu__time_duration.sum ... python bench.py: the launch list of ONE
        # replayed step of this very command (profiles/r2_bench_step_launches.*); numbers printed by such a run are
        # not bench values
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        graphed_step(inp)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    e0.record()
    for _ in range(args.steps):
        loss = graphed_step(inp)
    e1.record()
    barrier()
    clocks = sampler.stop()
    tms = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms = tms.item()
    value = world * B * T / ms * 1e3
    launches = gstep.library_launches * args.steps

    # ---- e2e: public API, host (pinned) inputs copied in every step, loss read back ----
    pinned = make_inputs(cfg, B, dev, pinned=True, seed=rank)
    h2d = sum(v.numel() * v.element_size() for v in pinned.values())
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        loss_val = graphed_step(pinned).item()
    barrier()
    e2e_s = torch.tensor([(time.perf_counter() - t0) / args.steps], device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e = {"value": round(world * B * T / e2e_s.item(), 1), "unit": UNIT, "h2d_bytes_per_step": h2d,
           "d2h_bytes_per_step": 4, "ms_per_step": round(e2e_s.item() * 1e3, 3),
           "api": "GraphedForwardBackward(decoder, ...)(pinned host tokens/text/z/targets) = "
                  "MambaTTSDecoder.forward + cross_entropy + backward replayed from a CUDA graph"
                  + (" + gradient all-reduce" if world > 1 else "") + ", loss.item() every step"}

    line = {
        "metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": round(ms, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": {"workload": "C2: 12-layer d_model 512 MambaTTSDecoder, bf16 autocast, teacher-forced "
                               "fwd+bwd, B=16 per GPU, T_audio=2048, T_text=256, cross-attn + FiLM",
                   "global_batch": world * B, "seq_len": T, "d_state": N, "vocab": cfg["vocab"],
                   "parallelism": f"dp{world}" if world > 1 else "single",
                   "l2": "working set (>10 GB of activations per step) is far larger than the 126 MB L2",
                   "execution": "CUDA-graph replay of forward + loss + backward (one graph launch per step)"
                                + ("" if world == 1 else
                                   (", bucketed NCCL all-reduce after the replay" if gstep.reduce_after_replay else
                                    ", bucketed NCCL all-reduces captured inside the graph ("
                                    + ("forked off the backward as each bucket completes" if gstep.overlap
                                       else "after the backward") + ")")),
                   "eager_ms_per_step": round(eager_ms, 3),
                   "loss": round(float(loss_val), 4)},
        "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline,
    }

    if rank == 0 and world == 1 and not args.no_extras:
        del model, inp
        torch.cuda.empty_cache()
        extra = {}
        try:
            extra["scan_isolated"] = extra_scan(peak)
        except Exception as ex:  # extras never invalidate the headline line
            extra["scan_isolated"] = {"error": repr(ex)}
        try:
            extra["decode_c3"] = extra_decode(cfg)
        except Exception as ex:
            extra["decode_c3"] = {"error": repr(ex)}
        line["extra"] = extra
        tps, dt, cores, sample = cpu_oracle_tokens_per_s(cfg, steps=2, warmup=1, t_sample=128)
        line["cpu_baseline"] = {"value": round(tps, 2), "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": sample}
    if rank == 0:
        json_out.write(json.dumps(line) + "\n")
        json_out.flush()
    _teardown(dist, world)



# ---------------------------------------------------------------------------------------------
# the other two multi-GPU workloads of BASELINE.json (not the driver's default line): --workload c3 / c5
# ---------------------------------------------------------------------------------------------
C5 = dict(vocab=1024, d_model=1024, n_layers=24, n_heads=16, d_ff=4096, d_style=256, d_state=16,
          batch=64, t_audio=4096, t_text=128, t_ref=128, micro_batch=8)


def _teardown(dist, world):
    """Leave without tearing the communicator down: a CUDA graph that captured NCCL work keeps it referenced, and
    destroy_process_group() then waits forever (observed at N = 2).  All ranks meet, drain the GPU and exit."""
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
        sys.stderr.flush()
        os._exit(0)


def _dist_setup():
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    sys.stdout.flush()
    json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: there is no CPU path for the product")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    return dist, world, rank, local, dev, json_out, barrier, max_over_ranks


def run_c3(args):
    """BASELINE configs[2]: batched greedy decode_step generation, the C2 model, T_kv 256, 1500 steps; the batch
    of 64 utterances is sharded over the GPUs (--decode-weak: 64 PER GPU instead), no collective on the data path."""
    dist, world, rank, local, dev, json_out, barrier, max_over_ranks = _dist_setup()
    from mamba_tts_project_b200 import _lib
    _lib.load()
    cfg = C2
    B = 64 if args.decode_weak else 64 // world
    steps = args.decode_steps
    model = build_decoder(cfg, dev).eval()
    g = torch.Generator().manual_seed(1 + rank)
    text = torch.randn(B, cfg["t_text"], cfg["d_model"], generator=g)
    z = torch.randn(B, cfg["d_style"], generator=g)
    text_p, z_p = text.pin_memory(), z.pin_memory()
    first = torch.ones(B, 1, dtype=torch.long, device=dev)
    text_d, z_d = text.to(dev), z.to(dev)
    for _ in range(max(1, min(args.warmup, 2))):
        model.generate(first, 16, text_d, z_d, dtype=torch.bfloat16)
    sampler = ClockSampler(local)
    sampler.start()
    walls, loops = [], []
    n0 = _lib.launch_count
    for _ in range(max(1, min(args.steps, 3))):
        barrier()
        t0 = time.perf_counter()
        toks = model.generate(first, steps, text_d, z_d, dtype=torch.bfloat16)
        torch.cuda.synchronize()
        walls.append(time.perf_counter() - t0)
        e0, e1, ns = model.last_generate_events
        loops.append(e0.elapsed_time(e1) / ns)
    launches = _lib.launch_count - n0
    # e2e: host (pinned) conditioning copied in, the generated ids copied back
    barrier()
    t0 = time.perf_counter()
    out_ids = model.generate(first, steps, text_p.to(dev, non_blocking=True), z_p.to(dev, non_blocking=True),
                             dtype=torch.bfloat16).cpu()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop()
    wall = max_over_ranks(min(walls))
    loop_ms = max_over_ranks(min(loops))
    e2e_s = max_over_ranks(e2e_s)
    e = 2
    D, L = cfg["d_model"], cfg["n_layers"]
    Di, N, W = 2 * D, cfg["d_state"], 4
    layer_params = sum(p.numel() for p in model.layers[0].parameters())
    step_bytes = L * (2 * B * Di * N * 4 + 2 * B * Di * W * e + 2 * B * cfg["t_text"] * D * e + e * layer_params) \
        + e * (cfg["vocab"] * D)
    peak, peak_src = measured_peaks()
    line = {
        "metric": "decoder_decode_tokens_per_sec", "value": round(world * B * steps / wall, 1), "unit": UNIT,
        "n_gpus": world, "steps": steps, "warmup": 16, "ms_per_step": round(wall / steps * 1e3, 4),
        "higher_is_better": True, "scaling": "weak" if args.decode_weak else "strong", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"C3: greedy decode_step generation, 12-layer d_model 512, T_kv 256, {steps} steps, "
                               f"B={B} per GPU ({'64 per GPU' if args.decode_weak else '64 utterances sharded'}), "
                               "per-layer conv / SSM state cache, CUDA-graph replay of one step",
                   "global_batch": world * B, "parallelism": f"shard{world}" if world > 1 else "single",
                   "collectives": "none (decode is shard-local)"},
        "clocks": clocks,
        "e2e": {"value": round(world * B * steps / e2e_s, 1), "unit": UNIT,
                "h2d_bytes_per_step": (text.numel() + z.numel()) * 4 // steps, "d2h_bytes_per_step": B * 8,
                "api": "MambaTTSDecoder.generate(first, steps, text, z) with pinned host conditioning copied in "
                       "and the generated ids copied back; includes K/V + FiLM precompute and graph capture"},
        "gpu_launches": launches,
        "roofline": {"kernel": "one replayed decode step (all kernels)", "bound": "hbm",
                     "achieved": round(step_bytes / loop_ms / 1e6, 1), "peak": peak, "unit": "GB/s",
                     "frac": round(step_bytes / loop_ms / 1e6 / peak, 4), "traffic": None, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": step_bytes, "ms_per_launch": round(loop_ms, 4),
                     "note": "steady-state step time from CUDA events around the replay loop, max over ranks"},
    }
    if rank == 0:
        json_out.write(json.dumps(line) + "\n")
        json_out.flush()
    _teardown(dist, world)


def run_c5(args):
    """BASELINE configs[4]: 24 x d_model 1024 training step (fwd + CE + bwd + all-reduce + clip + Adam) with
    ref_hidden built from the decoder's embeddings ([ref || text] cross-attention) and FiLM, GLOBAL batch 64 x T 4096
    sharded over the GPUs (strong scaling), micro-batches of 8 on every GPU."""
    dist, world, rank, local, dev, json_out, barrier, max_over_ranks = _dist_setup()
    from mamba_tts_project_b200 import MambaTTSDecoder, TrainStep, _lib
    from mamba_tts_project_b200.dp import GradAllReducer, broadcast_parameters
    _lib.load()
    cfg = C5
    Bg, T = cfg["batch"], cfg["t_audio"]
    B = Bg // world
    torch.manual_seed(0)
    dec = MambaTTSDecoder(cfg["vocab"], d_model=cfg["d_model"], n_layers=cfg["n_layers"], n_heads=cfg["n_heads"],
                          d_ff=cfg["d_ff"], d_style=cfg["d_style"], max_len=8192, num_quantizers=1,
                          d_state=cfg["d_state"]).to(dev)
    reducer = None
    if world > 1:
        broadcast_parameters(dec)
        reducer = GradAllReducer(dec, bucket_bytes=64 << 20)
    g = torch.Generator().manual_seed(100 + rank)
    host = dict(tok=torch.randint(1, cfg["vocab"], (B, T), generator=g),
                text=torch.randn(B, cfg["t_text"], cfg["d_model"], generator=g),
                z=torch.randn(B, cfg["d_style"], generator=g),
                voice=torch.randint(1, cfg["vocab"], (B, 1, cfg["t_ref"]), generator=g),
                tmask=torch.ones(B, cfg["t_text"], dtype=torch.bool))
    host = {k: v.pin_memory() for k, v in host.items()}
    devin = {k: v.to(dev) for k, v in host.items()}
    step = TrainStep(dec, micro_batch=min(cfg["micro_batch"], B), reducer=reducer, world_size=world)

    def one(src):
        return step(src["tok"], src["text"], src["z"], text_mask=src["tmask"], ref_tokens=src["voice"])

    for _ in range(max(1, min(args.warmup, 2))):
        one(devin)
    nsteps = max(1, min(args.steps, 3))
    sampler = ClockSampler(local)
    sampler.start()
    n0 = _lib.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(nsteps):
        loss = one(devin)
    e1.record()
    barrier()
    launches = _lib.launch_count - n0
    ms = max_over_ranks(e0.elapsed_time(e1) / nsteps)
    barrier()
    t0 = time.perf_counter()
    for _ in range(nsteps):
        loss_val = one({k: v.to(dev, non_blocking=True) for k, v in host.items()}).item()
    barrier()
    e2e_s = max_over_ranks((time.perf_counter() - t0) / nsteps)
    clocks = sampler.stop()
    h2d = sum(v.numel() * v.element_size() for v in host.values())
    line = {
        "metric": METRIC, "value": round(Bg * T / ms * 1e3, 1), "unit": UNIT, "n_gpus": world, "steps": nsteps,
        "warmup": max(1, min(args.warmup, 2)), "ms_per_step": round(ms, 2), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "C5: 24-layer d_model 1024 (16 heads, d_ff 4096) MambaTTSDecoder training step: fwd + CE "
                               "+ bwd + NCCL all-reduce + clip_grad_norm + Adam, ref_hidden from the decoder's own "
                               f"embeddings, [ref {cfg['t_ref']} || text {cfg['t_text']}] cross-attention + FiLM, "
                               f"global B={Bg} x T_audio={T}, {B} per GPU in micro-batches of {min(cfg['micro_batch'], B)}",
                   "global_batch": Bg, "seq_len": T, "parallelism": f"dp{world}" if world > 1 else "single",
                   "params": sum(p.numel() for p in dec.parameters()),
                   "peak_mem_GB": round(torch.cuda.max_memory_allocated() / 1e9, 1), "loss": round(float(loss_val), 4),
                   "all_reduce": "bucketed (64 MB), launched from the gradient hooks of the last micro-batch: "
                                 "overlaps the rest of that backward"},
        "clocks": clocks,
        "e2e": {"value": round(Bg * T / e2e_s, 1), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "ms_per_step": round(e2e_s * 1e3, 2),
                "api": "TrainStep(decoder, micro_batch=8, reducer)(pinned host tokens / text / z / voice-prompt ids), "
                       "loss.item() every step"},
        "gpu_launches": launches,
    }
    if rank == 0:
        json_out.write(json.dumps(line) + "\n")
        json_out.flush()
    _teardown(dist, world)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--workload", default="c2", choices=["c2", "c3", "c5"],
                    help="c2 (default, the driver's line): BASELINE configs[1]; c3: decode, configs[2]; c5: configs[4]")
    ap.add_argument("--decode-steps", type=int, default=1500)
    ap.add_argument("--decode-weak", action="store_true", help="c3: 64 utterances PER GPU instead of 64 in total")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    elif args.workload == "c3":
        run_c3(args)
    elif args.workload == "c5":
        run_c5(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
