"""Per-call device times of one eager C2 training step, keyed by entry point and (for mtts_gemm) by shape, layout
and epilogue: CUDA events around every library call (``_lib.call``) plus a total over the step, so the part
spent in ATen glue is the remainder.

    python tools/prof_ops.py [out.json] [--config C2|C5]
"""
import json
import os
import sys
from collections import defaultdict

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from mamba_tts_project_b200 import _lib  # noqa: E402

records = []
_orig = _lib.call


def traced(name, params, launches=1):
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    _orig(name, params, launches)
    e1.record()
    key = name
    if name == "mtts_gemm":
        p = params
        key = (f"gemm m{p.m} n{p.n} k{p.k} b{p.batch_outer}x{p.batch_inner} kb{p.k_batches} "
               f"{'KM'[p.a_major]}{'KN'[p.b_major]} epi{p.epilogue} out{p.out_dtype} acc{p.accumulate} sk{p.split_k}")
    records.append((key, e0, e1))


def main():
    out = next((a for a in sys.argv[1:] if a.endswith(".json")), None)
    cfg = bench.C2
    if "--c5" in sys.argv:       # one micro-batch of BASELINE configs[4]: 24 x d1024, 8 x 4096 tokens, 256 keys
        cfg = dict(bench.C5, batch=8, t_text=256)
    dev = torch.device("cuda", 0)
    model = bench.build_decoder(cfg, dev).train()
    inp = bench.make_inputs(cfg, cfg["batch"], dev)
    for _ in range(3):
        bench.train_step(model, inp)
    torch.cuda.synchronize()
    _lib.call = traced
    for mod in list(sys.modules.values()):
        if mod is not None and getattr(mod, "__name__", "").startswith("mamba_tts_project_b200"):
            if getattr(mod, "call", None) is _orig:
                mod.call = traced
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    steps = 3
    t0.record()
    for _ in range(steps):
        bench.train_step(model, inp)
    t1.record()
    torch.cuda.synchronize()
    total = t0.elapsed_time(t1) / steps
    agg = defaultdict(lambda: [0, 0.0])
    for key, e0, e1 in records:
        agg[key][0] += 1
        agg[key][1] += e0.elapsed_time(e1)
    rows = sorted(((k, n / steps, t / steps) for k, (n, t) in agg.items()), key=lambda r: -r[2])
    lib_ms = sum(r[2] for r in rows)
    print(f"eager step {total:.3f} ms; library calls {lib_ms:.3f} ms over {sum(r[1] for r in rows):.0f} calls; "
          f"rest (ATen + gaps) {total - lib_ms:.3f} ms")
    for k, n, t in rows:
        print(f"{t:8.3f} ms  x{n:<4.0f} {1e3 * t / n:8.1f} us  {k}")
    if out:
        json.dump({"eager_ms": total, "library_ms": lib_ms,
                   "rows": [{"key": k, "calls": n, "ms": t} for k, n, t in rows]}, open(out, "w"), indent=1)


if __name__ == "__main__":
    main()
