"""Kernel timings INSIDE the replayed decode-step CUDA graph (warm, back to back): per-kernel-name totals
per step, GPU busy time vs wall time per step.  python tools/prof_decode_graph.py"""
import os, sys, re, collections
import torch
from torch.profiler import profile, ProfilerActivity
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

cfg = bench.C2
dev = torch.device("cuda", 0)
model = bench.build_decoder(cfg, dev).eval()
B = 64
text = torch.randn(B, cfg["t_text"], cfg["d_model"], device=dev)
z = torch.randn(B, cfg["d_style"], device=dev)
first = torch.ones(B, 1, dtype=torch.long, device=dev)
model.generate(first, 64, text, z, dtype=torch.bfloat16)
torch.cuda.synchronize()
steps = 40
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    model.generate(first, steps, text, z, dtype=torch.bfloat16)
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
# the last `steps` repetitions of the graph: group by name
agg = collections.defaultdict(lambda: [0, 0.0])
for e in ev:
    name = re.sub(r"<.*", "", e.name)[:70]
    agg[name][0] += 1
    agg[name][1] += e.device_time
e0, e1, n = model.last_generate_events
wall = e0.elapsed_time(e1) / n * 1e3
rows = sorted(agg.items(), key=lambda kv: -kv[1][1])
busy = sum(t for _, (c, t) in rows if c >= steps) / steps
print(f"wall {wall:.1f} us/step (under profiler); kernels with >= {steps} launches: busy {busy:.1f} us/step")
for name, (c, t) in rows[:25]:
    if c >= steps:
        print(f"{t/steps:8.1f} us/step  x{c/steps:5.1f}  avg {t/c:6.2f} us  {name}")
