"""torch.profiler attribution of one C2 training step: which aten op launched which kernels."""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

cfg = bench.C2
dev = torch.device("cuda", 0)
model = bench.build_decoder(cfg, dev).train()
inp = bench.make_inputs(cfg, cfg["batch"], dev)
for _ in range(3):
    bench.train_step(model, inp)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
    bench.train_step(model, inp)
    torch.cuda.synchronize()
rows = [(e.key, e.self_device_time_total / 1e3, e.count, str(e.input_shapes)[:90])
        for e in prof.key_averages(group_by_input_shape=True)
        if e.self_device_time_total > 0 and (e.key.startswith("aten::") or e.key.startswith("_") or "Backward" in e.key)]
tot = sum(r[1] for r in rows)
print(f"ops with device time: {tot:.2f} ms")
for k, t, n, sh in sorted(rows, key=lambda r: -r[1])[:110]:
    print(f"{t:8.3f} ms x{n:<4d} {k:<40s} {sh}")
