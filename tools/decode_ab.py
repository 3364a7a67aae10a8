"""Steady-state decode step (C3: B 64, 12 x d512, 256 keys, bf16, CUDA-graph replay), fused bias + GELU epilogue on / off,
alternating in one process so that the box and its clocks are the same."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from mamba_tts_project_b200 import decoder as D
cfg = bench.C2
model = bench.build_decoder(cfg, "cuda").eval()
B = 64
g = torch.Generator().manual_seed(1)
text = torch.randn(B, cfg["t_text"], cfg["d_model"], generator=g).cuda()
z = torch.randn(B, cfg["d_style"], generator=g).cuda()
first = torch.ones(B, 1, dtype=torch.long, device="cuda")
res = {True: [], False: []}
for rep in range(4):
    for fused in (True, False):
        D._DECODE_FUSED_GELU = fused
        model.generate(first, 16, text, z, dtype=torch.bfloat16)
        model.generate(first, 1000, text, z, dtype=torch.bfloat16)
        torch.cuda.synchronize()
        e0, e1, ns = model.last_generate_events
        res[fused].append(e0.elapsed_time(e1) / ns)
for k, v in res.items():
    print("fused GELU epilogue" if k else "separate GELU kernel", [round(x, 4) for x in v])
