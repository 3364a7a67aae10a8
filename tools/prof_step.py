"""One C2 training step (or a few decode steps with --decode) between cudaProfilerStart/Stop, for
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv ...
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

cfg = bench.C2
dev = torch.device("cuda", 0)
if "--decode" in sys.argv:
    model = bench.build_decoder(cfg, dev).eval()
    B = 64
    text = torch.randn(B, cfg["t_text"], cfg["d_model"], device=dev)
    z = torch.randn(B, cfg["d_style"], device=dev)
    first = torch.ones(B, 1, dtype=torch.long, device=dev)
    model.generate(first, 4, text, z, dtype=torch.bfloat16, use_cuda_graph=False)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    model.generate(first, 2, text, z, dtype=torch.bfloat16, use_cuda_graph=False)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
else:
    model = bench.build_decoder(cfg, dev).train()
    inp = bench.make_inputs(cfg, cfg["batch"], dev)
    for _ in range(3):
        bench.train_step(model, inp)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    bench.train_step(model, inp)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print("ok")
