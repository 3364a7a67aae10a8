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
print(prof.key_averages(group_by_input_shape=True).table(sort_by="self_cuda_time_total", row_limit=45,
                                                           max_name_column_width=60, max_shapes_column_width=70))
