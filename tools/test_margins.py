import sys, torch
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
import torch.nn.functional as F
from test_reference_pin import _bf16_case, _load_bf16
from conftest import rel_err
from mamba_tts_project_b200 import MambaTTSDecoder
for name in ("c2_bf16", "c5_layer_bf16"):
    g, inp = _bf16_case(name)
    for rep in range(4):
        dec = _load_bf16(MambaTTSDecoder(**g["config"]), g).cuda().eval()
        mv = lambda t: None if t is None else t.cuda()
        V = g["config"]["vocab_size_audio"]
        with torch.autocast("cuda", dtype=torch.bfloat16):
            logits = dec(mv(inp["tokens"]), mv(inp["text_hidden"]), mv(inp["z_style"]), text_mask=mv(inp["text_mask"]), ref_hidden=mv(inp["ref_hidden"]), ref_mask=mv(inp["ref_mask"]))
        loss = F.cross_entropy(logits.reshape(-1, V).float(), mv(inp["target"]).reshape(-1), ignore_index=0)
        loss.backward()
        lg = logits.detach().float()
        e = rel_err(lg[:, ::g["every"]], g["logits_sub"])
        wl2, wmax, wn = 0, 0, 0
        for k, p in dec.named_parameters():
            n_ref = g["grad_norms"][k].item()
            wn = max(wn, abs(p.grad.float().norm().item() - n_ref) / max(n_ref, 1e-12))
            if k in g["grads_small"]:
                ref = g["grads_small"][k].float()
                wl2 = max(wl2, ((p.grad.float().cpu() - ref).norm() / ref.norm().clamp(min=1e-20)).item())
                wmax = max(wmax, rel_err(p.grad, ref))
        print(f"{name} rep{rep}: logits {e:.4f} (tol 0.02)  grad-norm {wn:.4f} (tol 0.04)  grad L2 {wl2:.4f} (tol 0.05)  grad max {wmax:.4f} (tol 0.10)")
