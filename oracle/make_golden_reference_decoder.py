"""Golden vectors produced by the REFERENCE's own decoder classes  (TEST INFRASTRUCTURE).

Run from the repo root in the build container (needs /root/reference, which the GPU box does not have):

    python -m oracle.make_golden_reference_decoder          # small + C1 (seconds)
    python -m oracle.make_golden_reference_decoder --big    # C2 / C3 / C5-layer shapes, bf16 convention (minutes)

``/root/reference/mamba_decoder.py`` is imported unmodified.  Its only missing dependency is the
third-party ``mamba_ssm`` package (``mamba_decoder.py:4``; un-vendored, un-pinned, not installable
offline), so ``sys.modules["mamba_ssm"]`` is stubbed with a module whose ``Mamba`` is the oracle's
``MambaRef`` -- the block with the ``(out, state)`` contract the reference documents at ``:9-15`` and calls
at ``:61,63`` -- itself pinned to HuggingFace's independent ``MambaMixer`` by ``oracle/make_golden.py``.
Everything else that runs is the reference's code: ``MambaTTSDecoderLayer`` (LayerNorms,
``nn.MultiheadAttention`` with ``key_padding_mask=~text_mask``, FiLM, FFN, ``:25-91``) and
``MambaTTSDecoder.forward`` / ``decode_step`` (embeddings, [ref || text] concatenation, masks, head,
``:94-256``).

Weights and inputs are name-keyed seeded tensors (``oracle/seeded.py``), so the fixtures hold outputs only:

* ``ref_decoder_small.pt``  2 layers x d_model 64, B 3, T 40, padded text mask, ref_hidden + ref_mask:
  logits, CE loss (ignore_index = 0, ``train.py:31-42``), every parameter gradient, and a 64-step greedy
  ``decode_step`` roll-out (logits + ids).
* ``ref_decoder_c1.pt``     BASELINE config C1 (2 layers x d_model 256, B 2, T 512, T_text 64, V 1024):
  logits at every 4th position, logsumexp + argmax at every position, loss, gradient norms of every
  parameter and the full gradient of every parameter with <= 8192 elements.

``tests/test_oracle_golden.py`` holds the oracle to these files on CPU, ``tests/test_gpu_model.py`` the
CUDA decoder on the B200.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import torch
import torch.nn.functional as F

from .mamba_ref import MambaRef
from .seeded import seeded_state_dict, seeded_tensor

REFERENCE = "/root/reference/mamba_decoder.py"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

SMALL = dict(vocab_size_audio=50, d_model=64, n_layers=2, n_heads=4, d_ff=128, d_style=16, max_len=128,
             num_quantizers=1)
C1 = dict(vocab_size_audio=1024, d_model=256, n_layers=2, n_heads=8, d_ff=2048, d_style=256, max_len=8192,
          num_quantizers=1)
# BASELINE configs[1] / [2] model (12 x d_model 512) and one layer of configs[4] (d_model 1024, 16 heads, d_ff 4096)
C2 = dict(vocab_size_audio=1024, d_model=512, n_layers=12, n_heads=8, d_ff=2048, d_style=256, max_len=2048,
          num_quantizers=1)
C5 = dict(vocab_size_audio=1024, d_model=1024, n_layers=1, n_heads=16, d_ff=4096, d_style=256, max_len=1024,
          num_quantizers=1)


def import_reference():
    """The reference module with ``mamba_ssm.Mamba`` := ``MambaRef`` (see the module docstring)."""
    stub = types.ModuleType("mamba_ssm")
    stub.Mamba = MambaRef
    saved = sys.modules.get("mamba_ssm")
    sys.modules["mamba_ssm"] = stub
    try:
        spec = importlib.util.spec_from_file_location("reference_mamba_decoder", REFERENCE)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        if saved is None:
            del sys.modules["mamba_ssm"]
        else:
            sys.modules["mamba_ssm"] = saved
    return mod


def make_inputs(case, cfg, B, T, T_text, T_ref, seed, masks, round_to=None):
    """Seeded inputs shared by the generator and the tests (they rebuild them from the fixture's meta).
    ``round_to``: float inputs are rounded through that dtype (the bf16 parity convention)."""
    g = torch.Generator().manual_seed(seed)
    V, D = cfg["vocab_size_audio"], cfg["d_model"]
    tokens = torch.randint(1, V, (B, T), generator=g)
    target = torch.randint(0, V, (B, T), generator=g)
    inp = dict(tokens=tokens, target=target,
               text_hidden=seeded_tensor(case + ".text_hidden", (B, T_text, D), seed),
               z_style=seeded_tensor(case + ".z_style", (B, cfg["d_style"]), seed),
               text_mask=None, ref_hidden=None, ref_mask=None)
    if T_ref:
        inp["ref_hidden"] = seeded_tensor(case + ".ref_hidden", (B, T_ref, D), seed)
    if masks:
        tm = torch.rand(B, T_text, generator=g) > 0.3
        tm[:, 0] = True                      # a fully masked row is NaN in nn.MultiheadAttention
        inp["text_mask"] = tm
        if T_ref:
            rm = torch.rand(B, T_ref, generator=g) > 0.3
            rm[:, 0] = True
            inp["ref_mask"] = rm
        target[0, -5:] = 0                   # codec padding id: ignored by the loss (train.py:38-42)
    if round_to is not None:
        for k in ("text_hidden", "z_style", "ref_hidden"):
            if inp[k] is not None:
                inp[k] = inp[k].to(round_to).float()
    return inp


def run_case(ref_mod, case, cfg, B, T, T_text, T_ref, seed, masks, decode_steps, round_to=None,
             zero_quant_embed=False, backward=True):
    dec = ref_mod.MambaTTSDecoder(**cfg).eval()
    sd = seeded_state_dict(dec.state_dict(), seed, round_to=round_to)
    if zero_quant_embed:                     # forward == decode_step position by position
        sd["quant_embed.weight"].zero_()
    dec.load_state_dict(sd)
    inp = make_inputs(case, cfg, B, T, T_text, T_ref, seed, masks, round_to=round_to)
    kw = dict(text_mask=inp["text_mask"], ref_hidden=inp["ref_hidden"], ref_mask=inp["ref_mask"])
    logits = dec(inp["tokens"], inp["text_hidden"], inp["z_style"], **kw)
    V = cfg["vocab_size_audio"]
    loss = F.cross_entropy(logits.reshape(-1, V), inp["target"].reshape(-1), ignore_index=0)
    grads = {}
    if backward:
        loss.backward()
        grads = {k: p.grad.detach().clone() for k, p in dec.named_parameters() if p.grad is not None}
    res = dict(case=case, config=cfg, seed=seed, B=B, T=T, T_text=T_text, T_ref=T_ref, masks=masks,
               round_to=round_to, zero_quant_embed=zero_quant_embed, loss=loss.detach(), source="/root/reference/mamba_decoder.py (mamba_ssm.Mamba := oracle MambaRef)")
    step = {}
    if decode_steps:
        with torch.no_grad():
            tok = torch.ones(B, 1, dtype=torch.long)
            states, lgs, ids = [None] * cfg["n_layers"], [], []
            for i in range(decode_steps):
                lg, states = dec.decode_step(tok, inp["text_hidden"], inp["z_style"], states, i, **kw)
                tok = lg.argmax(-1)
                lgs.append(lg)
                ids.append(tok)
            step = dict(step_logits=torch.cat(lgs, 1), greedy_ids=torch.cat(ids, 1))
    return res, logits.detach(), grads, step


def summarise(res, logits, grads, every):
    res.update(every=every, logits_sub=logits[:, ::every].clone(), logsumexp=torch.logsumexp(logits, -1),
               argmax=logits.argmax(-1), grad_norms={k: g.norm() for k, g in grads.items()},
               grad_absmax={k: g.abs().max() for k, g in grads.items()},
               grads_small={k: g for k, g in grads.items() if g.numel() <= 8192})
    return res


def main_big():
    """bf16-convention fixtures at the BASELINE shapes (weights and float inputs rounded through bf16, the
    reference evaluated in fp32): minutes of CPU time, run once here, compact summaries committed."""
    ref_mod = import_reference()
    bf = torch.bfloat16
    # C2: 12 x d512, T 2048, T_text 256 (B 2): logits + gradients
    res, logits, grads, _ = run_case(ref_mod, "c2", C2, B=2, T=2048, T_text=256, T_ref=0, seed=13, masks=False,
                                     decode_steps=0, round_to=bf)
    torch.save(summarise(res, logits, grads, 16), os.path.join(OUT, "ref_decoder_c2_bf16.pt"))
    print("c2: loss", float(res["loss"]))
    # C5 layer: d1024, 16 heads, d_ff 4096, [ref || text] = 256 + 128 with masks
    res, logits, grads, _ = run_case(ref_mod, "c5", C5, B=2, T=512, T_text=128, T_ref=256, seed=14, masks=True,
                                     decode_steps=0, round_to=bf)
    torch.save(summarise(res, logits, grads, 4), os.path.join(OUT, "ref_decoder_c5_layer_bf16.pt"))
    print("c5: loss", float(res["loss"]))
    # C3: the C2 model decoding B 64 against T_kv 256 for 64 steps, teacher forced.  The reference's
    # decode_step re-projects K/V of the whole memory every step (13 TFLOP on the CPU for this case), so the
    # fixture comes from its teacher-forced forward with quant_embed zeroed -- position t of forward IS step t
    # of decode_step then (D5; tests/test_oracle_golden.py holds step-vs-forward to 1e-6).
    res, logits, _, _ = run_case(ref_mod, "c3", C2, B=64, T=64, T_text=192, T_ref=64, seed=15, masks=True,
                                 decode_steps=0, round_to=bf, zero_quant_embed=True, backward=False)
    res.update(logits_rows=logits[::8].clone(), logsumexp=torch.logsumexp(logits, -1), argmax=logits.argmax(-1))
    torch.save(res, os.path.join(OUT, "ref_decoder_c3_bf16.pt"))
    print("c3: loss", float(res["loss"]))


def main():
    ref_mod = import_reference()
    os.makedirs(OUT, exist_ok=True)

    res, logits, grads, step = run_case(ref_mod, "small", SMALL, B=3, T=40, T_text=7, T_ref=5, seed=11,
                                        masks=True, decode_steps=64)
    res.update(logits=logits, grads=grads, **step)
    torch.save(res, os.path.join(OUT, "ref_decoder_small.pt"))
    print("small: loss", float(res["loss"]), "logits", tuple(logits.shape), "grads", len(grads))

    res, logits, grads, _ = run_case(ref_mod, "c1", C1, B=2, T=512, T_text=64, T_ref=0, seed=12, masks=False,
                                     decode_steps=0)
    res.update(logits_every4=logits[:, ::4].clone(), logsumexp=torch.logsumexp(logits, -1),
               argmax=logits.argmax(-1), grad_norms={k: g.norm() for k, g in grads.items()},
               grad_absmax={k: g.abs().max() for k, g in grads.items()},
               grads_small={k: g for k, g in grads.items() if g.numel() <= 8192})
    torch.save(res, os.path.join(OUT, "ref_decoder_c1.pt"))
    print("c1: loss", float(res["loss"]), "grads", len(grads), "small", len(res["grads_small"]))


if __name__ == "__main__":
    if "--big" in sys.argv:
        main_big()
    else:
        main()
