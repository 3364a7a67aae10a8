"""Generate the committed golden vectors under tests/golden/  (TEST INFRASTRUCTURE).

Run from the repo root in the build container:   python -m oracle.make_golden

Two families of fixtures:

1. ``hf_mixer_*.pt`` -- the PIN.  An independent implementation of the published Mamba-1
   algorithm that ships in this image (HuggingFace ``transformers`` ``MambaMixer.slow_forward``,
   ``models/mamba/modeling_mamba.py``) is run on seeded weights/inputs; weights, input and its
   output are stored.  ``tests/test_oracle_golden.py`` loads the same weights into
   ``oracle.mamba_ref.MambaRef`` and requires agreement.  (The reference's own SSM dependency,
   ``mamba_ssm``, is neither vendored, pinned nor installable offline -- SURVEY.md 8c -- and the
   reference has no test or fixture on this path, so this is the strongest pin available.)

2. ``oracle_*.pt`` -- seeded inputs and the oracle's outputs for every operator of the path
   (scan fwd + grads, conv fwd + grads, state update, block, decoder logits, greedy ids).  They
   freeze the oracle: the GPU parity tests compare the CUDA path against BOTH the live oracle
   and these files, so a later edit of the oracle cannot silently move the target.
"""
from __future__ import annotations

import math
import os

import torch

from .decoder_ref import MambaTTSDecoderRef
from .mamba_ref import MambaRef
from .ssm_ref import (causal_conv1d_ref, causal_conv1d_update_ref, selective_scan_ref,
                      selective_state_update_ref)

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def hf_mixer_case(name, d_model, T, batch, d_state=16, seed=0):
    from transformers.models.mamba.configuration_mamba import MambaConfig
    from transformers.models.mamba.modeling_mamba import MambaMixer

    torch.manual_seed(seed)
    cfg = MambaConfig(hidden_size=d_model, state_size=d_state, conv_kernel=4, expand=2,
                      time_step_rank=math.ceil(d_model / 16), use_bias=False, use_conv_bias=True,
                      hidden_act="silu", num_hidden_layers=1, vocab_size=8)
    mixer = MambaMixer(cfg, layer_idx=0).float().eval()
    with torch.no_grad():
        # HF leaves A_log / D at S4D init; perturb every tensor so no term is degenerate.
        mixer.A_log.add_(0.3 * torch.randn_like(mixer.A_log))
        mixer.D.add_(0.5 * torch.randn_like(mixer.D))
        mixer.dt_proj.bias.copy_(torch.randn_like(mixer.dt_proj.bias) - 3.0)
        mixer.conv1d.bias.copy_(0.2 * torch.randn_like(mixer.conv1d.bias))
        h = torch.randn(batch, T, d_model)
        out = mixer.slow_forward(h)
    sd = {k: v.clone() for k, v in mixer.state_dict().items()}
    torch.save({"d_model": d_model, "d_state": d_state, "state_dict": sd, "input": h,
                "output": out, "source": "transformers.MambaMixer.slow_forward"},
               os.path.join(OUT, f"hf_mixer_{name}.pt"))
    return sd, h, out


def scan_case(name, batch, dim, T, N, seed, with_z=True, with_init=False):
    g = torch.Generator().manual_seed(seed)
    u = torch.randn(batch, dim, T, generator=g)
    delta = 0.5 * torch.rand(batch, dim, T, generator=g)
    A = -0.5 * torch.rand(dim, N, generator=g) - 1e-3
    Bm = torch.randn(batch, N, T, generator=g)
    Cm = torch.randn(batch, N, T, generator=g)
    D = torch.randn(dim, generator=g)
    z = torch.randn(batch, dim, T, generator=g) if with_z else None
    dbias = 0.5 * torch.rand(dim, generator=g)
    h0 = torch.randn(batch, dim, N, generator=g) if with_init else None
    dout = torch.randn(batch, dim, T, generator=g)
    leaves = [t.requires_grad_() for t in (u, delta, A, Bm, Cm, D, dbias)]
    if z is not None:
        z.requires_grad_()
    out, last = selective_scan_ref(u, delta, A, Bm, Cm, D, z=z, delta_bias=dbias,
                                   delta_softplus=True, return_last_state=True,
                                   initial_state=h0)
    grads = torch.autograd.grad(out, leaves + ([z] if z is not None else []), dout)
    names = ["du", "ddelta", "dA", "dB", "dC", "dD", "ddelta_bias"] + (["dz"] if with_z else [])
    rec = {"u": u, "delta": delta, "A": A, "B": Bm, "C": Cm, "D": D, "z": z, "delta_bias": dbias,
           "initial_state": h0, "dout": dout, "out": out, "last_state": last}
    rec.update(dict(zip(names, grads)))
    rec = {k: (v.detach().clone() if torch.is_tensor(v) else v) for k, v in rec.items()}
    torch.save(rec, os.path.join(OUT, f"oracle_scan_{name}.pt"))


def conv_case(name, batch, dim, T, W, seed, activation="silu", with_init=False):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(batch, dim, T, generator=g).requires_grad_()
    w = (torch.randn(dim, W, generator=g) * 0.5).requires_grad_()
    b = torch.randn(dim, generator=g).requires_grad_()
    init = torch.randn(batch, dim, W - 1, generator=g) if with_init else None
    dout = torch.randn(batch, dim, T, generator=g)
    out, fin = causal_conv1d_ref(x, w, b, initial_states=init, return_final_states=True,
                                 activation=activation)
    dx, dw, db = torch.autograd.grad(out, [x, w, b], dout)
    rec = {"x": x, "weight": w, "bias": b, "initial_states": init, "activation": activation,
           "dout": dout, "out": out, "final_states": fin, "dx": dx, "dweight": dw, "dbias": db}
    rec = {k: (v.detach().clone() if torch.is_tensor(v) else v) for k, v in rec.items()}
    torch.save(rec, os.path.join(OUT, f"oracle_conv_{name}.pt"))


def update_case(name, batch, dim, N, W, seed):
    g = torch.Generator().manual_seed(seed)
    state = torch.randn(batch, dim, N, generator=g)
    conv_state = torch.randn(batch, dim, W, generator=g)
    x = torch.randn(batch, dim, generator=g)
    dt = torch.randn(batch, dim, generator=g)
    A = -torch.rand(dim, N, generator=g) - 0.1
    Bm = torch.randn(batch, N, generator=g)
    Cm = torch.randn(batch, N, generator=g)
    D = torch.randn(dim, generator=g)
    z = torch.randn(batch, dim, generator=g)
    dt_bias = torch.rand(dim, generator=g) - 4.0
    w = torch.randn(dim, W, generator=g) * 0.5
    b = torch.randn(dim, generator=g)
    st = state.clone()
    out = selective_state_update_ref(st, x, dt, A, Bm, Cm, D, z=z, dt_bias=dt_bias,
                                     dt_softplus=True)
    cs = conv_state.clone()
    cout = causal_conv1d_update_ref(x, cs, w, b, activation="silu")
    torch.save({"state": state, "x": x, "dt": dt, "A": A, "B": Bm, "C": Cm, "D": D, "z": z,
                "dt_bias": dt_bias, "out": out, "state_after": st,
                "conv_state": conv_state, "weight": w, "bias": b, "conv_out": cout,
                "conv_state_after": cs}, os.path.join(OUT, f"oracle_update_{name}.pt"))


def decoder_case():
    """C1-family decoder, shrunk so the fixture stays small: 2 layers, d_model 64."""
    torch.manual_seed(7)
    dec = MambaTTSDecoderRef(vocab_size_audio=64, d_model=64, n_layers=2, n_heads=4, d_ff=128,
                             d_style=32, max_len=256, num_quantizers=2).eval()
    with torch.no_grad():
        for layer in dec.layers:
            layer.mamba.A_log.add_(0.2 * torch.randn_like(layer.mamba.A_log))
    B, T, Tt, Tr = 2, 40, 10, 6
    tokens = torch.randint(0, 64, (B, T))
    text = torch.randn(B, Tt, 64)
    ref = torch.randn(B, Tr, 64)
    z = torch.randn(B, 32)
    tmask = torch.ones(B, Tt, dtype=torch.bool)
    tmask[1, -3:] = False
    with torch.no_grad():
        logits = dec(tokens, text, z, text_mask=tmask, ref_hidden=ref)
        # greedy decode, 24 steps, from token 1
        tok = torch.ones(B, 1, dtype=torch.long)
        states, ids, step_logits = None, [], []
        for i in range(24):
            lg, states = dec.decode_step(tok, text, z, states, i, text_mask=tmask, ref_hidden=ref)
            tok = lg.argmax(-1)
            ids.append(tok)
            step_logits.append(lg)
    torch.save({"config": dict(vocab_size_audio=64, d_model=64, n_layers=2, n_heads=4, d_ff=128,
                               d_style=32, max_len=256, num_quantizers=2),
                "state_dict": {k: v.clone() for k, v in dec.state_dict().items()},
                "tokens": tokens, "text_hidden": text, "ref_hidden": ref, "z_style": z,
                "text_mask": tmask, "logits": logits, "greedy_ids": torch.cat(ids, 1),
                "step_logits": torch.cat(step_logits, 1)},
               os.path.join(OUT, "oracle_decoder_small.pt"))


def block_case():
    torch.manual_seed(11)
    blk = MambaRef(64).eval()
    with torch.no_grad():
        blk.A_log.add_(0.2 * torch.randn_like(blk.A_log))
    h = torch.randn(2, 37, 64, requires_grad=True)
    out, (cs, ss) = blk(h)
    dout = torch.randn_like(out)
    params = dict(blk.named_parameters())
    grads = torch.autograd.grad(out, [h] + list(params.values()), dout)
    rec = {"state_dict": {k: v.detach().clone() for k, v in blk.state_dict().items()},
           "input": h.detach().clone(), "out": out.detach().clone(), "conv_state": cs.detach(),
           "ssm_state": ss.detach(), "dout": dout, "dinput": grads[0].clone(),
           "dparams": {k: g.clone() for k, g in zip(params.keys(), grads[1:])}}
    torch.save(rec, os.path.join(OUT, "oracle_block_d64.pt"))


def main():
    os.makedirs(OUT, exist_ok=True)
    hf_mixer_case("d64", 64, 48, 2, seed=0)
    hf_mixer_case("d128_n64", 128, 96, 1, d_state=64, seed=1)
    scan_case("n16", 2, 24, 300, 16, seed=1)
    scan_case("n64", 1, 16, 130, 64, seed=2)
    scan_case("n16_init_noz", 2, 8, 77, 16, seed=3, with_z=False, with_init=True)
    conv_case("w4_silu", 2, 24, 133, 4, seed=4)
    conv_case("w3_noact_init", 2, 8, 50, 3, seed=5, activation=None, with_init=True)
    conv_case("w2_short", 1, 8, 1, 2, seed=6)
    update_case("d96_n16", 3, 96, 16, 4, seed=7)
    update_case("d64_n64", 2, 64, 64, 4, seed=8)
    block_case()
    decoder_case()
    total = sum(os.path.getsize(os.path.join(OUT, f)) for f in os.listdir(OUT))
    print(f"wrote {len(os.listdir(OUT))} fixtures, {total / 1e6:.2f} MB -> {OUT}")


if __name__ == "__main__":
    main()
