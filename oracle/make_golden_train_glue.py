"""Golden vectors of the caller-side functions of the training step, produced by the REFERENCE's own code
(TEST INFRASTRUCTURE).  Run from the repo root in the build container (needs /root/reference):

    python -m oracle.make_golden_train_glue

``/root/reference/train.py`` cannot be imported (it pulls in the FACodec / text-encoder stack at module
level), so the two pure functions this path uses are taken from its SOURCE -- the ``def`` statements of
``codec_ce_loss`` (``train.py:31-42``) and ``embed_codec_tokens`` (``train.py:115-131``) are cut out with ``ast``
and executed unmodified -- against the reference ``MambaTTSDecoder``'s embedding tables
(``oracle/make_golden_reference_decoder.import_reference``).  The optimiser leg is the reference's literal call
sequence ``clip_grad_norm_(decoder.parameters(), 1.0); optim.step()`` with ``torch.optim.Adam(lr)``
(``train.py:152-159,233-234``) over seeded gradients.  Inputs are seeded, so the fixture
(``tests/golden/ref_train_glue.pt``) holds outputs only.
"""
from __future__ import annotations

import ast
import os

import torch
import torch.nn.functional as F

from .make_golden_reference_decoder import OUT, SMALL, import_reference
from .seeded import seeded_state_dict, seeded_tensor

TRAIN_PY = "/root/reference/train.py"
CFG = dict(SMALL, num_quantizers=3, vocab_size_audio=96)
ADAM_SHAPES = {"w_big": (70, 257), "w_mat": (33, 16), "b_vec": (19,), "scalar1": (1,)}   # 17990 > one 8192 chunk
ADAM_STEPS, ADAM_LR, SEED = 4, 1e-2, 21


def reference_functions(ref_mod):
    src = open(TRAIN_PY).read()
    tree = ast.parse(src)
    ns = {"torch": torch, "F": F, "MambaTTSDecoder": ref_mod.MambaTTSDecoder}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in ("codec_ce_loss", "embed_codec_tokens"):
            exec(compile(ast.Module(body=[node], type_ignores=[]), TRAIN_PY, "exec"), ns)
    return ns["codec_ce_loss"], ns["embed_codec_tokens"]


def glue_inputs():
    """Seeded inputs shared with the tests."""
    g = torch.Generator().manual_seed(SEED)
    B, Q, T, V = 3, CFG["num_quantizers"], 17, CFG["vocab_size_audio"]
    tokens_3d = torch.randint(0, V, (B, Q, T), generator=g)
    tokens_3d[0, :, -4:] = 0                                   # padding (id 0) -> mask True
    logits = 2.0 * seeded_tensor("glue.logits", (B, 24, V), SEED)
    targets = torch.randint(0, V, (B, 24), generator=g)
    targets[1, 5:11] = 0                                       # ignored by the loss
    targets[2, :] = 0
    d_hidden = seeded_tensor("glue.d_hidden", (B, Q * T, CFG["d_model"]), SEED)
    return dict(tokens_3d=tokens_3d, logits=logits, targets=targets, d_hidden=d_hidden)


def adam_inputs():
    params = {k: seeded_tensor("adam.p." + k, s, SEED) for k, s in ADAM_SHAPES.items()}
    # gradient scales: the first steps clip (norm >> 1), the last one does not (norm << 1)
    scales = [3.0, 1.0, 0.05, 1e-3]
    grads = [{k: scales[i] * seeded_tensor(f"adam.g{i}." + k, s, SEED) for k, s in ADAM_SHAPES.items()}
             for i in range(ADAM_STEPS)]
    return params, grads


def main():
    ref_mod = import_reference()
    codec_ce_loss, embed_codec_tokens = reference_functions(ref_mod)
    dec = ref_mod.MambaTTSDecoder(**CFG)
    dec.load_state_dict(seeded_state_dict(dec.state_dict(), SEED))
    inp = glue_inputs()

    ref_hidden, mask = embed_codec_tokens(inp["tokens_3d"], dec)
    ref_hidden.backward(inp["d_hidden"])
    res = dict(config=CFG, seed=SEED, ref_hidden=ref_hidden.detach(), mask=mask,
               d_token_embed=dec.token_embed.weight.grad.clone(), d_pos_embed=dec.pos_embed.weight.grad.clone(),
               d_quant_embed=dec.quant_embed.weight.grad.clone())

    logits = inp["logits"].clone().requires_grad_()
    loss = codec_ce_loss(logits, inp["targets"], pad_id=0)
    loss.backward()
    res.update(loss=loss.detach(), dlogits=logits.grad.clone())
    lb = inp["logits"].to(torch.bfloat16).float().requires_grad_()     # the bf16 convention: rounded inputs, fp32 math
    loss_b = codec_ce_loss(lb, inp["targets"], pad_id=0)
    loss_b.backward()
    res.update(loss_bf16_inputs=loss_b.detach(), dlogits_bf16_inputs=lb.grad.clone())

    p0, grads = adam_inputs()
    params = [torch.nn.Parameter(v.clone()) for v in p0.values()]
    optim = torch.optim.Adam(params, lr=ADAM_LR)
    norms = []
    for gstep in grads:
        optim.zero_grad()
        for p, gv in zip(params, gstep.values()):
            p.grad = gv.clone()
        norms.append(torch.nn.utils.clip_grad_norm_(params, 1.0).clone())
        optim.step()
    res.update(adam_params={k: p.detach().clone() for k, p in zip(p0, params)}, adam_norms=torch.stack(norms),
               source="/root/reference/train.py:31-42,115-131 (function source executed unmodified), :152-159,233-234")
    torch.save(res, os.path.join(OUT, "ref_train_glue.pt"))
    print("loss", float(loss), "ref_hidden", tuple(ref_hidden.shape), "norms", [round(float(n), 4) for n in norms])


if __name__ == "__main__":
    main()
