"""The caller-side pieces of the training step against vectors produced by the reference's own
functions (``oracle/make_golden_train_glue.py``: ``codec_ce_loss`` / ``embed_codec_tokens`` cut out of
``/root/reference/train.py`` and executed unmodified; ``clip_grad_norm_ + Adam`` as ``train.py:233-234`` calls them).

CPU: the host restatements in ``training.py``.  ``-m gpu``: the fused sm_100a kernels (``mtts_embed_sum_*``,
``mtts_ce_loss``, ``mtts_grad_sumsq`` + ``mtts_adam_step``) through the C ABI.
"""
import types

import pytest
import torch

from conftest import load_golden, rel_err
from oracle.make_golden_train_glue import ADAM_LR, adam_inputs, glue_inputs
from oracle.seeded import seeded_state_dict


def _tables(g, dev):
    """An object with the decoder's three embedding tables (the only thing embed_codec_tokens touches)."""
    import torch.nn as nn
    cfg = g["config"]
    m = nn.Module()
    m.token_embed = nn.Embedding(cfg["vocab_size_audio"], cfg["d_model"])
    m.pos_embed = nn.Embedding(cfg["max_len"], cfg["d_model"])
    m.quant_embed = nn.Embedding(cfg["num_quantizers"], cfg["d_model"])
    full = {"token_embed.weight": m.token_embed.weight, "pos_embed.weight": m.pos_embed.weight,
            "quant_embed.weight": m.quant_embed.weight}
    m.load_state_dict(seeded_state_dict({k: v.detach() for k, v in full.items()}, g["seed"]))
    return m.to(dev)


def _check_embed(dev, tol):
    import mamba_tts_project_b200.training as tr
    g, inp = load_golden("ref_train_glue.pt"), glue_inputs()
    dec = _tables(g, dev)
    hid, mask = tr.embed_codec_tokens(inp["tokens_3d"].to(dev), dec)
    assert torch.equal(mask.cpu(), g["mask"])
    assert torch.equal(hid.detach().cpu(), g["ref_hidden"])          # same additions in the same order: bit-exact
    hid.backward(inp["d_hidden"].to(dev))
    assert rel_err(dec.token_embed.weight.grad, g["d_token_embed"]) < tol
    assert rel_err(dec.pos_embed.weight.grad, g["d_pos_embed"]) < tol
    assert rel_err(dec.quant_embed.weight.grad, g["d_quant_embed"]) < tol


def _check_ce(dev, tol):
    import mamba_tts_project_b200.training as tr
    g, inp = load_golden("ref_train_glue.pt"), glue_inputs()
    lg = inp["logits"].to(dev).requires_grad_()
    loss = tr.codec_ce_loss(lg, inp["targets"].to(dev), pad_id=0)
    loss.backward()
    assert abs(loss.item() - g["loss"].item()) < tol * abs(g["loss"].item())
    assert rel_err(lg.grad, g["dlogits"]) < tol
    assert torch.count_nonzero(lg.grad[2]) == 0                      # a fully ignored sample gets no gradient


def test_host_embed_codec_tokens_matches_reference_function():
    _check_embed("cpu", 1e-6)


def test_host_codec_ce_loss_matches_reference_function():
    _check_ce("cpu", 1e-6)


@pytest.mark.gpu
def test_cuda_embed_sum_matches_reference_function():
    _check_embed("cuda", 1e-5)


@pytest.mark.gpu
def test_cuda_ce_loss_matches_reference_function():
    _check_ce("cuda", 1e-5)


@pytest.mark.gpu
def test_cuda_ce_loss_bf16_logits():
    """bf16 logits (the benched mode): the reference evaluated in fp32 on the same rounded values; the gradient
    comes back in bf16."""
    from mamba_tts_project_b200 import ops
    g, inp = load_golden("ref_train_glue.pt"), glue_inputs()
    lg = inp["logits"].to("cuda", torch.bfloat16).requires_grad_()
    loss = ops.ce_loss(lg, inp["targets"].cuda(), ignore_index=0)
    loss.backward()
    assert abs(loss.item() - g["loss_bf16_inputs"].item()) < 1e-5 * abs(g["loss_bf16_inputs"].item())
    assert lg.grad.dtype == torch.bfloat16
    assert rel_err(lg.grad, g["dlogits_bf16_inputs"]) < 4e-3          # one bf16 rounding of the result


@pytest.mark.gpu
def test_cuda_ce_loss_global_token_count_and_all_ignored():
    from mamba_tts_project_b200 import ops
    inp = glue_inputs()
    lg = inp["logits"].cuda().requires_grad_()
    tg = inp["targets"].cuda()
    n = (tg != 0).sum()
    a = ops.ce_loss(lg, tg, ignore_index=0, n_valid=2 * n)             # e.g. two data-parallel ranks
    b = ops.ce_loss(lg.detach(), tg, ignore_index=0)
    assert abs(a.item() * 2 - b.item()) < 1e-5 * abs(b.item())
    z = ops.ce_loss(lg, torch.zeros_like(tg), ignore_index=0)          # nothing to average over
    z.backward()
    assert z.item() == 0.0 and torch.count_nonzero(lg.grad) == 0


@pytest.mark.gpu
def test_cuda_fused_clip_adam_matches_torch_sequence():
    """clip_grad_norm_(params, 1.0) + torch.optim.Adam(lr).step(), four steps: two that clip, two that do not."""
    from mamba_tts_project_b200 import ops
    g = load_golden("ref_train_glue.pt")
    p0, grads = adam_inputs()
    params = [torch.nn.Parameter(v.clone().cuda()) for v in p0.values()]
    opt = ops.FusedClipAdam(params, lr=ADAM_LR, max_norm=1.0)
    for i, gstep in enumerate(grads):
        opt.zero_grad()
        for p, gv in zip(params, gstep.values()):
            p.grad = gv.clone().cuda()
        opt.step()
        assert abs(opt.grad_norm().item() - g["adam_norms"][i].item()) < 1e-5 * g["adam_norms"][i].item()
    for (k, ref), p in zip(g["adam_params"].items(), params):
        assert rel_err(p, ref) < 1e-5, k


@pytest.mark.gpu
def test_train_step_fused_equals_unfused():
    """TrainStep with the fused CE / clip / Adam kernels against the same step built from torch's functions."""
    from mamba_tts_project_b200 import MambaTTSDecoder, TrainStep
    cfg = dict(vocab_size_audio=64, d_model=64, n_layers=2, n_heads=4, d_ff=128, d_style=16, max_len=64)
    torch.manual_seed(0)
    a = MambaTTSDecoder(**cfg).cuda()
    b = MambaTTSDecoder(**cfg).cuda()
    b.load_state_dict(a.state_dict())
    tok = torch.randint(1, 64, (4, 32), device="cuda")
    tok[0, -6:] = 0
    text, z = torch.randn(4, 8, 64, device="cuda"), torch.randn(4, 16, device="cuda")
    sa = TrainStep(a, lr=1e-3, amp_dtype=None, fused_adam=True, micro_batch=2)
    sb = TrainStep(b, lr=1e-3, amp_dtype=None, fused_adam=False)
    for _ in range(3):
        la, lb = sa(tok, text, z), sb(tok, text, z)
        assert abs(la.item() - lb.item()) < 1e-4 * abs(lb.item())
    # Adam divides by sqrt(v): where the true gradient is zero (the key half of in_proj_bias -- softmax is shift
    # invariant -- and q / k sides fed by it) the update is +-lr times the SIGN of rounding noise, so those entries
    # may differ by 2 lr between any two correct implementations; everything else agrees closely.
    for (k, pa), pb in zip(a.named_parameters(), b.parameters()):
        e = rel_err(pa, pb)
        assert e < (5e-2 if k.endswith("in_proj_bias") else 1e-3), f"{k}: {e:.3e}"
