"""Parity pinned by the REFERENCE's own classes.

``tests/golden/ref_decoder_*.pt`` were produced by importing ``/root/reference/mamba_decoder.py`` unmodified
(``oracle/make_golden_reference_decoder.py``: only ``mamba_ssm.Mamba`` is stubbed, with the oracle's HF-pinned
block).  CPU tests hold the oracle restatement to them; ``-m gpu`` tests hold the CUDA decoder to them:
logits, CE loss, every parameter gradient, and a 64-step greedy ``decode_step`` roll-out (identical ids in
fp32) with padded text / reference masks and ``ref_hidden``.
"""
import pytest
import torch
import torch.nn.functional as F

from conftest import load_golden, rel_err
from oracle.decoder_ref import MambaTTSDecoderRef
from oracle.make_golden_reference_decoder import make_inputs
from oracle.seeded import seeded_state_dict

FP32_TOL = 1e-4     # BASELINE.json: logits within 1e-4 relative in fp32


def _case(name):
    g = load_golden(f"ref_decoder_{name}.pt")
    inp = make_inputs(g["case"], g["config"], g["B"], g["T"], g["T_text"], g["T_ref"], g["seed"], g["masks"])
    return g, inp


def _load(model, g):
    model.load_state_dict(seeded_state_dict(model.state_dict(), g["seed"]))
    return model


def _fwd_bwd(model, g, inp, dev):
    mv = lambda t: None if t is None else t.to(dev)
    V = g["config"]["vocab_size_audio"]
    logits = model(mv(inp["tokens"]), mv(inp["text_hidden"]), mv(inp["z_style"]), text_mask=mv(inp["text_mask"]),
                   ref_hidden=mv(inp["ref_hidden"]), ref_mask=mv(inp["ref_mask"]))
    loss = F.cross_entropy(logits.reshape(-1, V).float(), mv(inp["target"]).reshape(-1), ignore_index=0)
    loss.backward()
    return logits.detach(), loss.detach()


def _check_small(model, g, inp, dev, tol, gtol):
    logits, loss = _fwd_bwd(model, g, inp, dev)
    assert rel_err(logits, g["logits"]) < tol
    assert abs(loss.item() - g["loss"].item()) < tol * abs(g["loss"].item())
    seen = 0
    for k, p in model.named_parameters():
        if k not in g["grads"]:
            continue
        e = rel_err(p.grad, g["grads"][k])
        assert e < gtol, f"grad {k}: rel err {e:.3e}"
        seen += 1
    assert seen == len(g["grads"])


def _check_c1(model, g, inp, dev, tol, gtol):
    logits, loss = _fwd_bwd(model, g, inp, dev)
    assert rel_err(logits[:, ::4], g["logits_every4"]) < tol
    assert rel_err(torch.logsumexp(logits.float(), -1), g["logsumexp"]) < tol
    agree = (logits.argmax(-1).cpu() == g["argmax"]).float().mean().item()
    assert agree > 0.999, f"argmax agreement {agree}"
    assert abs(loss.item() - g["loss"].item()) < tol * abs(g["loss"].item())
    for k, p in model.named_parameters():
        n_ref = g["grad_norms"][k].item()
        assert abs(p.grad.float().norm().item() - n_ref) < 5 * gtol * max(n_ref, 1e-12), f"|grad {k}|"
        if k in g["grads_small"]:
            e = rel_err(p.grad, g["grads_small"][k])
            assert e < gtol, f"grad {k}: rel err {e:.3e}"


def _rollout(step_fn, g, n):
    tok = torch.ones(g["B"], 1, dtype=torch.long)
    states, lgs, ids = None, [], []
    for i in range(n):
        lg, states = step_fn(tok, states, i)
        tok = lg.argmax(-1).cpu()
        lgs.append(lg.float().cpu())
        ids.append(tok)
    return torch.cat(lgs, 1), torch.cat(ids, 1)


# ---- CPU: the oracle restatement against the reference's outputs -------------------------------------
def test_oracle_matches_reference_small():
    g, inp = _case("small")
    ref = _load(MambaTTSDecoderRef(**g["config"]).eval(), g)
    _check_small(ref, g, inp, "cpu", 1e-6, 1e-5)
    with torch.no_grad():
        kw = dict(text_mask=inp["text_mask"], ref_hidden=inp["ref_hidden"], ref_mask=inp["ref_mask"])
        lgs, ids = _rollout(lambda tok, st, i: ref.decode_step(tok, inp["text_hidden"], inp["z_style"], st, i, **kw),
                            g, 64)
    assert torch.equal(ids, g["greedy_ids"])
    assert rel_err(lgs, g["step_logits"]) < 1e-6


def test_oracle_matches_reference_c1():
    g, inp = _case("c1")
    ref = _load(MambaTTSDecoderRef(**g["config"]).eval(), g)
    _check_c1(ref, g, inp, "cpu", 1e-6, 1e-5)


def test_fixture_provenance():
    for name in ("small", "c1"):
        assert "/root/reference/mamba_decoder.py" in load_golden(f"ref_decoder_{name}.pt")["source"]


# ---- GPU: the CUDA decoder against the reference's outputs ------------------------------------------
@pytest.fixture
def no_tf32():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    yield


@pytest.mark.gpu
def test_cuda_decoder_matches_reference_small(no_tf32):
    from mamba_tts_project_b200 import MambaTTSDecoder
    g, inp = _case("small")
    dec = _load(MambaTTSDecoder(**g["config"]), g).cuda().eval()
    _check_small(dec, g, inp, "cuda", FP32_TOL, 5e-4)
    c = lambda k: None if inp[k] is None else inp[k].cuda()
    kw = dict(text_mask=c("text_mask"), ref_hidden=c("ref_hidden"), ref_mask=c("ref_mask"))
    lgs, ids = _rollout(lambda tok, st, i: dec.decode_step(tok.cuda(), c("text_hidden"), c("z_style"), st, i, **kw),
                        g, 64)
    assert torch.equal(ids, g["greedy_ids"]), "greedy ids differ from the reference's on the fp32 path"
    assert rel_err(lgs, g["step_logits"]) < FP32_TOL
    for graph in (False, True):
        out = dec.generate(torch.ones(g["B"], 1, dtype=torch.long, device="cuda"), 64, c("text_hidden"),
                           c("z_style"), use_cuda_graph=graph, **kw)
        assert torch.equal(out.cpu(), g["greedy_ids"]), f"generate(use_cuda_graph={graph})"


@pytest.mark.gpu
def test_cuda_decoder_matches_reference_c1(no_tf32):
    from mamba_tts_project_b200 import MambaTTSDecoder
    g, inp = _case("c1")
    dec = _load(MambaTTSDecoder(**g["config"]), g).cuda().eval()
    _check_c1(dec, g, inp, "cuda", FP32_TOL, 5e-4)
