"""CPU: the C-ABI library loads and exports every symbol the header declares; host-side logic."""
import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT, load_golden

import mamba_tts_project_b200 as mt
from mamba_tts_project_b200 import _lib
from oracle.decoder_ref import MambaTTSDecoderRef


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "mamba_tts_b200.h")).read()
    return sorted(set(re.findall(r"\b(mtts_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    syms = _header_symbols()
    assert len(syms) >= 13
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in the header but not exported"
    assert set(syms) == set(_lib.ENTRY_POINTS), "python binding and header disagree on entry points"


def test_struct_layouts_match_and_metadata():
    lib = _lib.load()  # raises on any sizeof mismatch
    assert lib.mtts_abi_version() == 1
    assert lib.mtts_target_sm() == 100
    assert lib.mtts_sizeof_params(len(_lib.PARAM_STRUCTS)) == -1
    assert lib.mtts_error_string(2).decode().startswith("a dimension")
    hdr = open(os.path.join(ROOT, "include", "mamba_tts_b200.h")).read()
    assert f"#define MTTS_SCAN_CHUNK {_lib.SCAN_CHUNK}" in hdr
    assert f"#define MTTS_MAX_DSTATE {_lib.MAX_DSTATE}" in hdr


def test_argument_errors_need_no_gpu():
    """Argument validation happens before any launch, so it is testable without a device."""
    lib = _lib.load()
    p = _lib.ScanFwdParams()  # all NULL
    assert lib.mtts_selective_scan_fwd(ctypes.byref(p), None) == 1  # MTTS_ERR_NULL
    c = _lib.Conv1dFwdParams(batch=1, dim=1, seqlen=1, width=7, x=1, weight=1, out=1)
    assert lib.mtts_causal_conv1d_fwd(ctypes.byref(c), None) == 2   # MTTS_ERR_SHAPE


def test_no_cpu_fallback():
    """The product path must fail loudly on CPU tensors instead of computing somewhere else."""
    blk = mt.Mamba(32)
    with pytest.raises(RuntimeError, match="CUDA-only"):
        blk(torch.randn(1, 4, 32))
    with pytest.raises(RuntimeError, match="CUDA-only|no CPU path"):
        mt.selective_scan_fn(torch.randn(1, 4, 8), torch.randn(1, 4, 8), -torch.rand(4, 2),
                             torch.randn(1, 2, 8), torch.randn(1, 2, 8))
    with pytest.raises(RuntimeError, match="no CPU path"):
        mt.causal_conv1d_fn(torch.randn(1, 4, 8), torch.randn(4, 4))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "mamba_tts_project_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle|oracle[./]", txt, re.M), \
                    f"{f} reaches into the oracle"


def test_state_dict_keys_match_reference_tree():
    g = load_golden("oracle_decoder_small.pt")
    dec = mt.MambaTTSDecoder(**g["config"])
    ref = MambaTTSDecoderRef(**g["config"])
    assert list(dec.state_dict().keys()) == list(ref.state_dict().keys())
    res = dec.load_state_dict(g["state_dict"], strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    # the names SURVEY.md 8b lists for the reference module tree
    keys = set(dec.state_dict())
    for k in ["token_embed.weight", "pos_embed.weight", "quant_embed.weight",
              "layers.0.norm_mamba.weight", "layers.0.mamba.in_proj.weight",
              "layers.0.mamba.conv1d.weight", "layers.0.mamba.conv1d.bias",
              "layers.0.mamba.x_proj.weight", "layers.0.mamba.dt_proj.weight",
              "layers.0.mamba.dt_proj.bias", "layers.0.mamba.A_log", "layers.0.mamba.D",
              "layers.0.mamba.out_proj.weight", "layers.0.cross_attn.in_proj_weight",
              "layers.0.cross_attn.in_proj_bias", "layers.0.cross_attn.out_proj.weight",
              "layers.0.cross_attn.out_proj.bias", "layers.0.ff.0.weight", "layers.0.ff.2.bias",
              "layers.0.style_mlp.0.weight", "norm_out.bias", "head.weight"]:
        assert k in keys
    mha = torch.nn.MultiheadAttention(64, 4, batch_first=True)
    assert set(mha.state_dict()) == {k.split("cross_attn.")[1] for k in keys
                                     if k.startswith("layers.0.cross_attn.")}


def test_mamba_init_matches_upstream_recipe():
    torch.manual_seed(0)
    m = mt.Mamba(64)
    assert m.d_inner == 128 and m.dt_rank == 4 and m.d_state == 16 and m.d_conv == 4
    assert torch.allclose(m.A_log[5], torch.log(torch.arange(1, 17.0)))
    assert torch.all(m.D == 1)
    dt = torch.nn.functional.softplus(m.dt_proj.bias)
    assert dt.min() >= 1e-4 - 1e-7 and dt.max() <= 0.1 + 1e-6
    assert m.dt_proj.weight.abs().max() <= 4 ** -0.5 + 1e-6


def test_decoder_argument_errors_match_reference():
    dec = mt.MambaTTSDecoder(16, d_model=32, n_layers=1, n_heads=2, d_ff=64, d_style=8, max_len=32)
    with pytest.raises(ValueError, match="audio_tokens must be"):
        dec(torch.zeros(3, dtype=torch.long), torch.randn(3, 2, 32), torch.randn(3, 8))
    with pytest.raises(AssertionError, match="text_mask"):
        dec(torch.zeros(3, 4, dtype=torch.long), torch.randn(3, 2, 32), torch.randn(3, 8),
            text_mask=torch.ones(2, 2, dtype=torch.bool))


def test_caller_contract_helpers_cpu():
    """embed_codec_tokens / codec_ce_loss restate train.py:115-131 / :31-42 (pure host logic)."""
    torch.manual_seed(0)
    dec = mt.MambaTTSDecoder(16, d_model=32, n_layers=1, n_heads=2, d_ff=64, d_style=8, max_len=32,
                             num_quantizers=3)
    tok = torch.randint(0, 16, (2, 3, 5))
    ref_hidden, mask = mt.embed_codec_tokens(tok, dec)
    assert ref_hidden.shape == (2, 15, 32) and mask.shape == (2, 15)
    # element (b, q, t) = token_embed[tok] + pos_embed[t] + quant_embed[q]
    b, q, t = 1, 2, 3
    want = dec.token_embed.weight[tok[b, q, t]] + dec.pos_embed.weight[t] + dec.quant_embed.weight[q]
    assert torch.allclose(ref_hidden[b, q * 5 + t], want)
    assert torch.equal(mask, (tok == 0).reshape(2, 15))
    logits, tgt = torch.randn(2, 7, 16), torch.randint(0, 16, (2, 7))
    tgt[0, :3] = 0
    want = torch.nn.functional.cross_entropy(logits.view(14, 16), tgt.view(14), ignore_index=0)
    assert torch.allclose(mt.codec_ce_loss(logits, tgt), want)


def test_flatten_unflatten_codes():
    from mamba_tts_project_b200 import flatten_codes, unflatten_codes
    codes = torch.arange(2 * 5 * 7).reshape(2, 5, 7)
    flat = flatten_codes(codes)
    assert flat.shape == (2, 35) and torch.equal(flat[0, :7], codes[0, 0]) and torch.equal(flat[1, 7:14], codes[1, 1])
    assert torch.equal(unflatten_codes(flat, 5), codes)
    # train.py:181-182 builds the same order from the codec's (B, T, Q) output
    assert torch.equal(codes.permute(0, 2, 1).permute(0, 2, 1).reshape(2, -1), flat)
    with pytest.raises(ValueError):
        unflatten_codes(flat[:, :34], 5)


def _write_reference_style(tmp_path, parallel):
    """Files exactly as the reference's two writers leave them (preprocess.py:272-305 pickles numpy arrays and
    lists through torch.save; preprocess_parallel.py:307-329 stores tensors)."""
    import json
    import numpy as np
    tdir = tmp_path / "tensors"
    tdir.mkdir()
    g = torch.Generator().manual_seed(3)
    meta, truth = [], {}
    for i, (name, T, Tt) in enumerate((("spk1/utt 01", 7, 5), ("spk2/utt_02", 4, 9), ("spk3/no audio", 0, 3))):
        safe = name.replace("/", "_").replace(" ", "_")
        ph = torch.randint(1, 80, (Tt,), generator=g).tolist()
        style = torch.randn(16, generator=g).numpy()
        codec = torch.randint(1, 1024, (1, T, 5), generator=g).numpy() if T else None
        spk = torch.randn(1, 8, generator=g).numpy()
        if parallel:
            torch.save(torch.tensor(ph, dtype=torch.long), tdir / f"{safe}_phonemes.pt")
            torch.save(torch.from_numpy(style), tdir / f"{safe}_style.pt")
            if codec is not None:
                torch.save(torch.from_numpy(codec), tdir / f"{safe}_codec.pt")
        else:
            torch.save(ph, tdir / f"{safe}_phonemes.pt")
            torch.save(style, tdir / f"{safe}_style.pt")
            if codec is not None:
                torch.save(codec, tdir / f"{safe}_codec.pt")
                torch.save(spk, tdir / f"{safe}_spk_emb.pt")
        meta.append({"item_name": name, "text": "t", "phonemes": [], "phoneme_str": "", "ph2word": [],
                     "style_prompt": "calm"})
        truth[name] = (ph, style, codec)
    with open(tmp_path / "metadata.json", "w") as f:
        json.dump(meta, f, indent=2)
    return truth


@pytest.mark.parametrize("parallel", [False, True], ids=["sequential_writer", "parallel_writer"])
def test_preprocessed_reader_and_collate(tmp_path, parallel):
    from mamba_tts_project_b200.data import PreprocessedItems, collate_codec_batch
    truth = _write_reference_style(tmp_path, parallel)
    ds = PreprocessedItems(str(tmp_path))
    assert len(ds) == 2                                   # the utterance without a codec file is dropped
    assert len(PreprocessedItems(str(tmp_path), require_codec=False)) == 3
    items = [ds[0], ds[1]]
    for it in items:
        ph, style, codec = truth[it["item_name"]]
        assert it["phoneme_ids"].tolist() == ph and it["codec"].shape == (codec.shape[1], 5)
        assert torch.equal(it["codec"], torch.from_numpy(codec)[0].long())
        assert torch.allclose(it["style"], torch.from_numpy(style))
        assert (it["spk_emb"] is None) == parallel
    batch = collate_codec_batch(items)
    # train.py:181-185 on the zero-padded (B, T, C) codec batch
    codec = torch.zeros(2, 7, 5, dtype=torch.long)
    codec[0] = items[0]["codec"]
    codec[1, :4] = items[1]["codec"]
    want_3d = codec.permute(0, 2, 1)
    assert torch.equal(batch["audio_tokens_3d"], want_3d)
    assert torch.equal(batch["audio_tokens"], want_3d.reshape(2, -1))
    assert torch.equal(batch["codec_pad_mask"], (want_3d == 0).reshape(2, -1))
    assert batch["codec_lengths"].tolist() == [7, 4]
    assert batch["phoneme_ids"].shape == (2, 9) and batch["text_mask"].sum(1).tolist() == [5, 9]
    assert torch.equal(mt.unflatten_codes(batch["audio_tokens"], 5), want_3d)
    assert batch["style"].shape == (2, 16) and (batch["spk_emb"] is None) == parallel


def test_film_terms_equal_the_per_layer_style_mlps():
    """MambaTTSDecoder.film_terms stacks every layer's style_mlp (mamba_decoder.py:45-48,81-83) into one contraction:
    pure host logic, the same (gamma, beta) as calling the layers' own Sequential(Linear, Tanh) one by one."""
    import torch
    from mamba_tts_project_b200 import MambaTTSDecoder
    torch.manual_seed(0)
    dec = MambaTTSDecoder(vocab_size_audio=32, d_model=48, n_layers=3, n_heads=4, d_ff=64, d_style=24, max_len=16)
    z = torch.randn(5, 24)
    films = dec.film_terms(z)
    assert len(films) == 3
    for layer, (gamma, beta) in zip(dec.layers, films):
        g_ref, b_ref = torch.chunk(layer.style_mlp(z), 2, dim=-1)
        assert gamma.is_contiguous() and beta.is_contiguous()
        assert torch.allclose(gamma, g_ref, atol=1e-6) and torch.allclose(beta, b_ref, atol=1e-6)
    # gradients reach every layer's own parameters through the stacked weight
    sum(g.sum() + b.sum() for g, b in films).backward()
    assert all(l.style_mlp[0].weight.grad is not None and l.style_mlp[0].bias.grad is not None for l in dec.layers)
