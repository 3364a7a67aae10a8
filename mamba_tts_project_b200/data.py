"""Reader for the reference's preprocessed on-disk format, feeding the decoder directly.

The reference's offline ETL writes, per utterance, ``tensors/{item}_codec.pt`` (FACodec ids, ``(1, T, Q)`` or
``(T, Q)``), ``{item}_phonemes.pt`` (phoneme ids), ``{item}_style.pt`` (BERT style embedding) and optionally
``{item}_spk_emb.pt``, plus one ``metadata.json`` (``data_utils/preprocess.py:272-305``; the parallel writer
``data_utils/preprocess_parallel.py:307-329`` stores the same files as tensors instead of numpy arrays / lists).
``train.py`` never reads them back: it re-encodes every wav through FACodec via temporary files on every step
(``train.py:99-112,177-184``).  This module is that missing reader: host-side I/O only (CPU tensors), the
collate function reproduces ``train.py:181-185`` so the batch drops into ``MambaTTSDecoder.forward`` /
``codec_ce_loss`` unchanged.
"""
from __future__ import annotations

import json
import os

import numpy as np
import torch


def _safe_name(item_name: str) -> str:
    return item_name.replace("/", "_").replace(" ", "_")   # preprocess.py:275 / preprocess_parallel.py:309


def _numpy_safe_globals():
    """The globals a pickled ``numpy.ndarray`` needs (what the sequential writer stores through
    ``torch.save``, ``preprocess.py:278-289``) -- allow-listed so ``weights_only=True`` can read them."""
    core = getattr(np, "_core", None) or np.core
    out = [core.multiarray._reconstruct, np.ndarray, np.dtype]
    out += [type(np.dtype(t)) for t in (np.int16, np.int32, np.int64, np.uint8, np.float16, np.float32,
                                        np.float64, np.bool_)]
    return out


def _load(path, allow_pickle: bool = False):
    """One ``tensors/*.pt`` file -> tensor.  Loaded with ``weights_only=True`` (tensors, lists and -- through
    an explicit allow-list -- numpy arrays): a dataset directory is untrusted input and must not be able to
    run code.  ``allow_pickle=True`` is the explicit opt-in for files that need the full unpickler."""
    if allow_pickle:
        obj = torch.load(path, map_location="cpu", weights_only=False)
    else:
        with torch.serialization.safe_globals(_numpy_safe_globals()):
            obj = torch.load(path, map_location="cpu", weights_only=True)
    if isinstance(obj, np.ndarray):
        obj = torch.from_numpy(obj)
    elif isinstance(obj, (list, tuple)):
        obj = torch.tensor(obj)
    if not isinstance(obj, torch.Tensor):
        raise ValueError(f"{path}: expected a tensor, numpy array or list, got {type(obj).__name__}")
    return obj


class PreprocessedItems(torch.utils.data.Dataset):
    """``output_dir`` as written by ``DatasetPreprocessor.preprocess`` / ``ParallelDatasetPreprocessor``.

    Item = dict(item_name, codec (T, Q) int64, phoneme_ids (T_text,) int64, style (d_style,) float32,
    spk_emb (d_spk,) float32 or None, meta = the utterance's metadata.json record).  Utterances whose codec file
    is missing (the writers skip it when the audio was not found) are dropped unless ``require_codec=False``.
    Files are read without the general unpickler unless ``allow_pickle=True`` (see ``_load``)."""

    def __init__(self, output_dir: str, require_codec: bool = True, allow_pickle: bool = False):
        self.allow_pickle = allow_pickle
        self.tensors_dir = os.path.join(output_dir, "tensors")
        with open(os.path.join(output_dir, "metadata.json")) as f:
            meta = json.load(f)
        if not isinstance(meta, list):
            raise ValueError("metadata.json must hold a list of utterance records")
        self.records = []
        for rec in meta:
            name = _safe_name(rec["item_name"])
            has_codec = os.path.exists(os.path.join(self.tensors_dir, f"{name}_codec.pt"))
            if has_codec or not require_codec:
                self.records.append(rec)

    def __len__(self):
        return len(self.records)

    def __getitem__(self, i):
        rec = self.records[i]
        name = _safe_name(rec["item_name"])
        path = lambda kind: os.path.join(self.tensors_dir, f"{name}_{kind}.pt")
        codec = None
        if os.path.exists(path("codec")):
            codec = _load(path("codec"), self.allow_pickle).long()
            if codec.dim() == 3 and codec.shape[0] == 1:      # FACodecEncoder.encode of one file: (1, T, Q)
                codec = codec[0]
            if codec.dim() != 2:
                raise ValueError(f"{path('codec')}: expected (T, Q) or (1, T, Q), got {tuple(codec.shape)}")
        spk = _load(path("spk_emb"), self.allow_pickle).float().reshape(-1) if os.path.exists(path("spk_emb")) else None
        return {"item_name": rec["item_name"], "codec": codec,
                "phoneme_ids": _load(path("phonemes"), self.allow_pickle).long().reshape(-1),
                "style": _load(path("style"), self.allow_pickle).float().reshape(-1), "spk_emb": spk, "meta": rec}


def collate_codec_batch(items, pad_id: int = 0):
    """Batch of ``PreprocessedItems`` -> the tensors ``train.py:177-188`` builds on the fly.

    Returns dict(audio_tokens (B, Q*T) int64 in the flattened quantizer-major order of ``train.py:181-182``,
    audio_tokens_3d (B, Q, T), codec_pad_mask (B, Q*T) bool (True = pad, ``:183``), codec_lengths (B,),
    phoneme_ids (B, T_text) zero-padded, text_mask (B, T_text) bool (True = valid, the decoder's convention),
    style (B, d_style), spk_emb (B, d_spk) or None, item_names).  Shorter utterances are padded with ``pad_id``
    (FACodec pads with zeros, ``train.py:184``), which ``codec_ce_loss`` ignores."""
    if not items:
        raise ValueError("empty batch")
    B = len(items)
    Q = items[0]["codec"].shape[1]
    T = max(it["codec"].shape[0] for it in items)
    codec = torch.full((B, T, Q), pad_id, dtype=torch.long)
    lengths = torch.zeros(B, dtype=torch.long)
    for b, it in enumerate(items):
        c = it["codec"]
        if c.shape[1] != Q:
            raise ValueError("all utterances must have the same number of quantizers")
        codec[b, : c.shape[0]] = c
        lengths[b] = c.shape[0]
    tokens_3d = codec.permute(0, 2, 1).contiguous()            # (B, Q, T)      train.py:181
    audio_tokens = tokens_3d.reshape(B, -1)                    # (B, Q*T)       train.py:182
    pad_mask = (tokens_3d == pad_id).reshape(B, -1)            #                train.py:183
    Tt = max(it["phoneme_ids"].numel() for it in items)
    ph = torch.zeros(B, Tt, dtype=torch.long)
    text_mask = torch.zeros(B, Tt, dtype=torch.bool)
    for b, it in enumerate(items):
        n = it["phoneme_ids"].numel()
        ph[b, :n] = it["phoneme_ids"]
        text_mask[b, :n] = True
    spk = None
    if all(it["spk_emb"] is not None for it in items):
        spk = torch.stack([it["spk_emb"] for it in items])
    return {"audio_tokens": audio_tokens, "audio_tokens_3d": tokens_3d, "codec_pad_mask": pad_mask,
            "codec_lengths": lengths, "phoneme_ids": ph, "text_mask": text_mask,
            "style": torch.stack([it["style"] for it in items]), "spk_emb": spk,
            "item_names": [it["item_name"] for it in items]}
