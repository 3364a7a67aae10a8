"""mamba_tts_project_b200 -- B200 (sm_100a) implementation of the MambaTTSDecoder hot path of
whcorkran/mamba-TTS-project (``mamba_decoder.py``): hand-written CUDA behind a C ABI
(``include/mamba_tts_b200.h``), PyTorch host code mirroring the reference's operator surface.

    from mamba_tts_project_b200 import Mamba                 # drop-in for mamba_ssm.Mamba as used
    from mamba_tts_project_b200 import MambaTTSDecoder       # drop-in for mamba_decoder.MambaTTSDecoder
    from mamba_tts_project_b200.ops import selective_scan_fn, selective_state_update, causal_conv1d_fn

CUDA-only.  The shared library must have been built (``python -m mamba_tts_project_b200.build``);
importing this package does not need a GPU, calling any op does.
"""
from . import _lib
from .style import (LengthRegulator, StyleConditioningPipeline, StyleDecoderCrossAttention, StyleProjection,
                    StyleTextCrossAttention)
from .data import PreprocessedItems, collate_codec_batch
from .decoder import (CrossAttention, GenerationContext, MambaTTSDecoder, MambaTTSDecoderLayer,
                      flatten_codes, unflatten_codes)
from .mamba import Mamba
from .training import GraphedForwardBackward, TrainStep, codec_ce_loss, embed_codec_tokens
from .ops import (causal_conv1d_fn, causal_conv1d_update, cross_attn_decode, cross_attn_block_decode, add_layernorm,
                  mamba_decode_step, mamba_inner_fn, selective_scan_fn, selective_state_update,
                  skinny_linear, gemm_bf16, bias_gelu, colsum)

__all__ = ["TrainStep", "GraphedForwardBackward", "codec_ce_loss", "embed_codec_tokens", "Mamba", "MambaTTSDecoder", "MambaTTSDecoderLayer", "CrossAttention", "GenerationContext", "flatten_codes", "unflatten_codes", "LengthRegulator", "StyleConditioningPipeline", "StyleProjection", "StyleTextCrossAttention",
           "StyleDecoderCrossAttention", "PreprocessedItems", "collate_codec_batch",
           "selective_scan_fn", "selective_state_update", "causal_conv1d_fn", "causal_conv1d_update",
           "mamba_inner_fn", "mamba_decode_step", "cross_attn_decode", "cross_attn_block_decode", "add_layernorm", "skinny_linear", "gemm_bf16", "bias_gelu", "colsum"]
__version__ = "0.1.0"


def library_path():
    return _lib.LIB_PATH
