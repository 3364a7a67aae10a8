"""ctypes binding of the C-ABI library (``include/mamba_tts_b200.h``).

This is the *only* way the Python host code reaches the kernels: POD structs of raw device
pointers, sizes and element strides plus the current CUDA stream.  There is no fallback: if the
shared library is missing or a call returns non-zero, a ``RuntimeError`` is raised (SURVEY.md 8b
"Errors"; upstream raises from ``TORCH_CHECK``).
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# MTTS_LIB: another build of the same library (A/B measurements of compile-time variants)
LIB_PATH = os.environ.get("MTTS_LIB") or os.path.join(_HERE, "libmamba_tts_b200.so")

F32, BF16 = 0, 1
SCAN_CHUNK = 32
MAX_DSTATE = 256
MAX_CONV_WIDTH = 4

_i32, _i64, _ptr, _f32 = C.c_int32, C.c_int64, C.c_void_p, C.c_float


def _struct(name, spec):
    """spec: whitespace separated 'type:name' tokens; type in i (int32) l (int64) p (pointer) f (float)."""
    kinds = {"i": _i32, "l": _i64, "p": _ptr, "f": _f32}
    fields = []
    for tok in spec.split():
        k, n = tok.split(":")
        fields.append((n, kinds[k]))
    return type(name, (C.Structure,), {"_fields_": fields})


Conv1dFwdParams = _struct("Conv1dFwdParams", """
    i:batch i:dim i:seqlen i:width i:io_dtype i:silu
    p:x l:x_batch_stride l:x_dim_stride p:weight p:bias
    p:initial_states l:init_batch_stride l:init_dim_stride
    p:out l:out_batch_stride l:out_dim_stride""")

Conv1dBwdParams = _struct("Conv1dBwdParams", """
    i:batch i:dim i:seqlen i:width i:io_dtype i:silu
    p:x l:x_batch_stride l:x_dim_stride p:weight p:bias
    p:initial_states l:init_batch_stride l:init_dim_stride
    p:dout l:dout_batch_stride l:dout_dim_stride
    p:dx l:dx_batch_stride l:dx_dim_stride p:dweight p:dbias""")

Conv1dUpdateParams = _struct("Conv1dUpdateParams", """
    i:batch i:dim i:width i:io_dtype i:silu
    p:x l:x_batch_stride p:conv_state p:weight p:bias p:out l:out_batch_stride""")

ScanFwdParams = _struct("ScanFwdParams", """
    i:batch i:dim i:seqlen i:dstate i:io_dtype i:delta_softplus
    p:u l:u_batch_stride l:u_dim_stride
    p:delta l:delta_batch_stride l:delta_dim_stride
    p:A
    p:B l:B_batch_stride l:B_state_stride
    p:C l:C_batch_stride l:C_state_stride
    p:D p:delta_bias
    p:z l:z_batch_stride l:z_dim_stride
    p:initial_state
    p:out l:out_batch_stride l:out_dim_stride
    p:last_state p:checkpoints
    p:y_pre l:y_batch_stride l:y_dim_stride""")

ScanBwdParams = _struct("ScanBwdParams", """
    i:batch i:dim i:seqlen i:dstate i:io_dtype i:delta_softplus
    p:u l:u_batch_stride l:u_dim_stride
    p:delta l:delta_batch_stride l:delta_dim_stride
    p:A
    p:B l:B_batch_stride l:B_state_stride
    p:C l:C_batch_stride l:C_state_stride
    p:D p:delta_bias
    p:z l:z_batch_stride l:z_dim_stride
    p:dout l:dout_batch_stride l:dout_dim_stride
    p:checkpoints
    p:du l:du_batch_stride l:du_dim_stride
    p:ddelta l:ddelta_batch_stride l:ddelta_dim_stride
    p:dz l:dz_batch_stride l:dz_dim_stride
    p:dA p:dB p:dC p:dD p:ddelta_bias
    p:y_pre l:y_batch_stride l:y_dim_stride""")

StateUpdateParams = _struct("StateUpdateParams", """
    i:batch i:dim i:dstate i:io_dtype i:dt_softplus
    p:state p:x l:x_batch_stride p:dt l:dt_batch_stride p:A
    p:B l:B_batch_stride p:C l:C_batch_stride p:D
    p:z l:z_batch_stride p:dt_bias p:out l:out_batch_stride""")

DecodeStepParams = _struct("DecodeStepParams", """
    i:batch i:dim i:dstate i:dt_rank i:width i:io_dtype
    p:xz l:xz_batch_stride p:conv_state p:ssm_state p:conv_weight p:conv_bias
    p:x_proj_w p:dt_proj_w p:dt_bias p:A p:D p:y l:y_batch_stride
    p:prefetch_a p:prefetch_b l:prefetch_bytes""")

CrossAttnDecodeParams = _struct("CrossAttnDecodeParams", """
    i:batch i:heads i:head_dim i:t_kv i:io_dtype
    p:q p:k p:v p:mask p:out""")

AddLayerNormFwdParams = _struct("AddLayerNormFwdParams", """
    i:rows i:dim i:rows_per_batch i:io_dtype f:eps
    p:x p:delta p:delta_bias p:x_out p:ln_weight p:ln_bias p:film_gamma p:film_beta p:out p:mean p:rstd""")

AddLayerNormBwdParams = _struct("AddLayerNormBwdParams", """
    i:rows i:dim i:rows_per_batch i:io_dtype
    p:x_out p:mean p:rstd p:ln_weight p:film_gamma p:dout p:dx_out p:dx p:ddelta p:colsum""")

SkinnyLinearParams = _struct("SkinnyLinearParams", """
    i:m i:n i:k i:io_dtype i:ln_mode i:gelu f:eps
    p:a p:x p:delta p:x_out p:ln_weight p:ln_bias p:film_gamma p:film_beta p:w p:bias p:out""")

GemmBf16Params = _struct("GemmBf16Params", """
    i:m i:n i:k i:gelu p:a l:lda p:w l:ldw p:bias p:out l:ldo p:pre_out""")

BiasGeluParams = _struct("BiasGeluParams", """
    i:rows i:cols i:io_dtype i:reserved l:ld p:x p:bias p:dout p:out p:colsum""")

CrossAttnBlockParams = _struct("CrossAttnBlockParams", """
    i:batch i:heads i:head_dim i:t_kv i:io_dtype f:eps_q f:eps_o
    p:x p:delta p:x_out p:lnq_weight p:lnq_bias p:wq p:bq p:k p:v p:mask p:wo p:bo
    p:lno_weight p:lno_bias p:film_gamma p:film_beta p:out""")

DecodeEmbedParams = _struct("DecodeEmbedParams", """
    i:batch i:dim p:tok p:pos p:tok_embed p:pos_embed p:x p:step""")

DecodeGreedyParams = _struct("DecodeGreedyParams", """
    i:batch i:vocab i:io_dtype i:reserved p:logits p:tok p:out l:out_stride p:step p:pos
    l:eos_id l:pad_id p:lengths""")

LengthRegulateFwdParams = _struct("LengthRegulateFwdParams", """
    i:batch i:t_text i:dim i:max_len i:io_dtype i:reserved
    p:hidden p:durations p:expanded p:output_lengths p:frame_index""")

LengthRegulateBwdParams = _struct("LengthRegulateBwdParams", """
    i:batch i:t_text i:dim i:max_len i:io_dtype i:reserved p:durations p:dexpanded p:dhidden""")

GemmParams = _struct("GemmParams", """
    i:m i:n i:k i:batch_outer i:batch_inner i:k_batches i:a_major i:b_major i:out_dtype i:epilogue
    i:accumulate i:split_k
    p:a l:lda l:a_bo_stride l:a_bi_stride
    p:b l:ldb l:b_bo_stride l:b_bi_stride
    p:out l:ldc l:c_bo_stride l:c_bi_stride
    p:bias_n p:bias_m
    p:aux l:ld_aux l:aux_bo_stride l:aux_bi_stride
    p:mask l:mask_bo_stride f:scale i:flags p:row_stat""")
GEMM_SINGLE_CTA, GEMM_AUX_GELU_GRAD = 1, 2
EPI_STORE, EPI_GELU, EPI_GELU_BWD, EPI_SOFTMAX, EPI_DSOFTMAX, EPI_MUL_AUX = range(6)

EmbedSumParams = _struct("EmbedSumParams", """
    i:batch i:seqlen i:dim i:reserved p:tokens p:pos_ids p:quant_ids p:token_embed p:pos_embed p:quant_embed p:x""")
CeLossParams = _struct("CeLossParams", """
    l:rows i:vocab i:io_dtype l:ld l:ignore_index p:logits p:targets p:n_valid f:grad_scale i:reserved
    p:loss_sum p:row_loss p:dlogits""")
AdamParams = _struct("AdamParams", """
    p:tensors p:chunks i:num_chunks f:max_norm p:grad_sumsq f:step_size f:beta1 f:beta2 f:eps
    f:bias_correction2_sqrt i:reserved""")
AdamTensor = _struct("AdamTensor", "p:param p:grad p:exp_avg p:exp_avg_sq l:numel")



class FilmFfnParams(C.Structure):
    _fields_ = [("tokens", _i64), ("d_model", _i32), ("d_ff", _i32), ("ln", AddLayerNormFwdParams),
                ("h", _ptr), ("w1", _ptr), ("b1", _ptr), ("w2", _ptr), ("act", _ptr), ("gprime", _ptr), ("f", _ptr),
                ("df", _ptr), ("dpre", _ptr), ("dw1", _ptr), ("db1", _ptr), ("dw2", _ptr), ("dh", _ptr)]


CrossAttnParams = _struct("CrossAttnParams", """
    i:batch i:t_q i:t_kv i:d_model i:heads i:reserved
    p:query p:memory p:w_in p:b_in p:w_out p:mask p:q p:kv p:p p:o p:out p:lse2
    p:dout p:d_o p:ds p:dq p:dkv p:dw_in p:db_in p:dw_out p:dquery p:dmemory""")

AddLayerNormFinishParams = _struct("AddLayerNormFinishParams", """
    i:batch i:dim p:colsum p:film_gamma p:ln_weight p:ln_bias p:dweight p:dbias p:dgamma p:dbeta p:ddelta_bias""")

AttnCoreBwdParams = _struct("AttnCoreBwdParams", """
    i:batch i:heads i:t_q i:t_kv i:d_model f:scale p:q p:kv p:o p:d_o p:lse2 p:mask p:dq p:dkv""")

AttnCoreFwdParams = _struct("AttnCoreFwdParams", """
    i:batch i:heads i:t_q i:t_kv i:d_model f:scale p:q p:kv p:mask p:o p:lse2""")

# argument of mtts_sizeof_params (declaration order of the header, later additions appended)
PARAM_STRUCTS = [Conv1dFwdParams, Conv1dBwdParams, Conv1dUpdateParams, ScanFwdParams,
                 ScanBwdParams, StateUpdateParams, DecodeStepParams, CrossAttnDecodeParams,
                 AddLayerNormFwdParams, AddLayerNormBwdParams, SkinnyLinearParams,
                 GemmBf16Params, BiasGeluParams, CrossAttnBlockParams, DecodeEmbedParams, DecodeGreedyParams,
                 LengthRegulateFwdParams, LengthRegulateBwdParams, GemmParams,
                 EmbedSumParams, CeLossParams, AdamParams, AdamTensor, FilmFfnParams, CrossAttnParams,
                 AddLayerNormFinishParams, AttnCoreBwdParams, AttnCoreFwdParams]

# every symbol include/mamba_tts_b200.h declares -> parameter struct (None: not a kernel call)
ENTRY_POINTS = {
    "mtts_error_string": None,
    "mtts_abi_version": None,
    "mtts_target_sm": None,
    "mtts_set_scan_impl": None,
    "mtts_sizeof_params": None,
    "mtts_causal_conv1d_fwd": Conv1dFwdParams,
    "mtts_causal_conv1d_bwd": Conv1dBwdParams,
    "mtts_causal_conv1d_update": Conv1dUpdateParams,
    "mtts_selective_scan_fwd": ScanFwdParams,
    "mtts_selective_scan_bwd": ScanBwdParams,
    "mtts_selective_state_update": StateUpdateParams,
    "mtts_mamba_decode_step": DecodeStepParams,
    "mtts_cross_attn_decode": CrossAttnDecodeParams,
    "mtts_cross_attn_block_decode": CrossAttnBlockParams,
    "mtts_decode_embed": DecodeEmbedParams,
    "mtts_decode_greedy": DecodeGreedyParams,
    "mtts_length_regulate_fwd": LengthRegulateFwdParams,
    "mtts_length_regulate_bwd": LengthRegulateBwdParams,
    "mtts_add_layernorm_fwd": AddLayerNormFwdParams,
    "mtts_add_layernorm_bwd": AddLayerNormBwdParams,
    "mtts_add_layernorm_bwd_finish": AddLayerNormFinishParams,
    "mtts_skinny_linear": SkinnyLinearParams,
    "mtts_gemm_bf16": GemmBf16Params,
    "mtts_gemm": GemmParams,
    "mtts_film_ffn_fwd": FilmFfnParams,
    "mtts_film_ffn_bwd": FilmFfnParams,
    "mtts_cross_attn_fwd": CrossAttnParams,
    "mtts_cross_attn_bwd": CrossAttnParams,
    "mtts_attn_core_bwd": AttnCoreBwdParams,
    "mtts_attn_core_fwd": AttnCoreFwdParams,
    "mtts_embed_sum_fwd": EmbedSumParams,
    "mtts_embed_sum_bwd": EmbedSumParams,
    "mtts_ce_loss": CeLossParams,
    "mtts_grad_sumsq": AdamParams,
    "mtts_adam_step": AdamParams,
    "mtts_adam_chunk_elems": None,
    "mtts_bias_gelu_fwd": BiasGeluParams,
    "mtts_bias_gelu_bwd": BiasGeluParams,
    "mtts_colsum": BiasGeluParams,
}

_lib = None
launch_count = 0  # kernels launched through this binding (bench.py's ``gpu_launches``)
# bench.py only: {entry point name: [(start_event, end_event), ...]} -- when a name is present, every
# call of that entry point is bracketed by CUDA events on the launching stream.
event_hook = {}


def load():
    """Load the shared library (once).  Raises if it has not been built -- no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m mamba_tts_project_b200.build` "
            "(nvcc, sm_100a).  There is no CPU or PyTorch fallback for this path.")
    lib = C.CDLL(LIB_PATH)
    lib.mtts_error_string.restype = C.c_char_p
    lib.mtts_error_string.argtypes = [C.c_int]
    lib.mtts_abi_version.restype = C.c_int
    lib.mtts_target_sm.restype = C.c_int
    lib.mtts_sizeof_params.restype = C.c_int
    lib.mtts_sizeof_params.argtypes = [C.c_int]
    for name, st in ENTRY_POINTS.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        if st is not None:
            fn.restype = C.c_int
            fn.argtypes = [C.POINTER(st), C.c_void_p]
    for i, st in enumerate(PARAM_STRUCTS):
        if lib.mtts_sizeof_params(i) != C.sizeof(st):
            raise RuntimeError(f"ABI mismatch: {st.__name__} is {C.sizeof(st)} bytes in Python, "
                               f"{lib.mtts_sizeof_params(i)} in the library")
    _lib = lib
    return lib


def set_scan_impl(name) -> None:
    """Force a selective-scan kernel family: None / "auto", "seq" (time-sequential) or "wide" (time-parallel)."""
    rc = load().mtts_set_scan_impl({None: 0, "auto": 0, "seq": 1, "wide": 2}[name])
    if rc != 0:
        raise RuntimeError("mtts_set_scan_impl failed")


def io_dtype(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise RuntimeError(f"unsupported activation dtype {t.dtype} (float32 or bfloat16)")


def ptr(t):
    return None if t is None else t.data_ptr()


def require_cuda(*tensors):
    """Every operand is a CUDA tensor on ONE device, and that device is the current one: ``call`` enqueues
    on the current device's current stream, so operands living elsewhere would be a fault (or a silent peer
    access with the wrong stream ordering).  Wrap calls for another GPU in ``torch.cuda.device(...)``."""
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("mamba_tts_project_b200 ops are CUDA-only (sm_100a); got a "
                               f"{t.device} tensor.  There is no CPU path.")
        if dev is None:
            dev = t.device
            if dev.index != torch.cuda.current_device():
                raise RuntimeError(f"operands live on {dev} but the current device is "
                                   f"cuda:{torch.cuda.current_device()}; use torch.cuda.device({dev.index})")
        elif t.device != dev:
            raise RuntimeError(f"operands on different devices: {dev} and {t.device}")


def call(name: str, params, launches: int = 1) -> None:
    """Enqueue one library call on torch's current stream of the current device.  ``launches``: kernels the call
    launches (the branch-level entry points launch several), for ``launch_count``."""
    global launch_count
    lib = load()
    stream = torch.cuda.current_stream().cuda_stream
    rec = event_hook.get(name) if event_hook else None
    if rec is not None:
        ev0 = torch.cuda.Event(enable_timing=True)
        ev1 = torch.cuda.Event(enable_timing=True)
        ev0.record()
    rc = getattr(lib, name)(C.byref(params), C.c_void_p(stream))
    if rc != 0:
        raise RuntimeError(f"{name} failed ({rc}): {lib.mtts_error_string(rc).decode()}")
    if rec is not None:
        ev1.record()
        rec.append((ev0, ev1))
    launch_count += launches
