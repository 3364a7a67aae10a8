// Library-level entry points of the C ABI (mamba_tts_b200.h).
#include <cuda_runtime.h>

#include "mamba_tts_b200.h"

extern "C" const char* mtts_error_string(int code) {
  if (code < 0) return cudaGetErrorString(static_cast<cudaError_t>(-code));
  switch (code) {
    case MTTS_OK: return "ok";
    case MTTS_ERR_NULL: return "a required pointer is NULL";
    case MTTS_ERR_SHAPE: return "a dimension is outside the supported range";
    case MTTS_ERR_DTYPE: return "unknown io dtype (expected MTTS_F32 or MTTS_BF16)";
    case MTTS_ERR_ALIGN: return "pointer or stride alignment not supported";
    case MTTS_ERR_UNSUPPORTED: return "unsupported variant";
    default: return "unknown error";
  }
}

// The one piece of process-wide state: which selective-scan kernel family serves a shape (0 = the library's
// choice).  A test / measurement hook, set explicitly -- the launch path reads no environment variables.
#include <atomic>
namespace mtts {
static std::atomic<int> g_scan_impl{0};
int scan_impl_override() { return g_scan_impl.load(std::memory_order_relaxed); }
}  // namespace mtts
extern "C" int mtts_set_scan_impl(int impl) {
  if (impl < 0 || impl > 2) return MTTS_ERR_UNSUPPORTED;
  mtts::g_scan_impl.store(impl, std::memory_order_relaxed);
  return MTTS_OK;
}

extern "C" int mtts_abi_version(void) { return 1; }

extern "C" int mtts_target_sm(void) { return 100; }

// sizeof() of every parameter struct, so a foreign-language binding (ctypes / cgo / JNI) can check
// its mirror of the layout at load time.  `which` follows the declaration order in the header.
extern "C" int mtts_sizeof_params(int which) {
  switch (which) {
    case 0: return (int)sizeof(mtts_conv1d_fwd_params);
    case 1: return (int)sizeof(mtts_conv1d_bwd_params);
    case 2: return (int)sizeof(mtts_conv1d_update_params);
    case 3: return (int)sizeof(mtts_scan_fwd_params);
    case 4: return (int)sizeof(mtts_scan_bwd_params);
    case 5: return (int)sizeof(mtts_state_update_params);
    case 6: return (int)sizeof(mtts_decode_step_params);
    case 7: return (int)sizeof(mtts_cross_attn_decode_params);
    case 8: return (int)sizeof(mtts_add_layernorm_fwd_params);
    case 9: return (int)sizeof(mtts_add_layernorm_bwd_params);
    case 10: return (int)sizeof(mtts_skinny_linear_params);
    case 11: return (int)sizeof(mtts_gemm_bf16_params);
    case 12: return (int)sizeof(mtts_bias_gelu_params);
    case 13: return (int)sizeof(mtts_cross_attn_block_params);
    case 14: return (int)sizeof(mtts_decode_embed_params);
    case 15: return (int)sizeof(mtts_decode_greedy_params);
    case 16: return (int)sizeof(mtts_length_regulate_fwd_params);
    case 17: return (int)sizeof(mtts_length_regulate_bwd_params);
    case 18: return (int)sizeof(mtts_gemm_params);
    case 19: return (int)sizeof(mtts_embed_sum_params);
    case 20: return (int)sizeof(mtts_ce_loss_params);
    case 21: return (int)sizeof(mtts_adam_params);
    case 22: return (int)sizeof(mtts_adam_tensor);
    case 23: return (int)sizeof(mtts_film_ffn_params);
    case 24: return (int)sizeof(mtts_cross_attn_params);
    case 25: return (int)sizeof(mtts_add_layernorm_finish_params);
    case 26: return (int)sizeof(mtts_attn_core_bwd_params);
    case 27: return (int)sizeof(mtts_attn_core_fwd_params);
    default: return -1;
  }
}
