// Branch-level entry points of the C ABI (SURVEY 8b "later" list): the FiLM'd FFN and the training-time
// cross-attention of a decoder layer (mamba_decoder.py:32-48,72-89) as ONE call per direction.  Host code only: each
// call enqueues its sequence of mtts_gemm / mtts_add_layernorm / mtts_colsum launches on the caller's stream, in
// the order a consumer needs them; all buffers are the caller's, nothing is allocated or retained.
#include <cmath>
#include <cstring>

#include "common.cuh"

namespace {

struct G : mtts_gemm_params {
  G(int m_, int n_, int k_) {
    std::memset(static_cast<mtts_gemm_params*>(this), 0, sizeof(mtts_gemm_params));
    m = m_; n = n_; k = k_;
    batch_outer = 1; batch_inner = 1; k_batches = 1; split_k = 1;
    out_dtype = MTTS_BF16; epilogue = MTTS_EPI_STORE; scale = 1.f;
  }
  G& A(const void* ptr, int major, long long ld, long long bo = 0, long long bi = 0) {
    a = ptr; a_major = major; lda = ld; a_bo_stride = bo; a_bi_stride = bi; return *this;
  }
  G& B(const void* ptr, int major, long long ld, long long bo = 0, long long bi = 0) {
    b = ptr; b_major = major; ldb = ld; b_bo_stride = bo; b_bi_stride = bi; return *this;
  }
  G& C(void* ptr, long long ld, long long bo = 0, long long bi = 0) {
    out = ptr; ldc = ld; c_bo_stride = bo; c_bi_stride = bi; return *this;
  }
  G& batches(int bo, int bi) { batch_outer = bo; batch_inner = bi; return *this; }
  G& f32_splitk() { out_dtype = MTTS_F32; split_k = -1; return *this; }
};

inline const unsigned char* at(const void* base, long long elems) {   // bf16 element offset
  return reinterpret_cast<const unsigned char*>(base) + 2 * elems;
}
inline unsigned char* at(void* base, long long elems) { return reinterpret_cast<unsigned char*>(base) + 2 * elems; }

int colsum_bf16(const void* x, long long rows, int cols, long long ld, float* out, mtts_stream_t s) {
  mtts_bias_gelu_params c;
  std::memset(&c, 0, sizeof(c));
  c.rows = (int)rows; c.cols = cols; c.io_dtype = MTTS_BF16; c.ld = ld; c.x = x; c.colsum = out;
  return mtts_colsum(&c, s);
}

}  // namespace

#define MTTS_TRY(expr)            \
  do {                            \
    const int rc_ = (expr);       \
    if (rc_ != MTTS_OK) return rc_; \
  } while (0)

extern "C" int mtts_film_ffn_fwd(const mtts_film_ffn_params* p, mtts_stream_t stream) {
  if (!p || !p->h || !p->w1 || !p->w2 || !p->act || !p->f) return MTTS_ERR_NULL;
  if (p->tokens < 0 || p->tokens > 0x7fffffffLL || p->d_model < 8 || p->d_ff < 8 || p->d_model % 8 || p->d_ff % 8)
    return MTTS_ERR_SHAPE;
  if (p->tokens == 0) return MTTS_OK;
  if (p->ln.x) {
    if (p->ln.out != p->h || p->ln.io_dtype != MTTS_BF16 || p->ln.rows != p->tokens || p->ln.dim != p->d_model)
      return MTTS_ERR_SHAPE;
    MTTS_TRY(mtts_add_layernorm_fwd(&p->ln, stream));
  }
  const int T = (int)p->tokens, D = p->d_model, F = p->d_ff;
  G g1(T, F, D);
  g1.A(p->h, 0, D).B(p->w1, 0, D).C(p->act, F);
  g1.bias_n = p->b1;
  g1.epilogue = MTTS_EPI_GELU;
  if (p->gprime) {
    g1.aux = p->gprime; g1.ld_aux = F; g1.flags = MTTS_GEMM_AUX_GELU_GRAD;
  }
  MTTS_TRY(mtts_gemm(&g1, stream));
  G g2(T, D, F);
  g2.A(p->act, 0, F).B(p->w2, 0, F).C(p->f, D);
  return mtts_gemm(&g2, stream);
}

extern "C" int mtts_film_ffn_bwd(const mtts_film_ffn_params* p, mtts_stream_t stream) {
  if (!p || !p->h || !p->w1 || !p->w2 || !p->act || !p->gprime || !p->df || !p->dpre || !p->dw1 || !p->dw2)
    return MTTS_ERR_NULL;
  if (p->tokens < 0 || p->tokens > 0x7fffffffLL || p->d_model < 8 || p->d_ff < 8 || p->d_model % 8 || p->d_ff % 8)
    return MTTS_ERR_SHAPE;
  if (p->tokens == 0) return MTTS_OK;
  const int T = (int)p->tokens, D = p->d_model, F = p->d_ff;
  // dpre = (df W2) o gelu'(pre): W2 (D, F) read as B[n = F, k = D] with n contiguous (MN-major)
  G gd(T, F, D);
  gd.A(p->df, 0, D).B(p->w2, 1, F).C(p->dpre, F);
  gd.epilogue = MTTS_EPI_MUL_AUX; gd.aux = p->gprime; gd.ld_aux = F;
  MTTS_TRY(mtts_gemm(&gd, stream));
  // dW2 (D, F) = df^T act;  dW1 (F, D) = dpre^T h: both operands token-major, i.e. MN-major for these products
  G gw2(D, F, T);
  gw2.A(p->df, 1, D).B(p->act, 1, F).C(p->dw2, F).f32_splitk();
  MTTS_TRY(mtts_gemm(&gw2, stream));
  G gw1(F, D, T);
  gw1.A(p->dpre, 1, F).B(p->h, 1, D).C(p->dw1, D).f32_splitk();
  MTTS_TRY(mtts_gemm(&gw1, stream));
  if (p->db1) MTTS_TRY(colsum_bf16(p->dpre, T, F, F, p->db1, stream));
  if (p->dh) {
    G gh(T, D, F);
    gh.A(p->dpre, 0, F).B(p->w1, 1, D).C(p->dh, D);
    MTTS_TRY(mtts_gemm(&gh, stream));
  }
  return MTTS_OK;
}

static int check_attn(const mtts_cross_attn_params* p, bool need_p) {
  if (!p || !p->query || !p->memory || !p->w_in || !p->b_in || !p->w_out || !p->q || !p->kv || (need_p && !p->p) ||
      !p->o)
    return MTTS_ERR_NULL;
  if (p->batch < 0 || p->t_q < 0 || p->t_kv < 1 || p->t_kv > 256 || p->heads < 1 || p->d_model < 8 ||
      p->d_model % p->heads || (p->d_model / p->heads) % 8)
    return MTTS_ERR_SHAPE;
  return MTTS_OK;
}

extern "C" int mtts_cross_attn_fwd(const mtts_cross_attn_params* p, mtts_stream_t stream) {
  MTTS_TRY(check_attn(p, !(p && p->lse2 != nullptr && p->d_model == p->heads * 64)));
  if (!p->out) return MTTS_ERR_NULL;
  if (p->batch == 0 || p->t_q == 0) return MTTS_OK;
  const int B = p->batch, T = p->t_q, Tk = p->t_kv, E = p->d_model, H = p->heads, dh = E / H;
  const int Tkp = (Tk + 7) / 8 * 8;
  // q = query Wq^T + bq;  kv = memory [Wk; Wv]^T + [bk; bv]
  G gq(B * T, E, E);
  gq.A(p->query, 0, E).B(p->w_in, 0, E).C(p->q, E);
  gq.bias_n = p->b_in;
  MTTS_TRY(mtts_gemm(&gq, stream));
  G gkv(B * Tk, 2 * E, E);
  gkv.A(p->memory, 0, E).B(at(p->w_in, (long long)E * E), 0, E).C(p->kv, 2 * E);
  gkv.bias_n = p->b_in + E;
  MTTS_TRY(mtts_gemm(&gkv, stream));
  if (p->lse2 != nullptr && p->p == nullptr && E == H * 64) {
    // scores, softmax and P V in one launch (attn_sm100.cu): only o and the row log-sum-exp reach memory
    mtts_attn_core_fwd_params c;
    std::memset(&c, 0, sizeof(c));
    c.batch = B; c.heads = H; c.t_q = T; c.t_kv = Tk; c.d_model = E;
    c.scale = 1.f / std::sqrt((float)dh);
    c.q = p->q; c.kv = p->kv; c.mask = p->mask; c.o = p->o; c.lse2 = p->lse2;
    MTTS_TRY(mtts_attn_core_fwd(&c, stream));
  } else {
    // P[b, h] = softmax(scale q_h k_h^T + mask): (batch, head) views of the projections, nothing is transposed
    G gs(T, Tk, dh);
    gs.batches(B, H)
        .A(p->q, 0, E, (long long)T * E, dh)
        .B(p->kv, 0, 2 * E, (long long)Tk * 2 * E, dh)
        .C(p->p, Tkp, (long long)H * T * Tkp, (long long)T * Tkp);
    gs.epilogue = MTTS_EPI_SOFTMAX; gs.mask = p->mask; gs.mask_bo_stride = Tk;
    gs.scale = 1.f / std::sqrt((float)dh);
    gs.row_stat = p->lse2;
    MTTS_TRY(mtts_gemm(&gs, stream));
    // o[b, :, h] = P[b, h] v_h: v read as B[n = dh, k = key] with n contiguous
    G go(T, dh, Tk);
    go.batches(B, H)
        .A(p->p, 0, Tkp, (long long)H * T * Tkp, (long long)T * Tkp)
        .B(at(p->kv, E), 1, 2 * E, (long long)Tk * 2 * E, dh)
        .C(p->o, E, (long long)T * E, dh);
    MTTS_TRY(mtts_gemm(&go, stream));
  }
  G gout(B * T, E, E);
  gout.A(p->o, 0, E).B(p->w_out, 0, E).C(p->out, E);
  return mtts_gemm(&gout, stream);
}

extern "C" int mtts_cross_attn_bwd(const mtts_cross_attn_params* p, mtts_stream_t stream) {
  MTTS_TRY(check_attn(p, !(p && p->lse2 != nullptr && p->d_model == p->heads * 64)));
  const bool fused = p->lse2 != nullptr && p->d_model == p->heads * 64;
  if (!p->dout || !p->d_o || (!fused && !p->ds) || !p->dq || !p->dkv || !p->dw_in || !p->db_in || !p->dw_out)
    return MTTS_ERR_NULL;
  if (p->batch == 0 || p->t_q == 0) return MTTS_OK;
  const int B = p->batch, T = p->t_q, Tk = p->t_kv, E = p->d_model, H = p->heads, dh = E / H;
  const int Tkp = (Tk + 7) / 8 * 8;
  const long long sq = (long long)T * E, skv = (long long)Tk * 2 * E, sp = (long long)H * T * Tkp, sph = (long long)T * Tkp;
  // d o = dout W_o;  dW_o = dout^T o
  G g1(B * T, E, E);
  g1.A(p->dout, 0, E).B(p->w_out, 1, E).C(p->d_o, E);
  MTTS_TRY(mtts_gemm(&g1, stream));
  G g2(E, E, B * T);
  g2.A(p->dout, 1, E).B(p->o, 1, E).C(p->dw_out, E).f32_splitk();
  MTTS_TRY(mtts_gemm(&g2, stream));
  if (fused) {
    // dV, dS, dQ, dK in one launch: P, dP and dS stay in tensor / shared memory (attn_sm100.cu)
    mtts_attn_core_bwd_params c;
    std::memset(&c, 0, sizeof(c));
    c.batch = B; c.heads = H; c.t_q = T; c.t_kv = Tk; c.d_model = E;
    c.scale = 1.f / std::sqrt((float)dh);
    c.q = p->q; c.kv = p->kv; c.o = p->o; c.d_o = p->d_o; c.lse2 = p->lse2; c.mask = p->mask;
    c.dq = p->dq; c.dkv = p->dkv;
    MTTS_TRY(mtts_attn_core_bwd(&c, stream));
  } else {
    // dV[b, h] (Tk, dh) = P^T dO
    G gv(Tk, dh, T);
    gv.batches(B, H).A(p->p, 1, Tkp, sp, sph).B(p->d_o, 1, E, sq, dh).C(at(p->dkv, E), 2 * E, skv, dh);
    MTTS_TRY(mtts_gemm(&gv, stream));
    // dS = scale P o (dP - rowsum(P o dP)),  dP = dO V^T
    G gs(T, Tk, dh);
    gs.batches(B, H).A(p->d_o, 0, E, sq, dh).B(at(p->kv, E), 0, 2 * E, skv, dh).C(p->ds, Tkp, sp, sph);
    gs.epilogue = MTTS_EPI_DSOFTMAX; gs.aux = p->p; gs.ld_aux = Tkp; gs.aux_bo_stride = sp; gs.aux_bi_stride = sph;
    gs.scale = 1.f / std::sqrt((float)dh);
    MTTS_TRY(mtts_gemm(&gs, stream));
    // dQ = dS K;  dK = dS^T Q
    G gq(T, dh, Tk);
    gq.batches(B, H).A(p->ds, 0, Tkp, sp, sph).B(p->kv, 1, 2 * E, skv, dh).C(p->dq, E, sq, dh);
    MTTS_TRY(mtts_gemm(&gq, stream));
    G gk(Tk, dh, T);
    gk.batches(B, H).A(p->ds, 1, Tkp, sp, sph).B(p->q, 1, E, sq, dh).C(p->dkv, 2 * E, skv, dh);
    MTTS_TRY(mtts_gemm(&gk, stream));
  }
  // packed projection gradients
  G gwq(E, E, B * T);
  gwq.A(p->dq, 1, E).B(p->query, 1, E).C(p->dw_in, E).f32_splitk();
  MTTS_TRY(mtts_gemm(&gwq, stream));
  G gwkv(2 * E, E, B * Tk);
  gwkv.A(p->dkv, 1, 2 * E).B(p->memory, 1, E).C(p->dw_in + (long long)E * E, E).f32_splitk();
  MTTS_TRY(mtts_gemm(&gwkv, stream));
  MTTS_TRY(colsum_bf16(p->dq, (long long)B * T, E, E, p->db_in, stream));
  MTTS_TRY(colsum_bf16(p->dkv, (long long)B * Tk, 2 * E, 2 * E, p->db_in + E, stream));
  if (p->dquery) {
    G g(B * T, E, E);
    g.A(p->dq, 0, E).B(p->w_in, 1, E).C(p->dquery, E);
    MTTS_TRY(mtts_gemm(&g, stream));
  }
  if (p->dmemory) {
    G g(B * Tk, E, 2 * E);
    g.A(p->dkv, 0, 2 * E).B(at(p->w_in, (long long)E * E), 1, E).C(p->dmemory, E);
    MTTS_TRY(mtts_gemm(&g, stream));
  }
  return MTTS_OK;
}
