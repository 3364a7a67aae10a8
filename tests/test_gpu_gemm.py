"""tcgen05 GEMM (csrc/gemm_sm100.cu, ``mtts_gemm``) against fp32 torch on the same bf16-rounded operands:
all four operand layouts, tails, batch / broadcast / (batch, head) views, the batch-reducing split-k weight
gradient, and every fused epilogue."""
import math

import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu
BF16_OUT = 6e-3      # one bf16 rounding of the result: 2^-9 relative to the element, here relative to the max
F32_OUT = 2e-5       # fp32 accumulation order only


def _rand(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(torch.bfloat16).cuda()


def _view(t, major):
    """(.., rows, k) logical view of t stored K-major (major 0) or MN-major (major 1)."""
    return t if major == 0 else t.transpose(-1, -2).contiguous().transpose(-1, -2)


@pytest.mark.parametrize("am,bm", [(0, 0), (0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize("m,n,k", [(300, 520, 200), (128, 256, 64), (64, 64, 32), (1000, 100, 4160), (130, 2048, 512)])
def test_gemm_layouts_and_tails(am, bm, m, n, k):
    from mamba_tts_project_b200.gemm import gemm
    # MN-major operands need their contiguous extent to be a multiple of 8 elements (16-byte TMA strides)
    m_ = m + (-m) % 8
    a = _view(_rand(m_, k + (-k) % 8, seed=1), am)[:m, :k]
    b = _view(_rand(n + (-n) % 8, k + (-k) % 8, seed=2), bm)[:n, :k]
    ref = a.float() @ b.float().t()
    out = gemm(a, b)
    assert out.dtype == torch.bfloat16 and rel_err(out, ref) < BF16_OUT
    out32 = gemm(a, b, out_dtype=torch.float32)
    assert rel_err(out32, ref) < F32_OUT


def test_gemm_bias_accumulate_and_strided_out():
    from mamba_tts_project_b200.gemm import gemm
    m, n, k = 260, 384, 192
    a, b = _rand(m, k, seed=3), _rand(n, k, seed=4)
    bn = torch.randn(n, device="cuda")
    bm_ = torch.randn(m, device="cuda")
    ref = a.float() @ b.float().t()
    assert rel_err(gemm(a, b, bias_n=bn), ref + bn) < BF16_OUT
    assert rel_err(gemm(a, b, bias_m=bm_, out_dtype=torch.float32), ref + bm_[:, None]) < F32_OUT
    big = torch.zeros(m, 2 * n, device="cuda")          # ldc > n: write into the right half only
    gemm(a, b, out=big[:, n:])
    assert rel_err(big[:, n:], ref) < F32_OUT and big[:, :n].abs().max() == 0
    acc32 = torch.ones(m, n, device="cuda")
    gemm(a, b, out=acc32, accumulate=True)
    assert rel_err(acc32, ref + 1) < F32_OUT
    acc16 = torch.ones(m, n, device="cuda", dtype=torch.bfloat16)
    gemm(a, b, out=acc16, accumulate=True)
    assert rel_err(acc16, ref + 1) < BF16_OUT


def test_gemm_batched_broadcast_and_head_views():
    from mamba_tts_project_b200.gemm import gemm
    B, H, T, dh, Tk = 3, 4, 200, 64, 136
    E = H * dh
    # in_proj-like: weight shared by the batch (stride 0), activations channel-major out
    w = _rand(2 * E, E, seed=5, scale=E ** -0.5)
    h = _rand(B, T, E, seed=6)
    xz = gemm(w.unsqueeze(0).expand(B, -1, -1), h)                     # (B, 2E, T)
    assert rel_err(xz, torch.einsum("ce,bte->bct", w.float(), h.float())) < BF16_OUT
    # out_proj-like: A is the transpose view of a channel-major tensor
    y = _rand(B, E, T, seed=7)
    wo = _rand(E, E, seed=8, scale=E ** -0.5)
    o = gemm(y.transpose(1, 2), wo.unsqueeze(0).expand(B, -1, -1))     # (B, T, E)
    assert rel_err(o, torch.einsum("bct,dc->btd", y.float(), wo.float())) < BF16_OUT
    # attention scores: (B, H, T, dh) views of (B, T, E) projections, two batch dimensions
    q, kk = _rand(B, T, E, seed=9), _rand(B, Tk, E, seed=10)
    qv = q.view(B, T, H, dh).transpose(1, 2)
    kv = kk.view(B, Tk, H, dh).transpose(1, 2)
    s = gemm(qv, kv, out_dtype=torch.float32)                          # (B, H, T, Tk)
    assert rel_err(s, qv.float() @ kv.float().transpose(-1, -2)) < F32_OUT
    # P V with V read N-major from the (B, Tk, E) tensor and O written straight into (B, T, E)
    p = _rand(B, H, T, Tk, seed=11, scale=0.1)
    vv = _rand(B, Tk, E, seed=12)
    o = torch.empty(B, T, E, device="cuda", dtype=torch.bfloat16)
    gemm(p, vv.view(B, Tk, H, dh).permute(0, 2, 3, 1), out=o.view(B, T, H, dh).transpose(1, 2))
    ref = (p.float() @ vv.view(B, Tk, H, dh).transpose(1, 2).float()).transpose(1, 2).reshape(B, T, E)
    assert rel_err(o, ref) < BF16_OUT


@pytest.mark.parametrize("split", [1, 5, -1])
def test_gemm_weight_gradient_reduces_over_batch(split):
    from mamba_tts_project_b200.gemm import gemm
    B, C, T, D = 4, 192, 520, 128
    dxz, h = _rand(B, C, T, seed=13), _rand(B, T, D, seed=14)
    # dW[c, d] = sum_b sum_t dxz[b, c, t] h[b, t, d]:  A K-major (k = t), B N-major
    dw = gemm(dxz, h.transpose(1, 2), out_dtype=torch.float32, reduce_batch=True, split_k=split)
    ref = torch.einsum("bct,btd->cd", dxz.float(), h.float())
    assert dw.shape == (C, D) and rel_err(dw, ref) < F32_OUT
    # token-major activations on both sides (FFN weight gradient): both operands MN-major, one long k
    dy, x = _rand(B * T, C, seed=15), _rand(B * T, D, seed=16)
    dw2 = gemm(dy.t(), x.t(), out_dtype=torch.float32, split_k=split)
    assert rel_err(dw2, dy.float().t() @ x.float()) < F32_OUT


def test_gemm_gelu_epilogues():
    from mamba_tts_project_b200.gemm import gemm
    m, n, k = 520, 768, 256
    x, w = _rand(m, k, seed=17), _rand(n, k, seed=18, scale=k ** -0.5)
    bias = torch.randn(n, device="cuda")
    pre = torch.empty(m, n, device="cuda", dtype=torch.bfloat16)
    act = gemm(x, w, bias_n=bias, epilogue="gelu", aux=pre)
    ref_pre = x.float() @ w.float().t() + bias
    assert rel_err(pre, ref_pre) < BF16_OUT
    assert rel_err(act, torch.nn.functional.gelu(ref_pre)) < BF16_OUT
    # backward: dpre = (dy @ W2) * gelu'(pre), W2 (d, n) read N-major
    d = 128
    dy, w2 = _rand(m, d, seed=19), _rand(d, n, seed=20, scale=d ** -0.5)
    dpre = gemm(dy, w2.t(), epilogue="gelu_bwd", aux=pre)
    pf = pre.float().requires_grad_()
    (g,) = torch.autograd.grad(torch.nn.functional.gelu(pf), pf, dy.float() @ w2.float())
    assert rel_err(dpre, g) < BF16_OUT


@pytest.mark.parametrize("tk", [256, 200, 72])
def test_gemm_softmax_epilogues(tk):
    from mamba_tts_project_b200.gemm import gemm
    B, H, T, dh = 2, 4, 300, 64
    E = H * dh
    q, kk = _rand(B, T, E, seed=21), _rand(B, tk, E, seed=22)
    mask = torch.rand(B, tk, generator=torch.Generator().manual_seed(1)) > 0.3
    mask[:, 0] = True
    qv = q.view(B, T, H, dh).transpose(1, 2)
    kv = kk.view(B, tk, H, dh).transpose(1, 2)
    scale = 1 / math.sqrt(dh)
    tkp = tk + (-tk) % 8
    P = torch.zeros(B, H, T, tkp, device="cuda", dtype=torch.bfloat16)
    gemm(qv, kv, out=P[..., :tk], epilogue="softmax", mask=mask.cuda(), scale=scale)
    s = (qv.float() @ kv.float().transpose(-1, -2)) * scale
    s = s.masked_fill(~mask.cuda()[:, None, None, :], float("-inf"))
    ref = torch.softmax(s, -1)
    assert rel_err(P[..., :tk], ref) < BF16_OUT
    assert tkp == tk or P[..., tk:].abs().max() == 0
    # softmax backward fused into dP = dO V^T
    do, vv = _rand(B, T, E, seed=23), _rand(B, tk, E, seed=24)
    dov = do.view(B, T, H, dh).transpose(1, 2)
    vvv = vv.view(B, tk, H, dh).transpose(1, 2)
    dS = torch.zeros_like(P)
    gemm(dov, vvv, out=dS[..., :tk], epilogue="dsoftmax", aux=P[..., :tk], scale=scale)
    Pf = P[..., :tk].float()
    dP = dov.float() @ vvv.float().transpose(-1, -2)
    ref_ds = scale * Pf * (dP - (Pf * dP).sum(-1, keepdim=True))
    assert rel_err(dS[..., :tk], ref_ds) < BF16_OUT


def test_gemm_persistent_many_tiles_and_legacy_entry():
    from mamba_tts_project_b200 import ops
    from mamba_tts_project_b200.gemm import gemm
    a, w = _rand(4096, 512, seed=25), _rand(2048, 512, seed=26, scale=512 ** -0.5)   # 256 tiles > 148 CTAs
    ref = a.float() @ w.float().t()
    assert rel_err(gemm(a, w), ref) < BF16_OUT
    bias = torch.randn(2048, device="cuda")
    out, pre = ops.gemm_bf16(a, w, bias, gelu=True, return_pre=True)
    assert rel_err(pre, ref + bias) < BF16_OUT and rel_err(out, torch.nn.functional.gelu(ref + bias)) < BF16_OUT
    assert rel_err(ops.gemm_bf16(a, w, bias), ref + bias) < BF16_OUT


def test_gemm_gelu_grad_aux_and_mul_aux_epilogues():
    """The forward can leave gelu'(pre) (same tanh as gelu(pre)); the backward then multiplies by it."""
    from mamba_tts_project_b200.gemm import gemm
    m, n, k, d = 392, 512, 192, 64
    x, w = _rand(m, k, seed=27), _rand(n, k, seed=28, scale=k ** -0.5)
    bias = torch.randn(n, device="cuda")
    gp = torch.empty(m, n, device="cuda", dtype=torch.bfloat16)
    act = gemm(x, w, bias_n=bias, epilogue="gelu", aux=gp, aux_gelu_grad=True)
    pre = (x.float() @ w.float().t() + bias).requires_grad_()
    y = torch.nn.functional.gelu(pre)
    (g_ref,) = torch.autograd.grad(y.sum(), pre)
    assert rel_err(act, y) < BF16_OUT
    assert rel_err(gp, g_ref) < BF16_OUT
    dy, w2 = _rand(m, d, seed=29), _rand(d, n, seed=30, scale=d ** -0.5)
    dpre = gemm(dy, w2.t(), epilogue="mul_aux", aux=gp)
    assert rel_err(dpre, (dy.float() @ w2.float()) * gp.float()) < BF16_OUT


def _ref_attention(query, memory, w_in, b_in, w_out, mask, H):
    E = query.shape[-1]
    q = query @ w_in[:E].t() + b_in[:E]
    k = memory @ w_in[E:2 * E].t() + b_in[E:2 * E]
    v = memory @ w_in[2 * E:].t() + b_in[2 * E:]
    B, T, _ = q.shape
    sh = lambda t: t.view(B, -1, H, E // H).transpose(1, 2)
    s = sh(q) @ sh(k).transpose(-1, -2) / math.sqrt(E // H)
    if mask is not None:
        s = s.masked_fill(~mask[:, None, None, :], float("-inf"))
    o = (torch.softmax(s, -1) @ sh(v)).transpose(1, 2).reshape(B, T, E)
    return o @ w_out.t()


@pytest.mark.parametrize("B,T,Tk,E,H,masked", [(2, 300, 72, 128, 4, True), (3, 128, 256, 256, 4, False),
                                                (1, 40, 13, 64, 8, True),
                                                # 64-wide heads: the backward core is the fused mtts_attn_core_bwd
                                                (2, 300, 72, 128, 2, True), (1, 40, 13, 64, 1, True),
                                                (2, 512, 200, 512, 8, True), (1, 1000, 129, 128, 2, False)])
def test_cross_attn_entry_points_forward_and_all_gradients(B, T, Tk, E, H, masked):
    """mtts_cross_attn_fwd / _bwd (one C-ABI call per direction) against fp32 nn.MultiheadAttention arithmetic."""
    from mamba_tts_project_b200 import dense
    g = torch.Generator().manual_seed(5)
    rnd = lambda *s, sc=1.0: (torch.randn(*s, generator=g) * sc).to(torch.bfloat16).float().cuda()
    query, memory = rnd(B, T, E).requires_grad_(), rnd(B, Tk, E).requires_grad_()
    w_in, w_out = rnd(3 * E, E, sc=E ** -0.5).requires_grad_(), rnd(E, E, sc=E ** -0.5).requires_grad_()
    b_in = (0.1 * torch.randn(3 * E, generator=g)).cuda().requires_grad_()
    mask = None
    if masked:
        mask = (torch.rand(B, Tk, generator=g) > 0.3).cuda()
        mask[:, 0] = True
    dout = rnd(B, T, E)
    ref = _ref_attention(query, memory, w_in, b_in, w_out, mask, H)
    g_ref = torch.autograd.grad(ref, [query, memory, w_in, b_in, w_out], dout)
    out = dense.cross_attention(query.detach().bfloat16().requires_grad_(), memory.detach().bfloat16().requires_grad_(),
                                w_in, b_in, w_out, mask, H)
    assert out.dtype == torch.bfloat16 and rel_err(out, ref) < 2e-2
    qb, mb = query.detach().bfloat16().requires_grad_(), memory.detach().bfloat16().requires_grad_()
    out = dense.cross_attention(qb, mb, w_in, b_in, w_out, mask, H)
    got = torch.autograd.grad(out, [qb, mb, w_in, b_in, w_out], dout.bfloat16())
    for name, a, b in zip(("dquery", "dmemory", "dw_in", "db_in", "dw_out"), got, g_ref):
        assert rel_err(a, b) < 3e-2, name


@pytest.mark.parametrize("T,D,Fd,bias", [(520, 128, 512, True), (96, 64, 128, False)])
def test_film_ffn_entry_points_forward_and_all_gradients(T, D, Fd, bias):
    """mtts_film_ffn_fwd / _bwd (one C-ABI call per direction) against fp32 Linear -> GELU -> Linear."""
    from mamba_tts_project_b200 import dense
    g = torch.Generator().manual_seed(6)
    rnd = lambda *s, sc=1.0: (torch.randn(*s, generator=g) * sc).to(torch.bfloat16).float().cuda()
    h = rnd(T, D).requires_grad_()
    w1, w2 = rnd(Fd, D, sc=D ** -0.5).requires_grad_(), rnd(D, Fd, sc=Fd ** -0.5).requires_grad_()
    b1 = (0.2 * torch.randn(Fd, generator=g)).cuda().requires_grad_() if bias else None
    df = rnd(T, D)
    ref = torch.nn.functional.gelu(torch.nn.functional.linear(h, w1, b1)) @ w2.t()
    ins = [h, w1, w2] + ([b1] if bias else [])
    g_ref = torch.autograd.grad(ref, ins, df)
    hb = h.detach().bfloat16().requires_grad_()
    f = dense.ffn(hb, w1, b1, w2)
    assert rel_err(f, ref) < 2e-2
    got = torch.autograd.grad(f, [hb, w1, w2] + ([b1] if bias else []), df.bfloat16())
    for name, a, b in zip(("dh", "dw1", "dw2", "db1"), got, g_ref):
        assert rel_err(a, b) < 3e-2, name


def test_film_ffn_forward_with_fused_layernorm_film():
    """The ln member: residual add + LayerNorm + FiLM feeding the FFN inside the same C-ABI call."""
    from mamba_tts_project_b200 import _lib
    from mamba_tts_project_b200._lib import ptr
    B, Tn, D, Fd = 2, 96, 128, 256
    g = torch.Generator().manual_seed(7)
    x = torch.randn(B * Tn, D, generator=g).cuda()
    delta = torch.randn(B * Tn, D, generator=g).bfloat16().cuda()
    lw, lb = (1 + 0.1 * torch.randn(D, generator=g)).cuda(), (0.1 * torch.randn(D, generator=g)).cuda()
    gam, bet = (1 + 0.2 * torch.randn(B, D, generator=g)).cuda(), (0.2 * torch.randn(B, D, generator=g)).cuda()
    w1, w2 = _rand(Fd, D, seed=31, scale=D ** -0.5), _rand(D, Fd, seed=32, scale=Fd ** -0.5)
    b1 = (0.1 * torch.randn(Fd, generator=g)).cuda()
    x_out = torch.empty_like(x)
    h = torch.empty(B * Tn, D, dtype=torch.bfloat16, device="cuda")
    act, gp = torch.empty(B * Tn, Fd, dtype=torch.bfloat16, device="cuda"), torch.empty(B * Tn, Fd, dtype=torch.bfloat16, device="cuda")
    f = torch.empty(B * Tn, D, dtype=torch.bfloat16, device="cuda")
    ln = _lib.AddLayerNormFwdParams(rows=B * Tn, dim=D, rows_per_batch=Tn, io_dtype=_lib.BF16, eps=1e-5, x=ptr(x),
                                    delta=ptr(delta), x_out=ptr(x_out), ln_weight=ptr(lw), ln_bias=ptr(lb),
                                    film_gamma=ptr(gam), film_beta=ptr(bet), out=ptr(h))
    p = _lib.FilmFfnParams(tokens=B * Tn, d_model=D, d_ff=Fd, ln=ln, h=ptr(h), w1=ptr(w1), b1=ptr(b1), w2=ptr(w2),
                           act=ptr(act), gprime=ptr(gp), f=ptr(f))
    _lib.call("mtts_film_ffn_fwd", p, launches=3)
    xo = x + delta.float()
    hr = torch.nn.functional.layer_norm(xo, (D,), lw, lb, 1e-5).view(B, Tn, D) * gam[:, None] + bet[:, None]
    assert rel_err(x_out, xo) < 1e-6
    assert rel_err(h, hr.view(-1, D)) < BF16_OUT
    ref = torch.nn.functional.gelu(h.float() @ w1.float().t() + b1) @ w2.float().t()
    assert rel_err(f, ref) < 2e-2


def test_fused_attention_fully_masked_rows_and_memory():
    """64-wide heads (fused mtts_attn_core_fwd / _bwd): a batch element whose keys are all masked gives zero attention
    output and finite (zero) gradients through it -- the reference's nn.MultiheadAttention gives NaN there;
    and the forward keeps no (B, H, T, T_kv) tensor for the backward."""
    from mamba_tts_project_b200 import dense
    B, T, Tk, E, H = 2, 200, 96, 128, 2
    g = torch.Generator().manual_seed(9)
    rnd = lambda *s, sc=1.0: (torch.randn(*s, generator=g) * sc).to(torch.bfloat16).cuda()
    query, memory = rnd(B, T, E).requires_grad_(), rnd(B, Tk, E).requires_grad_()
    w_in, w_out = rnd(3 * E, E, sc=E ** -0.5).float().requires_grad_(), rnd(E, E, sc=E ** -0.5).float().requires_grad_()
    b_in = torch.zeros(3 * E, device="cuda", requires_grad=True)
    mask = torch.ones(B, Tk, dtype=torch.bool, device="cuda")
    mask[1] = False
    out = dense.cross_attention(query, memory, w_in, b_in, w_out, mask, H)
    assert torch.count_nonzero(out[1]) == 0 and torch.isfinite(out).all()
    saved = [t for t in out.grad_fn.saved_tensors if t is not None]
    assert max(t.numel() for t in saved) <= max(B * T * E, B * Tk * 2 * E, 3 * E * E)   # no probability tensor
    gq, gm = torch.autograd.grad(out, [query, memory], torch.ones_like(out))
    assert torch.isfinite(gq).all() and torch.isfinite(gm).all()
    assert torch.count_nonzero(gq[1]) == 0
