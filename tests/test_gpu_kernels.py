"""sm_100a kernels vs a plain PyTorch fp32 reference of the same op."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from tiny_deepspeed_b200 import ops  # noqa: E402


def _dev():
    return torch.device("cuda", 0)


def _rand(*shape, scale=1.0):
    return (torch.randn(*shape, device=_dev(), dtype=torch.float32) * scale).to(torch.bfloat16)


def _close(got, ref, rtol=2e-2, atol=2e-2):
    torch.testing.assert_close(got.float(), ref.float(), rtol=rtol, atol=atol)


def _ref_gemm(a, b, a_mn, b_mn):
    A = a.float().transpose(-1, -2) if a_mn else a.float()
    B = b.float().transpose(-1, -2) if b_mn else b.float()
    return A @ B.transpose(-1, -2)


def test_extension_is_loaded_from_tree():
    import os
    mod = ops.ext()
    assert mod.__file__.endswith("tiny_deepspeed_b200/_C.so") and os.path.exists(mod.__file__)


GEMM_SHAPES = [(128, 128, 64), (128, 64, 64), (256, 256, 128), (1024, 768, 768), (1024, 2304, 768), (1024, 768, 3072),
               (1024, 3072, 768), (200, 136, 72), (1024, 1024, 64), (520, 50304, 256), (1000, 264, 1032)]


@pytest.mark.parametrize("M,N,K", GEMM_SHAPES)
@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, True), (True, False)])
def test_gemm_layouts(M, N, K, a_mn, b_mn):
    torch.manual_seed(M + N + K)
    a = _rand(K, M) if a_mn else _rand(M, K)
    b = _rand(K, N) if b_mn else _rand(N, K)
    ref = _ref_gemm(a, b, a_mn, b_mn)
    got = ops.gemm(a, b, a_mn=a_mn, b_mn=b_mn)
    tol = 0.02 * math.sqrt(K)
    torch.testing.assert_close(got.float(), ref, rtol=2e-2, atol=tol)
    rel = (got.float() - ref).norm() / ref.norm()
    assert rel < 5e-3, rel


@pytest.mark.parametrize("cfg", [0, 1, 2, 3])
def test_gemm_tile_configs_and_fp32_out(cfg):
    a, b = _rand(384, 512), _rand(640, 512)
    ref = a.float() @ b.float().t()
    got = ops.gemm(a, b, config=cfg, out_dtype=torch.float32)
    assert got.dtype == torch.float32
    torch.testing.assert_close(got, ref, rtol=1e-3, atol=1e-2)
    out = torch.ones(384, 640, device=_dev(), dtype=torch.float32)
    ops.gemm(a, b, out=out, accumulate=True, alpha=0.5, config=cfg)
    torch.testing.assert_close(out, 1 + 0.5 * ref, rtol=1e-3, atol=1e-2)


def test_gemm_epilogues():
    x, w, bias = _rand(1024, 768), _rand(3072, 768, scale=0.05), _rand(3072)
    lin = x.float() @ w.float().t() + bias.float()
    pre = torch.empty(1024, 3072, device=_dev(), dtype=torch.bfloat16)
    act = ops.gemm(x, w, bias=bias, aux=pre, epi=ops.EPI_GELU_SAVE)
    _close(pre, lin)
    _close(act, F.gelu(pre.float(), approximate="tanh"))
    dy = _rand(1024, 768)
    w2 = _rand(768, 3072, scale=0.05)
    dact = dy.float() @ w2.float()
    p = pre.float().requires_grad_()
    F.gelu(p, approximate="tanh").backward(dact)
    got = ops.gemm(dy, w2, b_mn=True, aux=pre, epi=ops.EPI_GELU_BWD)
    _close(got, p.grad, atol=5e-2)
    res = _rand(1024, 3072)
    got = ops.gemm(x, w, bias=bias, aux=res, epi=ops.EPI_RESIDUAL)
    _close(got, lin + res.float(), atol=5e-2)
    acc = _rand(1024, 3072)
    want = acc.float() + (x.float() @ w.float().t())
    ops.gemm(x, w, out=acc, accumulate=True)
    _close(acc, want, atol=5e-2)


def test_gemm_batched_strided_heads_and_tri():
    B, T, nh, hs = 2, 256, 4, 64
    C = nh * hs
    qkv = _rand(B, T, 3 * C, scale=0.5)
    q, k, v = (t.view(B, T, nh, hs).transpose(1, 2) for t in qkv.split(C, dim=2))
    S = ops.gemm(q, k)
    _close(S, q.float() @ k.float().transpose(-1, -2), atol=0.1)
    St = torch.zeros_like(S)
    ops.gemm(q, k, out=St, tri=1)
    mask = torch.ones(T, T, device=_dev(), dtype=torch.bool).tril()
    _close(St.float() * mask, (q.float() @ k.float().transpose(-1, -2)) * mask, atol=0.1)
    P = torch.softmax((q.float() @ k.float().transpose(-1, -2)).masked_fill(~mask, float("-inf")) / 8.0, -1).to(torch.bfloat16)
    y = torch.empty(B, T, C, device=_dev(), dtype=torch.bfloat16)
    ops.gemm(P, v, b_mn=True, out=y.view(B, T, nh, hs).transpose(1, 2), tri=2)
    _close(y.view(B, T, nh, hs).transpose(1, 2), P.float() @ v.float())
    dv = ops.gemm(P, q, a_mn=True, b_mn=True, tri=3)
    _close(dv, P.float().transpose(-1, -2) @ q.float(), atol=5e-2)


@pytest.mark.parametrize("N", [768, 1024, 1280, 1600, 200])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_layernorm(N, dtype):
    torch.manual_seed(N)
    x = torch.randn(1024, N, device=_dev()).to(dtype)
    w = (torch.rand(N, device=_dev()) + 0.5).to(dtype)
    b = torch.randn(N, device=_dev()).to(dtype)
    dy = torch.randn(1024, N, device=_dev()).to(dtype)
    res = torch.randn(1024, N, device=_dev()).to(dtype)
    xf = x.float().requires_grad_(); wf = w.float().requires_grad_(); bf = b.float().requires_grad_()
    yr = F.layer_norm(xf, (N,), wf, bf, 1e-5)
    yr.backward(dy.float())
    y, mean, rstd = ops.layernorm_fwd(x, w, b, 1e-5)
    tol = dict(rtol=2e-2, atol=3e-2) if dtype == torch.bfloat16 else dict(rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(y.float(), yr, **tol)
    torch.testing.assert_close(mean, x.float().mean(-1), rtol=1e-4, atol=1e-4)
    dx, dw, db = ops.layernorm_bwd(dy, x, w, mean, rstd, add_to_dx=res)
    torch.testing.assert_close(dx.float(), xf.grad + res.float(), **tol)
    big = dict(rtol=2e-2, atol=0.5) if dtype == torch.bfloat16 else dict(rtol=1e-3, atol=1e-3)
    torch.testing.assert_close(dw.float(), wf.grad, **big)
    torch.testing.assert_close(db.float(), bf.grad, **big)
    dw2, db2 = dw.clone(), db.clone()
    ops.layernorm_bwd(dy, x, w, mean, rstd, dw_out=dw2, db_out=db2, accumulate=True)
    torch.testing.assert_close(dw2.float(), 2 * wf.grad, **big)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_embedding(dtype):
    V, D = 50304, 768
    w = torch.randn(V, D, device=_dev()).to(dtype)
    idx = torch.randint(0, V, (2, 1024), device=_dev())
    idx[0, :8] = 5                                     # repeated rows exercise the atomics
    pos = torch.randn(1024, D, device=_dev()).to(dtype)
    out = ops.embedding_forward(idx, w, add=pos)
    torch.testing.assert_close(out.float(), F.embedding(idx, w).float() + pos.float(), rtol=1e-2, atol=2e-2)
    dy = torch.randn(2, 1024, D, device=_dev()).to(dtype)
    g = ops.embedding_weight_grad(idx, dy, w)
    ref = torch.zeros(V, D, device=_dev()).index_add_(0, idx.view(-1), dy.view(-1, D).float())
    torch.testing.assert_close(g.float(), ref, rtol=2e-2, atol=6e-2)


def test_causal_softmax_and_attention():
    B, T, nh, hs = 1, 1024, 12, 64
    C = nh * hs
    torch.manual_seed(0)
    qkv = _rand(B, T, 3 * C, scale=0.7)
    qf = qkv.float().requires_grad_()
    q, k, v = (t.view(B, T, nh, hs).transpose(1, 2) for t in qf.split(C, dim=2))
    ref = F.scaled_dot_product_attention(q, k, v, is_causal=True).transpose(1, 2).reshape(B, T, C)
    dy = _rand(B, T, C)
    ref.backward(dy.float())
    y, P = ops.causal_attention_forward(qkv, nh)
    _close(y, ref, atol=3e-2)
    dqkv = ops.causal_attention_backward(dy, qkv, P, nh, y=y)
    rel = (dqkv.float() - qf.grad).norm() / qf.grad.norm()
    assert rel < 2e-2, rel


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_cross_entropy(dtype):
    M, V = 1024, 50304
    l = (torch.randn(M, V, device=_dev()) * 2).to(dtype)
    t = torch.randint(0, V, (M,), device=_dev())
    lf = l.float().requires_grad_()
    ref = F.cross_entropy(lf, t)
    ref.backward()
    loss, lse = ops.cross_entropy_forward(l, t)
    torch.testing.assert_close(loss, ref, rtol=1e-4, atol=1e-4)
    g = ops.cross_entropy_backward(torch.tensor(1.0, device=_dev()), l, t, lse)
    torch.testing.assert_close(g.float(), lf.grad, rtol=2e-2, atol=1e-6 if dtype == torch.float32 else 1e-5)


def test_gelu_and_colsum():
    x = _rand(1024, 3072)
    dy = _rand(1024, 3072)
    _close(ops.gelu_forward(x), F.gelu(x.float(), approximate="tanh"))
    xf = x.float().requires_grad_(); F.gelu(xf, approximate="tanh").backward(dy.float())
    _close(ops.gelu_backward(dy, x), xf.grad)
    _close(ops.linear_bias_grad(dy.view(2, 512, 3072)), dy.float().sum(0), atol=0.5)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("decoupled", [False, True])
def test_fused_adam_multi_tensor(dtype, decoupled):
    import tiny_deepspeed_b200 as tds
    torch.manual_seed(0)
    shapes = [(768, 768), (768,), (50304, 64), (3, 5), (1027,)]
    ps = [torch.nn.Parameter(torch.randn(s, device=_dev()).to(dtype)) for s in shapes]
    ref = [torch.nn.Parameter(p.detach().float().clone()) for p in ps]
    opt = tds.AdamW([(f"p{i}", p) for i, p in enumerate(ps)], lr=1e-2, weight_decay=0.1, decoupled=decoupled)
    ropt = (torch.optim.AdamW if decoupled else torch.optim.Adam)(ref, lr=1e-2, weight_decay=0.1)
    for _ in range(4):
        for p, r in zip(ps, ref):
            g = torch.randn(p.shape, device=_dev())
            p.grad = g.to(dtype)
            r.grad = p.grad.float()
        opt.step(); ropt.step()
    for i, (p, r) in enumerate(zip(ps, ref)):
        w = opt.state[f"p{i}"]["master"] if dtype == torch.bfloat16 else p
        torch.testing.assert_close(w.float(), r, rtol=1e-4, atol=1e-5)
        assert p.grad is None


def test_fused_sgd_momentum():
    import tiny_deepspeed_b200 as tds
    p = torch.nn.Parameter(torch.randn(1000, 33, device=_dev()))
    r = torch.nn.Parameter(p.detach().clone())
    opt = tds.SGD([("p", p)], lr=0.05, momentum=0.9, weight_decay=0.01, nesterov=True)
    ropt = torch.optim.SGD([r], lr=0.05, momentum=0.9, weight_decay=0.01, nesterov=True)
    for _ in range(3):
        g = torch.randn_like(p)
        p.grad, r.grad = g.clone(), g.clone()
        opt.step(); ropt.step()
    torch.testing.assert_close(p, r, rtol=1e-5, atol=1e-6)


def test_splitk_lm_head_input_grad():
    dy = _rand(1024, 50304, scale=0.1)
    w = _rand(50304, 768, scale=0.1)
    assert ops._splitk_factor(1024, 768, 50304) > 1
    got = ops.linear_input_grad(dy, w)
    ref = dy.float() @ w.float()
    rel = (got.float() - ref).norm() / ref.norm()
    assert rel < 5e-3, rel


@pytest.mark.parametrize("cluster", [1, 2, 4, 8])
@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, True)])
@pytest.mark.parametrize("cfg", [0, 1, 2, 3])
def test_gemm_cluster_multicast(cluster, a_mn, b_mn, cfg):
    """B tile fetched once per cluster of M-tiles and TMA-multicast into all of them."""
    M, N, K = 1024, 1536, 832
    torch.manual_seed(cluster)
    a = _rand(K, M) if a_mn else _rand(M, K)
    b = _rand(K, N) if b_mn else _rand(N, K)
    ref = _ref_gemm(a, b, a_mn, b_mn)
    got = ops.gemm(a, b, a_mn=a_mn, b_mn=b_mn, config=cfg, cluster=cluster)
    rel = (got.float() - ref).norm() / ref.norm()
    assert rel < 5e-3, rel


@pytest.mark.parametrize("B,T,nh", [(1, 128, 2), (2, 384, 3), (1, 1024, 12)])
def test_flash_attention_kernels(B, T, nh):
    """Fused tcgen05 flash attention (fwd + bwd) vs fp32 SDPA autograd, and vs our materialised-score path."""
    C = nh * 64
    torch.manual_seed(T)
    qkv = _rand(B, T, 3 * C, scale=0.7)
    dy = _rand(B, T, C)
    qf = qkv.float().requires_grad_()
    q, k, v = (t.view(B, T, nh, 64).transpose(1, 2) for t in qf.split(C, dim=2))
    ref = F.scaled_dot_product_attention(q, k, v, is_causal=True).transpose(1, 2).reshape(B, T, C)
    ref.backward(dy.float())
    assert ops.flash_enabled() and ops.ext().flash_supported(T, 64)
    y, lse = ops.causal_attention_forward(qkv, nh)
    assert lse.dtype == torch.float32 and lse.shape == (B, nh, T)
    dqkv = ops.causal_attention_backward(dy, qkv, lse, nh, y=y)
    assert (y.float() - ref).norm() / ref.norm() < 5e-3
    assert (dqkv.float() - qf.grad).norm() / qf.grad.norm() < 8e-3
    ops.set_flash(False)
    try:
        y2, P = ops.causal_attention_forward(qkv, nh)
        d2 = ops.causal_attention_backward(dy, qkv, P, nh, y=y2)
    finally:
        ops.set_flash(True)
    assert P.dtype == torch.bfloat16
    assert (y.float() - y2.float()).norm() / ref.norm() < 8e-3
    assert (dqkv.float() - d2.float()).norm() / qf.grad.norm() < 2e-2


# ---- fp32 operands through kind::tf32 (fp32 models; reference trains fp32, example/ddp/train.py:22) ---------------------
def _rel(got, ref):
    return ((got.float() - ref.float()).norm() / ref.float().norm()).item()


@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, True), (True, False)])
@pytest.mark.parametrize("M,N,K", [(256, 384, 512), (1024, 768, 3072), (200, 136, 104), (128, 64, 32)])
def test_gemm_tf32_layouts(M, N, K, a_mn, b_mn):
    torch.manual_seed(M + N + K)
    a = torch.randn(K, M, device=_dev()) if a_mn else torch.randn(M, K, device=_dev())
    b = torch.randn(K, N, device=_dev()) if b_mn else torch.randn(N, K, device=_dev())
    A = a.t() if a_mn else a
    Bm = b.t() if b_mn else b
    ref = A.double() @ Bm.double().t()
    got = ops.gemm(a, b, a_mn=a_mn, b_mn=b_mn)
    assert got.dtype == torch.float32
    assert _rel(got, ref) < 1e-3, _rel(got, ref)      # TF32: 10-bit mantissa, fp32 accumulate


@pytest.mark.parametrize("cfg", [0, 1, 2, 3])
def test_gemm_tf32_configs_epilogues(cfg):
    torch.manual_seed(cfg)
    x, w, bias = torch.randn(512, 768, device=_dev()), torch.randn(1536, 768, device=_dev()) * 0.05, torch.randn(1536, device=_dev())
    lin = (x.double() @ w.double().t() + bias.double()).float()
    pre = torch.empty(512, 1536, device=_dev())
    act = ops.gemm(x, w, bias=bias, aux=pre, epi=ops.EPI_GELU_SAVE, config=cfg)
    assert _rel(pre, lin) < 1e-3
    assert _rel(act, F.gelu(lin, approximate="tanh")) < 2e-3
    res = torch.randn(512, 1536, device=_dev())
    got = ops.gemm(x, w, bias=bias, aux=res, epi=ops.EPI_RESIDUAL, config=cfg)
    assert _rel(got, lin + res) < 1e-3
    dy = torch.randn(512, 768, device=_dev())
    w2 = torch.randn(768, 1536, device=_dev()) * 0.05
    p = pre.clone().requires_grad_()
    F.gelu(p, approximate="tanh").backward(dy @ w2)
    got = ops.gemm(dy, w2, b_mn=True, aux=pre, epi=ops.EPI_GELU_BWD, config=cfg)
    assert _rel(got, p.grad) < 2e-3
    acc = torch.ones(512, 1536, device=_dev())
    ops.gemm(x, w, out=acc, accumulate=True, alpha=0.5, config=cfg)
    assert _rel(acc, 1 + 0.5 * (lin - bias)) < 1e-3
    # ragged N / unaligned bias path
    wr = torch.randn(100, 768, device=_dev()) * 0.05
    br = torch.randn(100, device=_dev())
    got = ops.gemm(x, wr, bias=br, config=cfg)
    assert _rel(got, x @ wr.t() + br) < 1e-3


def test_cast_kernel():
    x = torch.randn(1000, 37, device=_dev())
    y = ops.cast(x, torch.bfloat16)
    assert y.dtype == torch.bfloat16 and torch.equal(y, x.to(torch.bfloat16))
    z = ops.cast(y, torch.float32)
    assert z.dtype == torch.float32 and torch.equal(z, y.float())


def test_attention_fp32_model_dtype():
    B, T, nh, hs = 2, 256, 4, 64
    C = nh * hs
    qkv = (torch.randn(B, T, 3 * C, device=_dev()) * 0.5).requires_grad_()
    q, k, v = (t.view(B, T, nh, hs).transpose(1, 2) for t in qkv.split(C, dim=2))
    yr = F.scaled_dot_product_attention(q, k, v, is_causal=True).transpose(1, 2).reshape(B, T, C)
    dy = torch.randn_like(yr)
    (gr,) = torch.autograd.grad(yr, qkv, dy)
    y, P = ops.causal_attention_forward(qkv.detach(), nh)
    assert y.dtype == torch.float32
    assert _rel(y, yr) < 1e-2
    g = ops.causal_attention_backward(dy, qkv.detach(), P, nh, y=y)
    assert g.dtype == torch.float32
    assert _rel(g, gr) < 2e-2


@pytest.mark.parametrize("cfg", [0, 1, 2, 3])
def test_gemm_reduce_out_epilogue(cfg):
    """Experimental fused reduce epilogue: the fp32 product is ADDED into the destination by the TMA (the same op targets a
    peer's gradient shard in the multi-GPU reduce-scatter path); two launches accumulate, ragged edges are clipped."""
    for (M, N, K) in [(768, 3072, 1024), (200, 136, 256)]:
        dy, x = _rand(K, M), _rand(K, N)                       # dW = dY^T X : both operands MN-major
        ref = dy.float().t() @ x.float()
        out = torch.full((M, N), 0.5, device=_dev(), dtype=torch.float32)
        ops.gemm(dy, x, a_mn=True, b_mn=True, out=out, reduce_out=True, config=cfg)
        ops.gemm(dy, x, a_mn=True, b_mn=True, out=out, reduce_out=True, config=cfg, alpha=0.5)
        torch.testing.assert_close(out, 0.5 + 1.5 * ref, rtol=1e-3, atol=2e-2)


def test_gemm_cta_pair_kernel():
    """csrc/gemm2_sm100.cu (tcgen05 cta_group::2, 256 x BN tile over an SM pair): validated on hardware in round 2
    (profiles/r2_gemm2_check.log).  All three operand-major combinations, bias epilogue, both tile widths, vs an fp32 product."""
    ext = ops.ext()
    try:
        ext.set_gemm_pair(1)
        for (M, N, K) in [(1024, 2304, 768), (1024, 768, 3072), (512, 256, 128), (1024, 4096, 768)]:
            for a_mn, b_mn in [(False, False), (False, True), (True, True)]:
                a = _rand(K, M) if a_mn else _rand(M, K)
                b = _rand(K, N) if b_mn else _rand(N, K)
                bias = _rand(N)
                ref = _ref_gemm(a, b, a_mn, b_mn) + bias.float()
                got = ops.gemm(a, b, a_mn=a_mn, b_mn=b_mn, bias=bias)
                rel = ((got.float() - ref).norm() / ref.norm()).item()
                assert rel < 5e-3, (M, N, K, a_mn, b_mn, rel)
    finally:
        ext.set_gemm_pair(0)


@pytest.mark.parametrize("N", [768, 1600])
def test_layernorm_backward_variants_and_tuner(N):
    """Single-launch (L2 reductions + last-CTA finish) and two-kernel LayerNorm backward agree; the RuntimeAutoTuner picks
    between them per shape (reference threads its tuner through the LayerNorm ops, ops/layernorm.py:82-127)."""
    from tiny_deepspeed_b200.autotuner import RuntimeAutoTuner
    torch.manual_seed(N)
    x, dy = _rand(1024, N), _rand(1024, N)
    w = (torch.rand(N, device=_dev()) + 0.5).bfloat16()
    b = _rand(N)
    _, mean, rstd = ops.layernorm_fwd(x, w, b, 1e-5)
    outs = []
    for variant in (0, 1, 1):                      # the single-launch form twice: its accumulators must come back clean
        dw, db = torch.empty_like(w), torch.empty_like(w)
        dx = ops.ext().layernorm_bwd(dy, x, w, mean, rstd, dw, db, False, None, variant)
        outs.append((dx, dw, db))
    for dx, dw, db in outs[1:]:
        assert torch.equal(dx, outs[0][0])
        torch.testing.assert_close(dw.float(), outs[0][1].float(), rtol=2e-2, atol=0.25)
        torch.testing.assert_close(db.float(), outs[0][2].float(), rtol=2e-2, atol=0.25)
    tuner = RuntimeAutoTuner(enable=True, warmup_iterations=2, measure_iterations=5)
    dx, dw, db = ops.layernorm_bwd(dy, x, w, mean, rstd, runtime_tuner=tuner)
    assert torch.equal(dx, outs[0][0])
    assert any(k[0] == "layernorm_bwd" for k in tuner.cache)


def test_background_adam_kernel_matches_flooding_kernel():
    """adamw_multi_bg_kernel (fixed one-CTA-per-SM grid for optimizer-in-backward) == adamw_multi_kernel, bit for bit."""
    torch.manual_seed(3)
    shapes = [(768, 768), (50304, 64), (768,), (3, 5)]

    def run(bg):
        torch.manual_seed(4)
        ps = [_rand(*s) for s in shapes]
        gs = [_rand(*s, scale=0.1) for s in shapes]
        ms = [torch.zeros(s, device=_dev()) for s in shapes]
        vs = [torch.zeros(s, device=_dev()) for s in shapes]
        masters = [p.float().clone() for p in ps]
        step = torch.zeros(1, dtype=torch.int32, device=_dev())
        for _ in range(3):
            ops.ext().step_increment(step)
            ops.adamw_update(ps, gs, ms, vs, masters, lr=1e-2, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.1, step=1,
                             step_dev=step, background_ctas=bg)
        return ps, ms, vs, masters

    a, b = run(0), run(148)
    for ta, tb in zip(a, b):
        for u, v in zip(ta, tb):
            assert torch.equal(u, v)
