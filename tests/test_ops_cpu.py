"""CPU implementations of the op layer against autograd / F.* references (they are the oracle the GPU kernels are
tested against, so they are tested themselves first)."""

import pytest
import torch
import torch.nn.functional as F

from tiny_deepspeed_b200 import ops
import tiny_deepspeed_b200.nn as tnn


def test_linear_ops_3d_and_bias_grad_fix():
    torch.manual_seed(0)
    x = torch.randn(2, 5, 8, requires_grad=True)
    w = torch.randn(6, 8, requires_grad=True)
    b = torch.randn(6, requires_grad=True)
    y = F.linear(x, w, b)
    dy = torch.randn_like(y)
    y.backward(dy)
    torch.testing.assert_close(ops.linear_forward(x, w, b), y)
    torch.testing.assert_close(ops.linear_input_grad(dy, w), x.grad)
    torch.testing.assert_close(ops.linear_weight_grad(dy, x, w), w.grad)
    torch.testing.assert_close(ops.linear_bias_grad(dy, b), b.grad)      # reference raises here for 3-D input (Q9)


def test_gemm_layout_flags_and_epilogues():
    torch.manual_seed(1)
    A, B = torch.randn(3, 7, 5), torch.randn(3, 4, 5)
    ref = A @ B.transpose(-1, -2)
    torch.testing.assert_close(ops.gemm(A, B), ref)
    torch.testing.assert_close(ops.gemm(A.transpose(-1, -2).contiguous(), B, a_mn=True), ref)
    torch.testing.assert_close(ops.gemm(A, B.transpose(-1, -2).contiguous(), b_mn=True), ref)
    x, w = torch.randn(7, 5), torch.randn(4, 5)
    pre = torch.empty(7, 4)
    act = ops.gemm(x, w, aux=pre, epi=ops.EPI_GELU_SAVE)
    torch.testing.assert_close(pre, x @ w.t())
    torch.testing.assert_close(act, F.gelu(x @ w.t(), approximate="tanh"))
    p = (x @ w.t()).requires_grad_()
    F.gelu(p, approximate="tanh").backward(torch.ones_like(p))
    torch.testing.assert_close(ops.gemm(torch.ones(7, 3), torch.ones(4, 3) / 3, aux=p.detach(), epi=ops.EPI_GELU_BWD), p.grad)
    out = torch.ones(7, 4)
    ops.gemm(x, w, out=out, accumulate=True)
    torch.testing.assert_close(out, 1 + x @ w.t())


@pytest.mark.parametrize("N", [16, 100])
def test_layernorm(N):
    torch.manual_seed(2)
    x = torch.randn(3, 7, N, requires_grad=True)
    w = torch.randn(N, requires_grad=True)
    b = torch.randn(N, requires_grad=True)
    y = F.layer_norm(x, (N,), w, b, 1e-5)
    dy = torch.randn_like(y)
    y.backward(dy)
    y2, mean, rstd = ops.layernorm_fwd(x, w, b, 1e-5)
    torch.testing.assert_close(y2, y)
    res = torch.randn_like(x)
    dx, dw, db = ops.layernorm_bwd(dy, x, w, mean, rstd, add_to_dx=res)
    torch.testing.assert_close(dx, x.grad + res, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(dw, w.grad, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(db, b.grad, rtol=1e-4, atol=1e-5)
    dx2, parts = ops.layernorm_dx(dy, x, w, b, mean, rstd)
    dw2, db2 = ops.layernorm_dwdb(w, b, parts)
    torch.testing.assert_close(dw2, w.grad, rtol=1e-4, atol=1e-5)


def test_embedding():
    torch.manual_seed(3)
    w = torch.randn(11, 6, requires_grad=True)
    idx = torch.randint(0, 11, (2, 5))
    pos = torch.randn(5, 6)
    y = F.embedding(idx, w) + pos
    dy = torch.randn_like(y)
    y.backward(dy)
    torch.testing.assert_close(ops.embedding_forward(idx, w, add=pos), y)
    torch.testing.assert_close(ops.embedding_weight_grad(idx, dy, w), w.grad)


def test_attention_matches_sdpa():
    torch.manual_seed(4)
    B, T, nh, hs = 2, 16, 3, 8
    C = nh * hs
    qkv = torch.randn(B, T, 3 * C, requires_grad=True)
    q, k, v = (t.view(B, T, nh, hs).transpose(1, 2) for t in qkv.split(C, dim=2))
    ref = F.scaled_dot_product_attention(q, k, v, is_causal=True).transpose(1, 2).reshape(B, T, C)
    dy = torch.randn_like(ref)
    ref.backward(dy)
    y, P = ops.causal_attention_forward(qkv.detach(), nh)
    torch.testing.assert_close(y, ref, rtol=1e-4, atol=1e-5)
    dqkv = ops.causal_attention_backward(dy, qkv.detach(), P, nh)
    torch.testing.assert_close(dqkv, qkv.grad, rtol=1e-4, atol=1e-5)
    # model-file attention helpers (API parity) agree as well
    from tiny_deepspeed_b200.models.gpt2 import standard_attention, flash_attention
    q4, k4, v4 = (t.view(B, T, nh, hs) for t in qkv.detach().split(C, dim=2))
    torch.testing.assert_close(standard_attention(q4, k4, v4).reshape(B, T, C), ref, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(flash_attention(q4, k4, v4).reshape(B, T, C), ref, rtol=1e-4, atol=1e-5)


def test_cross_entropy():
    torch.manual_seed(5)
    l = torch.randn(12, 33, requires_grad=True)
    t = torch.randint(0, 33, (12,))
    ref = F.cross_entropy(l, t)
    ref.backward()
    loss, lse = ops.cross_entropy_forward(l.detach(), t)
    torch.testing.assert_close(loss, ref)
    torch.testing.assert_close(ops.cross_entropy_backward(torch.tensor(1.0), l.detach(), t, lse), l.grad)


def test_modules_autograd_end_to_end():
    """Our layers (policy = local) give the same grads as torch.nn layers with the same weights."""
    torch.manual_seed(6)
    lin, ln, emb = tnn.Linear(8, 12, bias=True), tnn.LayerNorm(8), tnn.Embedding(20, 8)
    rl, rn, re = torch.nn.Linear(8, 12), torch.nn.LayerNorm(8), torch.nn.Embedding(20, 8)
    for a, b in ((lin, rl), (ln, rn), (emb, re)):
        b.load_state_dict(a.state_dict())
    with torch.no_grad():
        ln.weight.uniform_(0.5, 1.5); ln.bias.uniform_(-1, 1)
        rn.load_state_dict(ln.state_dict())
    idx = torch.randint(0, 20, (3, 5))
    h, res = ln(emb(idx), with_residual=True)
    out = lin(h) .sum() + (res * 2).sum()
    out.backward()
    e = re(idx)
    ref = rl(rn(e)).sum() + (e * 2).sum()
    ref.backward()
    for a, b in ((lin, rl), (ln, rn), (emb, re)):
        for (n, p), (_, q) in zip(a.named_parameters(), b.named_parameters()):
            torch.testing.assert_close(p.grad, q.grad, rtol=1e-4, atol=1e-5, msg=n)
