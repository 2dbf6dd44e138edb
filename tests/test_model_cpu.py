"""GPT-2 built from our layers vs an independent plain-PyTorch GPT-2 (same weights): loss + all gradients."""

import torch
import torch.nn as nn
import torch.nn.functional as F

import tiny_deepspeed_b200 as tds
from tiny_deepspeed_b200.models.gpt2 import GPT2Model, gpt2_config


class _RefBlock(nn.Module):
    def __init__(s, C, nh):
        super().__init__()
        s.ln_1, s.ln_2 = nn.LayerNorm(C), nn.LayerNorm(C)
        s.c_attn, s.c_proj = nn.Linear(C, 3 * C, bias=False), nn.Linear(C, C, bias=False)
        s.c_fc, s.c_proj2 = nn.Linear(C, 4 * C, bias=False), nn.Linear(4 * C, C, bias=False)
        s.nh = nh

    def forward(s, x):
        B, T, C = x.shape
        q, k, v = s.c_attn(s.ln_1(x)).split(C, dim=2)
        q, k, v = (t.view(B, T, s.nh, C // s.nh).transpose(1, 2) for t in (q, k, v))
        y = F.scaled_dot_product_attention(q, k, v, is_causal=True).transpose(1, 2).reshape(B, T, C)
        x = x + s.c_proj(y)
        return x + s.c_proj2(F.gelu(s.c_fc(s.ln_2(x)), approximate="tanh"))


def _ref_forward(model: GPT2Model, idx, tgt):
    cfg = model.config
    sd = {k: v.detach().clone().requires_grad_() for k, v in model.state_dict().items()}
    x = F.embedding(idx, sd["transformer.wte.weight"]) + sd["transformer.wpe.weight"][: idx.shape[1]]
    for i in range(cfg.n_layer):
        p = f"transformer.h.{i}."
        blk = _RefBlock(cfg.n_embd, cfg.n_head)
        def lin(w):  # functional linears on the shared leaf tensors so grads land in `sd`
            return lambda t: F.linear(t, w)
        B, T, C = x.shape
        h = F.layer_norm(x, (C,), sd[p + "ln_1.weight"], sd[p + "ln_1.bias"])
        q, k, v = F.linear(h, sd[p + "attn.c_attn.weight"]).split(C, dim=2)
        q, k, v = (t.view(B, T, cfg.n_head, C // cfg.n_head).transpose(1, 2) for t in (q, k, v))
        y = F.scaled_dot_product_attention(q, k, v, is_causal=True).transpose(1, 2).reshape(B, T, C)
        x = x + F.linear(y, sd[p + "attn.c_proj.weight"])
        h = F.layer_norm(x, (C,), sd[p + "ln_2.weight"], sd[p + "ln_2.bias"])
        x = x + F.linear(F.gelu(F.linear(h, sd[p + "mlp.c_fc.weight"]), approximate="tanh"), sd[p + "mlp.c_proj.weight"])
    x = F.layer_norm(x, (cfg.n_embd,), sd["transformer.ln_f.weight"], sd["transformer.ln_f.bias"])
    logits = F.linear(x, sd["lm_head.weight"])
    loss = F.cross_entropy(logits.view(-1, logits.size(-1)), tgt.view(-1))
    loss.backward()
    return logits, loss, sd


def test_gpt2_matches_plain_pytorch():
    torch.manual_seed(0)
    cfg = gpt2_config("tiny")
    m = GPT2Model(cfg)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if "ln" in n:
                p.add_(torch.randn_like(p) * 0.1)
    idx = torch.randint(0, cfg.vocab_size, (2, 32))
    tgt = torch.randint(0, cfg.vocab_size, (2, 32))
    logits, loss = m(idx, tgt)
    loss.backward()
    rlogits, rloss, sd = _ref_forward(m, idx, tgt)
    torch.testing.assert_close(loss, rloss, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(logits, rlogits, rtol=1e-4, atol=1e-5)
    for n, p in m.named_parameters():
        torch.testing.assert_close(p.grad, sd[n].grad, rtol=2e-4, atol=2e-6, msg=n)


def test_single_device_training_reduces_loss_and_trainstep_eager():
    torch.manual_seed(0)
    cfg = gpt2_config("tiny")
    m = GPT2Model(cfg)
    opt = tds.AdamW(m.named_parameters(), lr=1e-3, weight_decay=0.1)
    step = tds.TrainStep(m, opt)          # CPU -> eager path
    idx = torch.randint(0, cfg.vocab_size, (1, 64))
    tgt = torch.randint(0, cfg.vocab_size, (1, 64))
    losses = [float(step(idx, tgt)) for _ in range(8)]
    assert losses[-1] < losses[0] - 0.05, losses
    assert opt.step_count == 8


def test_adopt_reuses_storage_and_error_handling():
    import pytest
    from tiny_deepspeed_b200.parallel import wrap_layers, error_handling
    import tiny_deepspeed_b200.nn as tnn
    seq = nn.Sequential(nn.Linear(4, 4), nn.LayerNorm(4), nn.Embedding(3, 4), nn.GELU(approximate="tanh"))
    ptrs = [p.data_ptr() for p in seq.parameters()]
    seq.train(False)
    wrap_layers(seq)
    assert isinstance(seq[0], tnn.Linear) and isinstance(seq[1], tnn.LayerNorm) and isinstance(seq[2], tnn.Embedding)
    assert [p.data_ptr() for p in seq.parameters()] == ptrs and not seq.training
    error_handling(seq)
    bad = nn.Sequential(nn.Linear(4, 4), nn.Conv1d(2, 2, 1))
    with pytest.raises(NotImplementedError):
        error_handling(wrap_layers(bad))
    with torch.device("meta"):
        mm = wrap_layers(nn.Sequential(nn.Linear(4, 4)))
    assert mm[0].weight.device.type == "meta"
