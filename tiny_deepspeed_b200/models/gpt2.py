"""GPT-2 model family (L4).  Same classes and parameter names as the reference's `example/model.py:15-157`
(``GPTConfig``, ``GPT2Model(config)(idx, targets) -> (logits, loss)``, ``standard_attention``,
``flash_attention``; registration order wte, wpe, h.i.{ln_1,attn.c_attn,attn.c_proj,ln_2,mlp.c_fc,
mlp.c_proj}, ln_f, lm_head; untied embeddings; ``bias=False``) so partition tables and checkpoints
line up, but built from this package's layers and fused the B200 way:

* ``ln -> (y, residual)`` / ``c_proj(+residual)`` : residual adds live in GEMM epilogues forward and in
  the LayerNorm-backward kernel backward — no stand-alone add kernels;
* ``c_fc -> GELU -> c_proj`` is one autograd node, GELU and GELU' are GEMM epilogues;
* attention consumes the packed qkv buffer through strided TMA maps (no transposes/copies);
* token + position embedding is one gather kernel.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch
import torch.nn as tnn
from torch.nn import functional as F

from .. import nn as tds
from ..nn.modules import fused_mlp, link_prefetch_chain

__all__ = ["GPTConfig", "GPT2Model", "standard_attention", "flash_attention", "CausalSelfAttention",
           "MLP", "Block", "PRESETS", "gpt2_config"]


@dataclass
class GPTConfig:
    block_size: int = 1024
    vocab_size: int = 50304
    max_vocab_size: int = 50257
    n_layer: int = 12
    n_head: int = 12
    n_embd: int = 768
    dropout: float = 0.0
    bias: bool = False
    attention = "standard_attention"  # class attribute, as in the reference; both run the same fused kernel path on GPU


PRESETS = {
    "small": dict(n_layer=12, n_head=12, n_embd=768),     # 163.0 M (untied)
    "medium": dict(n_layer=24, n_head=16, n_embd=1024),   # 406.2 M
    "large": dict(n_layer=36, n_head=20, n_embd=1280),    # 838.1 M
    "xl": dict(n_layer=48, n_head=25, n_embd=1600),       # 1637.5 M
    "tiny": dict(n_layer=2, n_head=2, n_embd=128, vocab_size=512, block_size=128),  # tests / smoke
}


def gpt2_config(name: str = "small", **overrides) -> GPTConfig:
    kw = dict(PRESETS[name])
    kw.update(overrides)
    return GPTConfig(**kw)


def standard_attention(q, k, v, dropout=True, dropout_p=0.0):
    """Materialised-score attention on ``[B,T,nh,hs]`` tensors (reference example/model.py:29-42)."""
    T = q.size(-3)
    scale = 1.0 / math.sqrt(k.size(-1))
    q, k, v = (t.transpose(1, 2) for t in (q, k, v))
    att = (q @ k.transpose(-2, -1)) * scale
    keep = torch.ones(T, T, device=v.device, dtype=torch.bool).tril()
    att = att.masked_fill(~keep, float("-inf"))
    att = F.softmax(att, dim=-1)
    return (att @ v).transpose(1, 2).contiguous()


def flash_attention(q, k, v, dropout=True, dropout_p=0.0):
    """SDPA-based attention on ``[B,T,nh,hs]`` tensors (reference example/model.py:44-51)."""
    scale = 1.0 / math.sqrt(k.size(-1))
    q, k, v = (t.transpose(1, 2) for t in (q, k, v))
    y = F.scaled_dot_product_attention(q, k, v, is_causal=True, scale=scale)
    return y.transpose(1, 2).contiguous()


class CausalSelfAttention(tnn.Module):
    def __init__(self, config):
        super().__init__()
        assert config.n_embd % config.n_head == 0
        self.attention = config.attention
        self.c_attn = tds.Linear(config.n_embd, 3 * config.n_embd, bias=config.bias)
        self.c_proj = tds.Linear(config.n_embd, config.n_embd, bias=config.bias)
        self.n_head = config.n_head
        self.n_embd = config.n_embd

    def forward(self, x, residual=None):
        qkv = self.c_attn(x)
        y = tds.causal_self_attention(qkv, self.n_head)
        return self.c_proj(y, residual=residual)


class MLP(tnn.Module):
    def __init__(self, config):
        super().__init__()
        self.c_fc = tds.Linear(config.n_embd, 4 * config.n_embd, bias=config.bias)
        self.gelu = tds.GELU(approximate="tanh")
        self.c_proj = tds.Linear(4 * config.n_embd, config.n_embd, bias=config.bias)

    def forward(self, x, residual=None):
        return fused_mlp(x, self.c_fc, self.c_proj, residual)


class Block(tnn.Module):
    def __init__(self, config):
        super().__init__()
        self.ln_1 = tds.LayerNorm(config.n_embd)
        self.attn = CausalSelfAttention(config)
        self.ln_2 = tds.LayerNorm(config.n_embd)
        self.mlp = MLP(config)

    def forward(self, x):
        h, res = self.ln_1(x, with_residual=True)
        x = self.attn(h, residual=res)
        h, res = self.ln_2(x, with_residual=True)
        return self.mlp(h, residual=res)


class GPT2Model(tnn.Module):
    def __init__(self, config):
        super().__init__()
        self.config = config
        self.transformer = tnn.ModuleDict(dict(
            wte=tds.Embedding(config.vocab_size, config.n_embd),
            wpe=tds.Embedding(config.block_size, config.n_embd),
            h=tnn.ModuleList([Block(config) for _ in range(config.n_layer)]),
            ln_f=tds.LayerNorm(config.n_embd),
        ))
        self.lm_head = tds.Linear(config.n_embd, config.vocab_size, bias=False)
        self._pos_cache = {}
        # L2 prefetch chain: every GEMM asks L2 for the weight of the GEMM that follows it in the pass (nn/modules.py)
        chain = []
        for blk in self.transformer.h:
            chain += [blk.attn.c_attn, blk.attn.c_proj, blk.mlp.c_fc, blk.mlp.c_proj]
        link_prefetch_chain(chain + [self.lm_head])

    def _positions(self, T, device):
        key = (T, str(device))
        if key not in self._pos_cache:
            self._pos_cache[key] = torch.arange(0, T, dtype=torch.long, device=device)
        return self._pos_cache[key]

    def forward(self, idx, targets=None):
        B, T = idx.size()
        assert T <= self.config.block_size, \
            f"Cannot forward sequence of length {T}, block size is only {self.config.block_size}"
        tr = self.transformer
        pos_emb = tr.wpe(self._positions(T, idx.device))          # [T, C]
        x = tr.wte(idx, add=pos_emb)                              # [B, T, C] gather + add in one kernel
        for block in tr.h:
            x = block(x)
        x = tr.ln_f(x)
        logits = self.lm_head(x)
        loss = None
        if targets is not None:
            loss = tds.cross_entropy(logits.view(-1, logits.size(-1)), targets.view(-1))
        return logits, loss

    def num_parameters(self):
        return sum(int(torch.Size(getattr(p, "_tds_shape", p.shape)).numel()) for p in self.parameters())
