"""Layers with policy-driven backward.  See package docstring."""
from __future__ import annotations

import os

import torch
import torch.nn as tnn

from .. import ops
from .policy import policy_of


# Tracing: TDS_NVTX=1 wraps every layer's forward/backward in an NVTX range (visible in nsys / ncu --nvtx);
# zero overhead when off because the decorator returns the function unchanged.
_NVTX = os.environ.get("TDS_NVTX", "0") == "1"


def _traced(name):
    def deco(fn):
        if not _NVTX:
            return fn

        def wrapped(*a, **k):
            torch.cuda.nvtx.range_push(name)
            try:
                return fn(*a, **k)
            finally:
                torch.cuda.nvtx.range_pop()
        return wrapped
    return deco


def _grad_of(policy, param, compute, rows=None):
    """Run ``compute(out, accumulate) -> grad`` into the policy's buffer and publish the result."""
    out, acc = policy.grad_out(param)
    g = compute(out, acc)
    if rows is None:
        policy.grad_ready(param, g)
    else:
        policy.grad_ready(param, g, rows=rows)


# dW || dX: the two GEMMs of a Linear's backward are independent and, at 1x1024 tokens, each is a single under-filled
# wave of latency-bound CTAs — so dW is launched on a side stream while dX runs on the current one (fork/join inside
# the captured graph).  TDS_DUAL_STREAM=0 restores the serial order.
_DUAL = os.environ.get("TDS_DUAL_STREAM", "1") != "0"
# TDS_SERIAL_BWD lists sites ("mlp_proj", "mlp_fc", "linear") that run dW then dX on one stream instead.  The kernel timeline
# suggested it for the MLP down-projection (its dX next to its dW shows as 23-28 us against ~8 + ~11 us back to back: the two
# single-wave kernels cannot share an SM), but the step says otherwise — none 3.115, mlp_proj 3.121, +mlp_fc 3.136,
# all 3.174 ms (profiles/r2_step_sweeps.md) — so the default is empty.
_SERIAL_SITES = set(filter(None, os.environ.get("TDS_SERIAL_BWD", "").split(",")))
_side_streams = {}


def _side_stream(t: torch.Tensor):
    if not _DUAL or not t.is_cuda or ops.is_forced_torch():
        return None
    key = t.device.index
    if key not in _side_streams:
        _side_streams[key] = torch.cuda.Stream(t.device)
    return _side_streams[key]


def _linear_backward(pol, weight, shape, dy, x, tuner, dx_fn, site="linear"):
    """``dW`` (into the policy's gradient buffer) and ``dx_fn()`` (the dX GEMM); concurrent when possible."""
    side = _side_stream(dy) if weight.requires_grad and site not in _SERIAL_SITES else None
    if side is None:
        if weight.requires_grad:
            _grad_of(pol, weight, lambda out, acc: ops.linear_weight_grad(
                dy, x, weight, tuner, out=out, accumulate=acc, out_dtype=weight.dtype))
        return dx_fn()
    out, acc = pol.grad_out(weight)
    if out is None:                                   # allocate on the CURRENT stream: the optimizer consumes it there
        out, acc = torch.empty(shape, dtype=weight.dtype, device=dy.device), False
    cur = torch.cuda.current_stream(dy.device)
    side.wait_stream(cur)                             # fork: dy (and x) are ready
    with torch.cuda.stream(side):
        g = ops.linear_weight_grad(dy, x, weight, tuner, out=out, accumulate=acc, out_dtype=weight.dtype)
    dx = dx_fn()
    cur.wait_stream(side)                             # join before anything downstream (collective, optimizer) touches dW
    pol.grad_ready(weight, g)
    return dx


def _hint(module, attr):
    """L2 prefetch hint for the GEMM about to be launched: the weight of the Linear whose GEMM comes next in this pass
    (``link_prefetch_chain``).  Stored as a module reference (not a Parameter) so nothing is registered twice."""
    nxt = module.__dict__.get(attr)
    if nxt is not None:
        ops.prefetch_next(nxt.weight.data)


def link_prefetch_chain(linears):
    """``linears``: the model's Linear layers in forward execution order.  Layer i's forward GEMM will prefetch layer i+1's
    weight into L2, its dX GEMM layer i-1's (ops.prefetch_next)."""
    for a, b in zip(linears, linears[1:]):
        object.__setattr__(a, "_pf_fwd", b)
        object.__setattr__(b, "_pf_bwd", a)


# ----------------------------------------------------------------------------------------
# Linear
# ----------------------------------------------------------------------------------------

class _LinearFn(torch.autograd.Function):
    @staticmethod
    @_traced("linear.fwd")
    def forward(ctx, x, weight, bias, module, residual):
        pol = policy_of(module)
        w = pol.acquire(weight)
        _hint(module, "_pf_fwd")
        y = ops.linear_forward(x, w, bias, getattr(module, "runtime_tuner", None), residual=residual)
        pol.release(weight, w)
        ctx.module = module
        ctx.has_res = residual is not None
        ctx.save_for_backward(x)
        return y

    @staticmethod
    @_traced("linear.bwd")
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        module = ctx.module
        pol = policy_of(module)
        weight, bias = module.weight, module.bias
        dy = dy.contiguous()
        tuner = getattr(module, "runtime_tuner", None)
        if bias is not None and bias.requires_grad:
            _grad_of(pol, bias, lambda out, acc: ops.linear_bias_grad(dy, bias, out=out, accumulate=acc))

        def dx_fn():
            if not ctx.needs_input_grad[0]:
                return None
            w = pol.acquire(weight, backward=True)
            _hint(module, "_pf_bwd")
            d = ops.linear_input_grad(dy, w, tuner)
            pol.release(weight, w)
            return d

        # dW and dX are issued together (dW on a side stream); the gradient's collective starts right after
        # (reference ordering idea, tiny_deepspeed/core/zero/ddp/module.py:36-66, minus the cuda.synchronize()).
        dx = _linear_backward(pol, weight, (module.out_features, module.in_features), dy, x, tuner, dx_fn)
        return dx, None, None, None, (dy if ctx.has_res else None)


class Linear(tnn.Linear):
    """``nn.Linear`` with our GEMMs and a comm policy (reference module/linear.py:16-92)."""

    policy = None
    runtime_tuner = None

    def forward(self, input, residual=None):
        return _LinearFn.apply(input, self.weight, self.bias, self, residual)


class _MLPFn(torch.autograd.Function):
    """c_fc -> GELU(tanh) -> c_proj (+residual) as one autograd node so that GELU forward/backward
    live in GEMM epilogues (EPI_GELU_SAVE on c_fc, EPI_GELU_BWD on c_proj's dX)."""

    @staticmethod
    @_traced("mlp.fwd")
    def forward(ctx, x, fc, proj, residual):
        pf, pp = policy_of(fc), policy_of(proj)
        pre = torch.empty(*x.shape[:-1], fc.out_features, device=x.device, dtype=x.dtype)
        w = pf.acquire(fc.weight)
        _hint(fc, "_pf_fwd")
        act = ops.linear_forward(x, w, fc.bias, gelu_aux=pre)
        pf.release(fc.weight, w)
        w = pp.acquire(proj.weight)
        _hint(proj, "_pf_fwd")
        y = ops.linear_forward(act, w, proj.bias, residual=residual)
        pp.release(proj.weight, w)
        ctx.fc, ctx.proj = fc, proj
        ctx.has_res = residual is not None
        ctx.save_for_backward(x, pre, act)
        return y

    @staticmethod
    @_traced("mlp.bwd")
    def backward(ctx, dy):
        x, pre, act = ctx.saved_tensors
        fc, proj = ctx.fc, ctx.proj
        pf, pp = policy_of(fc), policy_of(proj)
        dy = dy.contiguous()
        if proj.bias is not None:
            _grad_of(pp, proj.bias, lambda out, acc: ops.linear_bias_grad(dy, out=out, accumulate=acc))

        def dpre_fn():
            w = pp.acquire(proj.weight, backward=True)
            _hint(proj, "_pf_bwd")
            d = ops.linear_input_grad(dy, w, gelu_aux=pre)       # GELU' folded into the dX epilogue
            pp.release(proj.weight, w)
            return d

        dpre = _linear_backward(pp, proj.weight, (proj.out_features, proj.in_features), dy, act, None, dpre_fn, site="mlp_proj")
        if fc.bias is not None:
            _grad_of(pf, fc.bias, lambda out, acc: ops.linear_bias_grad(dpre, out=out, accumulate=acc))

        def dx_fn():
            if not ctx.needs_input_grad[0]:
                return None
            w = pf.acquire(fc.weight, backward=True)
            _hint(fc, "_pf_bwd")
            d = ops.linear_input_grad(dpre, w)
            pf.release(fc.weight, w)
            return d

        dx = _linear_backward(pf, fc.weight, (fc.out_features, fc.in_features), dpre, x, None, dx_fn, site="mlp_fc")
        return dx, None, None, (dy if ctx.has_res else None)


def fused_mlp(x, fc: Linear, proj: Linear, residual=None):
    return _MLPFn.apply(x, fc, proj, residual)


# ----------------------------------------------------------------------------------------
# LayerNorm
# ----------------------------------------------------------------------------------------

class _LayerNormFn(torch.autograd.Function):
    @staticmethod
    @_traced("layernorm.fwd")
    def forward(ctx, x, weight, bias, module, with_residual):
        pol = policy_of(module)
        w, b = pol.acquire(weight), pol.acquire(bias)
        y, mean, rstd = ops.layernorm_fwd(x, w, b, module.eps)
        pol.release(weight, w)
        pol.release(bias, b)
        ctx.module = module
        ctx.with_residual = with_residual
        ctx.save_for_backward(x, mean, rstd)
        if with_residual:
            # second output is x itself: the residual branch.  Its gradient comes back into this
            # node and is folded into the dx kernel (no separate add kernel in backward).
            return y, x.view_as(x)
        return y

    @staticmethod
    @_traced("layernorm.bwd")
    def backward(ctx, dy, dres=None):
        x, mean, rstd = ctx.saved_tensors
        module = ctx.module
        pol = policy_of(module)
        weight, bias = module.weight, module.bias
        w = pol.acquire(weight, backward=True)
        wo, wacc = pol.grad_out(weight)
        bo, bacc = pol.grad_out(bias)
        if wo is None or bo is None or wacc != bacc:
            wo = bo = None
            wacc = False
        dx, dw, db = ops.layernorm_bwd(dy.contiguous(), x, w, mean, rstd, dw_out=wo, db_out=bo,
                                       accumulate=wacc,
                                       add_to_dx=dres.contiguous() if dres is not None else None,
                                       runtime_tuner=getattr(module, "runtime_tuner", None))
        pol.release(weight, w)
        pol.grad_ready(weight, dw)
        pol.grad_ready(bias, db)
        return dx, None, None, None, None


class LayerNorm(tnn.LayerNorm):
    """Last-dim LayerNorm with affine weight+bias (same restrictions as reference
    module/normalization.py:35-38,62-63)."""

    policy = None
    runtime_tuner = None

    def _check(self):
        if not self.elementwise_affine or self.bias is None:
            raise NotImplementedError("LayerNorm requires elementwise_affine=True and bias=True")
        if len(self.normalized_shape) != 1:
            raise NotImplementedError("LayerNorm normalises the last dimension only")

    def forward(self, input, with_residual: bool = False):
        self._check()
        return _LayerNormFn.apply(input, self.weight, self.bias, self, with_residual)


# ----------------------------------------------------------------------------------------
# Embedding
# ----------------------------------------------------------------------------------------

class _EmbeddingFn(torch.autograd.Function):
    @staticmethod
    @_traced("embedding.fwd")
    def forward(ctx, idx, weight, module, add):
        pol = policy_of(module)
        w = pol.acquire(weight, sparse=True)          # ZeRO-3: a gather touches <= ntokens rows of a [V, D] table
        y = ops.embedding_forward(idx, w, module.padding_idx, module.max_norm, module.norm_type,
                                  module.scale_grad_by_freq, module.sparse, add=add)
        pol.release(weight, w)
        ctx.module = module
        ctx.add_shape = None if add is None else tuple(add.shape)
        ctx.save_for_backward(idx)
        return y

    @staticmethod
    @_traced("embedding.bwd")
    def backward(ctx, dy):
        (idx,) = ctx.saved_tensors
        module = ctx.module
        pol = policy_of(module)
        weight = module.weight
        dy = dy.contiguous()
        _grad_of(pol, weight, lambda out, acc: ops.embedding_weight_grad(
            idx, dy, weight, module.padding_idx, out=out, accumulate=acc,
            shape=(module.num_embeddings, module.embedding_dim)), rows=idx)
        dadd = None
        if ctx.add_shape is not None and ctx.needs_input_grad[3]:
            dadd = dy
            while dadd.dim() > len(ctx.add_shape):
                dadd = dadd[0] if dadd.shape[0] == 1 else dadd.sum(0)
        return None, None, None, dadd


class Embedding(tnn.Embedding):
    """Embedding with dense gradient (reference module/embedding.py:15-98); ``add`` fuses ``+ pos``."""

    policy = None
    runtime_tuner = None

    def forward(self, input, add=None):
        return _EmbeddingFn.apply(input, self.weight, self, add)


# ----------------------------------------------------------------------------------------
# GELU / attention / loss
# ----------------------------------------------------------------------------------------

class _GeluFn(torch.autograd.Function):
    @staticmethod
    @_traced("gelu.fwd")
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return ops.gelu_forward(x)

    @staticmethod
    @_traced("gelu.bwd")
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        return ops.gelu_backward(dy, x)


class GELU(tnn.GELU):
    """tanh-GELU (the model's only activation, reference example/model.py:94)."""

    def forward(self, input):
        if self.approximate != "tanh":
            return super().forward(input)
        return _GeluFn.apply(input)


class _AttnFn(torch.autograd.Function):
    @staticmethod
    @_traced("attention.fwd")
    def forward(ctx, qkv, n_head):
        y, aux = ops.causal_attention_forward(qkv, n_head)      # aux: LSE (flash kernels) or P (materialised path)
        ctx.n_head = n_head
        ctx.save_for_backward(qkv, aux, y)
        return y

    @staticmethod
    @_traced("attention.bwd")
    def backward(ctx, dy):
        qkv, aux, y = ctx.saved_tensors
        return ops.causal_attention_backward(dy.contiguous(), qkv, aux, ctx.n_head, y=y), None


def causal_self_attention(qkv: torch.Tensor, n_head: int) -> torch.Tensor:
    """Causal multi-head attention on the packed ``c_attn`` output ``[B,T,3C]`` -> ``[B,T,C]``."""
    return _AttnFn.apply(qkv, n_head)


class _XentFn(torch.autograd.Function):
    @staticmethod
    @_traced("cross_entropy.fwd")
    def forward(ctx, logits, targets):
        loss, lse = ops.cross_entropy_forward(logits, targets)
        ctx.save_for_backward(logits, targets, lse)
        return loss.to(torch.float32)

    @staticmethod
    @_traced("cross_entropy.bwd")
    def backward(ctx, g):
        logits, targets, lse = ctx.saved_tensors
        return ops.cross_entropy_backward(g, logits, targets, lse), None


def cross_entropy(logits: torch.Tensor, targets: torch.Tensor) -> torch.Tensor:
    """Mean token cross-entropy (fp32 scalar)."""
    return _XentFn.apply(logits, targets)


# ----------------------------------------------------------------------------------------
# In-place adoption of plain torch.nn layers
# ----------------------------------------------------------------------------------------

def supported_modules():
    """``{torch class: our class}`` (reference zero/ddp/wrapper.py:36-40 ``_supported_modules``)."""
    return {tnn.Linear: Linear, tnn.LayerNorm: LayerNorm, tnn.Embedding: Embedding, tnn.GELU: GELU}


def adopt(model: tnn.Module) -> tnn.Module:
    """Turn every supported ``torch.nn`` layer of ``model`` into ours *in place* by re-classing the
    instance: parameters keep their storage (works on meta tensors, costs nothing for XL) — the
    reference re-creates each layer on CPU with a full random init and copies the weights back
    (`tiny_deepspeed/core/zero/utils/wrapper.py:22-36`)."""
    table = supported_modules()
    for m in model.modules():
        for src, dst in table.items():
            if type(m) is src:
                m.__class__ = dst
                break
    return model
