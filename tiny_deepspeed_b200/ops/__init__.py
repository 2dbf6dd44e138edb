"""L0 op layer: free functions for every hot op of the engine.

API parity with the reference's nine exported ops (`tiny_deepspeed/core/module/ops/__init__.py:4-18`):
``linear_forward, linear_input_grad, linear_weight_grad, linear_bias_grad, layernorm_fwd,
layernorm_dx, layernorm_dwdb, embedding_forward, embedding_weight_grad`` — each still accepts an
optional ``runtime_tuner`` — plus the ops the reference leaves to ATen (attention, GELU,
cross-entropy, optimizer updates), which here are our own kernels as well.

CUDA tensors run on the sm_100a extension (tcgen05 GEMMs etc.); CPU tensors run the PyTorch
implementations below, which double as the fp32 oracle for the GPU numerics tests.
"""
from __future__ import annotations

import math
import os
from typing import Optional, Tuple

import torch
import torch.nn.functional as F

from . import _dispatch
from ._dispatch import (ext, on_gpu, count_launch, launches, reset_launches, force_torch,  # noqa: F401
                        is_forced_torch)

# epilogue selectors shared with csrc/gemm_sm100.cu
EPI_NONE = 0
EPI_GELU_SAVE = 1      # out = gelu(acc + bias); aux <- (acc + bias)          (c_fc forward)
EPI_GELU_BWD = 2       # out = acc * gelu'(aux)                               (mlp.c_proj dX)
EPI_RESIDUAL = 3       # out = acc + bias + aux                               (c_proj forward)

__all__ = [
    "linear_forward", "linear_input_grad", "linear_weight_grad", "linear_bias_grad",
    "layernorm_fwd", "layernorm_dx", "layernorm_dwdb", "layernorm_bwd",
    "embedding_forward", "embedding_weight_grad",
    "gelu_forward", "gelu_backward",
    "causal_attention_forward", "causal_attention_backward",
    "cross_entropy_forward", "cross_entropy_backward",
    "adamw_update", "sgd_update", "gemm",
]


# --------------------------------------------------------------------------------------
# GEMM
# --------------------------------------------------------------------------------------

def _gelu_tanh(x):
    return F.gelu(x, approximate="tanh")


def _gelu_tanh_grad(x):
    k0, k1 = 0.7978845608028654, 0.044715
    x = x.float()
    u = k0 * (x + k1 * x * x * x)
    t = torch.tanh(u)
    return 0.5 * (1.0 + t) + 0.5 * x * (1.0 - t * t) * k0 * (1.0 + 3.0 * k1 * x * x)


def gemm(a: torch.Tensor, b: torch.Tensor, *, a_mn: bool = False, b_mn: bool = False,
         out: Optional[torch.Tensor] = None, out_dtype: Optional[torch.dtype] = None,
         bias: Optional[torch.Tensor] = None, aux: Optional[torch.Tensor] = None,
         epi: int = EPI_NONE, accumulate: bool = False, alpha: float = 1.0,
         config: Optional[int] = None, tri: int = 0, cluster: int = 0, reduce_out: bool = False) -> torch.Tensor:
    """General (optionally batched) GEMM ``D[m,n] = alpha * sum_k A(m,k) * B(n,k)`` (+epilogue).

    ``a`` is stored ``[.., M, K]`` (K-major) or, with ``a_mn=True``, ``[.., K, M]`` (MN-major);
    likewise ``b`` is ``[.., N, K]`` or ``[.., K, N]``.  The inner stride must be 1, the other
    strides are free (so q/k/v head views of a packed qkv buffer are consumed in place).
    On CUDA this is one launch of the persistent tcgen05 kernel (csrc/gemm_sm100.cu).

    ``reduce_out=True`` (experimental): ``out`` is an fp32 ``[M,N]`` tensor — possibly another rank's copy of a symmetric
    buffer — and the product is *added* into it by the epilogue's TMA reduce (fused GEMM -> reduce-scatter).
    """
    if on_gpu(a, b):
        if reduce_out or getattr(out, "_tds_reduce", False):
            return _gemm_cuda(a, b, a_mn, b_mn, out, out_dtype, None, None, EPI_NONE, False, alpha, config, 0, 1, True)
        return _gemm_cuda(a, b, a_mn, b_mn, out, out_dtype, bias, aux, epi, accumulate, alpha, config, tri, cluster)
    if reduce_out:
        accumulate = True
    A = a.transpose(-1, -2) if a_mn else a
    Bm = b.transpose(-1, -2) if b_mn else b
    acc = torch.matmul(A.float(), Bm.float().transpose(-1, -2)) * alpha
    odt = out.dtype if out is not None else (out_dtype or a.dtype)
    if bias is not None:
        acc = acc + bias.float()
    if epi == EPI_GELU_SAVE:
        aux.copy_(acc.to(aux.dtype))
        acc = _gelu_tanh(aux.float())
    elif epi == EPI_GELU_BWD:
        acc = acc * _gelu_tanh_grad(aux)
    elif epi == EPI_RESIDUAL:
        acc = acc + aux.float()
    if out is None:
        return acc.to(odt)
    if accumulate:
        out.add_(acc.to(out.dtype))
    else:
        out.copy_(acc.to(out.dtype))
    return out


def set_pdl(on: bool, force: bool = False):
    """Programmatic-dependent-launch edges between our kernels.  Default: on (single GPU); the multi-GPU policies call
    ``set_pdl(False)`` because the early-launched CTAs take the SM slots their collectives need (csrc/common.cuh).  An explicit
    ``TDS_PDL`` in the environment wins unless ``force``.  Takes effect for kernels launched (or captured) afterwards."""
    if ("TDS_PDL" in os.environ and not force) or not torch.cuda.is_available() or is_forced_torch():
        return
    ext().set_pdl(bool(on))


_L2_PREFETCH = os.environ.get("TDS_L2_PREFETCH", "1") != "0"
_L2_PREFETCH_MAX = 16 << 20


def prefetch_next(t):
    """L2 hint: the next GEMM launch also asks L2 to fetch ``t`` (the weight the FOLLOWING kernel will stream from HBM) with its
    idle epilogue warps while its own main loop runs (``cp.async.bulk.prefetch.L2``; csrc/gemm_sm100.cu).  At 1 x 1024 tokens
    every weight of a pass comes from HBM (250 MB of them against 126 MB of L2) into a latency-bound single-wave GEMM;
    measured per GEMM: profiles/r2_gemm_rot.log.  No-op on CPU, for empty (non-resident ZeRO-3) or very large tensors."""
    if not _L2_PREFETCH or t is None or not t.is_cuda or is_forced_torch():
        return
    nbytes = t.numel() * t.element_size()
    if nbytes < 16 or nbytes > _L2_PREFETCH_MAX or not t.is_contiguous():
        return
    ext().gemm_set_prefetch(t)


def _gemm_cuda(a, b, a_mn, b_mn, out, out_dtype, bias, aux, epi, accumulate, alpha, config, tri=0, cluster=0, reduce_out=False):
    if getattr(b, "_tds_remote", False):
        # B aliases a peer GPU's memory (ZeRO-3 direct-fetch mode).  TMA *multicast* sourced from peer-mapped memory
        # wedged the GPU in testing (2xB200, r1), so the peer-fetch GEMM always runs with plain per-CTA TMA loads.
        cluster = 1
    if a.dim() == 2:
        M = a.shape[1] if a_mn else a.shape[0]
        N = b.shape[1] if b_mn else b.shape[0]
        shape = (M, N)
    else:
        M = a.shape[-1] if a_mn else a.shape[-2]
        N = b.shape[-1] if b_mn else b.shape[-2]
        shape = (*a.shape[:-2], M, N)
    if out is None:
        out = torch.empty(shape, device=a.device, dtype=out_dtype or a.dtype)
    ext().gemm(a, b, out, a_mn, b_mn, bias, aux, int(epi), bool(accumulate), float(alpha),
               -1 if config is None else int(config), int(tri), int(cluster), bool(reduce_out))
    count_launch()
    return out


def _flat2d(x):
    return x.reshape(-1, x.shape[-1])


def _splitk_factor(M, N, K, sms=148):
    """Split the reduction when a long-K GEMM has too few output tiles to fill the chip (lm_head dX: 48 tiles, K=50304)."""
    tiles = ((M + 127) // 128) * ((N + 127) // 128)
    if tiles * 2 > sms:
        return 1
    want = max(1, (2 * sms) // tiles)
    for s in range(min(want, 16), 1, -1):
        if K % (64 * s) == 0:
            return s
    return 1


def gemm_splitk(a, b, splits):
    """``D = A[M,K] . B[K,N]`` (A K-major, B MN-major) with K cut into ``splits`` batches: one batched tcgen05 launch
    into fp32 slices + one fold kernel.  Used where K is huge and M x N small."""
    M, K = a.shape
    N = b.shape[1]
    kc = K // splits
    a3 = a.view(M, splits, kc).transpose(0, 1)          # [S, M, K/S] strided view, no copy
    b3 = b.view(splits, kc, N)                          # [S, K/S, N]
    ws = torch.empty(splits, M, N, device=a.device, dtype=torch.float32)
    gemm(a3, b3, b_mn=True, out=ws)
    out = torch.empty(M, N, device=a.device, dtype=a.dtype)
    ext().sum_slices(ws, out)
    count_launch()
    return out


def _gemm_tuned(tuner, key, a, b, **kw):
    """Route a GEMM through the RuntimeAutoTuner: the candidates are the tile configurations of OUR kernel
    (heuristic, BN = 64 / 128 / 256), timed with CUDA events and cached per (op key, shapes).  Accumulating calls are
    never measured (re-running them would add the product several times)."""
    if tuner is None or not getattr(tuner, "enable", False) or not on_gpu(a, b) or kw.get("accumulate"):
        return gemm(a, b, **kw)
    import functools
    n_out = (b.shape[-1] if kw.get("b_mn") else b.shape[-2])
    configs = [None, 0, 1, 2] + ([3] if n_out % 192 == 0 else [])      # BN = heuristic, 64, 128, 256, 192
    cands = [functools.partial(gemm, config=c) for c in configs]
    return tuner.choose_function(cands, a, b, key=key, **kw)


def linear_forward(input, weight, bias=None, runtime_tuner=None, *, gelu_aux=None, residual=None):
    """``Y = X @ W^T (+ b)`` (reference ops/linear.py:50-54).

    ``gelu_aux`` (a ``[.., N]`` buffer) fuses GELU-tanh into the epilogue and receives the
    pre-activation; ``residual`` fuses a residual add.  Both are epilogue functors of the same
    tcgen05 kernel on GPU.
    """
    x2 = _flat2d(input)
    if gelu_aux is not None:
        y = _gemm_tuned(runtime_tuner, "linear_fwd_gelu", x2, weight, bias=bias, aux=_flat2d(gelu_aux), epi=EPI_GELU_SAVE)
    elif residual is not None:
        y = _gemm_tuned(runtime_tuner, "linear_fwd_res", x2, weight, bias=bias, aux=_flat2d(residual), epi=EPI_RESIDUAL)
    else:
        y = _gemm_tuned(runtime_tuner, "linear_fwd", x2, weight, bias=bias)
    return y.view(*input.shape[:-1], weight.shape[0])


def linear_input_grad(grad_output, weight, runtime_tuner=None, *, gelu_aux=None):
    """``dX = dY @ W`` (reference ops/linear.py:56-57); W ``[N,K]`` is consumed MN-major, no transpose copy.

    ``gelu_aux``: pre-activation of the *preceding* fused GELU — the epilogue multiplies by gelu'.
    """
    dy2 = _flat2d(grad_output)
    if gelu_aux is not None:
        dx = gemm(dy2, weight, b_mn=True, aux=_flat2d(gelu_aux), epi=EPI_GELU_BWD)
    elif on_gpu(dy2) and dy2.dtype == torch.bfloat16 and weight.shape[0] >= 8192 and _splitk_factor(dy2.shape[0], weight.shape[1], weight.shape[0]) > 1:
        dx = gemm_splitk(dy2, weight, _splitk_factor(dy2.shape[0], weight.shape[1], weight.shape[0]))
    else:
        dx = _gemm_tuned(runtime_tuner, "linear_dx", dy2, weight, b_mn=True)
    return dx.view(*grad_output.shape[:-1], weight.shape[1])


def linear_weight_grad(grad_output, input, weight=None, runtime_tuner=None, *, out=None,
                       accumulate=False, out_dtype=None):
    """``dW[N,K] = dY^T @ X`` with all leading dims flattened (reference ops/linear.py:59-68).

    ``out``/``accumulate`` let the GEMM epilogue write (or add) straight into a flat gradient
    buffer, which is what the comm policies hand in.
    """
    dy2, x2 = _flat2d(grad_output), _flat2d(input)
    return _gemm_tuned(runtime_tuner, "linear_dw", dy2, x2, a_mn=True, b_mn=True, out=out, accumulate=accumulate,
                       out_dtype=out_dtype or (weight.dtype if weight is not None else input.dtype))


def linear_bias_grad(grad_output, bias=None, runtime_tuner=None, *, out=None, accumulate=False):
    """``db = sum over all leading dims of dY`` — fixes the reference's 3-D bug (ops/linear.py:70-75)."""
    dy2 = _flat2d(grad_output)
    if on_gpu(dy2):
        if out is None:
            out = torch.empty(dy2.shape[1], device=dy2.device, dtype=dy2.dtype)
            accumulate = False
        ext().colsum(dy2, out, bool(accumulate))
        count_launch()
        return out
    db = dy2.float().sum(0)
    if out is None:
        return db.to(grad_output.dtype)
    out.add_(db.to(out.dtype)) if accumulate else out.copy_(db.to(out.dtype))
    return out


# --------------------------------------------------------------------------------------
# LayerNorm
# --------------------------------------------------------------------------------------

def layernorm_fwd(input, weight, bias, eps=1e-5, runtime_tuner=None):
    """Row LayerNorm over the last dim; returns ``(y, mean, rstd)`` with fp32 statistics
    (reference ops/layernorm.py:46-80 — three Triton passes; here one pass, row held in registers)."""
    x2 = _flat2d(input)
    if on_gpu(x2):
        y, mean, rstd = ext().layernorm_fwd(x2, weight, bias, float(eps))
        count_launch()
        return y.view_as(input), mean, rstd
    xf = x2.float()
    mean = xf.mean(-1)
    var = xf.var(-1, unbiased=False)
    rstd = torch.rsqrt(var + eps)
    y = (xf - mean[:, None]) * rstd[:, None] * weight.float() + bias.float()
    return y.to(input.dtype).view_as(input), mean, rstd


def layernorm_bwd(grad_output, input, weight, mean, rstd, *, dw_out=None, db_out=None,
                  accumulate=False, add_to_dx=None, runtime_tuner=None):
    """Fused LayerNorm backward: ``dx`` plus ``dw``/``db`` (two launches on GPU: row pass with
    per-CTA fp32 partials — no global spin-lock unlike reference ops/layernorm.py:257-269 — and a
    column reduce).  ``add_to_dx`` fuses the residual-branch gradient add."""
    dy2, x2 = _flat2d(grad_output), _flat2d(input)
    if on_gpu(dy2):
        if dw_out is None:
            dw_out = torch.empty_like(weight)
            db_out = torch.empty_like(weight)
            accumulate = False
        add2 = None if add_to_dx is None else _flat2d(add_to_dx)
        fast = dy2.dtype == torch.bfloat16 and dy2.shape[1] % 8 == 0 and dy2.shape[1] <= 2048
        variant = -1                                                   # -1: TDS_LN_SINGLE decides (default two kernels)
        tuner = runtime_tuner
        if tuner is not None and getattr(tuner, "enable", False) and fast and not accumulate:
            # the tuner picks between the two-kernel form (per-CTA partials + fold) and the single-launch form (L2 reductions
            # + last-CTA finish) per (rows, width) — reference ops/layernorm.py:82-127 threads its tuner through the same op
            import functools
            cands = [functools.partial(ext().layernorm_bwd, variant=v) for v in (0, 1)]
            dx = tuner.choose_function(cands, dy2, x2, weight, mean, rstd, dw_out, db_out, False, add2, key="layernorm_bwd")
            variant = tuner.best("layernorm_bwd", dy2, x2, weight, mean, rstd, dw_out, db_out, False, add2)
        else:
            dx = ext().layernorm_bwd(dy2, x2, weight, mean, rstd, dw_out, db_out, bool(accumulate), add2, -1)
        single = fast and (variant == 1 or (variant == -1 and os.environ.get("TDS_LN_SINGLE", "0") != "0"))
        count_launch(1 if single else 2)       # single launch (atomic column sums) or row pass + partial fold
        return dx.view_as(input), dw_out, db_out
    dyf, xf = dy2.float(), x2.float()
    xhat = (xf - mean[:, None]) * rstd[:, None]
    wdy = dyf * weight.float()
    c1 = (xhat * wdy).mean(-1, keepdim=True)
    c2 = wdy.mean(-1, keepdim=True)
    dx = (wdy - (xhat * c1 + c2)) * rstd[:, None]
    if add_to_dx is not None:
        dx = dx + _flat2d(add_to_dx).float()
    dw = (dyf * xhat).sum(0)
    db = dyf.sum(0)
    if dw_out is None:
        return dx.to(input.dtype).view_as(input), dw.to(weight.dtype), db.to(weight.dtype)
    if accumulate:
        dw_out.add_(dw.to(dw_out.dtype)); db_out.add_(db.to(db_out.dtype))
    else:
        dw_out.copy_(dw.to(dw_out.dtype)); db_out.copy_(db.to(db_out.dtype))
    return dx.to(input.dtype).view_as(input), dw_out, db_out


def layernorm_dx(grad_output, input, weight, bias, mean, rstd, args=None, runtime_tuner=None):
    """API-parity entry point (reference ops/layernorm.py:82-127): returns ``(dx, (dw, db))`` where
    the second element plays the role of the reference's partial-sum scratch."""
    dx, dw, db = layernorm_bwd(grad_output, input, weight, mean, rstd)
    return dx, (dw, db)


def layernorm_dwdb(weight, bias, partials, args=None, runtime_tuner=None):
    """API-parity entry point (reference ops/layernorm.py:130-156): finishes dw/db from ``layernorm_dx``."""
    dw, db = partials
    return dw, db


# --------------------------------------------------------------------------------------
# Embedding
# --------------------------------------------------------------------------------------

def embedding_forward(input, weight, padding_idx=None, max_norm=None, norm_type=2.0,
                      scale_grad_by_freq=False, sparse=False, runtime_tuner=None, *, add=None):
    """Row gather ``weight[idx]`` (reference ops/embedding.py:34-58); ``add`` fuses ``+ pos_emb``."""
    if max_norm is not None:
        with torch.no_grad():
            torch.embedding_renorm_(weight, input.reshape(-1), max_norm, norm_type)
    if on_gpu(weight):
        out = ext().embedding_fwd(input.reshape(-1), weight, None if add is None else _flat2d(add),
                                  0 if add is None else _flat2d(add).shape[0])
        count_launch()
        return out.view(*input.shape, weight.shape[1])
    out = weight.index_select(0, input.reshape(-1)).view(*input.shape, weight.shape[1])
    if add is not None:
        out = out + add
    return out


def embedding_weight_grad(input, grad_output, weight, padding_idx=None, runtime_tuner=None, *,
                          out=None, accumulate=False, shape=None):
    """Dense ``[V, D]`` gradient by scatter-add (reference ops/embedding.py:60-65)."""
    idx = input.reshape(-1)
    dy2 = _flat2d(grad_output)
    shape = tuple(shape) if shape is not None else tuple(weight.shape)
    if on_gpu(dy2):
        if out is None:
            out = torch.empty(shape, dtype=weight.dtype, device=dy2.device)
            accumulate = False
        ext().embedding_bwd(idx, dy2, out, bool(accumulate), -1 if padding_idx is None else int(padding_idx))
        count_launch(1 if accumulate else 2)
        return out
    g = torch.zeros(shape, dtype=torch.float32, device=dy2.device)
    if padding_idx is not None:
        keep = idx != padding_idx
        idx, dy2 = idx[keep], dy2[keep]
    g.index_add_(0, idx, dy2.float())
    if out is None:
        return g.to(weight.dtype)
    out.add_(g.to(out.dtype)) if accumulate else out.copy_(g.to(out.dtype))
    return out


# --------------------------------------------------------------------------------------
# GELU (standalone; the fused path is a GEMM epilogue)
# --------------------------------------------------------------------------------------

def gelu_forward(x):
    if on_gpu(x):
        y = ext().gelu_fwd(x.contiguous())
        count_launch()
        return y
    return _gelu_tanh(x)


def gelu_backward(grad_output, x):
    if on_gpu(x):
        dx = ext().gelu_bwd(grad_output.contiguous(), x.contiguous())
        count_launch()
        return dx
    return (grad_output.float() * _gelu_tanh_grad(x)).to(x.dtype)


# --------------------------------------------------------------------------------------
# Causal self-attention on a packed qkv buffer
# --------------------------------------------------------------------------------------

_flash = os.environ.get("TDS_FLASH", "1") != "0"


def flash_enabled() -> bool:
    return _flash


def set_flash(flag: bool) -> None:
    """Switch between the fused flash-attention kernels and the materialised-score GEMM path (both ours)."""
    global _flash
    _flash = bool(flag)


def cast(x, dtype):
    """bf16 <-> fp32 conversion (one kernel of ours on GPU)."""
    if x.dtype == dtype:
        return x
    if on_gpu(x) and {x.dtype, dtype} == {torch.float32, torch.bfloat16}:
        y = ext().cast(x.contiguous())
        count_launch()
        return y
    return x.to(dtype)


def _split_heads(qkv, n_head):
    B, T, C3 = qkv.shape
    C = C3 // 3
    hs = C // n_head
    q, k, v = qkv.split(C, dim=2)
    # [B, nh, T, hs] views — no copies; the GEMM consumes the strides through its TMA maps
    return (t.view(B, T, n_head, hs).transpose(1, 2) for t in (q, k, v))


def causal_attention_forward(qkv, n_head):
    """``softmax(mask(QK^T/sqrt(hs))) V`` on packed ``qkv [B,T,3C]``; returns ``(y [B,T,C], P)`` with the
    probabilities ``P [B,nh,T,T]`` kept for backward.  GPU: two batched tcgen05 GEMMs (strided
    head views, no transposes) + one causal-softmax kernel.  Replaces the reference's
    ``standard_attention`` (example/model.py:29-42)."""
    if on_gpu(qkv) and qkv.dtype == torch.float32:
        # fp32 model: the attention core runs on the bf16 tensor-core kernels (fp32 softmax statistics / accumulation)
        y, P = causal_attention_forward(cast(qkv, torch.bfloat16), n_head)
        return cast(y, torch.float32), P
    B, T, C3 = qkv.shape
    C = C3 // 3
    hs = C // n_head
    scale = 1.0 / math.sqrt(hs)
    if on_gpu(qkv) and flash_enabled() and qkv.is_contiguous() and ext().flash_supported(T, hs):
        # fused tcgen05 flash kernel: no T x T tensor; the second return value is the log2-domain LSE [B,nh,T] (fp32)
        y, lse = ext().flash_fwd(qkv, n_head)
        count_launch()
        return y, lse
    q, k, v = _split_heads(qkv, n_head)
    if on_gpu(qkv):
        S = torch.empty(B, n_head, T, T, device=qkv.device, dtype=qkv.dtype)
        gemm(q, k, out=S, tri=1)                            # S = Q K^T (tiles above the diagonal skipped)
        ext().softmax_causal_fwd(S.view(-1, T, T), float(scale))
        count_launch()
        y = torch.empty(B, T, C, device=qkv.device, dtype=qkv.dtype)
        gemm(S, v, b_mn=True, out=y.view(B, T, n_head, hs).transpose(1, 2), tri=2)   # Y = P V (V MN-major)
        return y, S
    att = torch.matmul(q.float(), k.float().transpose(-1, -2)) * scale
    mask = torch.ones(T, T, dtype=torch.bool, device=qkv.device).tril()
    att = att.masked_fill(~mask, float("-inf"))
    P = torch.softmax(att, dim=-1)
    y = torch.matmul(P, v.float()).transpose(1, 2).reshape(B, T, C)
    return y.to(qkv.dtype), P.to(qkv.dtype)


def causal_attention_backward(grad_y, qkv, P, n_head, y=None):
    """Backward of :func:`causal_attention_forward`; returns ``dqkv [B,T,3C]``.  ``P`` is whatever forward returned
    second: the fp32 LSE (flash path: one fused kernel + tiny prep/convert kernels; needs ``y``) or the bf16
    probabilities (materialised path: four batched GEMMs and one softmax-backward kernel)."""
    if on_gpu(qkv) and qkv.dtype == torch.float32:
        bf = torch.bfloat16
        dqkv = causal_attention_backward(cast(grad_y.contiguous(), bf), cast(qkv, bf), P, n_head,
                                         y=None if y is None else cast(y, bf))
        return cast(dqkv, torch.float32)
    if P.dtype == torch.float32 and P.dim() == 3 and on_gpu(qkv):
        dqkv = ext().flash_bwd(grad_y.contiguous(), qkv, y, P, n_head)
        count_launch(3)
        return dqkv
    B, T, C3 = qkv.shape
    C = C3 // 3
    hs = C // n_head
    scale = 1.0 / math.sqrt(hs)
    q, k, v = _split_heads(qkv, n_head)
    dy = grad_y.view(B, T, n_head, hs).transpose(1, 2)          # [B,nh,T,hs]
    if on_gpu(qkv):
        dqkv = torch.empty_like(qkv)
        dq, dk, dv = _split_heads(dqkv, n_head)
        dP = torch.empty_like(P)
        gemm(dy, v, out=dP, tri=1)                               # dP = dY V^T
        gemm(P, dy, a_mn=True, b_mn=True, out=dv, tri=3)        # dV = P^T dY
        ext().softmax_causal_bwd(P.view(-1, T, T), dP.view(-1, T, T), float(scale))   # dP <- dS
        count_launch()
        gemm(dP, k, b_mn=True, out=dq, tri=2)                    # dQ = dS K
        gemm(dP, q, a_mn=True, b_mn=True, out=dk, tri=3)         # dK = dS^T Q
        return dqkv
    Pf, dyf = P.float(), dy.float()
    dV = torch.matmul(Pf.transpose(-1, -2), dyf)
    dP = torch.matmul(dyf, v.float().transpose(-1, -2))
    dS = Pf * (dP - (dP * Pf).sum(-1, keepdim=True)) * scale
    dQ = torch.matmul(dS, k.float())
    dK = torch.matmul(dS.transpose(-1, -2), q.float())
    dqkv = torch.cat([t.transpose(1, 2).reshape(B, T, C) for t in (dQ, dK, dV)], dim=2)
    return dqkv.to(qkv.dtype)


# --------------------------------------------------------------------------------------
# Cross-entropy
# --------------------------------------------------------------------------------------

def cross_entropy_forward(logits, targets):
    """Mean token cross-entropy; returns ``(loss fp32 scalar, lse fp32 [M])``."""
    l2 = _flat2d(logits)
    t = targets.reshape(-1)
    if on_gpu(l2):
        loss, lse = ext().cross_entropy_fwd(l2, t)
        count_launch(2)
        return loss, lse
    lf = l2.float()
    lse = torch.logsumexp(lf, dim=-1)
    loss = (lse - lf.gather(1, t[:, None]).squeeze(1)).mean()
    return loss, lse


def cross_entropy_backward(grad_loss, logits, targets, lse, *, out=None):
    """``dlogits = grad_loss * (softmax(logits) - onehot) / M`` from the saved log-sum-exp."""
    l2 = _flat2d(logits)
    t = targets.reshape(-1)
    if on_gpu(l2):
        if out is None:
            out = torch.empty_like(l2)
        ext().cross_entropy_bwd(l2, t, lse, grad_loss.reshape(1).float(), _flat2d(out))
        count_launch()
        return out.view_as(logits)
    p = torch.exp(l2.float() - lse[:, None])
    p[torch.arange(p.shape[0], device=p.device), t] -= 1.0
    d = (p * (grad_loss.float() / l2.shape[0])).to(logits.dtype).view_as(logits)
    if out is not None:
        out.copy_(d)
        return out
    return d


# --------------------------------------------------------------------------------------
# Optimizer updates (multi-tensor, one launch)
# --------------------------------------------------------------------------------------

def adamw_update(params, grads, exp_avgs, exp_avg_sqs, masters, *, lr, beta1, beta2, eps,
                 weight_decay, step, decoupled=False, maximize=False, grad_scale=1.0,
                 max_exp_avg_sqs=None, step_dev=None, background_ctas=0):
    """One fused Adam step over a list of tensors.

    Update rule parity: the reference's "AdamW" is Adam with *coupled* L2 (``g += wd*p``,
    reference optim/adamw.py:37-38); ``decoupled=True`` gives true AdamW.  ``step`` is the
    per-*step* counter.  ``masters`` are optional fp32 master copies for bf16 params.
    """
    if on_gpu(*params):
        ext().adamw_multi(params, grads, exp_avgs, exp_avg_sqs,
                          masters if masters is not None else [],
                          max_exp_avg_sqs if max_exp_avg_sqs is not None else [],
                          float(lr), float(beta1), float(beta2), float(eps), float(weight_decay),
                          step_dev, bool(decoupled), bool(maximize), float(grad_scale), int(background_ctas))
        count_launch()
        return
    bc1 = 1.0 - beta1 ** step
    bc2 = 1.0 - beta2 ** step
    for i, p in enumerate(params):
        g = grads[i].float() * grad_scale
        if maximize:
            g = -g
        w = masters[i] if masters is not None else p
        wf = w.float()
        if weight_decay != 0.0:
            if decoupled:
                wf = wf * (1.0 - lr * weight_decay)
            else:
                g = g + weight_decay * wf
        m, v = exp_avgs[i], exp_avg_sqs[i]
        m.mul_(beta1).add_(g, alpha=1.0 - beta1)
        v.mul_(beta2).addcmul_(g, g, value=1.0 - beta2)
        vv = v
        if max_exp_avg_sqs is not None:
            torch.maximum(max_exp_avg_sqs[i], v, out=max_exp_avg_sqs[i])
            vv = max_exp_avg_sqs[i]
        denom = (vv / bc2).sqrt_().add_(eps)
        wf = wf - (lr / bc1) * (m / denom)
        if masters is not None:
            masters[i].copy_(wf)
        p.copy_(wf.to(p.dtype))


def sgd_update(params, grads, bufs, masters, *, lr, momentum, dampening, weight_decay, nesterov,
               maximize=False, first_step=False, grad_scale=1.0, step_dev=None):
    """One fused SGD(+momentum) step over a list of tensors (reference optim/sgd.py:28-46)."""
    if on_gpu(*params):
        ext().sgd_multi(params, grads, bufs if bufs is not None else [],
                        masters if masters is not None else [],
                        float(lr), float(momentum), float(dampening), float(weight_decay),
                        bool(nesterov), bool(maximize), step_dev, float(grad_scale))
        count_launch()
        return
    for i, p in enumerate(params):
        g = grads[i].float() * grad_scale
        if maximize:
            g = -g
        w = masters[i] if masters is not None else p
        wf = w.float()
        if weight_decay != 0.0:
            g = g + weight_decay * wf
        if momentum != 0.0:
            buf = bufs[i]
            if first_step:
                buf.copy_(g)
            else:
                buf.mul_(momentum).add_(g, alpha=1.0 - dampening)
            g = g + momentum * buf if nesterov else buf
        wf = wf - lr * g
        if masters is not None:
            masters[i].copy_(wf)
        p.copy_(wf.to(p.dtype))
