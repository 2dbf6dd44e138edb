// Fused multi-tensor optimizers for sm_100a: ONE launch updates every owned tensor.
// The tensor table travels by value in the kernel parameter space (CUDA >= 12.1 allows 32 KB), so the launch
// is CUDA-graph capturable without any host->device table copy; the Adam step counter is read from device
// memory so a captured graph keeps advancing bias correction on replay.
// Replaces the reference's ~10 elementwise kernels + 4 temporaries per tensor
// (tiny_deepspeed/core/optim/adamw.py:36-59, sgd.py:28-46).
#include "common.cuh"
#include "kernels.h"

namespace tds {

void set_pdl_enabled(bool on) { pdl_flag() = on ? 1 : 0; }

__global__ void step_inc_kernel(int* p) {
  pdl_launch(); pdl_wait(); *p += 1; }
void step_increment(int* step_ptr, cudaStream_t s) { launch_k(step_inc_kernel, dim3(1), dim3(1), 0, s, step_ptr); }

TDS_DEVICE int find_tensor(const int* blk_start, int count, int b) {
  int lo = 0, hi = count - 1;
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    if (blk_start[mid] <= b) lo = mid; else hi = mid - 1;
  }
  return lo;
}

template <typename T> struct Pk;  // 4-element packets
template <> struct Pk<float> {
  static TDS_DEVICE void ld(const float* p, float* f) { float4 a = *reinterpret_cast<const float4*>(p); f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; }
  static TDS_DEVICE void st(float* p, const float* f) { *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]); }
};
template <> struct Pk<__nv_bfloat16> {
  static TDS_DEVICE void ld(const __nv_bfloat16* p, float* f) {
    uint2 r = *reinterpret_cast<const uint2*>(p);
    float2 a = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&r.x)), b = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&r.y));
    f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y;
  }
  static TDS_DEVICE void st(__nv_bfloat16* p, const float* f) {
    uint2 r;
    *reinterpret_cast<__nv_bfloat162*>(&r.x) = __floats2bfloat162_rn(f[0], f[1]);
    *reinterpret_cast<__nv_bfloat162*>(&r.y) = __floats2bfloat162_rn(f[2], f[3]);
    *reinterpret_cast<uint2*>(p) = r;
  }
};

TDS_DEVICE void adam_math(float& w, float g, float& m, float& v, float* vmax, const AdamHyper& h, float bc1, float bc2_rsqrt) {
  g *= h.grad_scale;
  if (h.maximize) g = -g;
  if (h.weight_decay != 0.f) {
    if (h.decoupled) w *= (1.f - h.lr * h.weight_decay);
    else g += h.weight_decay * w;
  }
  m = h.beta1 * m + (1.f - h.beta1) * g;
  v = h.beta2 * v + (1.f - h.beta2) * g * g;
  float vv = v;
  if (vmax) { *vmax = fmaxf(*vmax, v); vv = *vmax; }
  const float denom = sqrtf(vv) * bc2_rsqrt + h.eps;
  w -= (h.lr / bc1) * (m / denom);
}

template <typename T>
__global__ void __launch_bounds__(256) adamw_multi_kernel(const __grid_constant__ TensorList tl,
                                                          const __grid_constant__ AdamHyper h) {
  pdl_launch(); pdl_wait();
  __shared__ float s_bc[2];
  __shared__ int s_t;
  if (threadIdx.x == 0) {
    const int step = *h.step_ptr;
    s_bc[0] = 1.f - powf(h.beta1, (float)step);
    s_bc[1] = rsqrtf(1.f - powf(h.beta2, (float)step));
    s_t = find_tensor(tl.blk_start, tl.count, blockIdx.x);
  }
  __syncthreads();
  const int t = s_t;
  const float bc1 = s_bc[0], bc2r = s_bc[1];
  T* p = reinterpret_cast<T*>(tl.p[t]);
  const T* g = reinterpret_cast<const T*>(tl.g[t]);
  float* m = tl.m[t];
  float* v = tl.v[t];
  float* master = tl.master[t];
  float* vmax = tl.vmax[t];
  const int64_t n = tl.numel[t];
  const int64_t base = (int64_t)(blockIdx.x - tl.blk_start[t]) * kOptChunk;
  const int64_t end = base + kOptChunk < n ? base + kOptChunk : n;
  const bool vec_ok = (n % 4 == 0) && ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g)) % 16 == 0);
  if (vec_ok) {
    for (int64_t i = base + threadIdx.x * 4; i < end; i += 256 * 4) {
      float w[4], gg[4], mm[4], vv[4], vx[4];
      Pk<T>::ld(g + i, gg);
      if (master) Pk<float>::ld(master + i, w); else Pk<T>::ld(p + i, w);
      Pk<float>::ld(m + i, mm);
      Pk<float>::ld(v + i, vv);
      if (vmax) Pk<float>::ld(vmax + i, vx);
#pragma unroll
      for (int j = 0; j < 4; ++j) adam_math(w[j], gg[j], mm[j], vv[j], vmax ? &vx[j] : nullptr, h, bc1, bc2r);
      Pk<float>::st(m + i, mm);
      Pk<float>::st(v + i, vv);
      if (vmax) Pk<float>::st(vmax + i, vx);
      if (master) Pk<float>::st(master + i, w);
      Pk<T>::st(p + i, w);
    }
  } else {
    for (int64_t i = base + threadIdx.x; i < end; i += 256) {
      float w = master ? master[i] : ldf(p + i), mm = m[i], vv = v[i];
      float vx = vmax ? vmax[i] : 0.f;
      adam_math(w, ldf(g + i), mm, vv, vmax ? &vx : nullptr, h, bc1, bc2r);
      m[i] = mm; v[i] = vv;
      if (vmax) vmax[i] = vx;
      if (master) master[i] = w;
      stf(p + i, w);
    }
  }
}

// Background variant for optimizer-in-backward: a FIXED, small grid (one CTA per SM, 256 threads) walks all chunks.  The
// flooding kernel above (one CTA per 8192 elements, ~20 k CTAs for GPT-2 small) fills every SM's register file, so a backward
// GEMM CTA (one per SM, ~32 k registers) launched next to it waits for several of those CTAs to retire — measured in round 1
// as "overlap buys nothing".  With at most `ctas` resident CTAs of 256 threads the GEMM CTAs always find room and the
// HBM-bound update streams underneath the latency-bound backward; two packets per thread keep enough bytes in flight.
template <typename T>
__global__ void __launch_bounds__(256) adamw_multi_bg_kernel(const __grid_constant__ TensorList tl,
                                                             const __grid_constant__ AdamHyper h) {
  pdl_launch(); pdl_wait();
  const int step = *h.step_ptr;
  const float bc1 = 1.f - powf(h.beta1, (float)step), bc2r = rsqrtf(1.f - powf(h.beta2, (float)step));
  const int total = tl.blk_start[tl.count];
  for (int blk = blockIdx.x; blk < total; blk += gridDim.x) {
    const int t = find_tensor(tl.blk_start, tl.count, blk);
    T* p = reinterpret_cast<T*>(tl.p[t]);
    const T* g = reinterpret_cast<const T*>(tl.g[t]);
    float* m = tl.m[t];
    float* v = tl.v[t];
    float* master = tl.master[t];
    float* vmax = tl.vmax[t];
    const int64_t n = tl.numel[t];
    const int64_t base = (int64_t)(blk - tl.blk_start[t]) * kOptChunk;
    const int64_t end = base + kOptChunk < n ? base + kOptChunk : n;
    const bool vec_ok = (n % 4 == 0) && ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g)) % 16 == 0) && !vmax && master;
    if (vec_ok) {
      for (int64_t i0 = base + threadIdx.x * 4; i0 < end; i0 += 256 * 4 * 2) {
        const int64_t i1 = i0 + 256 * 4;
        const bool two = i1 < end;
        float w0[4], g0[4], m0[4], v0[4], w1[4], g1[4], m1[4], v1[4];
        Pk<T>::ld(g + i0, g0); Pk<float>::ld(master + i0, w0); Pk<float>::ld(m + i0, m0); Pk<float>::ld(v + i0, v0);
        if (two) { Pk<T>::ld(g + i1, g1); Pk<float>::ld(master + i1, w1); Pk<float>::ld(m + i1, m1); Pk<float>::ld(v + i1, v1); }
#pragma unroll
        for (int j = 0; j < 4; ++j) adam_math(w0[j], g0[j], m0[j], v0[j], nullptr, h, bc1, bc2r);
        Pk<float>::st(m + i0, m0); Pk<float>::st(v + i0, v0); Pk<float>::st(master + i0, w0); Pk<T>::st(p + i0, w0);
        if (two) {
#pragma unroll
          for (int j = 0; j < 4; ++j) adam_math(w1[j], g1[j], m1[j], v1[j], nullptr, h, bc1, bc2r);
          Pk<float>::st(m + i1, m1); Pk<float>::st(v + i1, v1); Pk<float>::st(master + i1, w1); Pk<T>::st(p + i1, w1);
        }
      }
    } else {
      for (int64_t i = base + threadIdx.x; i < end; i += 256) {
        float w = master ? master[i] : ldf(p + i), mm = m[i], vv = v[i];
        float vx = vmax ? vmax[i] : 0.f;
        adam_math(w, ldf(g + i), mm, vv, vmax ? &vx : nullptr, h, bc1, bc2r);
        m[i] = mm; v[i] = vv;
        if (vmax) vmax[i] = vx;
        if (master) master[i] = w;
        stf(p + i, w);
      }
    }
  }
}

void adamw_multi(const TensorList& tl, const AdamHyper& h, int dtype, cudaStream_t s, int background_ctas) {
  const int blocks = tl.blk_start[tl.count];
  if (blocks == 0) return;
  if (background_ctas > 0) {
    const int grid = blocks < background_ctas ? blocks : background_ctas;
    if (dtype == kBF16) launch_k(adamw_multi_bg_kernel<__nv_bfloat16>, dim3(grid), dim3(256), 0, s, tl, h);
    else launch_k(adamw_multi_bg_kernel<float>, dim3(grid), dim3(256), 0, s, tl, h);
    return;
  }
  if (dtype == kBF16) launch_k(adamw_multi_kernel<__nv_bfloat16>, dim3(blocks), dim3(256), 0, s, tl, h);
  else launch_k(adamw_multi_kernel<float>, dim3(blocks), dim3(256), 0, s, tl, h);
}

template <typename T>
__global__ void __launch_bounds__(256) sgd_multi_kernel(const __grid_constant__ TensorList tl,
                                                        const __grid_constant__ SgdHyper h) {
  pdl_launch(); pdl_wait();
  __shared__ int s_t, s_first;
  if (threadIdx.x == 0) {
    s_t = find_tensor(tl.blk_start, tl.count, blockIdx.x);
    s_first = (*h.step_ptr == 1);
  }
  __syncthreads();
  const int t = s_t;
  const bool first = s_first != 0;
  T* p = reinterpret_cast<T*>(tl.p[t]);
  const T* g = reinterpret_cast<const T*>(tl.g[t]);
  float* buf = tl.m[t];
  float* master = tl.master[t];
  const int64_t n = tl.numel[t];
  const int64_t base = (int64_t)(blockIdx.x - tl.blk_start[t]) * kOptChunk;
  const int64_t end = base + kOptChunk < n ? base + kOptChunk : n;
  for (int64_t i = base + threadIdx.x; i < end; i += 256) {
    float w = master ? master[i] : ldf(p + i);
    float gg = ldf(g + i) * h.grad_scale;
    if (h.maximize) gg = -gg;
    if (h.weight_decay != 0.f) gg += h.weight_decay * w;
    if (h.momentum != 0.f) {
      float b = first ? gg : h.momentum * buf[i] + (1.f - h.dampening) * gg;
      buf[i] = b;
      gg = h.nesterov ? gg + h.momentum * b : b;
    }
    w -= h.lr * gg;
    if (master) master[i] = w;
    stf(p + i, w);
  }
}

void sgd_multi(const TensorList& tl, const SgdHyper& h, int dtype, cudaStream_t s) {
  const int blocks = tl.blk_start[tl.count];
  if (blocks == 0) return;
  if (dtype == kBF16) launch_k(sgd_multi_kernel<__nv_bfloat16>, dim3(blocks), dim3(256), 0, s, tl, h);
  else launch_k(sgd_multi_kernel<float>, dim3(blocks), dim3(256), 0, s, tl, h);
}

}  // namespace tds
