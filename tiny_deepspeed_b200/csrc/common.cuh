// Shared device helpers for the sm_100a kernels of tiny_deepspeed_b200.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

#define TDS_DEVICE __device__ __forceinline__

#include <stdlib.h>
#include <utility>

namespace tds {

// ---- Programmatic Dependent Launch ------------------------------------------------------------------------------
// Every kernel of ours begins with pdl_launch(); pdl_wait(): the NEXT kernel's CTAs may be scheduled (and run their
// prologue up to their own pdl_wait) while this grid is still draining, which hides launch latency between the ~330
// short kernels of a training step.  griddepcontrol.wait returns only when the prerequisite grid has fully completed
// and its writes are visible, so data hazards are exactly those of plain stream order.
TDS_DEVICE void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
TDS_DEVICE void pdl_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// Round 1 (4.87 ms step, kernels of 10-30 us): 4.98 ms with PDL edges.  Round 2, one GPU (3.2 ms step, GEMMs of 6-10 us whose
// barrier init / TMEM alloc / descriptor prefetch are ~1 us of prologue): 3.214 -> 3.190 ms, so the edges are ON by default.
// With collectives running next to backward they are a loss — 3.570 vs 3.349 ms/step at 2 GPUs (profiles/r2_step_sweeps.md):
// the early-launched CTAs of the next kernel sit on the SM slots the collective's CTAs need — so the multi-GPU policies
// switch them off (ops.set_pdl, unless TDS_PDL is set explicitly).  The griddepcontrol instructions are no-ops without the edge.
inline int& pdl_flag() {
  static int on = !(getenv("TDS_PDL") && atoi(getenv("TDS_PDL")) == 0);
  return on;
}
inline bool pdl_enabled() { return pdl_flag() != 0; }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = at; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

constexpr int kWarp = 32;

TDS_DEVICE float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
TDS_DEVICE float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// block-wide reductions through shared memory (blockDim.x multiple of 32, <= 1024)
TDS_DEVICE float block_sum(float v, float* smem /* >= 32 floats */) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) smem[wid] = v;
  __syncthreads();
  v = (lane < nw) ? smem[lane] : 0.f;
  return warp_sum(v);
}
TDS_DEVICE float block_max(float v, float* smem) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) smem[wid] = v;
  __syncthreads();
  v = (lane < nw) ? smem[lane] : -INFINITY;
  return warp_max(v);
}

// 8 x bf16 <-> 8 x float through one 16-byte transaction
struct alignas(16) bf16x8 {
  __nv_bfloat162 v[4];
};
TDS_DEVICE void unpack8(const bf16x8& p, float* f) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = __bfloat1622float2(p.v[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
TDS_DEVICE bf16x8 pack8(const float* f) {
  bf16x8 p;
#pragma unroll
  for (int i = 0; i < 4; ++i) p.v[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return p;
}
TDS_DEVICE bf16x8 ld8(const __nv_bfloat16* p) { return *reinterpret_cast<const bf16x8*>(p); }
TDS_DEVICE void st8(__nv_bfloat16* p, const bf16x8& v) { *reinterpret_cast<bf16x8*>(p) = v; }

// generic scalar load/store as float for {float, bf16}
template <typename T> TDS_DEVICE float ldf(const T* p);
template <> TDS_DEVICE float ldf<float>(const float* p) { return *p; }
template <> TDS_DEVICE float ldf<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <typename T> TDS_DEVICE void stf(T* p, float v);
template <> TDS_DEVICE void stf<float>(float* p, float v) { *p = v; }
template <> TDS_DEVICE void stf<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

// One MUFU instruction (max relative error ~2^-11, well inside bf16's 2^-8 rounding of every value it feeds), spelled out
// instead of relying on --use_fast_math to turn tanhf into it: the GELU epilogues run on the GEMM's 4 epilogue warps only,
// 192 elements per thread for a 128 x 192 tile, where every instruction is on the critical path (profiles/r2_gemm_epilogue.md).
TDS_DEVICE float tanh_fast(float x) { float y; asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
TDS_DEVICE float gelu_tanh(float x) {
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  float u = k0 * (x + k1 * x * x * x);
  return 0.5f * x * (1.f + tanh_fast(u));
}
TDS_DEVICE float gelu_tanh_grad(float x) {
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  float x2 = x * x;
  float u = k0 * (x + k1 * x * x2);
  float t = tanh_fast(u);
  return 0.5f * (1.f + t) + 0.5f * x * (1.f - t * t) * k0 * (1.f + 3.f * k1 * x2);
}

}  // namespace tds
