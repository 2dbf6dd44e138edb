// Standalone probe: which shape of the multi-tensor Adam kernel gets closest to the HBM copy roofline on a B200?
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/adam_probe tools/adam_probe.cu && tools/adam_probe
// One 124 M-element "tensor" (GPT-2 small): bf16 p, g; fp32 master, m, v.  28 bytes of HBM traffic per element.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

struct H { float lr, b1, b2, eps, wd, bc1, bc2r; };

template <bool FASTM>
__device__ __forceinline__ void adam1(float& w, float g, float& m, float& v, const H& h) {
  g += h.wd * w;
  m = h.b1 * m + (1.f - h.b1) * g;
  v = h.b2 * v + (1.f - h.b2) * g * g;
  if (FASTM) {
    float sq; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(sq) : "f"(v));
    w -= (h.lr / h.bc1) * __fdividef(m, sq * h.bc2r + h.eps);
  } else {
    w -= (h.lr / h.bc1) * (m / (sqrtf(v) * h.bc2r + h.eps));
  }
}

template <int HINT> __device__ __forceinline__ float4 ld4(const float* p) {
  if (HINT == 1) return __ldcs(reinterpret_cast<const float4*>(p));
  if (HINT == 2) { float4 r; asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p)); return r; }
  return *reinterpret_cast<const float4*>(p);
}
template <int HINT> __device__ __forceinline__ void st4(float* p, float4 v) {
  if (HINT == 1) __stcs(reinterpret_cast<float4*>(p), v); else *reinterpret_cast<float4*>(p) = v;
}
template <int HINT> __device__ __forceinline__ uint2 ld2(const __nv_bfloat16* p) {
  if (HINT == 1) return __ldcs(reinterpret_cast<const uint2*>(p));
  return *reinterpret_cast<const uint2*>(p);
}
template <int HINT> __device__ __forceinline__ void st2(__nv_bfloat16* p, uint2 v) {
  if (HINT == 1) __stcs(reinterpret_cast<uint2*>(p), v); else *reinterpret_cast<uint2*>(p) = v;
}

// PK packets (4 elements each) per thread per iteration; CHUNK elements per CTA (0 = grid-stride over everything)
template <int PK, int HINT, bool FASTM, int THREADS>
__global__ void __launch_bounds__(THREADS) adam_kernel(__nv_bfloat16* __restrict__ p, const __nv_bfloat16* __restrict__ g,
                                                       float* __restrict__ master, float* __restrict__ m, float* __restrict__ v,
                                                       long long n, long long chunk, H h) {
  long long base, end, stride;
  if (chunk) { base = (long long)blockIdx.x * chunk; end = base + chunk < n ? base + chunk : n; stride = (long long)THREADS * 4 * PK; }
  else { base = (long long)blockIdx.x * THREADS * 4 * PK; end = n; stride = (long long)gridDim.x * THREADS * 4 * PK; }
  for (long long i0 = base + threadIdx.x * 4; i0 < end; i0 += stride) {
    float4 w[PK], mm[PK], vv[PK]; uint2 gr[PK];
#pragma unroll
    for (int k = 0; k < PK; ++k) {
      const long long i = i0 + (long long)k * THREADS * 4;
      if (i < end) { gr[k] = ld2<HINT>(g + i); w[k] = ld4<HINT>(master + i); mm[k] = ld4<HINT>(m + i); vv[k] = ld4<HINT>(v + i); }
    }
#pragma unroll
    for (int k = 0; k < PK; ++k) {
      const long long i = i0 + (long long)k * THREADS * 4;
      if (i < end) {
        float2 a = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&gr[k].x)), b = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&gr[k].y));
        adam1<FASTM>(w[k].x, a.x, mm[k].x, vv[k].x, h); adam1<FASTM>(w[k].y, a.y, mm[k].y, vv[k].y, h);
        adam1<FASTM>(w[k].z, b.x, mm[k].z, vv[k].z, h); adam1<FASTM>(w[k].w, b.y, mm[k].w, vv[k].w, h);
        st4<HINT>(m + i, mm[k]); st4<HINT>(v + i, vv[k]); st4<HINT>(master + i, w[k]);
        uint2 r;
        *reinterpret_cast<__nv_bfloat162*>(&r.x) = __floats2bfloat162_rn(w[k].x, w[k].y);
        *reinterpret_cast<__nv_bfloat162*>(&r.y) = __floats2bfloat162_rn(w[k].z, w[k].w);
        st2<HINT>(p + i, r);
      }
    }
  }
}

__global__ void copy_kernel(const float4* __restrict__ a, float4* __restrict__ b, long long n4) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) b[i] = a[i];
}

template <typename F> float time_it(F f, int iters = 10) {
  for (int i = 0; i < 3; ++i) f();
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  CK(cudaDeviceSynchronize());
  cudaEventRecord(e0);
  for (int i = 0; i < iters; ++i) f();
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms * 1e3f / iters;
}

int main() {
  const long long n = 124439808LL;   // GPT-2 small
  __nv_bfloat16 *p, *g; float *master, *m, *v;
  CK(cudaMalloc(&p, n * 2)); CK(cudaMalloc(&g, n * 2)); CK(cudaMalloc(&master, n * 4)); CK(cudaMalloc(&m, n * 4)); CK(cudaMalloc(&v, n * 4));
  CK(cudaMemset(p, 0, n * 2)); CK(cudaMemset(g, 0, n * 2)); CK(cudaMemset(master, 0, n * 4)); CK(cudaMemset(m, 0, n * 4)); CK(cudaMemset(v, 0, n * 4));
  H h{1e-5f, 0.9f, 0.999f, 1e-8f, 0.1f, 0.1f, 31.6f};
  const double bytes = 28.0 * n;
  auto report = [&](const char* name, float us) { printf("%-58s %8.1f us  %6.0f GB/s\n", name, us, bytes / us / 1e3); };
  {
    float us = time_it([&] { copy_kernel<<<148 * 16, 512>>>((const float4*)master, (float4*)m, n / 4); });
    printf("%-58s %8.1f us  %6.0f GB/s (read+write)\n", "copy 498 MB (float4 grid-stride)", us, 8.0 * n / us / 1e3);
    us = time_it([&] { cudaMemcpyAsync(m, master, n * 4, cudaMemcpyDeviceToDevice); });
    printf("%-58s %8.1f us  %6.0f GB/s (read+write)\n", "cudaMemcpy D2D 498 MB", us, 8.0 * n / us / 1e3);
  }
#define RUN(PK, HINT, FASTM, THREADS, CHUNK, GRID, NAME) \
  report(NAME, time_it([&] { adam_kernel<PK, HINT, FASTM, THREADS><<<(GRID), THREADS>>>(p, g, master, m, v, n, CHUNK, h); CK(cudaGetLastError()); }))
  const int g8k = (int)((n + 8191) / 8192), g16k = (int)((n + 16383) / 16384), g32k = (int)((n + 32767) / 32768);
  RUN(1, 0, false, 256, 8192, g8k, "V0  pk1 256thr chunk8192 (current)");
  RUN(1, 0, true, 256, 8192, g8k, "V1  pk1 256thr chunk8192 fast-math");
  RUN(2, 0, false, 256, 8192, g8k, "V2  pk2 256thr chunk8192");
  RUN(2, 0, true, 256, 8192, g8k, "V2f pk2 256thr chunk8192 fast-math");
  RUN(2, 1, true, 256, 8192, g8k, "V3  pk2 256thr chunk8192 fast-math .cs");
  RUN(1, 1, true, 256, 8192, g8k, "V3a pk1 256thr chunk8192 fast-math .cs");
  RUN(2, 2, true, 256, 8192, g8k, "V3b pk2 256thr chunk8192 fast-math L1::no_allocate");
  RUN(4, 0, true, 256, 16384, g16k, "V4  pk4 256thr chunk16384 fast-math");
  RUN(2, 0, true, 512, 16384, g16k, "V5  pk2 512thr chunk16384 fast-math");
  RUN(2, 0, true, 256, 32768, g32k, "V6  pk2 256thr chunk32768 fast-math");
  RUN(2, 0, true, 256, 0, 148 * 8, "V7  pk2 256thr grid-stride 148x8 fast-math");
  RUN(2, 1, true, 256, 0, 148 * 8, "V7c pk2 256thr grid-stride 148x8 fast-math .cs");
  RUN(4, 0, true, 256, 0, 148 * 4, "V8  pk4 256thr grid-stride 148x4 fast-math");
  RUN(2, 0, true, 512, 0, 148 * 4, "V9  pk2 512thr grid-stride 148x4 fast-math");
  RUN(1, 0, true, 1024, 0, 148 * 2, "V10 pk1 1024thr grid-stride 148x2 fast-math");
  RUN(2, 0, true, 128, 4096, (int)((n + 4095) / 4096), "V11 pk2 128thr chunk4096 fast-math");
  printf("done\n");
  return 0;
}
