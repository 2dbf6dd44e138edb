// tcgen05.mma issue / completion cost probe (standalone).  One CTA per SM region under test; operands are zero-filled smem
// tiles in the GEMM's layout (128 x 64 bf16 K-major SWIZZLE_128B, B = N x 64), accumulators in TMEM.
//
//   mode 0: one issuer thread: t0, J MMAs (UMMA_K = 16 steps over the same 64-wide tile), t1, commit, t2, wait, t3
//   mode 1: steady state, one issuer: R rounds of (J MMAs + commit -> ring of 4 mbarriers, wait for round r-3)
//   mode 2: steady state, TWO issuer warps on disjoint accumulators (does the ~80-cycle issue cost parallelise?)
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I tiny_deepspeed_b200/csrc -o tools/mma_probe tools/mma_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string>
#include <vector>

#include "sm100_ptx.cuh"

using namespace tds;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

struct P { int mode, J, N, R, ctas_active; long long* out; };

__global__ void __launch_bounds__(128, 1) mma_probe(const __grid_constant__ P p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = base, sB = base + 16384, sBar = sB + 32768;
  __shared__ uint32_t tmem_slot;
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  // zero the operand tiles
  for (uint32_t i = threadIdx.x; i < (16384 + 32768) / 16; i += blockDim.x)
    ptx::st_shared_16(base + i * 16, make_uint4(0, 0, 0, 0));
  if (threadIdx.x == 0) {
    for (int i = 0; i < 16; ++i) ptx::mbar_init(sBar + 8 * i, 1);
    ptx::fence_mbar_init();
  }
  ptx::fence_proxy_async();
  if (warp == 0) { ptx::tmem_alloc(ptx::smem_u32(&tmem_slot), 512); ptx::tmem_relinquish(); }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t idesc = make_idesc_bf16(128, p.N, false, false);
  long long* out = p.out + blockIdx.x * 16;
  if (p.mode == 0) {
    if (warp == 1 && ptx::elect_one()) {
      for (int rep = 0; rep < 3; ++rep) {     // last repetition is reported (warm instruction cache)
        const long long t0 = clock64();
        for (int j = 0; j < p.J; ++j) {
          const uint64_t da = ptx::make_smem_desc(sA + (j & 3) * 32, 16, 1024), db = ptx::make_smem_desc(sB + (j & 3) * 32, 16, 1024);
          ptx::mma_f16_ss(tmem, da, db, idesc, j > 0 ? 1u : 0u);
        }
        const long long t1 = clock64();
        ptx::mma_commit(sBar);
        const long long t2 = clock64();
        ptx::mbar_wait_conv(sBar, rep & 1);
        const long long t3 = clock64();
        out[0] = t1 - t0; out[1] = t2 - t1; out[2] = t3 - t2;
      }
    }
  } else {
    const int nissuers = p.mode == 2 ? 2 : 1;
    if (warp >= 1 && warp <= nissuers) {
      const int w = warp - 1;
      const uint32_t d = tmem + w * 256;
      const uint32_t bar0 = sBar + 8 * (w * 4);
      const long long t0 = clock64();
      for (int r = 0; r < p.R; ++r) {
        if (r >= 4) ptx::mbar_wait_conv(bar0 + 8 * (r & 3), ((r >> 2) - 1) & 1);
        if (ptx::elect_one()) {
#pragma unroll 4
          for (int j = 0; j < p.J; ++j) {
            const uint64_t da = ptx::make_smem_desc(sA + (j & 3) * 32, 16, 1024), db = ptx::make_smem_desc(sB + (j & 3) * 32, 16, 1024);
            ptx::mma_f16_ss(d, da, db, idesc, (r | j) ? 1u : 0u);
          }
          ptx::mma_commit(bar0 + 8 * (r & 3));
        }
        __syncwarp();
      }
      // drain
      for (int r = p.R; r < p.R + 4; ++r) if (r >= 4) ptx::mbar_wait_conv(bar0 + 8 * (r & 3), ((r >> 2) - 1) & 1);
      const long long t1 = clock64();
      if (lane == 0) out[w] = t1 - t0;
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem, 512); }
}

int main() {
  long long* out;
  CK(cudaMalloc(&out, 148 * 16 * 8));
  std::vector<long long> h(148 * 16);
  CK(cudaFuncSetAttribute(mma_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  printf("# mode 0: single shot.  cycles: issue of J MMAs | commit | wait for completion\n");
  for (int N : {64, 128, 256})
    for (int J : {1, 2, 4, 8, 16, 32}) {
      P p{0, J, N, 0, 1, out};
      CK(cudaMemset(out, 0, 148 * 16 * 8));
      mma_probe<<<1, 128, 56 * 1024>>>(p);
      CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(h.data(), out, 148 * 16 * 8, cudaMemcpyDeviceToHost));
      printf("N=%3d J=%2d | issue %5lld (%5.1f / MMA) | commit %4lld | wait %5lld | total/MMA %6.1f\n", N, J, h[0], (double)h[0] / J, h[1], h[2],
             (double)(h[0] + h[1] + h[2]) / J);
    }
  printf("# mode 1/2: steady state, R = 64 rounds of J MMAs + commit (ring of 4).  cycles per MMA\n");
  for (int mode : {1, 2})
    for (int ctas : {1, 148})
      for (int N : {64, 128, 256})
        for (int J : {4, 8}) {
          P p{mode, J, N, 64, ctas, out};
          CK(cudaMemset(out, 0, 148 * 16 * 8));
          for (int it = 0; it < 3; ++it) mma_probe<<<ctas, 128, 56 * 1024>>>(p);
          CK(cudaDeviceSynchronize());
          CK(cudaMemcpy(h.data(), out, 148 * 16 * 8, cudaMemcpyDeviceToHost));
          const double c0 = (double)h[0] / (64.0 * J), c1 = (double)h[1] / (64.0 * J);
          printf("mode %d ctas %3d N=%3d J=%d | issuer0 %6.1f cyc/MMA (round %6.0f)%s\n", mode, ctas, N, J, c0, c0 * J,
                 mode == 2 ? (std::string(" | issuer1 ") + std::to_string(c1)).c_str() : "");
        }
  return 0;
}
