// TMA ingest probe (standalone, no torch): what limits the operand pipeline of a small tcgen05 GEMM on B200?
//
// Each CTA runs the GEMM's producer/consumer mbarrier ring WITHOUT the MMAs: one thread issues the A box (128 x 64 bf16) and
// the B box (BN x 64 bf16) of every k-block into an S-stage smem ring, a second thread waits for "full" and immediately
// releases the stage.  Sweeps: ring depth, B box height, CTA count, tile sharing pattern (GEMM-like: CTAs in a row share A and
// CTAs in a column share B; private: every CTA streams its own rows), box K extent.  Prints the steady-state period per k-block
// (ns, %globaltimer) and bytes/clk/SM, so the numbers can be compared with the per-k-block period of gemm_kernel
// (tools/gemm_timeline.py).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tma_probe tools/tma_probe.cu -lcuda && tools/tma_probe
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <algorithm>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  long long t0 = clock64();
  while (!ok) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (clock64() - t0 > 4000000000LL) asm volatile("trap;");
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"((uint64_t)m), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ long long gtime() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

struct Probe {
  int nkb;        // k-blocks per tile
  int tiles;      // tiles per CTA
  int stages;
  int bn;         // B box rows
  int kbox;       // K elements per box (64: one swizzle row)
  int share;      // 1: GEMM-like sharing (A row = blockIdx % 8, B row = blockIdx / 8), 0: private rows per CTA
  int only;       // 0 both, 1 only A, 2 only B
  long long* out; // per CTA: t_start, t_first_full, t_end, (then per-k-block stamps of CTA 0 from slot 8 on)
};

__global__ void __launch_bounds__(128, 1) probe_kernel(const __grid_constant__ CUtensorMap ta, const __grid_constant__ CUtensorMap tb,
                                                       const __grid_constant__ Probe p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_bytes = 128u * p.kbox * 2u, b_bytes = (uint32_t)p.bn * p.kbox * 2u;
  const uint32_t sA = base, sB = base + p.stages * a_bytes, sBar = sB + p.stages * b_bytes;
  auto full = [&](int s) { return sBar + 8u * s; };
  auto empty = [&](int s) { return sBar + 8u * (p.stages + s); };
  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&ta) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tb) : "memory");
    for (int s = 0; s < p.stages; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int arow = p.share ? (blockIdx.x % 8) * 128 : blockIdx.x * 128;
  const int brow = p.share ? (blockIdx.x / 8) * p.bn : blockIdx.x * p.bn;
  const uint32_t tx = (p.only == 2 ? 0u : a_bytes) + (p.only == 1 ? 0u : b_bytes);
  if (warp == 0 && lane == 0) {
    p.out[blockIdx.x * 4 + 0] = gtime();
    int stage = 0; uint32_t phase = 0;
    for (int t = 0; t < p.tiles; ++t)
      for (int kb = 0; kb < p.nkb; ++kb) {
        mbar_wait(empty(stage), phase ^ 1u);
        mbar_expect_tx(full(stage), tx);
        if (p.only != 2) tma_load_2d(sA + stage * a_bytes, &ta, full(stage), kb * p.kbox, arow);
        if (p.only != 1) tma_load_2d(sB + stage * b_bytes, &tb, full(stage), kb * p.kbox, brow);
        if (++stage == p.stages) { stage = 0; phase ^= 1u; }
      }
  } else if (warp == 1 && lane == 0) {
    int stage = 0; uint32_t phase = 0;
    int n = 0;
    for (int t = 0; t < p.tiles; ++t)
      for (int kb = 0; kb < p.nkb; ++kb) {
        mbar_wait(full(stage), phase);
        if (n == 0) p.out[blockIdx.x * 4 + 1] = gtime();
        if (blockIdx.x == 0 && n < 64) p.out[4096 + n] = gtime();
        ++n;
        mbar_arrive(empty(stage));
        if (++stage == p.stages) { stage = 0; phase ^= 1u; }
      }
    p.out[blockIdx.x * 4 + 2] = gtime();
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeFn get_encode() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  return (EncodeFn)fn;
}

static CUtensorMap make2d(EncodeFn enc, void* ptr, int rows, int K, int box_rows, int kbox, CUtensorMapL2promotion prom) {
  CUtensorMap m;
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)K * 2};
  cuuint32_t box[2] = {(cuuint32_t)kbox, (cuuint32_t)box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, ptr, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   kbox == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, prom, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
  return m;
}

int main(int argc, char** argv) {
  EncodeFn enc = get_encode();
  const int K = 3072;                 // row length (elements): 48 k-blocks of 64
  const int rowsA = 148 * 128, rowsB = 148 * 256;
  __nv_bfloat16 *A, *B;
  CK(cudaMalloc(&A, (size_t)rowsA * K * 2));
  CK(cudaMalloc(&B, (size_t)rowsB * K * 2));
  CK(cudaMemset(A, 0, (size_t)rowsA * K * 2));
  CK(cudaMemset(B, 0, (size_t)rowsB * K * 2));
  long long* out;
  CK(cudaMalloc(&out, 8192 * 8));
  std::vector<long long> h(8192);
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
  int clk_khz = 0;
  CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
  printf("# SM clock (attr) %d MHz\n", clk_khz / 1000);
  printf("%-7s %4s %3s %3s %5s %5s %4s | %9s %9s %9s | %8s %8s %10s\n", "pattern", "ctas", "S", "bn", "kbox", "only", "nkb",
         "first_ns", "period_ns", "total_ns", "B/ns/SM", "B/clk/SM", "chip GB/s");
  struct Case { int share, ctas, stages, bn, kbox, only, nkb, prom; };
  std::vector<Case> cases;
  for (int share : {1, 0})
    for (int ctas : {1, 8, 48, 96, 144})
      for (int bn : {64, 128, 256})
        cases.push_back({share, ctas, bn == 256 ? 4 : (bn == 128 ? 6 : 8), bn, 64, 0, 48, 2});
  for (int S : {1, 2, 3, 4, 6}) cases.push_back({1, 144, S, 128, 64, 0, 48, 2});   // ring depth
  for (int S : {1, 2, 4, 6}) cases.push_back({1, 1, S, 128, 64, 0, 48, 2});        // ring depth, one CTA alone
  for (int only : {1, 2}) { cases.push_back({1, 144, 6, 128, 64, only, 48, 2}); cases.push_back({1, 1, 6, 128, 64, only, 48, 2}); }
  for (int prom : {0, 1, 3}) cases.push_back({1, 144, 6, 128, 64, 0, 48, prom});  // L2 promotion none / 64B / 128B (2 = 256B above)
  cases.push_back({1, 144, 6, 128, 64, 0, 12, 2});                                 // the real c_attn length
  cases.push_back({1, 96, 8, 64, 64, 0, 12, 2});                                   // attn.c_proj
  for (const Case& c : cases) {
    CUtensorMapL2promotion prom = c.prom == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE
                                  : (c.prom == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
                                                 : (c.prom == 3 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_256B));
    CUtensorMap ta = make2d(enc, A, rowsA, K, 128, c.kbox, prom), tb = make2d(enc, B, rowsB, K, c.bn, c.kbox, prom);
    Probe p{c.nkb, 1, c.stages, c.bn, c.kbox, c.share, c.only, out};
    const size_t smem = (size_t)c.stages * (128 + c.bn) * c.kbox * 2 + 2048;
    std::vector<double> first, per, tot;
    for (int it = 0; it < 12; ++it) {
      // keep the GPU busy (clocks up) and the operands L2-warm like inside the training step: 6 back-to-back launches
      CK(cudaMemsetAsync(out, 0, 8192 * 8));
      for (int r = 0; r < 6; ++r) probe_kernel<<<c.ctas, 128, smem>>>(ta, tb, p);
      CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(h.data(), out, 8192 * 8, cudaMemcpyDeviceToHost));
      if (it < 2) continue;
      long long t0 = h[0], t1 = h[2];
      std::vector<double> f, pp;
      for (int b = 0; b < c.ctas; ++b) {
        t0 = std::min(t0, h[b * 4]); t1 = std::max(t1, h[b * 4 + 2]);
        f.push_back((double)(h[b * 4 + 1] - h[b * 4]));
        pp.push_back((double)(h[b * 4 + 2] - h[b * 4 + 1]) / (c.nkb - 1));
      }
      std::sort(f.begin(), f.end()); std::sort(pp.begin(), pp.end());
      first.push_back(f[f.size() / 2]); per.push_back(pp[pp.size() / 2]); tot.push_back((double)(t1 - t0));
    }
    std::sort(first.begin(), first.end()); std::sort(per.begin(), per.end()); std::sort(tot.begin(), tot.end());
    const double pns = per[per.size() / 2];
    const double bytes = (c.only == 2 ? 0 : 128.0 * c.kbox * 2) + (c.only == 1 ? 0 : (double)c.bn * c.kbox * 2);
    printf("%-7s %4d %3d %3d %5d %5d %4d | %9.0f %9.1f %9.0f | %8.1f %8.1f %10.0f\n", c.share ? "shared" : "private", c.ctas,
           c.stages, c.bn, c.kbox, c.only, c.nkb, first[first.size() / 2], pns, tot[tot.size() / 2], bytes / pns,
           bytes / pns / (clk_khz / 1e6), bytes / pns * c.ctas);
    fflush(stdout);
  }
  // cadence of CTA 0 in the last case (ns between consecutive "full" events)
  printf("# cadence CTA0 (ns since first):");
  for (int i = 0; i < 12; ++i) printf(" %lld", h[4096 + i] - h[4096]);
  printf("\n");
  return 0;
}
