// EXPERIMENTAL CTA-pair GEMM (tcgen05 cta_group::2) — opt-in (TDS_GEMM_2CTA=1), compiled for sm_100a but NOT yet run on
// hardware; the validated kernel is gemm_sm100.cu.
//
//   D[m][n] = alpha * sum_k A(m,k) * B(n,k) (+ bias[n])     bf16 operands (K- or MN-major), fp32 accumulate, bf16 output
//
// Two CTAs of a (2,1,1) cluster sit on the two SMs of one TPC and share one 256 x BN UMMA per k-step:
//   * CTA r loads A rows [m0 + 128 r, +128) and B rows [n0 + BN/2 r, +BN/2) into ITS smem (TMA .cta_group::2, the bytes are
//     accounted on the LEADER's full barrier) -> each SM ingests only half of the B tile: the measured limiter of the
//     1-CTA kernel at M = 1024 is per-SM L2->SMEM operand ingest;
//   * the leader's elected thread issues tcgen05.mma.cta_group::2 (M = 256): rows 0-127 accumulate in the leader's TMEM,
//     rows 128-255 in the peer's; tcgen05.commit.cta_group::2 multicasts "stage free" / "accumulator ready" to both CTAs;
//   * each CTA's four epilogue warps drain their own TMEM half through swizzled smem + TMA store and release the
//     accumulator stage on the leader's barrier (8 arrivals).
// Same role layout as gemm_sm100.cu: warp 0 TMA producer, warp 1 MMA issuer + TMEM owner, warps 2-5 epilogue.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"
#include "sm100_ptx_2cta.cuh"
#include "tma_util.h"

namespace tds {
namespace {

constexpr int kRows = 128;          // accumulator rows per CTA (pair tile = 256 rows)
constexpr int kBK = 64;             // one 128-byte swizzle row of bf16
constexpr int kUK = 16;
constexpr int kThreads2 = 192;
constexpr uint32_t kSlabBytes = 4096;

struct Gemm2Dev {
  const __nv_bfloat16* bias;
  float alpha;
  int M, N, K;
  int a_mn, b_mn;
  uint32_t idesc;                   // M = 256, N = BN
};

template <int BN> struct Cfg2 {
  static constexpr int kBH = BN / 2;                       // B rows (n) held by each CTA
  static constexpr int kStages = BN == 256 ? 5 : 8;
  static constexpr int kABytes = kRows * kBK * 2;          // 16 KB
  static constexpr int kBBytes = kBH * kBK * 2;            // 8 / 16 KB
  static constexpr int kSmem = kStages * (kABytes + kBBytes) + 4 * 2 * (int)kSlabBytes + 1024 + 256;
  static constexpr int kTmemCols = BN == 128 ? 256 : 512;  // two accumulator stages of BN columns
};

template <int BN>
__global__ void __launch_bounds__(kThreads2, 1)
gemm2_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
             const __grid_constant__ CUtensorMap tma_d, const __grid_constant__ Gemm2Dev g) {
  using C = Cfg2<BN>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = smem_base;
  const uint32_t sB = smem_base + C::kStages * C::kABytes;
  const uint32_t sStage = sB + C::kStages * C::kBBytes;
  const uint32_t sBar = sStage + 4u * 2u * kSlabBytes;
  auto full_bar = [&](int s) { return sBar + 8u * s; };                           // used on the leader only
  auto empty_bar = [&](int s) { return sBar + 8u * (C::kStages + s); };           // one per CTA
  auto tfull_bar = [&](int s) { return sBar + 8u * (2 * C::kStages + s); };       // one per CTA
  auto tempty_bar = [&](int s) { return sBar + 8u * (2 * C::kStages + 2 + s); };  // used on the leader only
  const uint32_t tmem_slot = sBar + 8u * (2 * C::kStages + 4);
  uint8_t* smem_gen = smem_raw + (smem_base - ptx::smem_u32(smem_raw));
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();        // 0 = leader
  const bool leader = rank == 0;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tma_a);
    ptx::prefetch_tmap(&tma_b);
    ptx::prefetch_tmap(&tma_d);
    for (int s = 0; s < C::kStages; ++s) { ptx::mbar_init(full_bar(s), 1); ptx::mbar_init(empty_bar(s), 1); }
    for (int s = 0; s < 2; ++s) { ptx::mbar_init(tfull_bar(s), 1); ptx::mbar_init(tempty_bar(s), 8); }   // 4 warps x 2 CTAs
    ptx::fence_mbar_init();
  }
  if (warp == 1) {                                     // the same warp of BOTH CTAs: columns reserved in both SMs
    ptx::tmem_alloc_2cta(tmem_slot, C::kTmemCols);
    ptx::tmem_relinquish_2cta();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();                                 // the peer's barriers exist before anything targets them
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  const int m_pairs = (g.M + 2 * kRows - 1) / (2 * kRows);
  const int n_tiles = (g.N + BN - 1) / BN;
  const int total = m_pairs * n_tiles;
  const int nkb = (g.K + kBK - 1) / kBK;
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int t = pair; t < total; t += npairs) {
        const int m0 = (t % m_pairs) * 2 * kRows + (int)rank * kRows;       // this CTA's 128 rows of A / D
        const int n0 = (t / m_pairs) * BN + (int)rank * C::kBH;             // this CTA's half of the B tile
        for (int kb = 0; kb < nkb; ++kb) {
          ptx::mbar_wait(empty_bar(stage), phase ^ 1u);
          if (leader) ptx::mbar_expect_tx(full_bar(stage), 2u * (C::kABytes + C::kBBytes));   // both CTAs' bytes land here
          const uint32_t a_dst = sA + stage * C::kABytes, b_dst = sB + stage * C::kBBytes;
          const int k0 = kb * kBK;
          if (!g.a_mn) {
            ptx::tma_load_4d_2cta(a_dst, &tma_a, full_bar(stage), k0, m0, 0, 0);
          } else {
#pragma unroll
            for (int i = 0; i < kRows / 64; ++i)
              ptx::tma_load_4d_2cta(a_dst + i * (kBK * 128), &tma_a, full_bar(stage), m0 + 64 * i, k0, 0, 0);
          }
          if (!g.b_mn) {
            ptx::tma_load_4d_2cta(b_dst, &tma_b, full_bar(stage), k0, n0, 0, 0);
          } else {
#pragma unroll
            for (int i = 0; i < C::kBH / 64; ++i)
              ptx::tma_load_4d_2cta(b_dst + i * (kBK * 128), &tma_b, full_bar(stage), n0 + 64 * i, k0, 0, 0);
          }
          if (++stage == C::kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (leader) {
      int stage = 0; uint32_t phase = 0;
      int local = 0;
      const uint32_t a_lbo = g.a_mn ? kBK * 128 : 16, b_lbo = g.b_mn ? kBK * 128 : 16;
      const uint32_t a_adv = g.a_mn ? kUK * 128 : kUK * 2, b_adv = g.b_mn ? kUK * 128 : kUK * 2;
      for (int t = pair; t < total; t += npairs) {
        const int as = local & 1;
        const uint32_t aphase = (local >> 1) & 1;
        ++local;
        ptx::mbar_wait(tempty_bar(as), aphase ^ 1u);   // both CTAs' epilogues drained this accumulator stage
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kb = 0; kb < nkb; ++kb) {
          ptx::mbar_wait(full_bar(stage), phase);
          ptx::tc_fence_after();
          if (lane == 0) {
            const uint32_t a_s = sA + stage * C::kABytes, b_s = sB + stage * C::kBBytes;   // same offsets in the peer CTA
#pragma unroll
            for (int k = 0; k < kBK / kUK; ++k) {
              const uint64_t da = ptx::make_smem_desc(a_s + k * a_adv, a_lbo, 1024);
              const uint64_t db = ptx::make_smem_desc(b_s + k * b_adv, b_lbo, 1024);
              ptx::mma_f16_ss_2cta(d_tmem, da, db, g.idesc, (kb > 0 || k > 0) ? 1u : 0u);
            }
            ptx::mma_commit_2cta(empty_bar(stage), 0b11);                  // stage reusable in BOTH CTAs
            if (kb == nkb - 1) ptx::mma_commit_2cta(tfull_bar(as), 0b11);  // accumulator complete in BOTH CTAs
          }
          __syncwarp();
          if (++stage == C::kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else {
    // ===================== epilogue warps (both CTAs, own TMEM half) =====================
    const int q = warp & 3;
    int local = 0;
    uint32_t toggle = 0;
    const uint32_t my_stage0 = sStage + (uint32_t)q * 2u * kSlabBytes;
    for (int t = pair; t < total; t += npairs) {
      const int m0 = (t % m_pairs) * 2 * kRows + (int)rank * kRows;
      const int n0 = (t / m_pairs) * BN;
      const int as = local & 1;
      const uint32_t aphase = (local >> 1) & 1;
      ++local;
      ptx::mbar_wait(tfull_bar(as), aphase);
      ptx::tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + as * BN;
#pragma unroll 1
      for (int slab = 0; slab < BN / 64; ++slab) {
        const int ns = n0 + slab * 64;
        if (ns >= g.N) break;
        const uint32_t dbuf = my_stage0 + toggle * kSlabBytes;
        toggle ^= 1u;
        if (lane == 0) ptx::bulk_wait_read<1>();
        __syncwarp();
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t raw[32];
          ptx::tmem_ld_32x32(t_row + slab * 64 + half * 32, raw);
          ptx::tmem_ld_wait();
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]) * g.alpha;
#pragma unroll
          for (int j8 = 0; j8 < 4; ++j8) {
            const int nc = ns + half * 32 + j8 * 8;
            if (g.bias && nc + 8 <= g.N) {
              float bf[8]; unpack8(ld8(g.bias + nc), bf);
#pragma unroll
              for (int j = 0; j < 8; ++j) v[j8 * 8 + j] += bf[j];
            }
            const uint32_t chunk = (uint32_t)(((half * 4 + j8) ^ (lane & 7)) * 16) + (uint32_t)lane * 128u;
            ptx::st_shared_16(dbuf + chunk, pack8(&v[j8 * 8]));
          }
        }
        ptx::fence_proxy_async();
        __syncwarp();
        if (lane == 0 && m0 + q * 32 < g.M) {
          ptx::tma_store_4d(&tma_d, dbuf, ns, m0 + q * 32, 0, 0);
          ptx::bulk_commit();
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive_leader(tempty_bar(as));    // the leader's MMA warp counts 8 of these
    }
    if (lane == 0) ptx::bulk_wait_read<0>();
    __syncwarp();
  }

  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();            // neither CTA frees TMEM / exits while the other may still signal it
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc_2cta(tmem_base, C::kTmemCols);
  }
}

template <int BN>
bool launch2(const GemmParams& p, cudaStream_t s, int num_sms) {
  CUtensorMap ta, tb, td;
  if (!make_map(&ta, p.a, p.M, p.K, 1, 1, kRows, kBK) || !make_map(&tb, p.b, p.N, p.K, 1, 1, Cfg2<BN>::kBH, kBK)) return false;
  GemmOperand od{p.d, p.ldd, 0, 0, false};
  if (!make_map(&td, od, p.M, p.N, 1, 1, 32)) return false;
  static bool attr_done = false;
  if (!attr_done) {
    if (cudaFuncSetAttribute(gemm2_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg2<BN>::kSmem) != cudaSuccess) {
      cudaGetLastError();
      return false;
    }
    attr_done = true;
  }
  Gemm2Dev g;
  g.bias = reinterpret_cast<const __nv_bfloat16*>(p.bias);
  g.alpha = p.alpha; g.M = p.M; g.N = p.N; g.K = p.K;
  g.a_mn = p.a.mn_major; g.b_mn = p.b.mn_major;
  g.idesc = make_idesc_bf16(2 * kRows, BN, p.a.mn_major, p.b.mn_major);
  const int tiles = ((p.M + 2 * kRows - 1) / (2 * kRows)) * ((p.N + BN - 1) / BN);
  int pairs = num_sms / 2;
  if (pairs > tiles) pairs = tiles;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3(kThreads2);
  cfg.dynamicSmemBytes = Cfg2<BN>::kSmem;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, gemm2_kernel<BN>, ta, tb, td, g) == cudaSuccess;
}

}  // namespace

// Returns false (nothing launched) when the problem is outside what the pair kernel covers; the caller then uses gemm_bf16.
bool gemm2_bf16(const GemmParams& p, cudaStream_t stream) {
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  const bool aligned = (p.ldd % 8 == 0) && ((reinterpret_cast<uintptr_t>(p.d) & 15) == 0) && (p.N % 8 == 0);
  if (p.in_dtype != kBF16 || p.d_dtype != kBF16 || p.io_dtype != kBF16 || p.batch != 1 || p.tri != 0 || p.accumulate ||
      p.reduce_out || p.aux != nullptr || !aligned || p.M < 2 * kRows || p.K <= 0)
    return false;
  const long long mp = (p.M + 2 * kRows - 1) / (2 * kRows);
  const long long pairs256 = mp * ((p.N + 255) / 256);
  // the widest tile that still gives (nearly) every SM pair work
  const int bn = (p.config == 2 || (p.config < 0 && pairs256 * 2 * 10 >= (long long)num_sms * 6)) ? 256 : 128;
  return bn == 256 ? launch2<256>(p, stream, num_sms) : launch2<128>(p, stream, num_sms);
}

}  // namespace tds
