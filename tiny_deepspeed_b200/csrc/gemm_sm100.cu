// Persistent warp-specialised tcgen05 GEMM for sm_100a (bf16 x bf16 -> fp32 in TMEM -> bf16/fp32, and fp32 x fp32 consumed as
// TF32 by kind::tf32 -> fp32 for fp32 models: same 128-byte-row smem layout, 32 instead of 64 K elements per stage).
//
//   D[b][m][n] = alpha * sum_k A(b,m,k) * B(b,n,k)      A: [M,K] K-major or [K,M] MN-major, same for B
//
// One kernel serves every GEMM of the GPT-2 step: Linear fwd (K-major/K-major), dX
// (B = W consumed MN-major, no transpose copy), dW (both MN-major), and the batched attention products on
// strided head views of the packed qkv buffer (4-D TMA maps: inner, row, head, batch).
//
// Structure (one CTA per SM, 192 threads):
//   warp 0      TMA producer: cp.async.bulk.tensor 128B-swizzled boxes -> smem ring, mbarrier expect_tx
//   warp 1      MMA issuer: one lane issues tcgen05.mma (UMMA 128 x BN x 16), tcgen05.commit frees smem
//               stages and publishes the accumulator; also owns the TMEM allocation (2 x BN columns)
//   warps 2-5   epilogue: tcgen05.ld 32x32b (thread == accumulator row), fused bias / GELU / GELU' /
//               residual, bf16 pack -> 128B-swizzled smem slab -> TMA store (coalesced 128-byte rows; the
//               first version stored 16 B per thread per row and spent half the kernel in partial-sector
//               writes, profiles/r1_ncu_summary.md); fp32 / accumulating outputs use direct stores.
//               Overlaps the next tile's MMAs through the double-buffered TMEM accumulator.
//   optional    thread-block cluster of cm CTAs = cm consecutive M tiles: each loads 1/cm of the B tile and
//               TMA-multicasts it to all (UTMALDG.MULTICAST), stage release by multicast tcgen05.commit.
// Ragged edges come for free: TMA zero-fills out-of-bounds loads and clips out-of-bounds stores.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <mutex>

#include "common.cuh"
#include "kernels.h"
#include "sm100_ptx.cuh"
#include "tma_util.h"

namespace tds {

constexpr int BM = 128;       // UMMA M (cta_group::1)
constexpr int BK = 64;        // one 128-byte swizzle row of bf16
constexpr int UK = 16;        // UMMA K for 16-bit inputs
constexpr int kThreads = 192;        // single pipeline: producer, issuer, 4 epilogue warps
// Epilogue warps: 4 (one per TMEM lane quadrant), or 8 on the FAST (bf16 TMA-store) kernels — two warps per quadrant, each
// taking 32 of a slab's 64 columns.  At one tile per CTA the epilogue is pure critical path, and with one warp per scheduler
// it ran at ~1/3 instruction per cycle (profiles/r2_gemm_epitrace_before.log: 1900 cycles for a 128 x 64 tile).
constexpr int epi_warps(int fast) { return fast == 2 ? 8 : 4; }
constexpr int gemm_threads(int np, int fast) { return 32 * (2 * np + epi_warps(fast)); }
constexpr uint32_t kStageBufBytes = 4096;   // one 32-row x 64-col bf16 slab, SWIZZLE_128B

enum { EPI_NONE = 0, EPI_GELU_SAVE = 1, EPI_GELU_BWD = 2, EPI_RESIDUAL = 3 };

struct GemmDev {
  void* d; int d_f32; long long ldd, dbs1, dbs2;
  const void* bias;       // [N], bf16 (fp32 when io_f32)
  void* aux; long long ld_aux;   // [M,N], bf16 (fp32 when io_f32)
  int io_f32;     // bias / aux are fp32 (fp32 model)
  int epi, accumulate; float alpha;
  int M, N, K, batch, nb2;
  int a_mn, b_mn;
  int tri;
  int cm;         // cluster size along M (1, 2, 4, 8): B tile fetched once per cluster and multicast
  int aux_tma;    // aux operand (residual / GELU pre-activation) is read through a tensor map
  int tma_store;  // epilogue goes registers -> swizzled smem -> TMA store (bf16 out, no accumulate, aligned)
  int dbg;      // TDS_GEMM_DBG bits (profiling only): 1 = no global stores, 2 = no MMA issue, 4 = no epilogue body
  uint32_t idesc;
  long long* prof;   // per-CTA phase timestamps (16 x int64 per CTA), nullptr outside tools/gemm_timeline.py
  const char* pf; long long pf_bytes;   // L2 prefetch hint for the next kernel's operand (GemmParams::prefetch)
};

// Phase stamps / debug bits exist only in builds with -DTDS_GEMM_PROF (tools/build_harness.sh): the production kernels carry
// none of that code — at M = 1024 the prologue and epilogue run once per CTA out of a cold instruction cache, so every
// kilobyte of dead branches costs cycles (profiles/r2_gemm_issue_path.md).
#ifdef TDS_GEMM_PROF
constexpr bool kProf = true;
#else
constexpr bool kProf = false;
#endif
__device__ __forceinline__ void prof_stamp(long long* prof, int slot) {
  if (prof) prof[(long long)blockIdx.x * 16 + slot] = clock64();
}
__device__ __forceinline__ long long globaltimer_ns() {
  long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t;
}

template <int BN> struct Cfg {
  static constexpr int kStages = BN == 256 ? 4 : (BN == 192 ? 4 : (BN == 128 ? 6 : 8));
  static constexpr int kABytes = BM * BK * 2;   // 16 KB
  static constexpr int kBBytes = BN * BK * 2;
  // epilogue staging: kEpiBufs 4 KB slabs per epilogue warp.  BN = 192 has the room for 4 (its 3 slabs of a residual / GELU'
  // aux tile are all prefetched during the main loop; GELU+save alternates two (output, pre-activation) pairs)
  static constexpr int kEpiBufs = BN == 192 ? 4 : 2;
  static constexpr int kSmem = kStages * (kABytes + kBBytes) + 4 * kEpiBufs * 4096 /*epilogue staging*/ + 1024 /*align*/ + 512 /*barriers*/;
  static constexpr int kTmemCols = BN == 64 ? 128 : (BN == 128 ? 256 : 512);   // power of two >= 2 * BN
  static constexpr int kTmemColsDual = BN == 64 ? 256 : 512;                   // 2 accumulator stages x 2 half-K accumulators
};

template <int KB>
__device__ __forceinline__ void tile_k_range(const GemmDev& g, int m0, int nkb, int& kb0, int& kb1) {
  kb0 = 0; kb1 = nkb;
  if (g.tri == 2) { int e = (m0 + BM + KB - 1) / KB; kb1 = e < nkb ? e : nkb; }
  else if (g.tri == 3) { kb0 = m0 / KB; }
}

// 8 consecutive elements of a bf16 or fp32 array as floats / back (epilogue operands of either model dtype)
__device__ __forceinline__ void ld8_any(const void* base, long long idx, int f32, float (&out)[8]) {
  if (f32) {
    const float4* p = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + idx);
    const float4 a = p[0], b = p[1];
    out[0] = a.x; out[1] = a.y; out[2] = a.z; out[3] = a.w; out[4] = b.x; out[5] = b.y; out[6] = b.z; out[7] = b.w;
  } else {
    unpack8(ld8(reinterpret_cast<const __nv_bfloat16*>(base) + idx), out);
  }
}
__device__ __forceinline__ float ld1_any(const void* base, long long idx, int f32) {
  return f32 ? reinterpret_cast<const float*>(base)[idx] : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(base)[idx]);
}

// F32 = fp32 operands (kind::tf32): compile-time so the bf16 instantiation keeps its fully unrolled producer / issue loops
// RED = reduce-scatter epilogue (EXPERIMENTAL, opt-in, see ops.gemm(reduce_out=True)): the fp32 tile is added into the
//       destination by the TMA (cp.reduce.async.bulk.tensor .add) instead of stored; the destination may be ANOTHER rank's
//       copy of a symmetric gradient shard, so the dW GEMM of every rank accumulates straight into the owner over NVLink,
//       tile by tile, while the GEMM is still running.  Compile-time so the default instantiations stay byte-identical.
// DUAL = two independent half-K pipelines inside the CTA (BN <= 128 only): (producer warp 0, issuer warp 1) stream the even
//       k-blocks through the lower half of the smem ring into accumulator D0, (warp 2, warp 3) the odd k-blocks through the
//       upper half into D1, and the epilogue adds D0 + D1.  Measured with tools/mma_probe.cu: ONE thread can issue a
//       tcgen05.mma only every ~80-100 cycles whatever its shape, i.e. 128 x 64/128 x 16 MMAs (32/64 cycles of tensor-pipe
//       work) leave the pipe idle most of the time; two issuing threads on disjoint accumulators reach the pipe's own rate
//       (49 / 65 cycles per MMA for N = 64 / 128).  The same holds for the single TMA-issuing thread (~340 cycles per k-block).
// EPI  = epilogue functor fixed at compile time (EPI_*), or -1: read g.epi at run time (generic instantiations)
// FAST = the output goes registers -> swizzled smem -> TMA store and nothing else is compiled in (bf16 out, aligned, no
//        accumulate: every GEMM of the bf16 training step); the generic instantiations keep the direct-store paths
template <int BN, bool F32, bool RED, int NP, int EPI, int FAST>   // FAST: 0 generic, 1 = bf16 TMA-store path, 2 = the same with 8 epilogue warps
__global__ void __launch_bounds__(gemm_threads(NP, FAST), 1)
gemm_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
            const __grid_constant__ CUtensorMap tma_d, const __grid_constant__ CUtensorMap tma_aux,
            const __grid_constant__ GemmDev g) {
  using C = Cfg<BN>;
  long long* const prof = kProf ? g.prof : nullptr;     // compile-time nullptr in production builds
  const int dbg = kProf ? g.dbg : 0;
  const int epi = EPI >= 0 ? EPI : g.epi;
  constexpr bool DUAL = NP > 1;
  constexpr int kHalves = NP;                           // k-interleaved pipelines ("halves" when NP = 2)
  static_assert(2 * NP * BN <= 512, "two accumulator stages of NP x BN fp32 columns must fit TMEM");
  constexpr int kSH = C::kStages / kHalves;            // ring stages per pipeline
  constexpr int kAccCols = kHalves * BN;               // TMEM columns of one accumulator stage
  constexpr uint32_t kTmemAlloc = 2 * kAccCols <= 128 ? 128u : (2 * kAccCols <= 256 ? 256u : 512u);   // power of two
  constexpr int kEpiWarp0 = 2 * NP;                    // first epilogue warp (4 consecutive warps cover the 4 TMEM lane quadrants)
  constexpr int kBK = F32 ? 32 : BK;     // K elements per stage = one 128-byte row
  constexpr int kGrp = F32 ? 32 : 64;    // MN elements per 128-byte row of an MN-major operand
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = smem_base;
  const uint32_t sB = smem_base + C::kStages * C::kABytes;
  constexpr int NB = C::kEpiBufs;
  const uint32_t sStage = sB + C::kStages * C::kBBytes;              // 4 warps x NB x (32 rows x 128 B), 1024-aligned
  const uint32_t sBar = sStage + 4u * (uint32_t)NB * kStageBufBytes;
  // barrier layout (8 B each): full[kStages], empty[kStages], tmem_full[2], tmem_empty[2], then tmem ptr
  auto full_bar = [&](int s) { return sBar + 8u * s; };
  auto empty_bar = [&](int s) { return sBar + 8u * (C::kStages + s); };
  auto tfull_bar = [&](int s) { return sBar + 8u * (2 * C::kStages + s); };
  auto tempty_bar = [&](int s) { return sBar + 8u * (2 * C::kStages + 2 + s); };
  const uint32_t tmem_slot = sBar + 8u * (2 * C::kStages + 4);
  auto aux_bar = [&](int w, int i) { return sBar + 8u * (2 * C::kStages + 6 + w * 4 + i); };   // per epilogue warp and staging buffer: aux slab landed
  uint8_t* smem_gen = smem_raw + (smem_base - ptx::smem_u32(smem_raw));
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));

  // shfl result = provably warp-uniform: role branches below are uniform branches and per-role state lives in uniform registers
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  pdl_launch();   // the next kernel may start its own prologue as soon as every CTA of this grid got here
  if (prof && threadIdx.x == 0) {
    uint32_t smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    prof[(long long)blockIdx.x * 16 + 0] = globaltimer_ns();
    prof[(long long)blockIdx.x * 16 + 12] = smid;
    prof_stamp(prof, 1);
  }

  if (warp == 0 && ptx::elect_one()) {
    ptx::prefetch_tmap(&tma_a);
    ptx::prefetch_tmap(&tma_b);
    if (FAST || g.tma_store) ptx::prefetch_tmap(&tma_d);
    // a smem stage is refilled by EVERY CTA of the cluster (B slices are multicast), so it is free only when all
    // cm MMA issuers have released it
    for (int s = 0; s < C::kStages; ++s) { ptx::mbar_init(full_bar(s), 1); ptx::mbar_init(empty_bar(s), (uint32_t)g.cm); }
    for (int s = 0; s < 2; ++s) { ptx::mbar_init(tfull_bar(s), kHalves); ptx::mbar_init(tempty_bar(s), (uint32_t)epi_warps(FAST)); }
    for (int w = 0; w < 4; ++w)
      for (int i = 0; i < NB; ++i) ptx::mbar_init(aux_bar(w, i), 1);
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, kTmemAlloc);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  const uint32_t cta_rank = g.cm > 1 ? ptx::cluster_ctarank() : 0u;
  const uint16_t cta_mask = (uint16_t)((1u << g.cm) - 1u);
  if (g.cm > 1) ptx::cluster_sync();   // peers' barriers exist before any multicast / remote commit targets them
  pdl_wait();     // everything above overlapped the previous kernel's tail; from here on we touch its outputs
  if (threadIdx.x == 0) prof_stamp(prof, 2);

  const int m_tiles = (g.M + BM - 1) / BM;
  const int n_tiles = (g.N + BN - 1) / BN;
  const int tiles_per_batch = m_tiles * n_tiles;
  const int total_tiles = tiles_per_batch * g.batch;
  const int nkb = (g.K + kBK - 1) / kBK;

  // half-K pipeline this warp belongs to (DUAL): warps 0/1 = half 0, warps 2/3 = half 1
  // warps 2p / 2p+1 = (TMA producer, MMA issuer) of pipeline p, which owns the k-blocks kb0 + p, kb0 + p + NP, ...
  const int half = warp < 2 * NP ? (warp >> 1) : 0;
  const bool is_producer = warp < 2 * NP && (warp & 1) == 0;
  const bool is_issuer = warp < 2 * NP && (warp & 1) == 1;
  const int s0 = half * kSH;                           // this pipeline's slice of the smem ring

  if (is_producer) {
    // ===================== TMA producer (one elected thread runs the whole loop) =====================
    if (ptx::elect_one()) {
      int stage = 0; uint32_t phase = 0;
      bool first = true;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const int b = t / tiles_per_batch, r = t % tiles_per_batch;
        const int m0 = (r % m_tiles) * BM, n0 = (r / m_tiles) * BN;   // m fastest: neighbours share the B tile in L2
        if (g.tri == 1 && n0 > m0 + BM - 1) continue;
        const int b1 = b / g.nb2, b2 = b % g.nb2;
        int kb0, kb1; tile_k_range<kBK>(g, m0, nkb, kb0, kb1);
        for (int kb = kb0 + half; kb < kb1; kb += kHalves) {
          const int sg = s0 + stage;
          ptx::mbar_wait_conv(empty_bar(sg), phase ^ 1u);
          ptx::mbar_expect_tx(full_bar(sg), C::kABytes + C::kBBytes);
          const uint32_t a_dst = sA + sg * C::kABytes, b_dst = sB + sg * C::kBBytes;
          const int k0 = kb * kBK;
          // MN-major operands arrive as one box per 128-byte-wide MN group (64 bf16 / 32 fp32), bk k-rows each
          constexpr uint32_t grp_bytes = (uint32_t)kBK * 128u;
          if (!g.a_mn) {
            ptx::tma_load_4d(a_dst, &tma_a, full_bar(sg), k0, m0, b2, b1);
          } else {
#pragma unroll
            for (int i = 0; i < BM / kGrp; ++i)
              ptx::tma_load_4d(a_dst + i * grp_bytes, &tma_a, full_bar(sg), m0 + kGrp * i, k0, b2, b1);
          }
          if (g.cm == 1) {
            if (!g.b_mn) {
              ptx::tma_load_4d(b_dst, &tma_b, full_bar(sg), k0, n0, b2, b1);
            } else {
#pragma unroll
              for (int i = 0; i < BN / kGrp; ++i)
                ptx::tma_load_4d(b_dst + i * grp_bytes, &tma_b, full_bar(sg), n0 + kGrp * i, k0, b2, b1);
            }
          } else {   // (bf16 only: pick_cluster never forms clusters for fp32 operands)
            // The cm CTAs of the cluster work on cm consecutive M tiles of the SAME N tile: each fetches 1/cm of
            // the B tile and the TMA multicasts it into every CTA's smem (one L2 / NVLink read per cluster).
            if (!g.b_mn) {
              const uint32_t rows = BN / g.cm;
              ptx::tma_load_4d_mc(b_dst + cta_rank * rows * 128u, &tma_b, full_bar(sg), k0, n0 + (int)(cta_rank * rows),
                                  b2, b1, cta_mask);
            } else {
              const uint32_t krows = BK / g.cm;
#pragma unroll
              for (int i = 0; i < BN / 64; ++i)
                ptx::tma_load_4d_mc(b_dst + i * (BK * 128) + cta_rank * krows * 128u, &tma_b, full_bar(sg),
                                    n0 + 64 * i, k0 + (int)(cta_rank * krows), b2, b1, cta_mask);
            }
          }
          if (prof && half == 0) { if (first) { prof_stamp(prof, 3); first = false; } prof_stamp(prof, 4); }
          if (++stage == kSH) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (is_issuer) {
    // ===================== MMA issuer (converged warp; one elected lane issues) =====================
    int stage = 0; uint32_t phase = 0;
    int local = 0;
    bool first = true;
    // descriptor strides: K-major: SBO = 1024 (8 rows x 128 B), per-UMMA_K advance 32 B;
    //                     MN-major: LBO = BK*128 (next 64-wide MN group), SBO = 1024, advance 16 k-rows = 2048 B
    //                     (fp32/TF32: UMMA K = 8 -> 32 B per k-step K-major as well, 8 k-rows = 1024 B MN-major)
    constexpr uint32_t uk_rows = F32 ? 8u : (uint32_t)UK;
    const uint32_t a_lbo = g.a_mn ? kBK * 128 : 16, b_lbo = g.b_mn ? kBK * 128 : 16;
    const uint32_t a_adv = g.a_mn ? uk_rows * 128 : 32, b_adv = g.b_mn ? uk_rows * 128 : 32;
    // MN-major fp32 operands sit in the 32-byte-atom swizzle (the only MN-major layout kind::tf32 accepts)
    const uint32_t a_lt = (F32 && g.a_mn) ? 1u : 2u, b_lt = (F32 && g.b_mn) ? 1u : 2u;
    const uint32_t a_sbo = a_lt == 1u ? 512u : 1024u, b_sbo = b_lt == 1u ? 512u : 1024u;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      const int r = t % tiles_per_batch;
      const int m0 = (r % m_tiles) * BM, n0 = (r / m_tiles) * BN;
      if (g.tri == 1 && n0 > m0 + BM - 1) continue;
      int kb0, kb1; tile_k_range<kBK>(g, m0, nkb, kb0, kb1);
      const int as = local & 1;
      const uint32_t aphase = (local >> 1) & 1;
      ++local;
      ptx::mbar_wait_conv(tempty_bar(as), aphase ^ 1u);
      ptx::tc_fence_after();
      const uint32_t d_tmem = tmem_base + as * kAccCols + half * BN;
      if (DUAL && kb0 + half >= kb1) {
        // this pipeline has no k-block in the tile (a one-k-block causal range): the epilogue still expects both arrivals
        if (ptx::elect_one()) ptx::mbar_arrive(tfull_bar(as));
        __syncwarp();
        continue;
      }
      for (int kb = kb0 + half; kb < kb1; kb += kHalves) {
        const int sg = s0 + stage;
        // fine-grained trace of the issuer loop (CTA 0, its first tile, this pipeline's k-blocks 2..5)
        const int it = (kb - kb0) / kHalves;
        const bool trace = prof && blockIdx.x == 0 && t == 0 && half == 0 && it >= 2 && it < 6;
        long long* tr = trace ? prof + 148 * 16 + (it - 2) * 8 : nullptr;
        if (trace && lane == 0) tr[0] = clock64();
        ptx::mbar_wait_conv(full_bar(sg), phase);         // all lanes poll: the warp stays converged
        ptx::tc_fence_after();
        if (trace && lane == 0) tr[1] = clock64();
        if (ptx::elect_one()) {
          const uint32_t a_s = sA + sg * C::kABytes, b_s = sB + sg * C::kBBytes;
#pragma unroll
          for (int k = 0; k < BK / UK; ++k) {
            const uint64_t da = ptx::make_smem_desc(a_s + k * a_adv, a_lbo, a_sbo, a_lt);
            const uint64_t db = ptx::make_smem_desc(b_s + k * b_adv, b_lbo, b_sbo, b_lt);
            const uint32_t acc = (kb > kb0 + half || k > 0) ? 1u : 0u;
            if (F32) ptx::mma_tf32_ss(d_tmem, da, db, g.idesc, acc);
            else if (!(dbg & 2)) ptx::mma_f16_ss(d_tmem, da, db, g.idesc, acc);
          }
          // smem stage reusable once these MMAs retire (told to every CTA of the cluster when B is multicast)
          if (g.cm == 1) ptx::mma_commit(empty_bar(sg));
          else ptx::mma_commit_mc(empty_bar(sg), cta_mask);
          if (kb + kHalves >= kb1) ptx::mma_commit(tfull_bar(as));   // this pipeline's accumulator complete -> epilogue
        }
        __syncwarp();
        if (prof && lane == 0 && half == 0) {
          if (first) { prof_stamp(prof, 5); first = false; }
          prof_stamp(prof, 6);
          if (trace) tr[2] = clock64();
        }
        if (++stage == kSH) { stage = 0; phase ^= 1u; }
      }
    }
  } else {
    // ===================== epilogue warps =====================
    const int q = warp & 3;                      // TMEM lane quadrant this warp may access
    constexpr int EW = epi_warps(FAST);
    constexpr bool kSplit = EW == 8;              // two warps per quadrant: this one takes column half `hf` of every 64-column slab
    const int hf = kSplit ? ((warp - kEpiWarp0) >> 2) : 0;
    auto pair_sync = [&]() { asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory"); };
    int local = 0;
    uint32_t sbuf_toggle = 0;
    const uint32_t my_stage0 = sStage + (uint32_t)q * (uint32_t)NB * kStageBufBytes;   // NB 4 KB staging buffers per warp
    const bool gelu_save = epi == EPI_GELU_SAVE;
    // aux tile (residual / pre-activation) arrives by TMA: always on the FAST kernels (the host only picks them when the aux
    // tensor map exists), a runtime property on the generic ones
    const bool aux_in = (FAST || g.aux_tma) && (epi == EPI_GELU_BWD || epi == EPI_RESIDUAL);
    uint32_t aux_phase = 0;                       // bit i: parity of aux_bar(q, i)
    if (g.pf) {
      // these four warps have nothing to do until the first accumulator is complete: spread the next kernel's weight
      // (or saved activation) over all CTAs in 4 KB requests and pull it from HBM into L2 underneath this main loop
      constexpr long long kChunk = 4096;
      const long long nchunk = (g.pf_bytes + kChunk - 1) / kChunk;
      const int e = (warp - kEpiWarp0) * 32 + lane;
      for (long long c = (long long)blockIdx.x + (long long)gridDim.x * e; c < nchunk; c += (long long)gridDim.x * (EW * 32)) {
        const long long left = g.pf_bytes - c * kChunk;
        ptx::prefetch_l2_bulk(g.pf + c * kChunk, (uint32_t)(left < kChunk ? left : kChunk));
      }
    }
    const bool vec_ok = (g.N % 8 == 0) && (g.ldd % 8 == 0) && ((reinterpret_cast<uintptr_t>(g.d) & 15) == 0) &&
                        (g.aux == nullptr || (g.ld_aux % 8 == 0 && (reinterpret_cast<uintptr_t>(g.aux) & 15) == 0));
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      const int b = t / tiles_per_batch, r = t % tiles_per_batch;
      const int m0 = (r % m_tiles) * BM, n0 = (r / m_tiles) * BN;
      if (g.tri == 1 && n0 > m0 + BM - 1) continue;
      const int b1 = b / g.nb2, b2 = b % g.nb2;
      const int as = local & 1;
      const uint32_t aphase = (local >> 1) & 1;
      ++local;
      // residual / pre-activation slabs (32 rows x 64 cols each) of this tile: fetched by TMA into the staging buffers NOW, while
      // the main loop is still running, instead of one exposed load latency per slab after it (slab s -> buffer s % NB;
      // the result is written back in place and TMA-stored from the same buffer)
      int n_slabs = 0;
      if ((FAST || g.tma_store) && !RED) {
        n_slabs = (g.N - n0 + 63) / 64;
        if (n_slabs > BN / 64) n_slabs = BN / 64;
        if (dbg & 4) n_slabs = 0;
      }
      if (aux_in && lane == 0 && hf == 0) {
        ptx::bulk_wait_read<0>();                 // the previous tile's stores have left the buffers
        for (int sl = 0; sl < n_slabs && sl < NB; ++sl) {
          ptx::mbar_expect_tx(aux_bar(q, sl), kStageBufBytes);
          ptx::tma_load_4d(my_stage0 + (uint32_t)sl * kStageBufBytes, &tma_aux, aux_bar(q, sl), n0 + sl * 64, m0 + q * 32, 0, 0);
        }
      }
      ptx::mbar_wait(tfull_bar(as), aphase);
      ptx::tc_fence_after();
      if (prof && warp == kEpiWarp0 && lane == 0) { if (t == (int)blockIdx.x) prof_stamp(prof, 8); prof_stamp(prof, 9); }
      const int m = m0 + q * 32 + lane;
      const bool row_ok = m < g.M;
      const long long d_off = (long long)b1 * g.dbs1 + (long long)b2 * g.dbs2 + (long long)m * g.ldd;
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + as * kAccCols;
      // NP > 1: D = sum of the pipelines' accumulators; pipeline p is untouched when the tile's K range has <= p k-blocks
      int npipe = 1;
      if (DUAL) { int kb0, kb1; tile_k_range<kBK>(g, m0, nkb, kb0, kb1); npipe = kb1 - kb0 < NP ? kb1 - kb0 : NP; }
      auto ld_acc = [&](uint32_t col, uint32_t (&raw)[32]) {
        ptx::tmem_ld_32x32(t_row + col, raw);
        if (DUAL && npipe > 1) {
#pragma unroll
          for (int p = 1; p < NP; ++p) {
            if (p < npipe) {
              uint32_t hi[32];
              ptx::tmem_ld_32x32(t_row + p * BN + col, hi);
              ptx::tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 32; ++j) raw[j] = __float_as_uint(__uint_as_float(raw[j]) + __uint_as_float(hi[j]));
            }
          }
        } else {
          ptx::tmem_ld_wait();
        }
      };
      if constexpr (RED) {
        // ---- reduce path: 32 x 32 fp32 slabs (128-byte rows, SWIZZLE_128B) -> TMA reduce-add into (peer) global memory ----
#pragma unroll 1
        for (int slab = 0; slab < BN / 32; ++slab) {
          const int ns = n0 + slab * 32;
          if (ns >= g.N) break;
          const uint32_t dbuf = my_stage0 + sbuf_toggle * kStageBufBytes;
          sbuf_toggle ^= 1u;
          if (lane == 0) ptx::bulk_wait_read<1>();
          __syncwarp();
          uint32_t raw[32];
          ld_acc(slab * 32, raw);
#pragma unroll
          for (int c16 = 0; c16 < 8; ++c16) {
            const float4 v = make_float4(__uint_as_float(raw[c16 * 4]) * g.alpha, __uint_as_float(raw[c16 * 4 + 1]) * g.alpha,
                                         __uint_as_float(raw[c16 * 4 + 2]) * g.alpha, __uint_as_float(raw[c16 * 4 + 3]) * g.alpha);
            ptx::st_shared_16(dbuf + (uint32_t)lane * 128u + ((((uint32_t)c16) ^ (uint32_t)(lane & 7)) << 4), v);
          }
          ptx::fence_proxy_async();
          __syncwarp();
          if (lane == 0 && m0 + q * 32 < g.M) {     // rows / columns past M, N are clipped by the tensor map
            ptx::tma_reduce_add_2d(&tma_d, dbuf, ns, m0 + q * 32);
            ptx::bulk_commit();
          }
        }
      } else if (FAST || g.tma_store) {
        // ---- coalesced path: registers -> 128B-swizzled smem slab (32 rows x 64 cols) -> TMA store ------------------
#pragma unroll 1
        for (int slab = 0; slab < n_slabs; ++slab) {
          const int ns = n0 + slab * 64;
          uint32_t dbuf, abuf = 0;
          // epilogue trace (tools/gemm_harness trace): first tile of CTA 0, first epilogue warp, slabs 0..1
          const bool etr = prof && blockIdx.x == 0 && t == 0 && warp == kEpiWarp0 && lane == 0 && slab < 2;
          long long* et = etr ? prof + 148 * 16 + 128 + slab * 8 : nullptr;
          if (etr) et[0] = clock64();
          if (aux_in) {
            const int bi = slab % NB;
            dbuf = abuf = my_stage0 + (uint32_t)bi * kStageBufBytes;
            if (etr) et[1] = clock64();
            ptx::mbar_wait(aux_bar(q, bi), (aux_phase >> bi) & 1u);
            aux_phase ^= 1u << bi;
          } else if (gelu_save) {
            // (output, pre-activation) pairs; one bulk group per slab holds both stores
            constexpr int kPairs = NB / 2;
            dbuf = my_stage0 + sbuf_toggle * 2u * kStageBufBytes;
            abuf = dbuf + kStageBufBytes;
            if (++sbuf_toggle == (uint32_t)kPairs) sbuf_toggle = 0;
            if (lane == 0) ptx::bulk_wait_read<kPairs - 1>();      // only the storing warp (hf = 0) has groups in flight
            if (kSplit) pair_sync(); else __syncwarp();
            if (etr) et[1] = clock64();
          } else {
            dbuf = my_stage0 + sbuf_toggle * kStageBufBytes;
            if (++sbuf_toggle == (uint32_t)NB) sbuf_toggle = 0;
            if (lane == 0) ptx::bulk_wait_read<NB - 1>();
            if (kSplit) pair_sync(); else __syncwarp();
            if (etr) et[1] = clock64();
          }
          // bias of the slab's 64 columns: 8 x 16 B at the same address in every lane (one broadcast transaction each), issued
          // BEFORE the accumulator loads so their latency hides under them.  Columns past N are clipped by the TMA store,
          // so the address is only clamped; no bias = add zeros.  (The previous form — a guarded 4 x 32-bit load per 8
          // columns, then a branch per operand source — serialised eight load latencies per slab: profiles/r2_gemm_epilogue.md)
          auto do_half = [&](const int half) {     // 32 of the slab's 64 columns (`half` is a compile-time constant at each call)
            uint4 braw[4];
#pragma unroll
            for (int j8 = 0; j8 < 4; ++j8) braw[j8] = make_uint4(0u, 0u, 0u, 0u);
            if (g.bias) {
#pragma unroll
              for (int j8 = 0; j8 < 4; ++j8) {
                int nc = ns + half * 32 + j8 * 8;
                nc = nc > g.N - 8 ? g.N - 8 : nc;
                braw[j8] = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(g.bias) + nc));
              }
            }
            // this thread's 4 x 16 B of the residual / pre-activation slab (same swizzled chunks the result goes back to)
            uint4 ax[4];
            if (aux_in) {
#pragma unroll
              for (int j8 = 0; j8 < 4; ++j8)
                ax[j8] = ptx::ld_shared_16<uint4>(abuf + (uint32_t)(((half * 4 + j8) ^ (lane & 7)) * 16) + (uint32_t)lane * 128u);
            }
            uint32_t raw[32];
            ld_acc(slab * 64 + half * 32, raw);
            if (etr) et[2 + half * 2] = clock64();
            const int nb = ns + half * 32;
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]) * g.alpha;
#pragma unroll
            for (int j8 = 0; j8 < 4; ++j8) {
              const int nc = nb + j8 * 8;
              {
                float bf[8]; unpack8(*reinterpret_cast<const bf16x8*>(&braw[j8]), bf);
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j8 * 8 + j] += bf[j];
              }
              const uint32_t chunk = (uint32_t)(((half * 4 + j8) ^ (lane & 7)) * 16) + (uint32_t)lane * 128u;
              if (gelu_save) {
                bf16x8 pk = pack8(&v[j8 * 8]);
                ptx::st_shared_16(abuf + chunk, pk);
                float af[8]; unpack8(pk, af);
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j8 * 8 + j] = gelu_tanh(af[j]);
              } else if (epi != EPI_NONE) {
                float af[8];
                if (aux_in) {
                  unpack8(*reinterpret_cast<const bf16x8*>(&ax[j8]), af);                  // OOB rows/cols were zero-filled
                } else {
#pragma unroll
                  for (int j = 0; j < 8; ++j) af[j] = 0.f;
                  if (row_ok && nc + 8 <= g.N) unpack8(ld8(reinterpret_cast<const __nv_bfloat16*>(g.aux) + (long long)m * g.ld_aux + nc), af);
                }
#pragma unroll
                for (int j = 0; j < 8; ++j)
                  v[j8 * 8 + j] = epi == EPI_GELU_BWD ? v[j8 * 8 + j] * gelu_tanh_grad(af[j]) : v[j8 * 8 + j] + af[j];
              }
              ptx::st_shared_16(dbuf + chunk, pack8(&v[j8 * 8]));
            }
            if (etr) et[3 + half * 2] = clock64();
          };
          if (kSplit) {
            if (hf == 0) do_half(0); else do_half(1);
          } else {
            do_half(0);
            do_half(1);
          }
          ptx::fence_proxy_async();     // generic-proxy smem writes -> visible to the TMA (async proxy)
          if (kSplit) pair_sync(); else __syncwarp();     // both column halves of the slab are in the staging buffer
          if (etr) et[6] = clock64();
          if (lane == 0 && hf == 0 && m0 + q * 32 < g.M && !(dbg & 1)) {
            ptx::tma_store_4d(&tma_d, dbuf, ns, m0 + q * 32, b2, b1);
            if (gelu_save) ptx::tma_store_4d(&tma_aux, abuf, ns, m0 + q * 32, 0, 0);
            ptx::bulk_commit();
          }
          if (aux_in && slab + NB < n_slabs && lane == 0 && hf == 0) {
            // more slabs than buffers (BN = 256, or the runtime-epilogue kernels): recycle this buffer for slab + NB
            ptx::bulk_wait_read<0>();
            ptx::mbar_expect_tx(aux_bar(q, slab % NB), kStageBufBytes);
            ptx::tma_load_4d(dbuf, &tma_aux, aux_bar(q, slab % NB), ns + NB * 64, m0 + q * 32, 0, 0);
          }
          if (etr) et[7] = clock64();
        }
      } else if constexpr (!FAST) {
        // ---- direct path (fp32 output, accumulate, or unaligned): per-thread row segments --------------------------
#pragma unroll 1
        for (int c = 0; c < ((dbg & 4) ? 0 : BN / 32); ++c) {
          uint32_t raw[32];
          ld_acc(c * 32, raw);
          const int nb = n0 + c * 32;
          if (row_ok && nb < g.N && !(dbg & 1)) {
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]) * g.alpha;
            if (vec_ok && nb + 32 <= g.N) {
              if (g.bias) {
#pragma unroll
                for (int j8 = 0; j8 < 4; ++j8) {
                  float bf[8]; ld8_any(g.bias, nb + j8 * 8, g.io_f32, bf);
#pragma unroll
                  for (int j = 0; j < 8; ++j) v[j8 * 8 + j] += bf[j];
                }
              }
              if (epi != EPI_NONE) {
                const long long a_off = (long long)m * g.ld_aux + nb;
#pragma unroll
                for (int j8 = 0; j8 < 4; ++j8) {
                  float af[8];
                  if (epi == EPI_GELU_SAVE) {
                    if (g.io_f32) {
                      float4* ap = reinterpret_cast<float4*>(reinterpret_cast<float*>(g.aux) + a_off + j8 * 8);
                      ap[0] = make_float4(v[j8 * 8], v[j8 * 8 + 1], v[j8 * 8 + 2], v[j8 * 8 + 3]);
                      ap[1] = make_float4(v[j8 * 8 + 4], v[j8 * 8 + 5], v[j8 * 8 + 6], v[j8 * 8 + 7]);
#pragma unroll
                      for (int j = 0; j < 8; ++j) af[j] = v[j8 * 8 + j];
                    } else {
                      bf16x8 pk = pack8(&v[j8 * 8]);
                      st8(reinterpret_cast<__nv_bfloat16*>(g.aux) + a_off + j8 * 8, pk);
                      unpack8(pk, af);
                    }
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[j8 * 8 + j] = gelu_tanh(af[j]);
                  } else {
                    ld8_any(g.aux, a_off + j8 * 8, g.io_f32, af);
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                      v[j8 * 8 + j] = epi == EPI_GELU_BWD ? v[j8 * 8 + j] * gelu_tanh_grad(af[j]) : v[j8 * 8 + j] + af[j];
                  }
                }
              }
              if (g.d_f32) {
                float* dp = reinterpret_cast<float*>(g.d) + d_off + nb;
#pragma unroll
                for (int j4 = 0; j4 < 8; ++j4) {
                  float4 o = make_float4(v[j4 * 4], v[j4 * 4 + 1], v[j4 * 4 + 2], v[j4 * 4 + 3]);
                  if (g.accumulate) { float4 p = reinterpret_cast<float4*>(dp)[j4]; o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w; }
                  reinterpret_cast<float4*>(dp)[j4] = o;
                }
              } else {
                __nv_bfloat16* dp = reinterpret_cast<__nv_bfloat16*>(g.d) + d_off + nb;
#pragma unroll
                for (int j8 = 0; j8 < 4; ++j8) {
                  if (g.accumulate) {
                    float pf[8]; unpack8(ld8(dp + j8 * 8), pf);
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[j8 * 8 + j] += pf[j];
                  }
                  st8(dp + j8 * 8, pack8(&v[j8 * 8]));
                }
              }
            } else {
              // ragged / unaligned tail: scalar, statically indexed so v[] stays in registers
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const int n = nb + j;
                if (n < g.N) {
                  float x = v[j];
                  if (g.bias) x += ld1_any(g.bias, n, g.io_f32);
                  const long long ai = (long long)m * g.ld_aux + n;
                  if (epi == EPI_GELU_SAVE) {
                    if (g.io_f32) {
                      reinterpret_cast<float*>(g.aux)[ai] = x;
                    } else {
                      __nv_bfloat16 pre = __float2bfloat16_rn(x);
                      reinterpret_cast<__nv_bfloat16*>(g.aux)[ai] = pre;
                      x = __bfloat162float(pre);
                    }
                    x = gelu_tanh(x);
                  } else if (epi == EPI_GELU_BWD) {
                    x *= gelu_tanh_grad(ld1_any(g.aux, ai, g.io_f32));
                  } else if (epi == EPI_RESIDUAL) {
                    x += ld1_any(g.aux, ai, g.io_f32);
                  }
                  if (g.d_f32) {
                    float* dp = reinterpret_cast<float*>(g.d) + d_off + n;
                    *dp = g.accumulate ? *dp + x : x;
                  } else {
                    __nv_bfloat16* dp = reinterpret_cast<__nv_bfloat16*>(g.d) + d_off + n;
                    *dp = __float2bfloat16_rn(g.accumulate ? __bfloat162float(*dp) + x : x);
                  }
                }
              }
            }
          }
        }
      }
      // all TMEM reads of this accumulator stage are complete (wait::ld above): hand it back to the MMA warp
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(tempty_bar(as));
      if (prof && warp == kEpiWarp0 && lane == 0) prof_stamp(prof, 10);
    }
    if (lane == 0) {
      if constexpr (RED) {
        // the adds must be PERFORMED (not just read out of smem) before this grid counts as complete: the consumer is a
        // kernel on another GPU that synchronises through flags only
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        __threadfence_system();
      } else {
        ptx::bulk_wait_read<0>();     // staging smem must outlive the last TMA store's reads
      }
    }
    if (prof && warp == kEpiWarp0 && lane == 0) prof_stamp(prof, 11);
    __syncwarp();
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (prof && threadIdx.x == 0) { prof_stamp(prof, 13); prof[(long long)blockIdx.x * 16 + 14] = globaltimer_ns(); }
  if (g.cm > 1) ptx::cluster_sync();   // nobody leaves while a peer may still multicast into / signal this CTA
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, kTmemAlloc);
  }
}

// =====================================================================================================
// Host side
// =====================================================================================================
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// 4-D map over one bf16 operand: dims (inner, rows, nb2, nb1)
bool make_map(CUtensorMap* out, const GemmOperand& op, int rows_mn, int K, int nb1, int nb2, int box_rows_kmajor,
              int box_krows_mnmajor, bool f32) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return false;
  cuuint64_t dims[4];
  cuuint32_t box[4];
  const cuuint32_t inner = f32 ? 32 : 64;     // one 128-byte swizzle row
  const cuuint64_t es = f32 ? 4 : 2;
  if (!op.mn_major) { dims[0] = (cuuint64_t)K; dims[1] = (cuuint64_t)rows_mn; box[0] = inner; box[1] = (cuuint32_t)box_rows_kmajor; }
  else              { dims[0] = (cuuint64_t)rows_mn; dims[1] = (cuuint64_t)K; box[0] = inner; box[1] = (cuuint32_t)box_krows_mnmajor; }
  dims[2] = (cuuint64_t)nb2; dims[3] = (cuuint64_t)nb1;
  box[2] = 1; box[3] = 1;
  const cuuint64_t row_bytes = (cuuint64_t)op.ld * es;
  cuuint64_t strides[3];
  strides[0] = row_bytes;
  strides[1] = nb2 > 1 ? (cuuint64_t)op.batch_stride2 * es : row_bytes * dims[1];
  strides[2] = nb1 > 1 ? (cuuint64_t)op.batch_stride * es : strides[1] * (nb2 > 1 ? (cuuint64_t)nb2 : 1);
  for (int i = 1; i < 3; ++i) if (strides[i] == 0) strides[i] = row_bytes;
  cuuint32_t estr[4] = {1, 1, 1, 1};
  // TFLOAT32: the TMA rounds fp32 to TF32 on the way into smem (the MMA would otherwise truncate the low mantissa bits);
  // TDS_GEMM_TF32_MAP=0 loads the raw fp32 bits instead
  static const int tf32_map = getenv("TDS_GEMM_TF32_MAP") ? atoi(getenv("TDS_GEMM_TF32_MAP")) : 1;
  const CUtensorMapDataType dt = !f32 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                      : (tf32_map ? CU_TENSOR_MAP_DATA_TYPE_TFLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32);
  const CUtensorMapSwizzle sw = (f32 && op.mn_major) ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B;
  CUresult r = enc(out, dt, 4, const_cast<void*>(op.ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    fprintf(stderr, "[tds] cuTensorMapEncodeTiled failed (%d): dims=(%llu,%llu,%llu,%llu) strides=(%llu,%llu,%llu) ptr=%p\n",
            (int)r, (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)dims[2],
            (unsigned long long)dims[3], (unsigned long long)strides[0], (unsigned long long)strides[1],
            (unsigned long long)strides[2], op.ptr);
    return false;
  }
  return true;
}

// 2-D fp32 map (cols contiguous), SWIZZLE_128B, box (box_cols <= 32, box_rows): used for TMA reduce-add of fp32 tiles
bool make_map_f32_2d(CUtensorMap* out, void* ptr, int64_t rows, int64_t cols, int64_t ld, int box_cols, int box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return false;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { fprintf(stderr, "[tds] cuTensorMapEncodeTiled(f32) failed (%d)\n", (int)r); return false; }
  return true;
}

static int g_num_sms = 0;
static long long* g_prof = nullptr;
void gemm_set_prof(long long* buf) { g_prof = buf; }
static int g_dbg = -1;       // TDS_GEMM_DBG bits; tools/gemm_harness.cu overrides them per run
void gemm_set_debug(int bits) { g_dbg = bits; }
static int g_variant = 0;    // experimental kernel variants compared by tools/gemm_harness.cu
void gemm_set_variant(int v) { g_variant = v; }

int gemm_num_configs() { return 4; }   // BN = 64, 128, 256, 192

static int pick_config(const GemmParams& p) {
  if (p.config >= 0 && p.config < 4) return p.config;
  const long long mt = (p.M + BM - 1) / BM;
  auto tiles = [&](int bn) { return mt * ((p.N + bn - 1) / bn) * p.batch; };
  // largest tile that still gives (nearly) every SM work; small problems take the narrow tile for parallelism
  if (tiles(256) >= g_num_sms) return 2;
  // just over one wave with 128-wide tiles but exactly fits with 192-wide ones (c_fc: 1024 x 3072 -> 128 tiles)
  if (tiles(128) > g_num_sms && tiles(192) <= g_num_sms && p.N % 192 == 0) return 3;
  if (tiles(128) * 10 >= (long long)g_num_sms * 6) return 1;
  return tiles(64) > tiles(128) ? 0 : 1;
}

template <int BN>
static int max_clusters(int cm) {
  static int cache[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  if (cache[cm]) return cache[cm];
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(g_num_sms / cm * cm);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = Cfg<BN>::kSmem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = cm; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, gemm_kernel<BN, false, false, 1, -1, false>, &cfg) != cudaSuccess || n <= 0) { cudaGetLastError(); n = g_num_sms / cm / 2; }
  cache[cm] = n;
  return n;
}

template <int BN, bool F32, bool RED = false, int NP = 1, int EPI = -1, int FAST = 0>
static void launch(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& td, const CUtensorMap& tx,
                   const GemmDev& g, int tiles, cudaStream_t s) {
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(gemm_kernel<BN, F32, RED, NP, EPI, FAST>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<BN>::kSmem);
    attr_done = true;
  }
  if (g.cm == 1) {
    const int grid = tiles < g_num_sms ? tiles : g_num_sms;
    launch_k(gemm_kernel<BN, F32, RED, NP, EPI, FAST>, dim3(grid), dim3(gemm_threads(NP, FAST)), Cfg<BN>::kSmem, s, ta, tb, td, tx, g);
    return;
  }
  // cluster launch: cm consecutive CTAs = cm consecutive M tiles of one N tile; grid is a whole number of clusters
  int clusters = tiles / g.cm;
  const int cap = max_clusters<BN>(g.cm);
  if (clusters > cap) clusters = cap;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(clusters * g.cm);
  cfg.blockDim = dim3(gemm_threads(NP, FAST));
  cfg.dynamicSmemBytes = Cfg<BN>::kSmem;
  cfg.stream = s;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = g.cm; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[1].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = at; cfg.numAttrs = 2;
  cudaLaunchKernelEx(&cfg, gemm_kernel<BN, F32, RED, NP, EPI, FAST>, ta, tb, td, tx, g);
}

// cluster size along M: B-tile multicast divides L2->SM (or NVLink, for a ZeRO-3 peer weight) operand traffic by cm
static int pick_cluster(const GemmParams& p, int bn) {
  static const int env = getenv("TDS_GEMM_CM") ? atoi(getenv("TDS_GEMM_CM")) : -1;
  // measured on B200 (profiles/r1_overlap_pdl.md): at M = 1024 the GPT-2 GEMMs are latency-, not L2-bound, and
  // cluster launch costs more than the multicast saves -> off unless requested
  int want = p.cluster_m > 0 ? p.cluster_m : (env >= 0 ? env : 1);
  if (want <= 1 || p.tri != 0 || p.batch != 1 || p.in_dtype == kF32 || p.reduce_out) return 1;
  const int m_tiles = (p.M + BM - 1) / BM;
  int cm = 8;
  while (cm > 1 && (cm > want || m_tiles % cm != 0 || (bn / cm) % 8 != 0)) cm >>= 1;
  return cm;
}

void gemm_bf16(const GemmParams& p, cudaStream_t stream) {
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  if (p.M <= 0 || p.N <= 0 || p.batch <= 0) return;
  const int cfg = pick_config(p);
  const int bn = cfg == 0 ? 64 : (cfg == 1 ? 128 : (cfg == 2 ? 256 : 192));
  const int nb2 = p.nbatch2 > 0 ? p.nbatch2 : 1;
  const int nb1 = p.batch / nb2;
  CUtensorMap ta, tb;
  const int cm = pick_cluster(p, bn);
  const bool f32 = p.in_dtype == kF32;
  const int bk = f32 ? 32 : BK;
  if (!make_map(&ta, p.a, p.M, p.K, nb1, nb2, BM, bk, f32) || !make_map(&tb, p.b, p.N, p.K, nb1, nb2, bn / cm, bk / cm, f32)) {
    fprintf(stderr, "[tds] gemm: tensor map creation failed (M=%d N=%d K=%d)\n", p.M, p.N, p.K);
    abort();
  }
  GemmDev g;
  g.d = p.d; g.d_f32 = p.d_dtype == kF32; g.ldd = p.ldd; g.dbs1 = p.d_batch_stride; g.dbs2 = p.d_batch_stride2;
  g.bias = p.bias;
  g.aux = p.aux; g.ld_aux = p.ld_aux;
  g.io_f32 = p.io_dtype == kF32 ? 1 : 0;
  g.epi = p.aux ? p.epi : EPI_NONE; g.accumulate = p.accumulate ? 1 : 0; g.alpha = p.alpha;
  g.M = p.M; g.N = p.N; g.K = p.K; g.batch = p.batch; g.nb2 = nb2;
  g.a_mn = p.a.mn_major; g.b_mn = p.b.mn_major; g.tri = p.tri;
  g.pf = nullptr; g.pf_bytes = 0;
  if (p.prefetch && p.prefetch_bytes >= 16 && (reinterpret_cast<uintptr_t>(p.prefetch) & 15) == 0) {
    g.pf = reinterpret_cast<const char*>(p.prefetch); g.pf_bytes = p.prefetch_bytes & ~15LL;
  }
  if (g_dbg < 0) g_dbg = getenv("TDS_GEMM_DBG") ? atoi(getenv("TDS_GEMM_DBG")) : 0;
  g.dbg = g_dbg;
  g.prof = g_prof;
  g.idesc = f32 ? make_idesc_tf32(BM, bn, p.a.mn_major, p.b.mn_major) : make_idesc_bf16(BM, bn, p.a.mn_major, p.b.mn_major);
  g.cm = cm;
  const long long tiles = (long long)((p.M + BM - 1) / BM) * ((p.N + bn - 1) / bn) * p.batch;
  // output through TMA (coalesced 128-byte rows) whenever the layout allows it
  CUtensorMap td = ta, tx = ta;
  g.tma_store = 0; g.aux_tma = 0;
  static const int no_tma_store = getenv("TDS_GEMM_DIRECT_STORE") ? atoi(getenv("TDS_GEMM_DIRECT_STORE")) : 0;
  const bool aligned = (p.ldd % 8 == 0) && ((reinterpret_cast<uintptr_t>(p.d) & 15) == 0) && (p.N % 8 == 0) &&
                       (p.d_batch_stride % 8 == 0) && (p.d_batch_stride2 % 8 == 0);
  if (p.reduce_out) {
    // fp32 [M, N] destination (possibly peer-mapped), 32 x 32 boxes; bf16 operands only, no batch, no epilogue functor
    if (f32 || p.d_dtype != kF32 || p.batch != 1 || p.bias || p.aux || (p.ldd % 4) != 0 ||
        (reinterpret_cast<uintptr_t>(p.d) & 15) != 0 || !make_map_f32_2d(&td, p.d, p.M, p.N, p.ldd, 32, 32)) {
      fprintf(stderr, "[tds] gemm: reduce_out needs bf16 operands and a 16-byte aligned fp32 [M,N] destination\n");
      abort();
    }
  } else if (!no_tma_store && p.d_dtype == kBF16 && !p.accumulate && aligned && !g.io_f32 && p.N >= 8 &&
             (reinterpret_cast<uintptr_t>(p.bias) & 15) == 0) {
    GemmOperand od{p.d, p.ldd, p.d_batch_stride, p.d_batch_stride2, false};
    bool ok = make_map(&td, od, p.M, p.N, nb1, nb2, 32);
    if (ok && g.epi != EPI_NONE) {
      GemmOperand oa{p.aux, p.ld_aux, 0, 0, false};
      ok = (p.ld_aux % 8 == 0) && ((reinterpret_cast<uintptr_t>(p.aux) & 15) == 0) && make_map(&tx, oa, p.M, p.N, 1, 1, 32);
      static const int no_aux_tma = getenv("TDS_GEMM_AUX_DIRECT") ? atoi(getenv("TDS_GEMM_AUX_DIRECT")) : 0;
      g.aux_tma = (ok && !no_aux_tma) ? 1 : 0;
    }
    g.tma_store = ok ? 1 : 0;
  }
  static const int dual_env = getenv("TDS_GEMM_DUAL") ? atoi(getenv("TDS_GEMM_DUAL")) : 1;   // 0: one pipeline per CTA
  const int T = (int)tiles;
#define TDS_L(BN_, F32_, RED_, NP_, EPI_, FAST_) launch<BN_, F32_, RED_, NP_, EPI_, FAST_>(ta, tb, td, tx, g, T, stream)
  // pipelines per CTA: BN = 64 -> 4 (its MMAs are 32-48 cycles of tensor work against ~80-100 cycles of issue per thread),
  // BN = 128 -> 2 (64 cycles: two issuers already saturate the pipe), wider tiles and fp32 -> 1.  g_variant (harness only):
  // 1 = single pipeline everywhere, 2 = at most two pipelines.
#define TDS_BY_BN(F32_, RED_, EPI_, FAST_)                                                        \
  do {                                                                                            \
    const int np64 = (F32_) || g_variant == 1 || !dual_env ? 1 : (g_variant == 2 || (RED_) ? 2 : 4);   \
    const int np128 = (F32_) || g_variant == 1 || !dual_env ? 1 : 2;                              \
    if (cfg == 0) {                                                                               \
      if (np64 == 4) TDS_L(64, F32_, RED_, (F32_) || (RED_) ? 1 : 4, EPI_, FAST_);               \
      else if (np64 == 2) TDS_L(64, F32_, RED_, (F32_) ? 1 : 2, EPI_, FAST_);                    \
      else TDS_L(64, F32_, RED_, 1, EPI_, FAST_);                                                \
    } else if (cfg == 1) {                                                                        \
      if (np128 == 2) TDS_L(128, F32_, RED_, (F32_) ? 1 : 2, EPI_, FAST_);                       \
      else TDS_L(128, F32_, RED_, 1, EPI_, FAST_);                                               \
    } else if (cfg == 2) TDS_L(256, F32_, RED_, 1, EPI_, FAST_);                                 \
    else TDS_L(192, F32_, RED_, 1, EPI_, FAST_);                                                 \
  } while (0)
  if (p.reduce_out) {
    TDS_BY_BN(false, true, -1, 0);
  } else if (f32) {
    TDS_BY_BN(true, false, -1, 0);
  } else if (g.tma_store && (g.aux_tma || g.epi == EPI_NONE || g.epi == EPI_GELU_SAVE)) {
    // the bf16 training step lives here: epilogue functor and store path fixed at compile time
    // TDS_GEMM_EPI_WARPS=8: two epilogue warps per TMEM lane quadrant.  Faster per GEMM (profiles/r2_gemm_epi_8warps.log) but the
    // CTA then owns ~62 K of the SM's 64 K registers, so no collective CTA can share the SM: default 4
    static const bool ew8 = getenv("TDS_GEMM_EPI_WARPS") && atoi(getenv("TDS_GEMM_EPI_WARPS")) == 8;
    if (ew8) {
      if (g.epi == EPI_NONE) TDS_BY_BN(false, false, EPI_NONE, 2);
      else if (g.epi == EPI_GELU_SAVE) TDS_BY_BN(false, false, EPI_GELU_SAVE, 2);
      else if (g.epi == EPI_GELU_BWD) TDS_BY_BN(false, false, EPI_GELU_BWD, 2);
      else TDS_BY_BN(false, false, EPI_RESIDUAL, 2);
    } else {
      if (g.epi == EPI_NONE) TDS_BY_BN(false, false, EPI_NONE, 1);
      else if (g.epi == EPI_GELU_SAVE) TDS_BY_BN(false, false, EPI_GELU_SAVE, 1);
      else if (g.epi == EPI_GELU_BWD) TDS_BY_BN(false, false, EPI_GELU_BWD, 1);
      else TDS_BY_BN(false, false, EPI_RESIDUAL, 1);
    }
  } else {
    TDS_BY_BN(false, false, -1, 0);
  }
#undef TDS_BY_BN
#undef TDS_L
}

}  // namespace tds
