// Fused causal attention for sm_100a (head size 64, T % 128 == 0): tcgen05 + TMEM + TMA, no T x T matrix in HBM.
//
// Forward, one CTA per (128-query block, head, batch):
//   warp 0    TMA producer: Q once, then the K_j / V_j blocks of 128 keys through a 2-stage ring
//   warp 1    MMA issuer:   S_j = Q K_j^T  (UMMA 128x128x16, 4 k-steps)  -> TMEM S[j&1]
//                           PV_j = P_j V_j (UMMA 128x64x16, 8 k-steps, V consumed MN-major) -> TMEM PV[j&1]
//             S_{j+1} is issued before PV_j so the tensor core works underneath the softmax of block j
//   warps 2-9 softmax / correction / epilogue: TWO threads per query row (warps w and w + 4 share a TMEM lane quadrant; each
//             takes 64 of the block's 128 score columns and 32 of the 64 output columns), S read from TMEM once, row maximum
//             agreed through smem + a 64-thread named barrier, online softmax in the log2 domain, P_j written as bf16 into
//             128B-swizzled smem (the A operand of PV_j), running output in registers (O = O * alpha + PV_j), final O / l
//             through swizzled smem + TMA store, LSE to HBM.  (One thread per row was a single warp per scheduler running a
//             ~3000-instruction dependent stream per key block: 17.8 us per launch at T = 1024; profiles/r2_timeline_small.md)
// Only the diagonal block is masked; blocks above the diagonal are never touched.
//
// Replaces the reference's standard_attention (example/model.py:29-42: QK^T, mask, softmax, PV as four ATen ops
// with a materialised [B,nh,T,T] score tensor) and the SDPA fallback.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdio.h>

#include "common.cuh"
#include "kernels.h"
#include "sm100_ptx.cuh"
#include "tma_util.h"

namespace tds {

namespace {

constexpr int FB = 128;          // query rows per CTA == keys per KV block
constexpr int HS = 64;           // head size
constexpr int kFThreads = 320;       // forward: producer, issuer, 8 softmax warps (two per TMEM lane quadrant)
constexpr int kFBThreads = 352;      // backward: producer, issuer A, 8 softmax warps, issuer B (warp 10)
constexpr uint32_t kTileQK = FB * HS * 2;        // 16 KB: one [128 x 64] bf16 tile
constexpr uint32_t kTileP = FB * FB * 2;         // 32 KB: P as two K-major 64-column slabs
constexpr uint32_t kFwdSmem = kTileQK * 5 + kTileP * 2 + 1024 + 256 + 2048 /*row-max / row-sum exchange*/;

struct FlashDev {
  int T, nh;
  float cs;        // softmax scale * log2(e)
  float* lse;      // [B, nh, T]  (log2 domain: m * cs + log2(l))
  uint32_t idesc_s, idesc_pv;
};

__global__ void __launch_bounds__(kFThreads, 1)
flash_fwd_kernel(const __grid_constant__ CUtensorMap tma_q, const __grid_constant__ CUtensorMap tma_k,
                 const __grid_constant__ CUtensorMap tma_v, const __grid_constant__ CUtensorMap tma_o,
                 const __grid_constant__ FlashDev g) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = base;
  const uint32_t sK = sQ + kTileQK;               // 2 stages
  const uint32_t sV = sK + 2 * kTileQK;           // 2 stages
  const uint32_t sP = sV + 2 * kTileQK;           // 2 buffers
  const uint32_t sBar = sP + 2 * kTileP;
  enum { Q_FULL = 0, KV_FULL = 1, KV_EMPTY = 3, S_FULL = 5, S_FREE = 7, P_FULL = 9, P_FREE = 11, PV_FULL = 13, PV_FREE = 15, NBAR = 17 };
  auto bar = [&](int i) { return sBar + 8u * i; };
  const uint32_t tmem_slot = sBar + 8u * NBAR;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - ptx::smem_u32(smem_raw)));
  float* xch = reinterpret_cast<float*>(smem_raw + (sBar + 256u - ptx::smem_u32(smem_raw)));   // [2 parities][2 halves][128 rows]

  // warp index through a shuffle: provably warp-uniform, so role branches are uniform and issue code can use elect.sync
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int qb = (int)gridDim.x - 1 - (int)blockIdx.x;    // heaviest (last) query blocks are scheduled first
  const int h = blockIdx.y, b = blockIdx.z;
  const int n_kv = qb + 1;
  pdl_launch();

  if (warp == 0 && ptx::elect_one()) {
    ptx::prefetch_tmap(&tma_q); ptx::prefetch_tmap(&tma_k); ptx::prefetch_tmap(&tma_v); ptx::prefetch_tmap(&tma_o);
    ptx::mbar_init(bar(Q_FULL), 1);
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(bar(KV_FULL + s), 1);  ptx::mbar_init(bar(KV_EMPTY + s), 1);
      ptx::mbar_init(bar(S_FULL + s), 1);   ptx::mbar_init(bar(S_FREE + s), 8);
      ptx::mbar_init(bar(P_FULL + s), 8);   ptx::mbar_init(bar(P_FREE + s), 1);
      ptx::mbar_init(bar(PV_FULL + s), 1);  ptx::mbar_init(bar(PV_FREE + s), 8);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 1) { ptx::tmem_alloc(tmem_slot, 512); ptx::tmem_relinquish(); }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot_ptr;
  const uint32_t tS = tmem, tPV = tmem + 256;      // S[2] at columns 0/128, PV[2] at 256/320
  pdl_wait();

  if (warp == 0) {
    if (ptx::elect_one()) {
      ptx::mbar_expect_tx(bar(Q_FULL), kTileQK);
      ptx::tma_load_4d(sQ, &tma_q, bar(Q_FULL), 0, qb * FB, h, b);
      for (int j = 0; j < n_kv; ++j) {
        const int st = j & 1; const uint32_t ph = (j >> 1) & 1;
        ptx::mbar_wait_conv(bar(KV_EMPTY + st), ph ^ 1u);
        ptx::mbar_expect_tx(bar(KV_FULL + st), 2 * kTileQK);
        ptx::tma_load_4d(sK + st * kTileQK, &tma_k, bar(KV_FULL + st), 0, j * FB, h, b);
        ptx::tma_load_4d(sV + st * kTileQK, &tma_v, bar(KV_FULL + st), 0, j * FB, h, b);
      }
    }
  } else if (warp == 1) {
    auto issue_s = [&](int j) {
      const int st = j & 1; const uint32_t ph = (j >> 1) & 1;
      ptx::mbar_wait_conv(bar(KV_FULL + st), ph);
      ptx::mbar_wait_conv(bar(S_FREE + st), ph ^ 1u);
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
#pragma unroll
        for (int k = 0; k < HS / 16; ++k) {
          const uint64_t da = ptx::make_smem_desc(sQ + k * 32, 16, 1024);
          const uint64_t db = ptx::make_smem_desc(sK + st * kTileQK + k * 32, 16, 1024);
          ptx::mma_f16_ss(tS + st * 128, da, db, g.idesc_s, k > 0 ? 1u : 0u);
        }
        ptx::mma_commit(bar(S_FULL + st));
      }
      __syncwarp();
    };
    ptx::mbar_wait_conv(bar(Q_FULL), 0);
    issue_s(0);
    for (int j = 0; j < n_kv; ++j) {
      if (j + 1 < n_kv) issue_s(j + 1);
      const int st = j & 1; const uint32_t ph = (j >> 1) & 1;
      ptx::mbar_wait_conv(bar(P_FULL + st), ph);
      ptx::mbar_wait_conv(bar(PV_FREE + st), ph ^ 1u);
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
#pragma unroll
        for (int ks = 0; ks < FB / 16; ++ks) {
          const uint64_t da = ptx::make_smem_desc(sP + st * kTileP + (ks >> 2) * (FB * 128) + (ks & 3) * 32, 16, 1024);
          const uint64_t db = ptx::make_smem_desc(sV + st * kTileQK + ks * 2048, FB * 128, 1024);   // MN-major
          ptx::mma_f16_ss(tPV + st * 64, da, db, g.idesc_pv, ks > 0 ? 1u : 0u);
        }
        ptx::mma_commit(bar(PV_FULL + st));
        ptx::mma_commit(bar(KV_EMPTY + st));
        ptx::mma_commit(bar(P_FREE + st));
      }
      __syncwarp();
    }
  } else {
    const int q = warp & 3;                        // TMEM lane quadrant of this warp
    const int hf = (warp - 2) >> 2;                // which half of the score / output columns this thread owns
    const int row = q * 32 + lane;                 // row inside the 128-query block
    const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
    auto pair_sync = [&]() { asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory"); };   // the two warps of a quadrant
    float m = -1e30f, l = 0.f;
    float o[HS / 2];
#pragma unroll
    for (int i = 0; i < HS / 2; ++i) o[i] = 0.f;
    float alpha_prev = 1.f;
    // O = O * alpha_j + P_j V_j   (alpha_j rescales everything accumulated BEFORE block j); this thread: 32 of the 64 columns
    auto o_update = [&](int jj, float a) {
      const int st2 = jj & 1; const uint32_t ph2 = (jj >> 1) & 1;
      ptx::mbar_wait(bar(PV_FULL + st2), ph2);
      ptx::tc_fence_after();
      uint32_t raw[32];
      ptx::tmem_ld_32x32(tPV + st2 * 64 + lane_sel + hf * 32, raw);
      ptx::tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) o[i] = o[i] * a + __uint_as_float(raw[i]);
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(bar(PV_FREE + st2));
    };
    for (int j = 0; j < n_kv; ++j) {
      const int st = j & 1; const uint32_t ph = (j >> 1) & 1;
      const bool diag = (j == qb);
      ptx::mbar_wait(bar(S_FULL + st), ph);
      ptx::tc_fence_after();
      // this thread's 64 scores, read ONCE; the accumulator stage goes back to the issuer right away
      const uint32_t t_s = tS + st * 128 + lane_sel + hf * 64;
      uint32_t raw[64];
      ptx::tmem_ld_32x32(t_s, &raw[0]);
      ptx::tmem_ld_32x32(t_s + 32, &raw[32]);
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(bar(S_FREE + st));
      const int lim = row - hf * 64;               // diagonal block: column i of this half is visible iff i <= lim
      float mx = -1e30f;
      if (diag) {
#pragma unroll
        for (int i = 0; i < 64; ++i) if (i <= lim) mx = fmaxf(mx, __uint_as_float(raw[i]));
      } else {
#pragma unroll
        for (int i = 0; i < 64; ++i) mx = fmaxf(mx, __uint_as_float(raw[i]));
      }
      // agree on the row maximum with the thread that owns the other 64 columns (double-buffered by block parity)
      float* xm = xch + (j & 1) * 256;
      xm[hf * 128 + row] = mx;
      pair_sync();
      const float m_new = fmaxf(m, fmaxf(mx, xm[(hf ^ 1) * 128 + row]));
      const float alpha = exp2f((m - m_new) * g.cs);
      const float mcs = m_new * g.cs;
      // probabilities -> bf16 -> this half's 64-column K-major slab of P (A operand of P V)
      ptx::mbar_wait(bar(P_FREE + st), ph ^ 1u);
      float rs = 0.f;
      const uint32_t slab = sP + st * kTileP + (uint32_t)hf * (FB * 128) + (uint32_t)row * 128u;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        float p[32];
        if (diag) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float e = exp2f(__uint_as_float(raw[c * 32 + i]) * g.cs - mcs);
            p[i] = (c * 32 + i <= lim) ? e : 0.f;
            rs += p[i];
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            p[i] = exp2f(__uint_as_float(raw[c * 32 + i]) * g.cs - mcs);
            rs += p[i];
          }
        }
#pragma unroll
        for (int j8 = 0; j8 < 4; ++j8) {
          const uint32_t idx = (uint32_t)(c * 4 + j8);
          ptx::st_shared_16(slab + ((idx ^ (uint32_t)(row & 7)) << 4), pack8(&p[j8 * 8]));
        }
      }
      ptx::fence_proxy_async();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(bar(P_FULL + st));
      l = l * alpha + rs;                          // partial row sum over this thread's columns (same alpha in both threads)
      m = m_new;
      // O update of the PREVIOUS block (software pipelining: P_{j-1} V_{j-1} ran on the tensor core while this
      // thread was busy with the softmax of block j), then remember this block's rescale factor
      if (j > 0) o_update(j - 1, alpha_prev);
      alpha_prev = alpha;
    }
    o_update(n_kv - 1, alpha_prev);
    // row sum = the two partial sums
    float* xl = xch + (n_kv & 1) * 256;
    xl[hf * 128 + row] = l;
    pair_sync();
    l += xl[(hf ^ 1) * 128 + row];
    // epilogue: O / l -> bf16 -> swizzled staging (sP[0] is free: every P V has completed) -> TMA store
    const float inv = 1.f / l;
    const uint32_t stg = sP + (uint32_t)q * 4096u + (uint32_t)lane * 128u;
#pragma unroll
    for (int j8 = 0; j8 < 4; ++j8) {
      float t[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) t[i] = o[j8 * 8 + i] * inv;
      ptx::st_shared_16(stg + (((uint32_t)(hf * 4 + j8) ^ (uint32_t)(lane & 7)) << 4), pack8(t));
    }
    ptx::fence_proxy_async();
    pair_sync();                                   // both halves of the 32 x 64 staging slab are written
    if (hf == 0) {
      if (lane == 0) {
        ptx::tma_store_4d(&tma_o, sP + (uint32_t)q * 4096u, 0, qb * FB + q * 32, h, b);
        ptx::bulk_commit();
      }
      g.lse[((size_t)b * g.nh + h) * g.T + qb * FB + row] = m * g.cs + log2f(l);
      if (lane == 0) ptx::bulk_wait_read<0>();
      __syncwarp();
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem, 512); }
}

// =====================================================================================================
// Backward.  One CTA per (128-key block j, head, batch) owns dK_j and dV_j (accumulated in TMEM over the query
// blocks i >= j) and streams Q_i / dO_i through a 2-stage TMA ring.  Per (i, j):
//     S  = Q_i K_j^T                (UMMA 128x128, K-major / K-major)         -> TMEM
//     dP = dO_i V_j^T               (UMMA 128x128, K-major / K-major)         -> TMEM
//     P  = exp2(S*cs - lse_i),  dS = P o (dP - D_i) * scale                   (thread == query row; bf16 -> smem)
//     dV_j += P^T dO_i,  dK_j += dS^T Q_i   (UMMA 128x64, BOTH operands MN-major: the P/dS/dO/Q tiles as stored)
//     dQ_i  = dS K_j                (UMMA 128x64, K-major / MN-major) -> fp32 -> TMA reduce-add into the dQ workspace
// =====================================================================================================
constexpr uint32_t kBwdSmem = kTileQK * 2 /*K,V*/ + kTileQK * 4 /*Q,dO x2*/ + kTileP * 2 /*P,dS*/ + 32768 /*dQ staging*/ + 1024 + 256;

struct FlashBwdDev {
  int T, nh;
  float cs, scale;
  const float* lse;    // [B, nh, T] log2 domain
  const float* dsum;   // [B, nh, T]  D = rowsum(dO o O)
  int n_split;         // key blocks jb < n_split are processed by two CTAs (half of the query range each)
  uint32_t idesc_s, idesc_dkv, idesc_dq;
};

__global__ void __launch_bounds__(kFBThreads, 1)
flash_bwd_kernel(const __grid_constant__ CUtensorMap tma_q, const __grid_constant__ CUtensorMap tma_k,
                 const __grid_constant__ CUtensorMap tma_v, const __grid_constant__ CUtensorMap tma_do,
                 const __grid_constant__ CUtensorMap tma_dk, const __grid_constant__ CUtensorMap tma_dv,
                 const __grid_constant__ CUtensorMap tma_dq, const __grid_constant__ CUtensorMap tma_dkw,
                 const __grid_constant__ CUtensorMap tma_dvw, const __grid_constant__ FlashBwdDev g) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sK = base, sV = sK + kTileQK;
  const uint32_t sQ = sV + kTileQK;                 // 2 stages
  const uint32_t sDO = sQ + 2 * kTileQK;            // 2 stages
  const uint32_t sP = sDO + 2 * kTileQK;
  const uint32_t sDS = sP + kTileP;
  const uint32_t sStg = sDS + kTileP;               // 8 warps x 4 KB (fp32 dQ quarter / bf16 dV or dK slab)
  const uint32_t sBar = sStg + 32768;
  enum { KV_FULL = 0, QDO_FULL = 1, QDO_EMPTY = 3, SDP_FULL = 5, SDP_FREE = 6, PDS_FULL = 7, PDS_FREE = 8, DQ_FULL = 9, DQ_FREE = 10, DKV_DONE = 11, NBAR = 12 };
  auto bar = [&](int i) { return sBar + 8u * i; };
  const uint32_t tmem_slot = sBar + 8u * NBAR;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - ptx::smem_u32(smem_raw)));

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  // Work items.  Key block jb meets the query blocks i >= jb: nq - jb steps, 8 for jb = 0 but 1 for the last one, and the launch
  // is as long as its longest CTA.  The key blocks with more than ceil(nq / 2) steps are therefore split into two CTAs that take
  // half of the query range each and add their fp32 dK / dV partials into a workspace (TMA reduce-add, converted together with
  // dQ); the rest store bf16 dK / dV directly.  T = 1024: 12 CTAs per head with at most 4 steps instead of 8 with up to 8.
  const int h = blockIdx.y, b = blockIdx.z;
  const int nq = g.T / FB;
  const int n_split = g.n_split;                  // key blocks jb < n_split are split (host: nq - ceil(nq / 2))
  int jb, it0, n_it;
  bool split;
  if ((int)blockIdx.x < 2 * n_split) {
    jb = (int)blockIdx.x >> 1;
    const int n_all = nq - jb, h0 = (n_all + 1) / 2;
    split = true;
    if (blockIdx.x & 1) { it0 = h0; n_it = n_all - h0; } else { it0 = 0; n_it = h0; }
  } else {
    jb = (int)blockIdx.x - n_split; it0 = 0; n_it = nq - jb; split = false;
  }
  const int i_base = jb + it0;                    // first query block of this CTA
  pdl_launch();

  if (warp == 0 && ptx::elect_one()) {
    ptx::prefetch_tmap(&tma_q); ptx::prefetch_tmap(&tma_k); ptx::prefetch_tmap(&tma_v); ptx::prefetch_tmap(&tma_do);
    ptx::mbar_init(bar(KV_FULL), 1);
    // Q / dO stages and the P / dS tiles are read by BOTH issuers: two commits release them
    for (int s = 0; s < 2; ++s) { ptx::mbar_init(bar(QDO_FULL + s), 1); ptx::mbar_init(bar(QDO_EMPTY + s), 2); }
    ptx::mbar_init(bar(SDP_FULL), 1); ptx::mbar_init(bar(SDP_FREE), 8);
    ptx::mbar_init(bar(PDS_FULL), 8); ptx::mbar_init(bar(PDS_FREE), 2);
    ptx::mbar_init(bar(DQ_FULL), 1);  ptx::mbar_init(bar(DQ_FREE), 8);
    ptx::mbar_init(bar(DKV_DONE), 1);
    ptx::fence_mbar_init();
  }
  if (warp == 1) { ptx::tmem_alloc(tmem_slot, 512); ptx::tmem_relinquish(); }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot_ptr;
  const uint32_t tS = tmem, tDP = tmem + 128, tDV = tmem + 256, tDK = tmem + 320, tDQ = tmem + 384;
  pdl_wait();

  if (warp == 0) {
    if (ptx::elect_one()) {
      ptx::mbar_expect_tx(bar(KV_FULL), 2 * kTileQK);
      ptx::tma_load_4d(sK, &tma_k, bar(KV_FULL), 0, jb * FB, h, b);
      ptx::tma_load_4d(sV, &tma_v, bar(KV_FULL), 0, jb * FB, h, b);
      for (int it = 0; it < n_it; ++it) {
        const int st = it & 1; const uint32_t ph2 = (it >> 1) & 1;
        ptx::mbar_wait_conv(bar(QDO_EMPTY + st), ph2 ^ 1u);
        ptx::mbar_expect_tx(bar(QDO_FULL + st), 2 * kTileQK);
        ptx::tma_load_4d(sQ + st * kTileQK, &tma_q, bar(QDO_FULL + st), 0, (i_base + it) * FB, h, b);
        ptx::tma_load_4d(sDO + st * kTileQK, &tma_do, bar(QDO_FULL + st), 0, (i_base + it) * FB, h, b);
      }
    }
  } else if (warp == 1) {
    // ===== issuer A: S = Q K^T, dP = dO V^T (8 MMAs) and dQ = dS K (8 MMAs) =====
    // One thread issues a tcgen05.mma only every ~80-100 cycles (tools/mma_probe.cu): the 32 MMAs of an (i, j) step cost one
    // issuer ~2900 cycles against ~1300 cycles of tensor-pipe work, so the step is split between two issuing warps.
    // Order: as soon as P_i / dS_i exist, S / dP of step i+1 go FIRST (the softmax warps wait for nothing else), dQ_i second
    // (its read-out by the softmax warps is deferred by one step, off the critical path).
    ptx::mbar_wait_conv(bar(KV_FULL), 0);
    auto issue_sdp = [&](int k) {
      const int st = k & 1; const uint32_t ph2 = (k >> 1) & 1, ph = k & 1;
      const uint32_t q_s = sQ + st * kTileQK, do_s = sDO + st * kTileQK;
      ptx::mbar_wait_conv(bar(QDO_FULL + st), ph2);
      ptx::mbar_wait_conv(bar(SDP_FREE), ph ^ 1u);       // S / dP of step k-1 have been read out of TMEM
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
#pragma unroll
        for (int kk = 0; kk < HS / 16; ++kk)
          ptx::mma_f16_ss(tS, ptx::make_smem_desc(q_s + kk * 32, 16, 1024), ptx::make_smem_desc(sK + kk * 32, 16, 1024),
                          g.idesc_s, kk > 0 ? 1u : 0u);
#pragma unroll
        for (int kk = 0; kk < HS / 16; ++kk)
          ptx::mma_f16_ss(tDP, ptx::make_smem_desc(do_s + kk * 32, 16, 1024), ptx::make_smem_desc(sV + kk * 32, 16, 1024),
                          g.idesc_s, kk > 0 ? 1u : 0u);
        ptx::mma_commit(bar(SDP_FULL));
      }
      __syncwarp();
    };
    issue_sdp(0);
    for (int it = 0; it < n_it; ++it) {
      const int st = it & 1; const uint32_t ph = it & 1;
      ptx::mbar_wait_conv(bar(PDS_FULL), ph);
      if (it + 1 < n_it) issue_sdp(it + 1);
      ptx::mbar_wait_conv(bar(DQ_FREE), ph ^ 1u);       // dQ of step it-1 has been read out of TMEM
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
#pragma unroll
        for (int k = 0; k < FB / 16; ++k)      // reduction over the 128 keys of block j
          ptx::mma_f16_ss(tDQ, ptx::make_smem_desc(sDS + (k >> 2) * (FB * 128) + (k & 3) * 32, 16, 1024),
                          ptx::make_smem_desc(sK + k * 2048, FB * 128, 1024), g.idesc_dq, k > 0 ? 1u : 0u);
        ptx::mma_commit(bar(DQ_FULL));
        ptx::mma_commit(bar(QDO_EMPTY + st));     // this thread's S / dP reads of Q, dO have retired (1 of 2 arrivals)
        ptx::mma_commit(bar(PDS_FREE));           // ... and its dQ read of dS (1 of 2 arrivals)
      }
      __syncwarp();
    }
  } else if (warp == 10) {
    // ===== issuer B: dV += P^T dO, dK += dS^T Q (16 MMAs, both operands MN-major) =====
    ptx::mbar_wait_conv(bar(KV_FULL), 0);
    for (int it = 0; it < n_it; ++it) {
      const int st = it & 1; const uint32_t ph2 = (it >> 1) & 1, ph = it & 1;
      const uint32_t q_s = sQ + st * kTileQK, do_s = sDO + st * kTileQK;
      ptx::mbar_wait_conv(bar(QDO_FULL + st), ph2);     // TMA-written Q / dO visible to this warp
      ptx::mbar_wait_conv(bar(PDS_FULL), ph);           // P, dS tiles written by the softmax warps
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
#pragma unroll
        for (int k = 0; k < FB / 16; ++k) {    // reduction over the 128 queries of block i
          const uint32_t acc = (it > 0 || k > 0) ? 1u : 0u;
          ptx::mma_f16_ss(tDV, ptx::make_smem_desc(sP + k * 2048, FB * 128, 1024), ptx::make_smem_desc(do_s + k * 2048, FB * 128, 1024),
                          g.idesc_dkv, acc);
          ptx::mma_f16_ss(tDK, ptx::make_smem_desc(sDS + k * 2048, FB * 128, 1024), ptx::make_smem_desc(q_s + k * 2048, FB * 128, 1024),
                          g.idesc_dkv, acc);
        }
        ptx::mma_commit(bar(QDO_EMPTY + st));     // second arrival: the stage may be refilled
        ptx::mma_commit(bar(PDS_FREE));           // second arrival: P / dS may be overwritten
        if (it == n_it - 1) ptx::mma_commit(bar(DKV_DONE));   // dV_j, dK_j complete -> epilogue
      }
      __syncwarp();
    }
  } else {
    // ===== warps 2-9: two threads per query row (warps w, w + 4 share a TMEM lane quadrant); each owns 64 of the 128 key
    // columns of S / dP / P / dS, 32 of the 64 dQ columns, and one of dV / dK in the epilogue.  lse and D are per-row inputs,
    // so the two threads never have to talk to each other. =====
    const int q = warp & 3;
    const int hf = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
    const uint32_t stg = sStg + (uint32_t)(hf * 4 + q) * 4096u;       // 4 KB per warp: 32 x 32 fp32 (dQ) or 32 x 64 bf16 (dV / dK)
    // dQ_i contribution of this key block: TMEM -> fp32 staging -> TMA reduce-add (this warp: rows q*32.., columns hf*32..)
    auto dq_readout = [&](int k) {
      ptx::mbar_wait(bar(DQ_FULL), (uint32_t)(k & 1));
      ptx::tc_fence_after();
      if (lane == 0) ptx::bulk_wait_read<0>();
      __syncwarp();
      uint32_t raw[32];
      ptx::tmem_ld_32x32(tDQ + lane_sel + hf * 32, raw);
      ptx::tmem_ld_wait();
#pragma unroll
      for (int c16 = 0; c16 < 8; ++c16) {
        const uint4 v = make_uint4(raw[c16 * 4], raw[c16 * 4 + 1], raw[c16 * 4 + 2], raw[c16 * 4 + 3]);
        ptx::st_shared_16(stg + (uint32_t)lane * 128u + ((((uint32_t)c16) ^ (uint32_t)(lane & 7)) << 4), v);
      }
      ptx::tc_fence_before();
      ptx::fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        ptx::mbar_arrive(bar(DQ_FREE));
        const int grow = (int)(((size_t)b * g.nh + h) * g.T) + (i_base + k) * FB + q * 32;
        ptx::tma_reduce_add_2d(&tma_dq, stg, hf * 32, grow);
        ptx::bulk_commit();
      }
    };
    for (int it = 0; it < n_it; ++it) {
      const uint32_t ph = it & 1;
      const int i = i_base + it;
      const bool diag = (i == jb);
      const size_t ridx = ((size_t)b * g.nh + h) * g.T + (size_t)i * FB + row;
      const float lse_r = g.lse[ridx], d_r = g.dsum[ridx];
      const int lim = row - hf * 64;               // diagonal block: column k of this half is visible iff k <= lim
      ptx::mbar_wait(bar(SDP_FULL), ph);
      ptx::tc_fence_after();
      bf16x8 pp[8], dd[8];                         // this thread's 64 P and 64 dS values, packed: computed before P / dS are free
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t rs[32], rp[32];
        ptx::tmem_ld_32x32(tS + lane_sel + hf * 64 + c * 32, rs);
        ptx::tmem_ld_32x32(tDP + lane_sel + hf * 64 + c * 32, rp);
        ptx::tmem_ld_wait();
        float p[32], ds[32];
        if (diag) {
#pragma unroll
          for (int k = 0; k < 32; ++k) {
            const float e = exp2f(__uint_as_float(rs[k]) * g.cs - lse_r);
            p[k] = (c * 32 + k <= lim) ? e : 0.f;
            ds[k] = p[k] * (__uint_as_float(rp[k]) - d_r) * g.scale;
          }
        } else {
#pragma unroll
          for (int k = 0; k < 32; ++k) {
            p[k] = exp2f(__uint_as_float(rs[k]) * g.cs - lse_r);
            ds[k] = p[k] * (__uint_as_float(rp[k]) - d_r) * g.scale;
          }
        }
#pragma unroll
        for (int j8 = 0; j8 < 4; ++j8) { pp[c * 4 + j8] = pack8(&p[j8 * 8]); dd[c * 4 + j8] = pack8(&ds[j8 * 8]); }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(bar(SDP_FREE));        // S / dP may be overwritten by step it+1
      ptx::mbar_wait(bar(PDS_FREE), ph ^ 1u);                // the 24 MMAs that read P / dS of step it-1 have retired
      const uint32_t off = (uint32_t)hf * (FB * 128) + (uint32_t)row * 128u;
#pragma unroll
      for (int j8 = 0; j8 < 8; ++j8) {
        const uint32_t sw = (((uint32_t)j8) ^ (uint32_t)(row & 7)) << 4;
        ptx::st_shared_16(sP + off + sw, pp[j8]);
        ptx::st_shared_16(sDS + off + sw, dd[j8]);
      }
      ptx::fence_proxy_async();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(bar(PDS_FULL));
      if (it > 0) dq_readout(it - 1);              // deferred by one step: dQ_{it-1} was issued after S / dP of this step
    }
    dq_readout(n_it - 1);
    // epilogue: dV_j (hf = 0) or dK_j (hf = 1), rows = keys
    ptx::mbar_wait(bar(DKV_DONE), 0);
    ptx::tc_fence_after();
    const uint32_t src = hf == 0 ? tDV : tDK;
    if (!split) {
      // whole key block in this CTA: -> bf16 -> staging -> TMA store into the V / K slice of dqkv
      if (lane == 0) ptx::bulk_wait_read<0>();
      __syncwarp();
      const uint32_t dst = stg + (uint32_t)lane * 128u;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t raw[32];
        ptx::tmem_ld_32x32(src + lane_sel + c * 32, raw);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int j8 = 0; j8 < 4; ++j8) {
          float t[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) t[k] = __uint_as_float(raw[j8 * 8 + k]);
          ptx::st_shared_16(dst + ((((uint32_t)(c * 4 + j8)) ^ (uint32_t)(lane & 7)) << 4), pack8(t));
        }
      }
      ptx::tc_fence_before();
      ptx::fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        if (hf == 0) ptx::tma_store_4d(&tma_dv, stg, 0, jb * FB + q * 32, h, b);
        else ptx::tma_store_4d(&tma_dk, stg, 0, jb * FB + q * 32, h, b);
        ptx::bulk_commit();
        ptx::bulk_wait_read<0>();
      }
    } else {
      // half of the query range: fp32 partial -> staging (32 x 32 slabs) -> TMA reduce-add into the dK / dV workspace
      const int grow = (int)(((size_t)b * g.nh + h) * g.T) + jb * FB + q * 32;
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        if (lane == 0) ptx::bulk_wait_read<0>();
        __syncwarp();
        uint32_t raw[32];
        ptx::tmem_ld_32x32(src + lane_sel + c * 32, raw);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int c16 = 0; c16 < 8; ++c16) {
          const uint4 v = make_uint4(raw[c16 * 4], raw[c16 * 4 + 1], raw[c16 * 4 + 2], raw[c16 * 4 + 3]);
          ptx::st_shared_16(stg + (uint32_t)lane * 128u + ((((uint32_t)c16) ^ (uint32_t)(lane & 7)) << 4), v);
        }
        ptx::fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          ptx::tma_reduce_add_2d(hf == 0 ? &tma_dvw : &tma_dkw, stg, c * 32, grow);
          ptx::bulk_commit();
        }
      }
      ptx::tc_fence_before();
      if (lane == 0) ptx::bulk_wait_read<0>();
    }
    __syncwarp();
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem, 512); }
}

// D[b,h,t] = sum_d dO[b,t,h,d] * O[b,t,h,d]   (one thread per (b,t,h): 8 x 16-byte loads from each tensor)
// Also clears the fp32 dQ / dK / dV workspaces (3 x 8 floats per thread: the grids coincide) — one graph node less per layer
// than a memset.
__global__ void flash_dsum_kernel(const __nv_bfloat16* __restrict__ dout, const __nv_bfloat16* __restrict__ out,
                                  float* __restrict__ dsum, float* __restrict__ dq_ws, int B, int T, int nh, int t_split) {
  pdl_launch(); pdl_wait();
  {
    const size_t gid0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t n8 = (size_t)B * T * nh * 8;
    if (gid0 < n8) {
      // workspace rows are (b, h, t): dK / dV partial sums exist only for the split key blocks, t < t_split
      const int nw = (int)((gid0 >> 3) % (size_t)T) < t_split ? 3 : 1;
      for (int w = 0; w < nw; ++w) {
        float4* z = reinterpret_cast<float4*>(dq_ws) + ((size_t)w * n8 + gid0) * 2;
        z[0] = make_float4(0.f, 0.f, 0.f, 0.f);
        z[1] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  }
  // 8 lanes per (b,t,h) row: one 16-byte load from each tensor per lane (fully coalesced), 3-step shuffle reduce
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  const int idx = gid >> 3, v8 = gid & 7;
  const bool ok = idx < B * T * nh;
  float acc = 0.f;
  int hh = 0, t = 0, bb = 0;
  if (ok) {
    hh = idx % nh; t = (idx / nh) % T; bb = idx / (nh * T);
    const size_t off = ((size_t)bb * T + t) * (size_t)(nh * HS) + (size_t)hh * HS + v8 * 8;
    float a[8], c[8];
    unpack8(ld8(dout + off), a);
    unpack8(ld8(out + off), c);
#pragma unroll
    for (int k = 0; k < 8; ++k) acc += a[k] * c[k];
  }
  acc += __shfl_xor_sync(0xffffffffu, acc, 1);
  acc += __shfl_xor_sync(0xffffffffu, acc, 2);
  acc += __shfl_xor_sync(0xffffffffu, acc, 4);
  if (ok && v8 == 0) dsum[((size_t)bb * nh + hh) * T + t] = acc;
}

// workspaces fp32 [3][B, nh, T, 64] (dQ, dK, dV) -> bf16 into the Q / K / V slices of dqkv: dQ for every row, dK / dV for the
// key rows t < t_split whose key blocks were split over two CTAs (the others were stored by the backward kernel itself)
__global__ void flash_dq_convert_kernel(const float* __restrict__ ws, __nv_bfloat16* __restrict__ dqkv, int B, int T, int nh, int t_split) {
  pdl_launch(); pdl_wait();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;       // one thread per 8 output elements
  const int per_row = HS / 8;
  if (idx >= B * nh * T * per_row) return;
  const int v8 = idx % per_row;
  const int t = (idx / per_row) % T, hh = (idx / (per_row * T)) % nh, bb = idx / (per_row * T * nh);
  const size_t n_ws = (size_t)B * nh * T * HS;
  const int C = nh * HS;
  const int nw = t < t_split ? 3 : 1;
  for (int w = 0; w < nw; ++w) {
    const float* src = ws + (size_t)w * n_ws + (((size_t)bb * nh + hh) * T + t) * HS + v8 * 8;
    const float4 a = reinterpret_cast<const float4*>(src)[0], c = reinterpret_cast<const float4*>(src)[1];
    float f[8] = {a.x, a.y, a.z, a.w, c.x, c.y, c.z, c.w};
    st8(dqkv + ((size_t)bb * T + t) * (size_t)(3 * C) + (size_t)w * C + (size_t)hh * HS + v8 * 8, pack8(f));
  }
}

}  // namespace

bool flash_supported(int T, int hs) { return hs == HS && T % FB == 0 && T >= FB; }

// dqkv receives dK, dV (TMA stores, or converted fp32 partial sums for the split key blocks) and dQ (converted from the fp32
// workspace).  `dq_ws`: 3 x [B, nh, T, 64] floats (dQ, dK, dV), zeroed by this call.
void flash_bwd(const void* qkv, const void* y, const void* dy, const float* lse, float* dsum, float* dq_ws, void* dqkv,
               int B, int T, int nh, float scale, cudaStream_t stream) {
  const int C = nh * HS;
  const __nv_bfloat16* base = reinterpret_cast<const __nv_bfloat16*>(qkv);
  __nv_bfloat16* dbase = reinterpret_cast<__nv_bfloat16*>(dqkv);
  CUtensorMap tq, tk, tv, tdo, tdk, tdv, tdq, tdkw, tdvw;
  const size_t n_ws = (size_t)B * nh * T * HS;
  auto view = [&](const void* p, int64_t ld, int64_t bs) { return GemmOperand{p, ld, bs, (int64_t)HS, false}; };
  bool ok = make_map(&tq, view(base, 3 * C, (int64_t)T * 3 * C), T, HS, B, nh, FB) &&
            make_map(&tk, view(base + C, 3 * C, (int64_t)T * 3 * C), T, HS, B, nh, FB) &&
            make_map(&tv, view(base + 2 * C, 3 * C, (int64_t)T * 3 * C), T, HS, B, nh, FB) &&
            make_map(&tdo, view(dy, C, (int64_t)T * C), T, HS, B, nh, FB) &&
            make_map(&tdk, view(dbase + C, 3 * C, (int64_t)T * 3 * C), T, HS, B, nh, 32) &&
            make_map(&tdv, view(dbase + 2 * C, 3 * C, (int64_t)T * 3 * C), T, HS, B, nh, 32) &&
            make_map_f32_2d(&tdq, dq_ws, (int64_t)B * nh * T, HS, HS, 32, 32) &&
            make_map_f32_2d(&tdkw, dq_ws + n_ws, (int64_t)B * nh * T, HS, HS, 32, 32) &&
            make_map_f32_2d(&tdvw, dq_ws + 2 * n_ws, (int64_t)B * nh * T, HS, HS, 32, 32);
  if (!ok) { fprintf(stderr, "[tds] flash_bwd: tensor map creation failed\n"); abort(); }
  static_assert(HS == 64, "flash_dsum_kernel clears 8 floats of the dQ workspace per thread (8 threads per row of 64)");
  // TDS_FLASH_SPLIT=0: one CTA per key block (no dK / dV workspace traffic); default: split the long key blocks in two
  static const bool want_split = !(getenv("TDS_FLASH_SPLIT") && atoi(getenv("TDS_FLASH_SPLIT")) == 0);
  const int nq = T / FB, n_split = want_split ? nq - (nq + 1) / 2 : 0;
  const int nthreads = B * T * nh * 8;
  launch_k(flash_dsum_kernel, dim3((nthreads + 255) / 256), dim3(256), 0, stream, (const __nv_bfloat16*)dy,
           (const __nv_bfloat16*)y, dsum, dq_ws, B, T, nh, n_split * FB);
  FlashBwdDev g;
  g.n_split = n_split;
  g.T = T; g.nh = nh; g.cs = scale * 1.4426950408889634f; g.scale = scale; g.lse = lse; g.dsum = dsum;
  g.idesc_s = make_idesc_bf16(FB, FB, false, false);
  g.idesc_dkv = make_idesc_bf16(FB, HS, true, true);
  g.idesc_dq = make_idesc_bf16(FB, HS, false, true);
  static bool attr = false;
  if (!attr) { cudaFuncSetAttribute(flash_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBwdSmem); attr = true; }
  launch_k(flash_bwd_kernel, dim3(nq + n_split, nh, B), dim3(kFBThreads), kBwdSmem, stream, tq, tk, tv, tdo, tdk, tdv, tdq, tdkw, tdvw, g);
  const int nconv = B * nh * T * (HS / 8);
  launch_k(flash_dq_convert_kernel, dim3((nconv + 255) / 256), dim3(256), 0, stream, (const float*)dq_ws, dbase, B, T, nh, n_split * FB);
}

void flash_fwd(const void* qkv, void* y, float* lse, int B, int T, int nh, float scale, cudaStream_t stream) {
  const int C = nh * HS;
  const __nv_bfloat16* base = reinterpret_cast<const __nv_bfloat16*>(qkv);
  CUtensorMap tq, tk, tv, to;
  auto view = [&](const void* p, int64_t ld, int64_t bs) { return GemmOperand{p, ld, bs, (int64_t)HS, false}; };
  bool ok = make_map(&tq, view(base, 3 * C, (int64_t)T * 3 * C), T, HS, B, nh, FB) &&
            make_map(&tk, view(base + C, 3 * C, (int64_t)T * 3 * C), T, HS, B, nh, FB) &&
            make_map(&tv, view(base + 2 * C, 3 * C, (int64_t)T * 3 * C), T, HS, B, nh, FB) &&
            make_map(&to, view(y, C, (int64_t)T * C), T, HS, B, nh, 32);
  if (!ok) { fprintf(stderr, "[tds] flash_fwd: tensor map creation failed\n"); abort(); }
  FlashDev g;
  g.T = T; g.nh = nh; g.cs = scale * 1.4426950408889634f; g.lse = lse;
  g.idesc_s = make_idesc_bf16(FB, FB, false, false);
  g.idesc_pv = make_idesc_bf16(FB, HS, false, true);
  static bool attr = false;
  if (!attr) { cudaFuncSetAttribute(flash_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFwdSmem); attr = true; }
  launch_k(flash_fwd_kernel, dim3(T / FB, nh, B), dim3(kFThreads), kFwdSmem, stream, tq, tk, tv, to, g);
}

}  // namespace tds
