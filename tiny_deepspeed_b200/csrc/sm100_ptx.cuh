// Thin inline-PTX wrappers for the Blackwell (sm_100a) async machinery: mbarrier, TMA (cp.async.bulk.tensor),
// tcgen05 (alloc / mma / commit / ld / fences) and the shared-memory + instruction descriptors.
// Bit layouts follow the PTX ISA "tcgen05 matrix/instruction descriptor" tables (same fields as
// cute::UMMA::SmemDescriptor / InstrDescriptor).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace tds { namespace ptx {

#define TDS_PTX __device__ __forceinline__

TDS_PTX uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// ---- mbarrier -----------------------------------------------------------------------------------------
TDS_PTX void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
TDS_PTX void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
TDS_PTX void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
TDS_PTX void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
TDS_PTX void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
TDS_PTX bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must trap (and surface as a CUDA error) instead of hanging the GPU.
TDS_PTX void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 8000000000LL) { asm volatile("trap;"); }
  }
}

// Wait as ONE asm block: the spin loop is invisible to the compiler, so a warp whose lanes all call it stays converged and
// warp-uniform values (stage, phase, descriptors) stay in uniform registers.  With the C++-level loop above, every
// tcgen05.mma / TMA issue that followed under `if (lane == 0)` was compiled into an ELECT + R2UR + BRA.U.ANY sequence
// (~70-120 cycles per instruction, measured with tools/gemm_harness trace) and the single issuing thread became the
// bottleneck of every M = 1024 GEMM.  Bounded like mbar_wait: ~2^22 polls with a 1 ms suspend hint, then trap.
TDS_PTX void mbar_wait_conv(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .u32 n;\n\t"
      "mov.u32 n, 0;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, 1000000;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "add.u32 n, n, 1;\n\t"
      "setp.lt.u32 p, n, 4194304;\n\t"
      "@p bra WAIT_LOOP;\n\t"
      "trap;\n\t"
      "WAIT_DONE:\n\t}"
      ::"r"(bar), "r"(parity)
      : "memory");
}
// Exactly one lane of the (converged) warp gets `true`; the compiler knows the guarded region is single-threaded, so
// uniform-datapath instructions (UTMALDG / UTCHMMA / UTCBAR / UTMASTG) are emitted straight, without an election loop.
TDS_PTX bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ---- TMA ----------------------------------------------------------------------------------------------
TDS_PTX void prefetch_tmap(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
TDS_PTX void tma_load_4d(uint32_t dst_smem, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst_smem), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// multicast variant: the box lands at the same smem offset of every CTA in `cta_mask`, each CTA's mbarrier (same offset)
// receives the complete_tx bytes
TDS_PTX void tma_load_4d_mc(uint32_t dst_smem, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2, int c3,
                            uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%3, %4, %5, %6}], [%2], %7;"
      ::"r"(dst_smem), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "h"(cta_mask)
      : "memory");
}

// ---- thread-block clusters ---------------------------------------------------------------------------------
TDS_PTX uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
TDS_PTX void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

TDS_PTX void tma_store_4d(const CUtensorMap* m, uint32_t src_smem, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(src_smem), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
// element-wise add of a smem tile into global memory, performed by the TMA/L2 (no register traffic, no atomics in SASS)
TDS_PTX void tma_reduce_add_2d(const CUtensorMap* m, uint32_t src_smem, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(src_smem), "r"(c0), "r"(c1)
               : "memory");
}
TDS_PTX void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// ask L2 to fetch `bytes` (multiple of 16) at a 16-byte aligned global address; no destination, no completion to wait for
TDS_PTX void prefetch_l2_bulk(const void* gptr, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gptr), "r"(bytes) : "memory");
}
template <int N> TDS_PTX void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <typename T16> TDS_PTX void st_shared_16(uint32_t addr, const T16& v) {
  static_assert(sizeof(T16) == 16, "16-byte payload");
  const uint4 u = *reinterpret_cast<const uint4*>(&v);
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(u.x), "r"(u.y), "r"(u.z), "r"(u.w) : "memory");
}

template <typename T16> TDS_PTX T16 ld_shared_16(uint32_t addr) {
  static_assert(sizeof(T16) == 16, "16-byte payload");
  uint4 u;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "r"(addr) : "memory");
  T16 v;
  *reinterpret_cast<uint4*>(&v) = u;
  return v;
}

// ---- tcgen05 ------------------------------------------------------------------------------------------
TDS_PTX void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
}
TDS_PTX void tmem_relinquish() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory"); }
TDS_PTX void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
TDS_PTX void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
TDS_PTX void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc]; bf16/fp16 inputs, fp32 accumulate.
TDS_PTX void mma_f16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// fp32 operands consumed as TF32 (10-bit mantissa), UMMA K = 8, fp32 accumulate.
TDS_PTX void mma_tf32_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// All previously issued tcgen05.mma of this thread arrive on the mbarrier when they complete.
TDS_PTX void mma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// ... and on the same barrier of every CTA in `cta_mask` (stage release when the operand tile was multicast)
TDS_PTX void mma_commit_mc(uint32_t bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(cta_mask) : "memory");
}

// 32 lanes x 32 consecutive fp32 columns: thread i of the warp receives row (lane base + i), columns [c, c+32)
TDS_PTX void tmem_ld_32x32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
TDS_PTX void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- descriptors ----------------------------------------------------------------------------------------
// Shared-memory matrix descriptor, SWIZZLE_128B, Blackwell version field = 1.
//   K-major  tile [rows][64 x bf16]: 8-row groups are 1024 B apart  -> SBO = 1024, LBO unused.
//   MN-major tile [k rows][64 x bf16] per 64-wide MN group: 8-k-row groups 1024 B apart -> SBO = 1024,
//            consecutive 64-wide MN groups LBO bytes apart.
//   MN-major fp32 (TF32) tile: the only legal layout is SWIZZLE_128B_BASE32B (layout type 1: 32-byte chunks XOR-ed with the
//            k-row index mod 4, = TMA CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B): 4-k-row groups 512 B apart -> SBO = 512.
TDS_PTX uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type = 2) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;   // descriptor version (sm_100)
  d |= static_cast<uint64_t>(layout_type) << 61;   // 2 = SWIZZLE_128B, 1 = SWIZZLE_128B_BASE32B
  return d;
}

}  // namespace ptx

// Instruction descriptor for kind::f16 with bf16 A/B and fp32 accumulate (host+device).
__host__ __device__ inline uint32_t make_idesc_bf16(int M, int N, bool a_mn, bool b_mn) {
  uint32_t d = 0;
  d |= 1u << 4;                      // c_format  = F32
  d |= 1u << 7;                      // a_format  = BF16
  d |= 1u << 10;                     // b_format  = BF16
  d |= (a_mn ? 1u : 0u) << 15;       // a_major   (0 = K-major, 1 = MN-major)
  d |= (b_mn ? 1u : 0u) << 16;       // b_major
  d |= static_cast<uint32_t>(N >> 3) << 17;
  d |= static_cast<uint32_t>(M >> 4) << 24;
  return d;
}

// Instruction descriptor for kind::tf32 (fp32 storage read as TF32) with fp32 accumulate.
__host__ __device__ inline uint32_t make_idesc_tf32(int M, int N, bool a_mn, bool b_mn) {
  uint32_t d = 0;
  d |= 1u << 4;                      // c_format  = F32
  d |= 2u << 7;                      // a_format  = TF32
  d |= 2u << 10;                     // b_format  = TF32
  d |= (a_mn ? 1u : 0u) << 15;
  d |= (b_mn ? 1u : 0u) << 16;
  d |= static_cast<uint32_t>(N >> 3) << 17;
  d |= static_cast<uint32_t>(M >> 4) << 24;
  return d;
}

}  // namespace tds
