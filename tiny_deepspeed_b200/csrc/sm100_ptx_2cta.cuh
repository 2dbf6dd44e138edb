// PTX wrappers for CTA-pair (cta_group::2) tcgen05 kernels: two CTAs of a (2,1,1) cluster on the two SMs of one TPC
// cooperate on one 256-row UMMA; only the even ("leader") CTA issues MMAs and owns the full / tmem-empty barriers.
// EXPERIMENTAL: used by gemm2_sm100.cu only (opt-in), compiled but not yet run on hardware.
#pragma once
#include "sm100_ptx.cuh"

namespace tds {
namespace ptx {

// Shared-window addresses of the two CTAs of a pair differ in bit 24; clearing it names the leader's copy of the same
// offset (both for own-CTA and for peer addresses).
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;

// TMA load issued by EITHER CTA of the pair into its OWN smem; the complete_tx bytes go to the LEADER's mbarrier.
TDS_PTX void tma_load_4d_2cta(uint32_t dst_smem, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst_smem), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// D[tmem of both CTAs, 128 lanes each] (+)= A[256 x 16: 128 rows from each CTA's smem] . B[N x 16: N/2 rows from each CTA]
TDS_PTX void mma_f16_ss_2cta(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// All MMAs issued so far by this thread arrive (once) on the mbarrier at this offset in every CTA of cta_mask.
TDS_PTX void mma_commit_2cta(uint32_t bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(cta_mask) : "memory");
}
// Executed by the same warp index of BOTH CTAs: the same columns are reserved in both SMs' TMEM.
TDS_PTX void tmem_alloc_2cta(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
}
TDS_PTX void tmem_relinquish_2cta() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory"); }
TDS_PTX void tmem_dealloc_2cta(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// Arrive on the LEADER's copy of a barrier (from either CTA).
TDS_PTX void mbar_arrive_leader(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar & kPeerBitMask) : "memory");
}

}  // namespace ptx
}  // namespace tds
