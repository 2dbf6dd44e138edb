// Device-initiated collectives over NVLink 5 / NVSwitch symmetric memory (comm_sm100.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.h"

namespace tds {

constexpr int kMaxRanks = 8;
constexpr int kCommMaxBlocks = 128;

// One symmetric allocation as seen from this process: the same buffer on every rank, peer-mapped, plus (when the
// fabric supports it) one multicast address that aliases all of them (NVLS).
struct SymmBuf {
  void* peer[kMaxRanks];   // peer[r] = rank r's copy mapped in THIS address space (peer[rank] is local)
  void* mc;                // multicast address or nullptr
};
struct CommCtx {
  uint32_t* flags[kMaxRanks];  // per-rank flag pads (symmetric, zero-initialised), >= kCommMaxBlocks*kMaxRanks*4 words
  int rank, world;
  int* error_flag;             // local sticky error word (timeouts)
  long long spin_limit;        // SM cycles a flag wait may spin before it reports + traps (TDS_COMM_TIMEOUT_S, default 600 s)
};

// in-place sum all-reduce of buf[off, off+numel) (bf16, or fp32 when is_f32) across ranks; two-shot, one kernel
void allreduce(const CommCtx& c, const SymmBuf& buf, int64_t elem_off, int64_t numel, bool is_f32, float scale,
               int blocks, int channel, cudaStream_t s);
// sum-reduce buf[off, off+numel) onto rank `dst` only (other ranks' copies untouched)
void reduce_to(const CommCtx& c, const SymmBuf& buf, int64_t elem_off, int64_t numel, bool is_f32, int dst, float scale,
               int blocks, int channel, cudaStream_t s);
// replicate rank src's buf[off, off+numel) into every rank
void broadcast_from(const CommCtx& c, const SymmBuf& buf, int64_t byte_off, int64_t nbytes, int src, int blocks,
                    int channel, cudaStream_t s);
// rank src pushes nbytes from its LOCAL pointer into dst[dst_byte_off ...) on every rank (ZeRO-3 parameter fetch)
void push_from(const CommCtx& c, const void* src_local, const SymmBuf& dst, int64_t dst_byte_off, int64_t nbytes,
               int src_rank, int blocks, int channel, cudaStream_t s);
// every rank multicasts ITS slot [base + rank * slot_bytes, + slot_bytes) of buf to all ranks (slot_bytes % 16 == 0)
void allgather_slots(const CommCtx& c, const SymmBuf& buf, int64_t base_byte_off, int64_t slot_bytes, int blocks, int channel,
                     cudaStream_t s);
// row-sparse bf16 all-reduce of the [vocab, row_bytes] table at table_byte_off: only the rows named in ids[0, nids) (duplicates
// allowed; identical list on every rank), each reduced once by rank row % world.  epoch_of_row: int[vocab], never cleared;
// *epoch_ptr must differ from the previous call's value (the optimizer's device step counter).
void allreduce_rows(const CommCtx& c, const SymmBuf& buf, int64_t table_byte_off, int64_t row_bytes, const int64_t* ids, int nids,
                    int64_t vocab, int* epoch_of_row, const int* epoch_ptr, int blocks, int channel, cudaStream_t s);
// cross-GPU barrier (all blocks of all ranks)
void barrier(const CommCtx& c, int channel, cudaStream_t s);

// ZeRO-1/2 fused step on the ranges this rank owns: switch-reduced gradient (multimem.ld_reduce / P2P sum) ->
// scale -> Adam on local fp32 master + moments -> new bf16 parameter multicast to every rank (multimem.st / P2P).
constexpr int kMaxRanges = 320;
struct OwnedRanges {
  int64_t elem_off[kMaxRanges];   // offset in the flat GRAD buffer (elements, multiple of 8)
  int64_t pelem_off[kMaxRanges];  // offset in the PARAM buffer (differs from elem_off for ZeRO-3's owner-only layout)
  int64_t numel[kMaxRanges];      // multiple of 8 (padded)
  int64_t state_off[kMaxRanges];  // offset in the compact local fp32 state arrays
  int blk_start[kMaxRanges + 1];
  int count;
};
struct OwnedRangesRS {            // fused reduce-scatter path only (separate type: the default kernel's parameter layout is untouched)
  OwnedRanges r;
  uint8_t rs[kMaxRanges];         // 1 = gradient already summed in the local fp32 buffer
};
void zero_fused_adam(const CommCtx& c, const SymmBuf& grads, const SymmBuf& params, const OwnedRanges& r, float* master,
                     float* exp_avg, float* exp_avg_sq, const AdamHyper& h, bool bcast_params, int channel,
                     cudaStream_t s);
// EXPERIMENTAL variant for the GEMM -> reduce-scatter epilogue path (see comm_sm100.cu)
void zero_fused_adam_rs(const CommCtx& c, const SymmBuf& grads, const SymmBuf& params, const SymmBuf& rs, const OwnedRangesRS& r,
                        float* master, float* exp_avg, float* exp_avg_sq, const AdamHyper& h, bool bcast_params, int channel,
                        cudaStream_t s);
constexpr int kZeroChunk = 256 * 8 * 4;   // elements per CTA-iteration in zero_fused_adam

}  // namespace tds
