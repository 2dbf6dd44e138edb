// Device-initiated collectives for sm_100a over NVLink 5 / NVSwitch.
//
// Data path: every kernel here addresses peer memory directly — `multimem.ld_reduce` / `multimem.st` on the
// NVLS multicast mapping (the switch reduces / replicates in-fabric), or plain 16-byte loads/stores on the
// peer-mapped unicast pointers when no multicast object exists.  No NCCL on this path; torch.distributed is
// used only once, to exchange the memory handles.  Synchronisation is device-side: per-block flag words in a
// symmetric pad, set with release / consumed with acquire CAS at .sys scope (bounded spins: a dead peer turns
// into a trap + diagnostic, never a silent GPU hang — SURVEY §5 "failure detection").
//
// Replaces the reference's per-tensor dist.all_reduce / dist.reduce / dist.broadcast + cuda.synchronize()
// (tiny_deepspeed/core/zero/ddp/module.py:17-24, zero1/module.py:17-24, zero1/optim.py:20-34).
#include <stdio.h>
#include <stdlib.h>

#include "comm.h"
#include "common.cuh"

namespace tds {

constexpr int kCommThreads = 512;
// Flag waits are bounded by CommCtx::spin_limit (SM cycles; host default 600 s — long enough for rank-skewed host work such as
// a rank-0 checkpoint or a first-step compile): on expiry the rank prints a diagnostic, raises the sticky error word
// the host watchdog polls, and traps so the failure surfaces as a CUDA error instead of a silent hang.

// ---- .sys-scope flag primitives -----------------------------------------------------------------------
TDS_DEVICE uint32_t cas_release_sys(uint32_t* p, uint32_t cmp, uint32_t val) {
  uint32_t old;
  asm volatile("atom.global.release.sys.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "l"(p), "r"(cmp), "r"(val) : "memory");
  return old;
}
TDS_DEVICE uint32_t cas_acquire_sys(uint32_t* p, uint32_t cmp, uint32_t val) {
  uint32_t old;
  asm volatile("atom.global.acquire.sys.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "l"(p), "r"(cmp), "r"(val) : "memory");
  return old;
}

TDS_DEVICE uint32_t* flag_slot(const CommCtx& c, int owner, int channel, int block, int peer) {
  return c.flags[owner] + ((size_t)channel * kCommMaxBlocks + block) * kMaxRanks + peer;
}

TDS_DEVICE void spin_fail(const CommCtx& c, const char* what, int peer) {
  printf("[tds comm] rank %d block %d: timeout in %s waiting on rank %d\n", c.rank, blockIdx.x, what, peer);
  if (c.error_flag) *c.error_flag = 1;
  asm volatile("trap;");
}

// Barrier between block b of every rank.  Thread r talks to rank r: raises the peer's slot, then consumes its own.
TDS_DEVICE void block_barrier(const CommCtx& c, int channel) {
  __syncthreads();
  if (threadIdx.x < c.world) {
    const int peer = threadIdx.x;
    uint32_t* theirs = flag_slot(c, peer, channel, blockIdx.x, c.rank);
    uint32_t* mine = flag_slot(c, c.rank, channel, blockIdx.x, peer);
    long long t0 = clock64();
    while (cas_release_sys(theirs, 0u, 1u) != 0u)
      if (clock64() - t0 > c.spin_limit) spin_fail(c, "barrier(signal)", peer);
    t0 = clock64();
    while (cas_acquire_sys(mine, 1u, 0u) != 1u)
      if (clock64() - t0 > c.spin_limit) spin_fail(c, "barrier(wait)", peer);
  }
  __syncthreads();
}

// ---- 16-byte multimem / peer accessors ---------------------------------------------------------------------
TDS_DEVICE uint4 mm_ld_reduce_bf16(const void* mc) {
  uint4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(mc) : "memory");
  return v;
}
TDS_DEVICE uint4 mm_ld_reduce_f32(const void* mc) {
  uint4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(mc) : "memory");
  return v;
}
TDS_DEVICE void mm_st(void* mc, uint4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};"
               ::"l"(mc), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
TDS_DEVICE uint4 ld16(const void* p) {
  uint4 v;
  asm volatile("ld.global.relaxed.sys.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
TDS_DEVICE void st16(void* p, uint4 v) {
  asm volatile("st.global.relaxed.sys.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

TDS_DEVICE void acc_bf16x8(float* acc, uint4 v) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) { float2 f = __bfloat1622float2(h[i]); acc[2 * i] += f.x; acc[2 * i + 1] += f.y; }
}
TDS_DEVICE uint4 pack_bf16x8(const float* f) {
  uint4 v;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return v;
}

// sum over ranks of one 16-byte vector at byte offset `boff` — through the switch when possible
template <bool F32>
TDS_DEVICE void reduce_vec(const CommCtx& c, const SymmBuf& b, size_t boff, float* out /*8 (bf16) or 4 (f32)*/) {
  if (b.mc) {
    uint4 v = F32 ? mm_ld_reduce_f32((const char*)b.mc + boff) : mm_ld_reduce_bf16((const char*)b.mc + boff);
    if (F32) { out[0] = __uint_as_float(v.x); out[1] = __uint_as_float(v.y); out[2] = __uint_as_float(v.z); out[3] = __uint_as_float(v.w); }
    else { for (int i = 0; i < 8; ++i) out[i] = 0.f; acc_bf16x8(out, v); }
  } else {
    for (int i = 0; i < (F32 ? 4 : 8); ++i) out[i] = 0.f;
    for (int r = 0; r < c.world; ++r) {
      uint4 v = ld16((const char*)b.peer[(c.rank + r) % c.world] + boff);
      if (F32) { out[0] += __uint_as_float(v.x); out[1] += __uint_as_float(v.y); out[2] += __uint_as_float(v.z); out[3] += __uint_as_float(v.w); }
      else acc_bf16x8(out, v);
    }
  }
}
TDS_DEVICE void bcast_vec(const CommCtx& c, const SymmBuf& b, size_t boff, uint4 v) {
  if (b.mc) mm_st((char*)b.mc + boff, v);
  else for (int r = 0; r < c.world; ++r) st16((char*)b.peer[(c.rank + r) % c.world] + boff, v);
}

// =====================================================================================================
// all-reduce (two-shot: reduce-scatter slice -> all-gather slice), in place
// =====================================================================================================
template <bool F32>
__global__ void __launch_bounds__(kCommThreads) allreduce_kernel(const __grid_constant__ CommCtx c,
                                                                const __grid_constant__ SymmBuf b, long long boff,
                                                                long long nvec, float scale, int channel) {
  block_barrier(c, channel);                        // every rank's contribution is in place
  const long long per = (nvec + c.world - 1) / c.world;
  const long long v0 = per * c.rank, v1 = (v0 + per < nvec) ? v0 + per : nvec;
  // NVLink round trips are ~3 us: keep U independent 16-byte requests in flight per thread (Little's law:
  // blocks x threads x U x 16 B must cover ~770 GB/s x latency ~ 2.3 MB)
  constexpr int U = 8;
  const long long stride = (long long)gridDim.x * blockDim.x * U;
  for (long long base = v0 + (long long)blockIdx.x * blockDim.x * U + threadIdx.x; base < v1; base += stride) {
    if (b.mc && scale == 1.f) {
      uint4 v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long i = base + (long long)u * blockDim.x;
        if (i < v1) v[u] = F32 ? mm_ld_reduce_f32((const char*)b.mc + boff + i * 16) : mm_ld_reduce_bf16((const char*)b.mc + boff + i * 16);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) { const long long i = base + (long long)u * blockDim.x; if (i < v1) mm_st((char*)b.mc + boff + i * 16, v[u]); }
    } else {
#pragma unroll 2
      for (int u = 0; u < U; ++u) {
        const long long i = base + (long long)u * blockDim.x;
        if (i >= v1) break;
        float f[8];
        reduce_vec<F32>(c, b, (size_t)(boff + i * 16), f);
        uint4 o;
        if (F32) { o.x = __float_as_uint(f[0] * scale); o.y = __float_as_uint(f[1] * scale); o.z = __float_as_uint(f[2] * scale); o.w = __float_as_uint(f[3] * scale); }
        else { for (int j = 0; j < 8; ++j) f[j] *= scale; o = pack_bf16x8(f); }
        bcast_vec(c, b, (size_t)(boff + i * 16), o);
      }
    }
  }
  __threadfence_system();
  block_barrier(c, channel);                        // results visible everywhere before anyone returns
}

void allreduce(const CommCtx& c, const SymmBuf& buf, int64_t elem_off, int64_t numel, bool is_f32, float scale,
               int blocks, int channel, cudaStream_t s) {
  const int esz = is_f32 ? 4 : 2;
  const long long boff = (long long)elem_off * esz, nvec = ((long long)numel * esz + 15) / 16;
  if (blocks > kCommMaxBlocks) blocks = kCommMaxBlocks;
  if (is_f32) allreduce_kernel<true><<<blocks, kCommThreads, 0, s>>>(c, buf, boff, nvec, scale, channel);
  else allreduce_kernel<false><<<blocks, kCommThreads, 0, s>>>(c, buf, boff, nvec, scale, channel);
}

// =====================================================================================================
// reduce to one rank / broadcast from one rank
// =====================================================================================================
template <bool F32>
__global__ void __launch_bounds__(kCommThreads) reduce_to_kernel(const __grid_constant__ CommCtx c,
                                                                const __grid_constant__ SymmBuf b, long long boff,
                                                                long long nvec, int dst, float scale, int channel) {
  block_barrier(c, channel);
  if (c.rank == dst) {
    char* local = (char*)b.peer[c.rank];
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
      float f[8];
      reduce_vec<F32>(c, b, (size_t)(boff + i * 16), f);
      uint4 o;
      if (F32) { o.x = __float_as_uint(f[0] * scale); o.y = __float_as_uint(f[1] * scale); o.z = __float_as_uint(f[2] * scale); o.w = __float_as_uint(f[3] * scale); }
      else { for (int j = 0; j < 8; ++j) f[j] *= scale; o = pack_bf16x8(f); }
      *reinterpret_cast<uint4*>(local + boff + i * 16) = o;
    }
  }
  __threadfence_system();
  block_barrier(c, channel);   // peers may now overwrite their (consumed) contributions
}

void reduce_to(const CommCtx& c, const SymmBuf& buf, int64_t elem_off, int64_t numel, bool is_f32, int dst, float scale,
               int blocks, int channel, cudaStream_t s) {
  const int esz = is_f32 ? 4 : 2;
  const long long boff = (long long)elem_off * esz, nvec = ((long long)numel * esz + 15) / 16;
  if (blocks > kCommMaxBlocks) blocks = kCommMaxBlocks;
  if (is_f32) reduce_to_kernel<true><<<blocks, kCommThreads, 0, s>>>(c, buf, boff, nvec, dst, scale, channel);
  else reduce_to_kernel<false><<<blocks, kCommThreads, 0, s>>>(c, buf, boff, nvec, dst, scale, channel);
}

__global__ void __launch_bounds__(kCommThreads) broadcast_kernel(const __grid_constant__ CommCtx c,
                                                                const __grid_constant__ SymmBuf b, long long boff,
                                                                long long nvec, int src, int channel) {
  block_barrier(c, channel);   // nobody is still reading the old contents
  if (c.rank == src) {
    const char* local = (const char*)b.peer[c.rank];
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
      uint4 v = *reinterpret_cast<const uint4*>(local + boff + i * 16);
      if (b.mc) mm_st((char*)b.mc + boff + i * 16, v);
      else for (int r = 1; r < c.world; ++r) st16((char*)b.peer[(c.rank + r) % c.world] + boff + i * 16, v);
    }
  }
  __threadfence_system();
  block_barrier(c, channel);
}

void broadcast_from(const CommCtx& c, const SymmBuf& buf, int64_t byte_off, int64_t nbytes, int src, int blocks,
                    int channel, cudaStream_t s) {
  if (blocks > kCommMaxBlocks) blocks = kCommMaxBlocks;
  broadcast_kernel<<<blocks, kCommThreads, 0, s>>>(c, buf, (long long)byte_off, (long long)((nbytes + 15) / 16), src, channel);
}

// Owner-push of a tensor into a symmetric staging slot on EVERY rank (ZeRO-3 parameter fetch): the owner streams its
// local copy through `multimem.st` (the switch replicates: owner egress = 1x the tensor, whatever the world size) or,
// without multicast, through per-peer stores.  Consumers only take part in the two barriers.
__global__ void __launch_bounds__(kCommThreads) push_kernel(const __grid_constant__ CommCtx c, const char* __restrict__ src,
                                                           const __grid_constant__ SymmBuf dst, long long dst_boff,
                                                           long long nvec, int src_rank, int channel) {
  block_barrier(c, channel);   // every rank has released the slot's previous contents
  if (c.rank == src_rank) {
    constexpr int U = 4;
    // Two ranks: plain stores into the one peer.  A multicast store is replicated to EVERY member of the group, the source
    // included, so at world = 2 half of the owner's inbound NVLink bandwidth would carry its own data back (measured on GPT-2 XL
    // ZeRO-3: the echo doubles the 3.3 GB of parameter traffic each rank receives per step); from 4 ranks up the switch-side
    // replication (owner egress 1x instead of (N-1)x) wins.
    const bool unicast = dst.mc == nullptr || c.world == 2;
    const long long stride = (long long)gridDim.x * blockDim.x * U;
    for (long long base = (long long)blockIdx.x * blockDim.x * U + threadIdx.x; base < nvec; base += stride) {
      uint4 v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) { const long long i = base + (long long)u * blockDim.x; if (i < nvec) v[u] = *reinterpret_cast<const uint4*>(src + i * 16); }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long i = base + (long long)u * blockDim.x;
        if (i >= nvec) continue;
        if (unicast) {
          for (int r = 1; r < c.world; ++r) st16((char*)dst.peer[(c.rank + r) % c.world] + dst_boff + i * 16, v[u]);   // the owner reads its own parameters in place
        } else {
          mm_st((char*)dst.mc + dst_boff + i * 16, v[u]);
        }
      }
    }
  }
  __threadfence_system();
  block_barrier(c, channel);   // slot contents visible on every rank
}

void push_from(const CommCtx& c, const void* src_local, const SymmBuf& dst, int64_t dst_byte_off, int64_t nbytes,
               int src_rank, int blocks, int channel, cudaStream_t s) {
  if (blocks > kCommMaxBlocks) blocks = kCommMaxBlocks;
  push_kernel<<<blocks, kCommThreads, 0, s>>>(c, (const char*)src_local, dst, (long long)dst_byte_off,
                                             (long long)((nbytes + 15) / 16), src_rank, channel);
}

// =====================================================================================================
// Row-sparse all-reduce of an embedding gradient (DDP).  Backward touches at most `ntok` rows per rank of the [V, D]
// table (1024 of 50304 for GPT-2 small): instead of all-reducing 77 MB of mostly zeros at the very end of backward, the
// ranks exchange their token ids (kernel 1: every rank multicasts its ids into its slot of a symmetric buffer) and then
// all-reduce only the touched rows (kernel 2).  Every row is reduced EXACTLY once, by the rank `row % world`, with the
// same switch reduction + multicast store as the dense kernel, so all replicas receive identical bits; rows nobody
// touched stay zero on every rank (each rank zero-fills its own dense gradient before the scatter).
// A per-row epoch word (never cleared: the epoch is the optimizer step) deduplicates ids that occur several times.
// =====================================================================================================
__global__ void __launch_bounds__(256) allgather_slots_kernel(const __grid_constant__ CommCtx c,
                                                              const __grid_constant__ SymmBuf b, long long base_boff,
                                                              long long slot_vecs, int channel) {
  block_barrier(c, channel);                       // every rank has filled its own slot (stream order) and nobody still reads
  const long long my0 = base_boff + (long long)c.rank * slot_vecs * 16;
  const char* local = (const char*)b.peer[c.rank];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < slot_vecs; i += (long long)gridDim.x * blockDim.x) {
    const uint4 v = *reinterpret_cast<const uint4*>(local + my0 + i * 16);
    bcast_vec(c, b, (size_t)(my0 + i * 16), v);
  }
  __threadfence_system();
  block_barrier(c, channel);                       // all slots visible everywhere
}

void allgather_slots(const CommCtx& c, const SymmBuf& buf, int64_t base_byte_off, int64_t slot_bytes, int blocks, int channel,
                     cudaStream_t s) {
  if (blocks > kCommMaxBlocks) blocks = kCommMaxBlocks;
  allgather_slots_kernel<<<blocks, 256, 0, s>>>(c, buf, (long long)base_byte_off, (long long)(slot_bytes / 16), channel);
}

// ids: world * ntok gathered token ids (int64, identical on every rank); one warp per gathered id
__global__ void __launch_bounds__(256) allreduce_rows_kernel(const __grid_constant__ CommCtx c,
                                                             const __grid_constant__ SymmBuf b, long long table_boff,
                                                             int row_vecs, const long long* __restrict__ ids, int nids,
                                                             long long vocab, int* __restrict__ epoch_of_row,
                                                             const int* __restrict__ epoch_ptr, int channel) {
  block_barrier(c, channel);                       // every rank's scatter-add into its dense gradient is complete
  const int epoch = *epoch_ptr;
  const int lane = threadIdx.x & 31;
  const int warp = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5);
  const int nwarps = (int)((gridDim.x * (long long)blockDim.x) >> 5);
  for (int e = warp; e < nids; e += nwarps) {
    const long long id = ids[e];
    if (id < 0 || id >= vocab || (int)(id % c.world) != c.rank) continue;       // rows are dealt round-robin to the ranks
    int won = 0;
    if (lane == 0) won = atomicExch(epoch_of_row + id, epoch) != epoch;          // first occurrence of this row in this step
    won = __shfl_sync(0xffffffffu, won, 0);
    if (!won) continue;
    const size_t row = (size_t)table_boff + (size_t)id * (size_t)row_vecs * 16;
    for (int v = lane; v < row_vecs; v += 32) {
      float f[8];
      reduce_vec<false>(c, b, row + (size_t)v * 16, f);
      bcast_vec(c, b, row + (size_t)v * 16, pack_bf16x8(f));
    }
  }
  __threadfence_system();
  block_barrier(c, channel);                       // reduced rows visible on every rank
}

void allreduce_rows(const CommCtx& c, const SymmBuf& buf, int64_t table_byte_off, int64_t row_bytes, const int64_t* ids, int nids,
                    int64_t vocab, int* epoch_of_row, const int* epoch_ptr, int blocks, int channel, cudaStream_t s) {
  if (blocks > kCommMaxBlocks) blocks = kCommMaxBlocks;
  allreduce_rows_kernel<<<blocks, 256, 0, s>>>(c, buf, (long long)table_byte_off, (int)(row_bytes / 16),
                                               reinterpret_cast<const long long*>(ids), nids, (long long)vocab, epoch_of_row,
                                               epoch_ptr, channel);
}

__global__ void barrier_kernel(const __grid_constant__ CommCtx c, int channel) { block_barrier(c, channel); }
void barrier(const CommCtx& c, int channel, cudaStream_t s) { barrier_kernel<<<1, 32, 0, s>>>(c, channel); }

// =====================================================================================================
// ZeRO-1/2 fused step: reduce(grad) -> scale -> Adam(master, m, v) -> multicast(param)
// =====================================================================================================
TDS_DEVICE int find_range(const int* blk_start, int count, int b) {
  int lo = 0, hi = count - 1;
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    if (blk_start[mid] <= b) lo = mid; else hi = mid - 1;
  }
  return lo;
}

__global__ void __launch_bounds__(256) zero_fused_adam_kernel(const __grid_constant__ CommCtx c,
                                                              const __grid_constant__ SymmBuf grads,
                                                              const __grid_constant__ SymmBuf params,
                                                              const __grid_constant__ OwnedRanges R,
                                                              float* __restrict__ master, float* __restrict__ exp_avg,
                                                              float* __restrict__ exp_avg_sq,
                                                              const __grid_constant__ AdamHyper h, int bcast, int channel,
                                                              int work_blocks) {
  block_barrier(c, channel);                       // all ranks' gradients are complete
  __shared__ float s_bc[2];
  if (threadIdx.x == 0) {
    const int step = *h.step_ptr;
    s_bc[0] = 1.f - powf(h.beta1, (float)step);
    s_bc[1] = rsqrtf(1.f - powf(h.beta2, (float)step));
  }
  __syncthreads();
  const float bc1 = s_bc[0], bc2r = s_bc[1];
  char* local_param = (char*)params.peer[c.rank];
  for (int wb = blockIdx.x; wb < work_blocks; wb += gridDim.x) {
    const int t = find_range(R.blk_start, R.count, wb);
    const long long base = (long long)(wb - R.blk_start[t]) * kZeroChunk;
    const long long n = R.numel[t];
    const long long end = base + kZeroChunk < n ? base + kZeroChunk : n;
    constexpr int U = 4;               // gradient fetches in flight per thread (NVLink latency hiding)
    for (long long i0 = base + threadIdx.x * 8; i0 < end; i0 += 256 * 8 * U) {
      uint4 graw[U];
      float gsum[U][8];
      if (grads.mc) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const long long i = i0 + (long long)u * 256 * 8;
          if (i < end) graw[u] = mm_ld_reduce_bf16((const char*)grads.mc + (size_t)(R.elem_off[t] + i) * 2);
        }
      } else {
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const long long i = i0 + (long long)u * 256 * 8;
          if (i < end) reduce_vec<false>(c, grads, (size_t)(R.elem_off[t] + i) * 2, gsum[u]);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long i = i0 + (long long)u * 256 * 8;
        if (i >= end) break;
        float g[8];
        if (grads.mc) { for (int j = 0; j < 8; ++j) g[j] = 0.f; acc_bf16x8(g, graw[u]); }
        else { for (int j = 0; j < 8; ++j) g[j] = gsum[u][j]; }
        const long long so = R.state_off[t] + i;
        float w[8], m[8], v[8];
        *reinterpret_cast<float4*>(w) = *reinterpret_cast<const float4*>(master + so);
        *reinterpret_cast<float4*>(w + 4) = *reinterpret_cast<const float4*>(master + so + 4);
        *reinterpret_cast<float4*>(m) = *reinterpret_cast<const float4*>(exp_avg + so);
        *reinterpret_cast<float4*>(m + 4) = *reinterpret_cast<const float4*>(exp_avg + so + 4);
        *reinterpret_cast<float4*>(v) = *reinterpret_cast<const float4*>(exp_avg_sq + so);
        *reinterpret_cast<float4*>(v + 4) = *reinterpret_cast<const float4*>(exp_avg_sq + so + 4);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float gg = g[j] * h.grad_scale;
          if (h.maximize) gg = -gg;
          if (h.weight_decay != 0.f) {
            if (h.decoupled) w[j] *= (1.f - h.lr * h.weight_decay);
            else gg += h.weight_decay * w[j];
          }
          m[j] = h.beta1 * m[j] + (1.f - h.beta1) * gg;
          v[j] = h.beta2 * v[j] + (1.f - h.beta2) * gg * gg;
          w[j] -= (h.lr / bc1) * (m[j] / (sqrtf(v[j]) * bc2r + h.eps));
        }
        *reinterpret_cast<float4*>(master + so) = *reinterpret_cast<float4*>(w);
        *reinterpret_cast<float4*>(master + so + 4) = *reinterpret_cast<float4*>(w + 4);
        *reinterpret_cast<float4*>(exp_avg + so) = *reinterpret_cast<float4*>(m);
        *reinterpret_cast<float4*>(exp_avg + so + 4) = *reinterpret_cast<float4*>(m + 4);
        *reinterpret_cast<float4*>(exp_avg_sq + so) = *reinterpret_cast<float4*>(v);
        *reinterpret_cast<float4*>(exp_avg_sq + so + 4) = *reinterpret_cast<float4*>(v + 4);
        const uint4 o = pack_bf16x8(w);
        const size_t poff = (size_t)(R.pelem_off[t] + i) * 2;
        if (bcast) bcast_vec(c, params, poff, o);
        else *reinterpret_cast<uint4*>(local_param + poff) = o;
      }
    }
  }
  __threadfence_system();
  block_barrier(c, channel);                       // new parameters visible on every rank
}

static int step_blocks() {
  static const int n = [] {
    const char* e = getenv("TDS_STEP_BLOCKS");
    int v = e ? atoi(e) : kCommMaxBlocks;
    return v < 1 ? 1 : (v > kCommMaxBlocks ? kCommMaxBlocks : v);
  }();
  return n;
}

void zero_fused_adam(const CommCtx& c, const SymmBuf& grads, const SymmBuf& params, const OwnedRanges& r, float* master,
                     float* exp_avg, float* exp_avg_sq, const AdamHyper& h, bool bcast_params, int channel,
                     cudaStream_t s) {
  const int work = r.blk_start[r.count];
  // every rank launches the SAME grid (the barrier is per block index), whatever it owns
  // every rank launches the SAME grid (the barrier is per block index), whatever it owns; TDS_STEP_BLOCKS (<= 128) trades the
  // step kernel's bandwidth against the SM slots it takes from the backward GEMMs it runs under
  zero_fused_adam_kernel<<<step_blocks(), 256, 0, s>>>(c, grads, params, r, master, exp_avg, exp_avg_sq, h,
                                                        bcast_params ? 1 : 0, channel, work);
}


// =====================================================================================================
// EXPERIMENTAL (opt-in TDS_FUSED_RS=1, not yet run on multi-GPU hardware): fused step for the GEMM -> reduce-scatter path.
// Ranges flagged in R.rs were already summed over all ranks INTO this rank's fp32 reduction buffer `rs` (every rank's dW GEMM
// epilogue TMA-reduce-adds its tile into the owner, gemm_sm100.cu RED variant), at the same offset as the optimizer state;
// they are consumed from local memory and zeroed for the next step.  Unflagged ranges (embeddings, LayerNorm, biases) keep
// the switch-reduced multimem path.  Kept as a separate kernel so the default one above stays byte-identical.
// =====================================================================================================
__global__ void __launch_bounds__(256) zero_fused_adam_rs_kernel(const __grid_constant__ CommCtx c,
                                                                 const __grid_constant__ SymmBuf grads,
                                                                 const __grid_constant__ SymmBuf params,
                                                                 const __grid_constant__ SymmBuf rs,
                                                                 const __grid_constant__ OwnedRangesRS RR,
                                                                 float* __restrict__ master, float* __restrict__ exp_avg,
                                                                 float* __restrict__ exp_avg_sq,
                                                                 const __grid_constant__ AdamHyper h, int bcast, int channel,
                                                                 int work_blocks) {
  const OwnedRanges& R = RR.r;
  block_barrier(c, channel);                       // all ranks' backward kernels (and their remote reduce-adds) are complete
  __shared__ float s_bc[2];
  if (threadIdx.x == 0) {
    const int step = *h.step_ptr;
    s_bc[0] = 1.f - powf(h.beta1, (float)step);
    s_bc[1] = rsqrtf(1.f - powf(h.beta2, (float)step));
  }
  __syncthreads();
  const float bc1 = s_bc[0], bc2r = s_bc[1];
  char* local_param = (char*)params.peer[c.rank];
  float* local_rs = (float*)rs.peer[c.rank];
  for (int wb = blockIdx.x; wb < work_blocks; wb += gridDim.x) {
    const int t = find_range(R.blk_start, R.count, wb);
    const long long base = (long long)(wb - R.blk_start[t]) * kZeroChunk;
    const long long n = R.numel[t];
    const long long end = base + kZeroChunk < n ? base + kZeroChunk : n;
    const bool from_rs = RR.rs[t] != 0;
    for (long long i = base + threadIdx.x * 8; i < end; i += 256 * 8) {
      float g[8];
      const long long so = R.state_off[t] + i;
      if (from_rs) {
        // written by remote atomics at this GPU's L2: read around L1, then clear for the next step
        const float4 a = __ldcg(reinterpret_cast<const float4*>(local_rs + so));
        const float4 b = __ldcg(reinterpret_cast<const float4*>(local_rs + so + 4));
        g[0] = a.x; g[1] = a.y; g[2] = a.z; g[3] = a.w; g[4] = b.x; g[5] = b.y; g[6] = b.z; g[7] = b.w;
        __stcg(reinterpret_cast<float4*>(local_rs + so), make_float4(0.f, 0.f, 0.f, 0.f));
        __stcg(reinterpret_cast<float4*>(local_rs + so + 4), make_float4(0.f, 0.f, 0.f, 0.f));
      } else if (grads.mc) {
        const uint4 raw = mm_ld_reduce_bf16((const char*)grads.mc + (size_t)(R.elem_off[t] + i) * 2);
#pragma unroll
        for (int j = 0; j < 8; ++j) g[j] = 0.f;
        acc_bf16x8(g, raw);
      } else {
        reduce_vec<false>(c, grads, (size_t)(R.elem_off[t] + i) * 2, g);
      }
      float w[8], m[8], v[8];
      *reinterpret_cast<float4*>(w) = *reinterpret_cast<const float4*>(master + so);
      *reinterpret_cast<float4*>(w + 4) = *reinterpret_cast<const float4*>(master + so + 4);
      *reinterpret_cast<float4*>(m) = *reinterpret_cast<const float4*>(exp_avg + so);
      *reinterpret_cast<float4*>(m + 4) = *reinterpret_cast<const float4*>(exp_avg + so + 4);
      *reinterpret_cast<float4*>(v) = *reinterpret_cast<const float4*>(exp_avg_sq + so);
      *reinterpret_cast<float4*>(v + 4) = *reinterpret_cast<const float4*>(exp_avg_sq + so + 4);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float gg = g[j] * h.grad_scale;
        if (h.maximize) gg = -gg;
        if (h.weight_decay != 0.f) {
          if (h.decoupled) w[j] *= (1.f - h.lr * h.weight_decay);
          else gg += h.weight_decay * w[j];
        }
        m[j] = h.beta1 * m[j] + (1.f - h.beta1) * gg;
        v[j] = h.beta2 * v[j] + (1.f - h.beta2) * gg * gg;
        w[j] -= (h.lr / bc1) * (m[j] / (sqrtf(v[j]) * bc2r + h.eps));
      }
      *reinterpret_cast<float4*>(master + so) = *reinterpret_cast<float4*>(w);
      *reinterpret_cast<float4*>(master + so + 4) = *reinterpret_cast<float4*>(w + 4);
      *reinterpret_cast<float4*>(exp_avg + so) = *reinterpret_cast<float4*>(m);
      *reinterpret_cast<float4*>(exp_avg + so + 4) = *reinterpret_cast<float4*>(m + 4);
      *reinterpret_cast<float4*>(exp_avg_sq + so) = *reinterpret_cast<float4*>(v);
      *reinterpret_cast<float4*>(exp_avg_sq + so + 4) = *reinterpret_cast<float4*>(v + 4);
      const uint4 o = pack_bf16x8(w);
      const size_t poff = (size_t)(R.pelem_off[t] + i) * 2;
      if (bcast) bcast_vec(c, params, poff, o);
      else *reinterpret_cast<uint4*>(local_param + poff) = o;
    }
  }
  __threadfence_system();
  block_barrier(c, channel);                       // new parameters visible everywhere; every rank's reduction buffer is clear
}

void zero_fused_adam_rs(const CommCtx& c, const SymmBuf& grads, const SymmBuf& params, const SymmBuf& rs, const OwnedRangesRS& r,
                        float* master, float* exp_avg, float* exp_avg_sq, const AdamHyper& h, bool bcast_params, int channel,
                        cudaStream_t s) {
  const int work = r.r.blk_start[r.r.count];
  zero_fused_adam_rs_kernel<<<step_blocks(), 256, 0, s>>>(c, grads, params, rs, r, master, exp_avg, exp_avg_sq, h,
                                                           bcast_params ? 1 : 0, channel, work);
}

}  // namespace tds
