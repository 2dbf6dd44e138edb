// Python bindings of the symmetric-memory collectives (comm_sm100.cu).
#include <torch/extension.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>

#include <stdlib.h>
#include <vector>

#include "comm.h"

using torch::Tensor;
using namespace tds;

namespace {

cudaStream_t cur_stream() { return at::cuda::getCurrentCUDAStream().stream(); }

void check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  TORCH_CHECK(e == cudaSuccess, "tiny_deepspeed_b200 comm launch failed in ", what, ": ", cudaGetErrorString(e));
}

// Python-side handle: pointers exchanged once through torch's symmetric-memory rendezvous (handles only).
struct PyComm {
  CommCtx ctx{};
  Tensor error;   // int32[1] on device
  PyComm(std::vector<int64_t> flag_ptrs, int64_t rank, int64_t world, Tensor err) : error(err) {
    TORCH_CHECK(world <= kMaxRanks && (int64_t)flag_ptrs.size() == world);
    for (int r = 0; r < world; ++r) ctx.flags[r] = reinterpret_cast<uint32_t*>(flag_ptrs[r]);
    ctx.rank = (int)rank; ctx.world = (int)world;
    ctx.error_flag = err.data_ptr<int>();
    const char* t = getenv("TDS_COMM_TIMEOUT_S");
    const double secs = t ? atof(t) : 600.0;
    ctx.spin_limit = (long long)((secs > 0.01 ? secs : 0.01) * 2.0e9);   // ~2 GHz SM clock
  }
};
struct PyBuf {
  SymmBuf buf{};
  PyBuf(std::vector<int64_t> peer_ptrs, int64_t mc_ptr) {
    TORCH_CHECK(peer_ptrs.size() <= (size_t)kMaxRanks);
    for (size_t r = 0; r < peer_ptrs.size(); ++r) buf.peer[r] = reinterpret_cast<void*>(peer_ptrs[r]);
    buf.mc = reinterpret_cast<void*>(mc_ptr);
  }
  bool has_multicast() const { return buf.mc != nullptr; }
};

void py_allreduce(const PyComm& c, const PyBuf& b, int64_t elem_off, int64_t numel, bool is_f32, double scale,
                  int64_t blocks, int64_t channel) {
  allreduce(c.ctx, b.buf, elem_off, numel, is_f32, (float)scale, (int)blocks, (int)channel, cur_stream());
  check_launch("allreduce");
}
void py_reduce_to(const PyComm& c, const PyBuf& b, int64_t elem_off, int64_t numel, bool is_f32, int64_t dst, double scale,
                  int64_t blocks, int64_t channel) {
  reduce_to(c.ctx, b.buf, elem_off, numel, is_f32, (int)dst, (float)scale, (int)blocks, (int)channel, cur_stream());
  check_launch("reduce_to");
}
void py_broadcast(const PyComm& c, const PyBuf& b, int64_t byte_off, int64_t nbytes, int64_t src, int64_t blocks,
                  int64_t channel) {
  broadcast_from(c.ctx, b.buf, byte_off, nbytes, (int)src, (int)blocks, (int)channel, cur_stream());
  check_launch("broadcast");
}
void py_push(const PyComm& c, int64_t src_ptr, const PyBuf& dst, int64_t dst_byte_off, int64_t nbytes, int64_t src_rank,
             int64_t blocks, int64_t channel) {
  push_from(c.ctx, reinterpret_cast<const void*>(src_ptr), dst.buf, dst_byte_off, nbytes, (int)src_rank, (int)blocks,
            (int)channel, cur_stream());
  check_launch("push");
}
void py_allgather_slots(const PyComm& c, const PyBuf& b, int64_t base_byte_off, int64_t slot_bytes, int64_t blocks, int64_t channel) {
  TORCH_CHECK(slot_bytes % 16 == 0 && base_byte_off % 16 == 0, "allgather_slots: 16-byte aligned slots");
  allgather_slots(c.ctx, b.buf, base_byte_off, slot_bytes, (int)blocks, (int)channel, cur_stream());
  check_launch("allgather_slots");
}
void py_allreduce_rows(const PyComm& c, const PyBuf& b, int64_t table_byte_off, int64_t row_bytes, const Tensor& ids, int64_t vocab,
                       Tensor epoch_of_row, const Tensor& epoch, int64_t blocks, int64_t channel) {
  TORCH_CHECK(ids.is_cuda() && ids.scalar_type() == at::kLong && ids.is_contiguous(), "ids: contiguous int64 CUDA tensor");
  TORCH_CHECK(epoch_of_row.scalar_type() == at::kInt && epoch_of_row.numel() >= vocab && epoch.scalar_type() == at::kInt);
  TORCH_CHECK(row_bytes % 16 == 0 && table_byte_off % 16 == 0, "allreduce_rows: 16-byte aligned rows");
  allreduce_rows(c.ctx, b.buf, table_byte_off, row_bytes, ids.data_ptr<int64_t>(), (int)ids.numel(), vocab,
                 epoch_of_row.data_ptr<int>(), epoch.data_ptr<int>(), (int)blocks, (int)channel, cur_stream());
  check_launch("allreduce_rows");
}
void py_barrier(const PyComm& c, int64_t channel) {
  barrier(c.ctx, (int)channel, cur_stream());
  check_launch("barrier");
}

// ranges: list of (grad_elem_off, numel, state_off, param_elem_off); all multiples of 8
int64_t py_zero_fused_adam(const PyComm& c, const PyBuf& grads, const PyBuf& params,
                           const std::vector<std::vector<int64_t>>& ranges, Tensor master, Tensor exp_avg,
                           Tensor exp_avg_sq, double lr, double b1, double b2, double eps, double wd, const Tensor& step,
                           bool decoupled, bool maximize, double grad_scale, bool bcast, int64_t channel,
                           int64_t min_launches) {
  AdamHyper h{(float)lr, (float)b1, (float)b2, (float)eps, (float)wd, (float)grad_scale, decoupled ? 1 : 0,
              maximize ? 1 : 0, step.data_ptr<int>()};
  float* mp = master.numel() ? master.data_ptr<float>() : nullptr;
  float* m1 = exp_avg.numel() ? exp_avg.data_ptr<float>() : nullptr;
  float* m2 = exp_avg_sq.numel() ? exp_avg_sq.data_ptr<float>() : nullptr;
  int64_t launches = 0;
  size_t i = 0;
  do {  // at least one launch even with no owned ranges: every rank takes part in the barriers
    OwnedRanges R{};
    int cnt = 0, blk = 0;
    while (i < ranges.size() && cnt < kMaxRanges) {
      const auto& r = ranges[i];
      TORCH_CHECK(r.size() == 4 && r[0] % 8 == 0 && r[1] % 8 == 0 && r[2] % 8 == 0 && r[3] % 8 == 0, "ranges must be 8-element aligned");
      R.elem_off[cnt] = r[0]; R.numel[cnt] = r[1]; R.state_off[cnt] = r[2]; R.pelem_off[cnt] = r[3];
      R.blk_start[cnt] = blk;
      blk += (int)((r[1] + kZeroChunk - 1) / kZeroChunk);
      ++cnt; ++i;
    }
    R.blk_start[cnt] = blk;
    R.count = cnt > 0 ? cnt : 1;
    if (cnt == 0) { R.blk_start[0] = 0; R.blk_start[1] = 0; }
    zero_fused_adam(c.ctx, grads.buf, params.buf, R, mp, m1, m2, h, bcast, (int)channel, cur_stream());
    ++launches;
    // every rank must launch the SAME number of (barrier-carrying) kernels: pad up to the count of the rank owning most
  } while (i < ranges.size() || launches < min_launches);
  check_launch("zero_fused_adam");
  return launches;
}

// EXPERIMENTAL: ranges are [grad_off, numel, state_off, param_off, from_rs]
int64_t py_zero_fused_adam_rs(const PyComm& c, const PyBuf& grads, const PyBuf& params, const PyBuf& rs,
                              const std::vector<std::vector<int64_t>>& ranges, Tensor master, Tensor exp_avg,
                              Tensor exp_avg_sq, double lr, double b1, double b2, double eps, double wd, const Tensor& step,
                              bool decoupled, bool maximize, double grad_scale, bool bcast, int64_t channel,
                              int64_t min_launches) {
  AdamHyper h{(float)lr, (float)b1, (float)b2, (float)eps, (float)wd, (float)grad_scale, decoupled ? 1 : 0,
              maximize ? 1 : 0, step.data_ptr<int>()};
  float* mp = master.numel() ? master.data_ptr<float>() : nullptr;
  float* m1 = exp_avg.numel() ? exp_avg.data_ptr<float>() : nullptr;
  float* m2 = exp_avg_sq.numel() ? exp_avg_sq.data_ptr<float>() : nullptr;
  int64_t launches = 0;
  size_t i = 0;
  do {
    OwnedRangesRS RR{};
    OwnedRanges& R = RR.r;
    int cnt = 0, blk = 0;
    while (i < ranges.size() && cnt < kMaxRanges) {
      const auto& r = ranges[i];
      TORCH_CHECK(r.size() == 5 && r[0] % 8 == 0 && r[1] % 8 == 0 && r[2] % 8 == 0 && r[3] % 8 == 0, "ranges must be 8-element aligned");
      R.elem_off[cnt] = r[0]; R.numel[cnt] = r[1]; R.state_off[cnt] = r[2]; R.pelem_off[cnt] = r[3]; RR.rs[cnt] = r[4] ? 1 : 0;
      R.blk_start[cnt] = blk;
      blk += (int)((r[1] + kZeroChunk - 1) / kZeroChunk);
      ++cnt; ++i;
    }
    R.blk_start[cnt] = blk;
    R.count = cnt > 0 ? cnt : 1;
    if (cnt == 0) { R.blk_start[0] = 0; R.blk_start[1] = 0; }
    zero_fused_adam_rs(c.ctx, grads.buf, params.buf, rs.buf, RR, mp, m1, m2, h, bcast, (int)channel, cur_stream());
    ++launches;
  } while (i < ranges.size() || launches < min_launches);
  check_launch("zero_fused_adam_rs");
  return launches;
}

}  // namespace

void bind_comm(pybind11::module_& m) {
  pybind11::class_<PyComm>(m, "CommCtx")
      .def(pybind11::init<std::vector<int64_t>, int64_t, int64_t, Tensor>())
      .def_property_readonly("rank", [](const PyComm& c) { return c.ctx.rank; })
      .def_property_readonly("world", [](const PyComm& c) { return c.ctx.world; });
  pybind11::class_<PyBuf>(m, "SymmBuf")
      .def(pybind11::init<std::vector<int64_t>, int64_t>())
      .def_property_readonly("has_multicast", &PyBuf::has_multicast);
  m.def("comm_allreduce", &py_allreduce);
  m.def("comm_reduce_to", &py_reduce_to);
  m.def("comm_broadcast", &py_broadcast);
  m.def("comm_barrier", &py_barrier);
  m.def("comm_push", &py_push);
  m.def("comm_allgather_slots", &py_allgather_slots);
  m.def("comm_allreduce_rows", &py_allreduce_rows);
  m.def("comm_zero_fused_adam", &py_zero_fused_adam);
  m.def("comm_zero_fused_adam_rs", &py_zero_fused_adam_rs);
  m.attr("COMM_MAX_BLOCKS") = kCommMaxBlocks;
  m.attr("COMM_MAX_RANKS") = kMaxRanks;
  m.attr("COMM_MAX_RANGES") = kMaxRanges;
}
