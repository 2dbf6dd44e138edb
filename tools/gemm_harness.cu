// Standalone (no torch, starts in a second) correctness + timing harness for the tcgen05 GEMMs in csrc/gemm_sm100.cu and
// csrc/gemm2_sm100.cu.  Used to iterate on the kernels with short gpurun calls:
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -I tiny_deepspeed_b200/csrc \
//        -o tools/gemm_harness tools/gemm_harness.cu tiny_deepspeed_b200/csrc/gemm_sm100.cu tiny_deepspeed_b200/csrc/gemm2_sm100.cu -lcuda
//   tools/gemm_harness [check] [sweep]
//
// Every shape is verified against a naive fp32 CUDA GEMM (max relative error over the full output) and timed two ways:
// back-to-back launches between CUDA events (warm L2, the situation inside the captured training step, where the activation
// operand was just written) and one launch after an L2 flush (cold).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <string>
#include <vector>

#include "kernels.h"

namespace tds { void gemm_set_debug(int bits); void gemm_set_variant(int v); void gemm_set_prof(long long* buf); }

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__global__ void fill_kernel(__nv_bfloat16* p, size_t n, uint32_t seed) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    uint32_t x = (uint32_t)i * 2654435761u + seed;
    x ^= x >> 16; x *= 2246822519u; x ^= x >> 13; x *= 3266489917u; x ^= x >> 16;
    p[i] = __float2bfloat16(((int)(x & 0xFFFF) - 32768) / 32768.0f);
  }
}
// D[m][n] = sum_k A(m,k) B(n,k) (+ bias[n]); a_mn: A stored [K][M]; b_mn: B stored [K][N]
__global__ void ref_gemm(const __nv_bfloat16* A, const __nv_bfloat16* B, const __nv_bfloat16* bias, float* D, int M, int N, int K,
                         int a_mn, int b_mn) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x, m = blockIdx.y;
  if (n >= N) return;
  float acc = 0.f;
  for (int k = 0; k < K; ++k) {
    const float a = __bfloat162float(a_mn ? A[(size_t)k * M + m] : A[(size_t)m * K + k]);
    const float b = __bfloat162float(b_mn ? B[(size_t)k * N + n] : B[(size_t)n * K + k]);
    acc += a * b;
  }
  if (bias) acc += __bfloat162float(bias[n]);
  D[(size_t)m * N + n] = acc;
}
__global__ void cmp_kernel(const __nv_bfloat16* got, const float* ref, size_t n, float* out /*[2]: max abs err, max abs ref*/) {
  float e = 0.f, r = 0.f;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    e = fmaxf(e, fabsf(__bfloat162float(got[i]) - ref[i]));
    r = fmaxf(r, fabsf(ref[i]));
  }
  atomicMax(reinterpret_cast<int*>(out), __float_as_int(e));
  atomicMax(reinterpret_cast<int*>(out + 1), __float_as_int(r));
}

struct Shape { const char* name; int M, N, K, a_mn, b_mn; };
static const Shape kShapes[] = {
    {"c_attn fwd", 1024, 2304, 768, 0, 0},   {"attn.c_proj fwd", 1024, 768, 768, 0, 0}, {"c_fc fwd", 1024, 3072, 768, 0, 0},
    {"mlp.c_proj fwd", 1024, 768, 3072, 0, 0}, {"c_attn dX", 1024, 768, 2304, 0, 1},    {"attn.c_proj dX", 1024, 768, 768, 0, 1},
    {"c_fc dX", 1024, 768, 3072, 0, 1},      {"mlp.c_proj dX", 1024, 3072, 768, 0, 1},  {"c_attn dW", 2304, 768, 1024, 1, 1},
    {"attn.c_proj dW", 768, 768, 1024, 1, 1}, {"c_fc dW", 3072, 768, 1024, 1, 1},       {"mlp.c_proj dW", 768, 3072, 1024, 1, 1},
    {"lm_head fwd", 1024, 50304, 768, 0, 0}, {"lm_head dW", 50304, 768, 1024, 1, 1},    {"square 4096", 4096, 4096, 4096, 0, 0},
};

static void* g_flush = nullptr;
static void l2_flush() {
  if (!g_flush) CK(cudaMalloc(&g_flush, 256u << 20));
  CK(cudaMemsetAsync(g_flush, 0, 256u << 20));
}

struct Bufs { __nv_bfloat16 *a, *b, *bias, *d; float* ref; float* err; };

static tds::GemmParams make_params(const Shape& s, const Bufs& bf, int config, bool bias) {
  tds::GemmParams p{};
  p.a = {bf.a, s.a_mn ? s.M : s.K, 0, 0, (bool)s.a_mn};
  p.b = {bf.b, s.b_mn ? s.N : s.K, 0, 0, (bool)s.b_mn};
  p.d = bf.d; p.d_dtype = tds::kBF16; p.ldd = s.N; p.d_batch_stride = 0; p.d_batch_stride2 = 0;
  p.in_dtype = tds::kBF16; p.io_dtype = tds::kBF16;
  p.bias = bias ? bf.bias : nullptr; p.aux = nullptr; p.ld_aux = 0; p.epi = 0; p.accumulate = false; p.reduce_out = false;
  p.alpha = 1.f; p.M = s.M; p.N = s.N; p.K = s.K; p.batch = 1; p.nbatch2 = 1; p.config = config; p.cluster_m = 0; p.tri = 0;
  return p;
}

static void run(const tds::GemmParams& p, int pair) {
  if (pair && tds::gemm2_bf16(p, 0)) return;
  tds::gemm_bf16(p, 0);
}

static float time_warm(const tds::GemmParams& p, int pair, int iters = 30) {
  for (int i = 0; i < 5; ++i) run(p, pair);
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  CK(cudaEventRecord(e0));
  for (int i = 0; i < iters; ++i) run(p, pair);
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  CK(cudaEventDestroy(e0)); CK(cudaEventDestroy(e1));
  return ms * 1e3f / iters;
}
// the same 30 launches as ONE CUDA graph (how the training step actually runs them): per-launch time inside the graph
static float time_graph(const tds::GemmParams& p, int pair, int iters = 30) {
  cudaStream_t st;
  CK(cudaStreamCreate(&st));
  tds::GemmParams q = p;
  cudaGraph_t graph; cudaGraphExec_t exec;
  CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeGlobal));
  for (int i = 0; i < iters; ++i) { if (!(pair && tds::gemm2_bf16(q, st))) tds::gemm_bf16(q, st); }
  CK(cudaStreamEndCapture(st, &graph));
  CK(cudaGraphInstantiate(&exec, graph, 0));
  for (int i = 0; i < 3; ++i) CK(cudaGraphLaunch(exec, st));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  CK(cudaEventRecord(e0, st));
  for (int i = 0; i < 5; ++i) CK(cudaGraphLaunch(exec, st));
  CK(cudaEventRecord(e1, st));
  CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  CK(cudaEventDestroy(e0)); CK(cudaEventDestroy(e1));
  CK(cudaGraphExecDestroy(exec)); CK(cudaGraphDestroy(graph)); CK(cudaStreamDestroy(st));
  return ms * 1e3f / (5 * iters);
}
// The training step never finds a weight in L2 (250 MB of them per pass): launch i uses B buffer i % nbuf (pool > L2), A stays hot
// like the activation a previous kernel just wrote.  prefetch: launch i also asks for launch i+1's B through the kernel's L2 hint.
static float time_graph_rot(const tds::GemmParams& p, const std::vector<__nv_bfloat16*>& pool, size_t bbytes, bool prefetch, int iters = 60) {
  cudaStream_t st;
  CK(cudaStreamCreate(&st));
  cudaGraph_t graph; cudaGraphExec_t exec;
  CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeGlobal));
  for (int i = 0; i < iters; ++i) {
    tds::GemmParams q = p;
    q.b.ptr = pool[i % pool.size()];
    if (prefetch) { q.prefetch = pool[(i + 1) % pool.size()]; q.prefetch_bytes = (int64_t)bbytes; }
    tds::gemm_bf16(q, st);
  }
  CK(cudaStreamEndCapture(st, &graph));
  CK(cudaGraphInstantiate(&exec, graph, 0));
  for (int i = 0; i < 3; ++i) CK(cudaGraphLaunch(exec, st));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  CK(cudaEventRecord(e0, st));
  for (int i = 0; i < 5; ++i) CK(cudaGraphLaunch(exec, st));
  CK(cudaEventRecord(e1, st));
  CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  CK(cudaEventDestroy(e0)); CK(cudaEventDestroy(e1));
  CK(cudaGraphExecDestroy(exec)); CK(cudaGraphDestroy(graph)); CK(cudaStreamDestroy(st));
  return ms * 1e3f / (5 * iters);
}
static float time_cold(const tds::GemmParams& p, int pair, int iters = 9) {
  std::vector<float> v;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int i = 0; i < iters; ++i) {
    l2_flush();
    CK(cudaEventRecord(e0)); run(p, pair); CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    v.push_back(ms * 1e3f);
  }
  CK(cudaEventDestroy(e0)); CK(cudaEventDestroy(e1));
  std::sort(v.begin(), v.end());
  return v[v.size() / 2];
}

static double check(const Shape& s, const Bufs& bf, const tds::GemmParams& p, int pair) {
  CK(cudaMemset(bf.d, 0xFF, (size_t)s.M * s.N * 2));
  run(p, pair);
  CK(cudaMemset(bf.err, 0, 8));
  cmp_kernel<<<296, 256>>>(bf.d, bf.ref, (size_t)s.M * s.N, bf.err);
  float h[2];
  CK(cudaMemcpy(h, bf.err, 8, cudaMemcpyDeviceToHost));
  return h[1] > 0 ? h[0] / h[1] : h[0];
}

int main(int argc, char** argv) {
  bool do_check = false, sweep = false, dbg = false, trace = false, nobias = true, rot = false, epi_mode = false;
  std::string only;
  for (int i = 1; i < argc; ++i) {
    if (!strcmp(argv[i], "check")) do_check = true;
    else if (!strcmp(argv[i], "sweep")) sweep = true;
    else if (!strcmp(argv[i], "dbg")) dbg = true;
    else if (!strcmp(argv[i], "trace")) trace = true;
    else if (!strcmp(argv[i], "rot")) rot = true;
    else if (!strcmp(argv[i], "epi")) epi_mode = true;
    else if (!strcmp(argv[i], "bias")) nobias = false;
    else only = argv[i];
  }
  size_t maxel = 0;
  for (const Shape& s : kShapes) maxel = std::max(maxel, std::max((size_t)s.M * s.K, std::max((size_t)s.N * s.K, (size_t)s.M * s.N)));
  Bufs bf;
  CK(cudaMalloc(&bf.a, maxel * 2)); CK(cudaMalloc(&bf.b, maxel * 2)); CK(cudaMalloc(&bf.d, maxel * 2));
  CK(cudaMalloc(&bf.bias, 65536 * 2)); CK(cudaMalloc(&bf.ref, maxel * 4)); CK(cudaMalloc(&bf.err, 8));
  fill_kernel<<<592, 256>>>(bf.a, maxel, 1u);
  fill_kernel<<<592, 256>>>(bf.b, maxel, 77u);
  fill_kernel<<<64, 256>>>(bf.bias, 65536, 5u);
  CK(cudaDeviceSynchronize());
  const int nvar = getenv("HARNESS_VARIANTS") ? atoi(getenv("HARNESS_VARIANTS")) : 1;
  printf("%-16s %6s %6s %6s lay | %-44s | %-22s\n", "shape", "M", "N", "K", "variant x {auto,cfg0,cfg1,cfg2,cfg3} warm us (cold auto)", "pair warm/cold");
  for (const Shape& s : kShapes) {
    if (!only.empty() && !strstr(s.name, only.c_str())) continue;
    const bool big = (size_t)s.M * s.N > (4u << 20);
    if (do_check && !(big && !sweep)) {
      dim3 g((s.N + 127) / 128, s.M);
      ref_gemm<<<g, 128>>>(bf.a, bf.b, nobias ? nullptr : bf.bias, bf.ref, s.M, s.N, s.K, s.a_mn, s.b_mn);
      CK(cudaDeviceSynchronize());
    }
    if (epi_mode && trace) {
      // per-phase cycle medians of the epilogue variants (prof build only)
      if (big) continue;
      __nv_bfloat16* aux; CK(cudaMalloc(&aux, (size_t)s.M * s.N * 2)); CK(cudaMemset(aux, 0, (size_t)s.M * s.N * 2));
      static long long* prof = nullptr;
      if (!prof) CK(cudaMalloc(&prof, (148 * 16 + 256) * 8));
      printf("%-16s %6d %6d %6d %d%d\n", s.name, s.M, s.N, s.K, s.a_mn, s.b_mn);
      const char* vn[] = {"plain", "bias", "bias+residual", "bias+GELU-save", "GELU-bwd"};
      for (int variant = 0; variant < 5; ++variant) {
        tds::GemmParams p = make_params(s, bf, -1, variant >= 1 && variant <= 3);
        if (variant >= 2) { p.aux = aux; p.ld_aux = s.N; p.epi = variant == 2 ? 3 : (variant == 3 ? 1 : 2); }
        for (int i = 0; i < 10; ++i) run(p, 0);
        CK(cudaMemset(prof, 0, (148 * 16 + 256) * 8));
        tds::gemm_set_prof(prof);
        run(p, 0);
        tds::gemm_set_prof(nullptr);
        CK(cudaDeviceSynchronize());
        std::vector<long long> h(148 * 16 + 256);
        CK(cudaMemcpy(h.data(), prof, h.size() * 8, cudaMemcpyDeviceToHost));
        long long t0 = h[0], t1 = h[14];
        std::vector<long long> main_, epi, wr, tail, tot;
        for (int b = 0; b < 148; ++b) if (h[b * 16 + 1]) {
          t0 = std::min(t0, h[b * 16]); t1 = std::max(t1, h[b * 16 + 14]);
          main_.push_back(h[b * 16 + 9] - h[b * 16 + 2]); epi.push_back(h[b * 16 + 10] - h[b * 16 + 9]);
          wr.push_back(h[b * 16 + 11] - h[b * 16 + 10]); tail.push_back(h[b * 16 + 13] - h[b * 16 + 11]); tot.push_back(h[b * 16 + 14] - h[b * 16 + 0]);
        }
        auto med = [](std::vector<long long>& v) { std::sort(v.begin(), v.end()); return v.empty() ? 0LL : v[v.size() / 2]; };
        printf("  %-16s wall %6lld ns | CTAs %3zu: until accumulator %6lld cyc, epilogue body %6lld, store-read wait %5lld, exit sync %5lld, CTA lifetime %6lld ns | slab0:",
               vn[variant], t1 - t0, tot.size(), med(main_), med(epi), med(wr), med(tail), med(tot));
        const long long* tr = &h[148 * 16 + 128];
        if (tr[0]) for (int i = 1; i <= 7; ++i) printf(" %5lld", tr[i] - tr[0]);
        printf("\n");
      }
      CK(cudaFree(aux));
      continue;
    }
    if (epi_mode) {
      if (big) continue;
      __nv_bfloat16* aux; CK(cudaMalloc(&aux, (size_t)s.M * s.N * 2)); CK(cudaMemset(aux, 0, (size_t)s.M * s.N * 2));
      printf("%-16s %6d %6d %6d %d%d  | in-graph us (B hot):", s.name, s.M, s.N, s.K, s.a_mn, s.b_mn);
      for (int cfg : {-1, 0, 1, 3}) {
        if (cfg == 3 && s.N % 192) continue;
        printf("  cfg %2d:", cfg);
        for (int variant = 0; variant < 5; ++variant) {   // plain, +bias, +bias+residual, +bias+GELU save, GELU' (no bias)
          tds::GemmParams p = make_params(s, bf, cfg, variant >= 1 && variant <= 3);
          if (variant >= 2) { p.aux = aux; p.ld_aux = s.N; p.epi = variant == 2 ? 3 : (variant == 3 ? 1 : 2); }
          printf(" %5.2f", time_graph(p, 0, 60));
        }
      }
      printf("   (plain, bias, bias+residual, bias+GELU-save, GELU-bwd)\n");
      CK(cudaFree(aux));
      continue;
    }
    if (rot) {
      if (big) continue;
      const size_t bbytes = (size_t)s.N * s.K * 2;
      size_t nbuf = ((size_t)400 << 20) / bbytes + 1; if (nbuf > 120) nbuf = 120;
      std::vector<__nv_bfloat16*> pool(nbuf);
      for (auto& q : pool) { CK(cudaMalloc(&q, bbytes)); CK(cudaMemcpy(q, bf.b, bbytes, cudaMemcpyDeviceToDevice)); }
      tds::GemmParams p = make_params(s, bf, -1, !nobias);
      const float hot = time_graph(p, 0, 60), cold = time_graph_rot(p, pool, bbytes, false), pf = time_graph_rot(p, pool, bbytes, true);
      printf("%-16s %6d %6d %6d %d%d  | in-graph us: B hot %5.2f   B from HBM (pool of %zu) %5.2f   + L2 prefetch by the previous launch %5.2f\n",
             s.name, s.M, s.N, s.K, s.a_mn, s.b_mn, hot, nbuf, cold, pf);
      for (auto q : pool) CK(cudaFree(q));
      continue;
    }
    printf("%-16s %6d %6d %6d %d%d  |", s.name, s.M, s.N, s.K, s.a_mn, s.b_mn);
    for (int v = 0; v < nvar; ++v) {
      tds::gemm_set_variant(v);
      printf(" v%d:", v);
      for (int cfg = -1; cfg < 4; ++cfg) {
        if (cfg == 3 && s.N % 192) { printf("    -"); continue; }
        tds::GemmParams p = make_params(s, bf, cfg, !nobias);
        if (do_check && !(big && !sweep)) {
          const double e = check(s, bf, p, 0);
          if (e > 8e-3) printf(" [ERR %.1e]", e);
        }
        printf(" %5.1f", time_warm(p, 0));
      }
      tds::GemmParams p = make_params(s, bf, -1, !nobias);
      printf(" (%5.1f) graph %5.2f", time_cold(p, 0), time_graph(p, 0));
    }
    {
      tds::gemm_set_variant(0);
      tds::GemmParams p = make_params(s, bf, -1, !nobias);
      double e = 0;
      if (do_check && !(big && !sweep)) e = check(s, bf, p, 1);
      printf(" | %5.1f / %5.1f%s", time_warm(p, 1), time_cold(p, 1), e > 8e-3 ? " [PAIR ERR]" : "");
    }
    if (dbg) {
      printf(" | dbg2/4/6:");
      for (int d : {2, 4, 6}) {
        tds::gemm_set_debug(d);
        tds::GemmParams p = make_params(s, bf, -1, !nobias);
        printf(" %5.1f", time_warm(p, 0));
      }
      tds::gemm_set_debug(0);
    }
    if (trace) {
      static long long* prof = nullptr;
      if (!prof) CK(cudaMalloc(&prof, (148 * 16 + 256) * 8));
      tds::GemmParams p = make_params(s, bf, -1, !nobias);
      for (int i = 0; i < 10; ++i) run(p, 0);
      CK(cudaMemset(prof, 0, (148 * 16 + 256) * 8));
      tds::gemm_set_prof(prof);
      run(p, 0);
      tds::gemm_set_prof(nullptr);
      CK(cudaDeviceSynchronize());
      std::vector<long long> h(148 * 16 + 256);
      CK(cudaMemcpy(h.data(), prof, h.size() * 8, cudaMemcpyDeviceToHost));
      long long t0 = h[0], t1 = h[14];
      for (int b = 0; b < 148; ++b) if (h[b * 16 + 1]) { t0 = std::min(t0, h[b * 16]); t1 = std::max(t1, h[b * 16 + 14]); }
      printf("\n    kernel wall %lld ns; issuer trace (cycles rel. to loop top: wait_done, fence+desc, mma0, mma3, commit, syncwarp | next top):", t1 - t0);
      for (int kb = 0; kb < 4; ++kb) {
        const long long* tr = &h[148 * 16 + kb * 8];
        printf("\n      kb%d:", kb + 4);
        for (int i = 1; i <= 6; ++i) printf(" %5lld", tr[i] - tr[0]);
        if (kb < 3) printf(" | %5lld", h[148 * 16 + (kb + 1) * 8] - tr[0]);
      }
      printf("\n    producer trace (wait_done, expect_tx, tma_issued | next top):");
      for (int kb = 0; kb < 4; ++kb) {
        const long long* tr = &h[148 * 16 + 64 + kb * 8];
        printf("\n      kb%d:", kb + 6);
        for (int i = 1; i <= 3; ++i) printf(" %5lld", tr[i] - tr[0]);
        if (kb < 3) printf(" | %5lld", h[148 * 16 + 64 + (kb + 1) * 8] - tr[0]);
      }
      printf("\n    epilogue trace per 64-col slab (cycles: wait_read+sync, ld0, st0, ld1, st1, fence+sync, tma_store):");
      for (int sl = 0; sl < 2; ++sl) {
        const long long* tr = &h[148 * 16 + 128 + sl * 8];
        if (!tr[0]) continue;
        printf("\n      slab%d:", sl);
        for (int i = 1; i <= 7; ++i) printf(" %5lld", tr[i] - tr[0]);
      }
      // phase medians over CTAs in ns-equivalent cycles of each warp's own clock
      {
        std::vector<long long> epi, tot;
        for (int b = 0; b < 148; ++b) if (h[b * 16 + 1]) { epi.push_back(h[b * 16 + 10] - h[b * 16 + 9]); tot.push_back(h[b * 16 + 14] - h[b * 16 + 0]); }
        std::sort(epi.begin(), epi.end()); std::sort(tot.begin(), tot.end());
        printf("\n    median over CTAs: epilogue body %lld cycles, CTA lifetime %lld ns", epi[epi.size() / 2], tot[tot.size() / 2]);
      }
    }
    printf("\n");
    fflush(stdout);
  }
  CK(cudaDeviceSynchronize());
  printf("done\n");
  return 0;
}
