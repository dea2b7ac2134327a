// C-ABI glue: error text, device query, and the GEMM entry points of include/nrf_b200.h.
#include <stdarg.h>
#include <atomic>
#include <mutex>
#include <vector>
#include "gemm_common.cuh"

namespace nrf {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

struct TimedLaunch { int cat; cudaEvent_t a, b; };
static std::atomic<int64_t> g_launches{0};
static std::atomic<bool> g_timing{false};
static std::mutex g_timing_mu;
static std::vector<TimedLaunch> g_timed;

LaunchScope::LaunchScope(int category, cudaStream_t s) : idx(-1), stream(s) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  if (!g_timing.load(std::memory_order_relaxed)) return;
  TimedLaunch t;
  t.cat = category;
  if (cudaEventCreate(&t.a) != cudaSuccess || cudaEventCreate(&t.b) != cudaSuccess) return;
  cudaEventRecord(t.a, s);
  std::lock_guard<std::mutex> lk(g_timing_mu);
  g_timed.push_back(t);
  idx = (int)g_timed.size() - 1;
}

LaunchScope::~LaunchScope() {
  if (idx < 0) return;
  std::lock_guard<std::mutex> lk(g_timing_mu);
  if (idx < (int)g_timed.size()) cudaEventRecord(g_timed[idx].b, stream);
}

}  // namespace nrf

using namespace nrf;

extern "C" int64_t nrf_launch_count(void) { return g_launches.load(); }

extern "C" int nrf_timing_begin(void) {
  std::lock_guard<std::mutex> lk(g_timing_mu);
  for (auto& t : g_timed) { cudaEventDestroy(t.a); cudaEventDestroy(t.b); }
  g_timed.clear();
  g_timing.store(true);
  return NRF_OK;
}

extern "C" int nrf_timing_end(double* ms, int64_t* launches) {
  NRF_REQUIRE(ms && launches, NRF_EINVAL, "nrf_timing_end: null output");
  g_timing.store(false);
  NRF_CUDA_OK(cudaDeviceSynchronize());
  std::lock_guard<std::mutex> lk(g_timing_mu);
  for (int c = 0; c < NRF_TIMING_CATEGORIES; ++c) { ms[c] = 0.0; launches[c] = 0; }
  for (auto& t : g_timed) {
    float dt = 0.f;
    if (cudaEventElapsedTime(&dt, t.a, t.b) == cudaSuccess && t.cat >= 0 && t.cat < NRF_TIMING_CATEGORIES) {
      ms[t.cat] += dt;
      launches[t.cat] += 1;
    }
    cudaEventDestroy(t.a);
    cudaEventDestroy(t.b);
  }
  g_timed.clear();
  return NRF_OK;
}

extern "C" const char* nrf_version(void) { return "nrf_b200 0.2 (sm_100a; tcgen05/TMEM/TMA)"; }
extern "C" const char* nrf_last_error(void) { return g_err; }

extern "C" int nrf_gemm(const NrfGemm* g, int precision, void* stream) {
  NRF_REQUIRE(g && g->A[0] && g->B && g->M > 0 && g->N > 0 && g->K[0] > 0 && g->K[1] >= 0 && g->K[2] >= 0,
              NRF_EINVAL, "nrf_gemm: bad arguments");
  NRF_REQUIRE((g->K[1] == 0 || g->A[1]) && (g->K[2] == 0 || g->A[2]), NRF_EINVAL, "nrf_gemm: K[i] > 0 needs A[i]");
  NRF_REQUIRE(g->n_store > 0 && g->n_store <= g->N, NRF_EINVAL, "nrf_gemm: n_store out of range");
  NRF_REQUIRE(g->out_f32 || g->out_act, NRF_EINVAL, "nrf_gemm: no output");
  if (precision == NRF_PREC_BF16 || precision == NRF_PREC_FP16)
    return gemm_tc_launch(*g, op_fmt(precision), as_stream(stream));
  if (precision == NRF_PREC_FP32) return gemm_simt_launch(*g, as_stream(stream));
  set_error("nrf_gemm: unknown precision %d", precision);
  return NRF_EINVAL;
}

extern "C" int64_t nrf_wgrad_workspace_bytes(int N, int K) {
  (void)N; (void)K;
  // (output tiles) x (sample splits) <= #SMs; a split's partial tile is 128 x (256 + 1 bias column) fp32
  return (int64_t)sm_count() * 128 * (256 + 8) * 4 + 4096;
}

extern "C" int nrf_wgrad(const void* G, int ldg, const void* A, int lda, int M, int N, int K, int n_valid,
                         int k_valid, float* dW, int ldw, float* dbias, void* workspace, int precision,
                         void* stream) {
  NRF_REQUIRE(G && A && dW && M > 0 && N > 0 && K > 0, NRF_EINVAL, "nrf_wgrad: bad arguments");
  NRF_REQUIRE(n_valid > 0 && n_valid <= N && k_valid > 0 && k_valid <= K, NRF_EINVAL,
              "nrf_wgrad: n_valid/k_valid out of range");
  if (precision == NRF_PREC_BF16 || precision == NRF_PREC_FP16)
    return wgrad_tc_launch(G, ldg, A, lda, M, N, K, n_valid, k_valid, dW, ldw, dbias, workspace,
                           op_fmt(precision), as_stream(stream));
  if (precision == NRF_PREC_FP32)
    return wgrad_simt_launch(G, ldg, A, lda, M, N, K, n_valid, k_valid, dW, ldw, dbias, as_stream(stream));
  set_error("nrf_wgrad: unknown precision %d", precision);
  return NRF_EINVAL;
}
