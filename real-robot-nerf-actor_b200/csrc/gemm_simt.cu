// fp32 SIMT GEMMs: the parity-grade precision mode of the field MLP (NRF_PREC_FP32).
// Plain FFMA accumulation in fp32, shapes unrestricted.  Not the fast path -- the tcgen05 kernels in
// gemm_tc.cu are -- but the mode in which every stage is checked against the fp32 oracle at 1e-5.
#include "gemm_common.cuh"

namespace nrf {

constexpr int BM = 64, BN = 64, BK = 16;

struct SimtA {
  const float* A[3];
  int K[3];
  int lda[3];
};

// C = epilogue([A0|A1|A2] . B^T).  256 threads, each a 4x4 micro-tile.
__global__ void __launch_bounds__(256) gemm_simt_kernel(SimtA sa, const float* __restrict__ B, int ldb, int N,
                                                        Epilogue<float> ep) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int M = ep.M, K01 = sa.K[0] + sa.K[1], K = K01 + sa.K[2];
  int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  int tid = threadIdx.x;
  int tx = tid % 16, ty = tid / 16;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += BK) {
    // 64x16 tile of A and of B: 1024 elements each, 4 per thread
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int e = tid + i * 256;
      int row = e / BK, kk = e % BK;
      int k = k0 + kk;
      float av = 0.f, bv = 0.f;
      int m = m0 + row;
      if (m < M && k < K) {
        if (k < sa.K[0]) av = sa.A[0][(int64_t)m * sa.lda[0] + k];
        else if (k < K01) av = sa.A[1][(int64_t)m * sa.lda[1] + (k - sa.K[0])];
        else av = sa.A[2][(int64_t)m * sa.lda[2] + (k - K01)];
      }
      int n = n0 + row;
      if (n < N && k < K) bv = B[(int64_t)n * ldb + k];
      As[kk][row] = av;
      Bs[kk][row] = bv;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n >= ep.n_store || n >= N) continue;
      float v = acc[i][j];
      if (ep.bias) v += ep.bias[n];
      if (ep.mask_src && !(ep.mask_src[(int64_t)m * ep.ldmask + n] > 0.0f)) v = 0.0f;
      if (ep.resid) v += ep.resid[(int64_t)m * ep.ldr + n];
      if (ep.out_act) ep.out_act[(int64_t)m * ep.ldact + n] = ep.relu_act ? fmaxf(v, 0.0f) : v;
      if (ep.out_act2) ep.out_act2[(int64_t)m * ep.ldact2 + n] = ep.relu_act2 ? fmaxf(v, 0.0f) : v;
      if (ep.out_f32) ep.out_f32[(int64_t)m * ep.ldo + n] = v;
    }
  }
}

// dW[n,k] += sum_m G[m,n] A[m,k]; grid (k tiles, n tiles, M splits); fp32 atomics across splits.
__global__ void __launch_bounds__(256) wgrad_simt_kernel(const float* __restrict__ G, int ldg,
                                                         const float* __restrict__ A, int lda, int M,
                                                         int n_valid, int k_valid, float* __restrict__ dW,
                                                         int ldw, float* __restrict__ dbias, int m_per) {
  __shared__ float Gs[BK][BN + 4];
  __shared__ float As[BK][BM + 4];
  int n0 = blockIdx.y * BN, k0 = blockIdx.x * BM;
  int m_begin = blockIdx.z * m_per, m_end = min(M, m_begin + m_per);
  int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
  float acc[4][4] = {};
  float bsum = 0.f;   // threads 0..63 (blockIdx.x == 0) accumulate the bias gradient
  for (int mb = m_begin; mb < m_end; mb += BK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int e = tid + i * 256;
      int mm = e / 64, col = e % 64;
      int m = mb + mm;
      float gv = 0.f, av = 0.f;
      if (m < m_end) {
        if (n0 + col < n_valid) gv = G[(int64_t)m * ldg + n0 + col];
        if (k0 + col < k_valid) av = A[(int64_t)m * lda + k0 + col];
      }
      Gs[mm][col] = gv;
      As[mm][col] = av;
    }
    __syncthreads();
#pragma unroll
    for (int mm = 0; mm < BK; ++mm) {
      float g[4], a[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) g[i] = Gs[mm][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) a[j] = As[mm][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(g[i], a[j], acc[i][j]);
    }
    if (dbias && blockIdx.x == 0 && tid < 64) {
#pragma unroll
      for (int mm = 0; mm < BK; ++mm) bsum += Gs[mm][tid];
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int n = n0 + ty * 4 + i;
    if (n >= n_valid) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int k = k0 + tx * 4 + j;
      if (k < k_valid) atomicAdd(dW + (int64_t)n * ldw + k, acc[i][j]);
    }
  }
  if (dbias && blockIdx.x == 0 && tid < 64 && n0 + tid < n_valid) atomicAdd(dbias + n0 + tid, bsum);
}

int gemm_simt_launch(const NrfGemm& g, cudaStream_t stream) {
  Epilogue<float> ep = make_epilogue<float>(g);
  // grid.y is limited to 65535: fold large M into chunks
  int64_t rows_per_launch = (int64_t)65535 * BM;
  for (int64_t m_off = 0; m_off < g.M; m_off += rows_per_launch) {
    int64_t rows = g.M - m_off < rows_per_launch ? g.M - m_off : rows_per_launch;
    Epilogue<float> e = ep;
    e.M = (int)rows;
    if (e.mask_src) e.mask_src += m_off * e.ldmask;
    if (e.resid) e.resid += m_off * e.ldr;
    if (e.out_f32) e.out_f32 += m_off * e.ldo;
    if (e.out_act) e.out_act += m_off * e.ldact;
    if (e.out_act2) e.out_act2 += m_off * e.ldact2;
    SimtA sa;
    for (int i = 0; i < 3; ++i) {
      sa.K[i] = g.K[i];
      sa.lda[i] = g.lda[i];
      sa.A[i] = g.K[i] > 0 ? reinterpret_cast<const float*>(g.A[i]) + m_off * g.lda[i] : nullptr;
    }
    dim3 gr((g.N + BN - 1) / BN, (unsigned)((rows + BM - 1) / BM));
    { LaunchScope ls_(NRF_CAT_SIMT, stream);
    gemm_simt_kernel<<<gr, 256, 0, stream>>>(sa, reinterpret_cast<const float*>(g.B), g.ldb, g.N, e);
    }
    NRF_LAUNCH_OK();
  }
  return NRF_OK;
}

int wgrad_simt_launch(const void* G, int ldg, const void* A, int lda, int M, int N, int K, int n_valid,
                      int k_valid, float* dW, int ldw, float* dbias, cudaStream_t stream) {
  (void)N; (void)K;
  int n_tiles = (n_valid + BN - 1) / BN, k_tiles = (k_valid + BM - 1) / BM;
  int want_splits = (sm_count() * 4 + n_tiles * k_tiles - 1) / (n_tiles * k_tiles);
  int max_splits = (M + 255) / 256;
  int splits = want_splits < max_splits ? want_splits : max_splits;
  if (splits < 1) splits = 1;
  int m_per = ((M + splits - 1) / splits + BK - 1) / BK * BK;
  splits = (M + m_per - 1) / m_per;
  dim3 grid(k_tiles, n_tiles, splits);
  { LaunchScope ls_(NRF_CAT_SIMT, stream);
  wgrad_simt_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const float*>(G), ldg,
                                               reinterpret_cast<const float*>(A), lda, M, n_valid, k_valid,
                                               dW, ldw, dbias, m_per);
  }
  NRF_LAUNCH_OK();
  return NRF_OK;
}

}  // namespace nrf
