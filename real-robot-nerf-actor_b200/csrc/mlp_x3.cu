// NRF_PREC_BF16X3: the ResnetFC field MLP (resnetfc.py:55-64,146-195) on the tensor cores at fp32-grade accuracy.
//
// Every GEMM operand is split into two bf16 terms, a = hi + lo with hi = bf16(a), lo = bf16(a - hi) (16 significant
// bits together), and every product is evaluated as hi.hi + lo.hi + hi.lo with fp32 accumulation in TMEM - three
// tcgen05 MMAs per k-step instead of one; the dropped lo.lo term is 2^-18 relative.  SURVEY.md section 10 measured
// 1.5e-5 / 7e-6 / 1e-6 relative error on rgb / embed / depth for this scheme against the fp32 reference (bf16: 8e-3,
// fp16: 1.4e-3): it is the tensor-core mode that meets north_star's <= 1e-3 on every output, at a third of the
// bf16 rate.
//
// The split is expressed as CONCATENATION ALONG K, which the existing GEMM kernel (gemm_tc_kernel, csrc/gemm_tc.cu)
// already supports through its up-to-three A sources:
//     activations are stored as rows [hi(K) | lo(K)]                               (N, 2K) bf16
//     weights are packed as rows     [hi(K) | hi(K) | lo(K)]                       (n, 3K) bf16
//     source 0 = the whole row (K0 = 2K: hi.hi + lo.hi), source 1 = the hi half again (K1 = K: hi.lo)
// so a layer is ONE GEMM with K tripled and an fp32 output, followed by a small element-wise kernel that applies the
// ReLU gate / residual stream (kept in fp32) and writes the next operand as a [hi | lo] row.  Weight gradients are three
// launches of wgrad_tc_kernel (G_hi^T A_hi + G_hi^T A_lo + G_lo^T A_hi) accumulating into the same fp32 dW.
// To the caller the mode looks like NRF_PREC_FP32: fp32 field input, fp32 d_field, fp32 outputs.
#include <stdlib.h>
#include <string.h>
#include "gemm_common.cuh"

namespace nrf {

namespace {

inline int64_t rup(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

struct X3Layout {
  int H, C, Din, Dout, nb, nz, kin_pad, dout_pad, nout_pad, cpad;
  int64_t W0, bias0, Wout, bias_out, WoutT;
  int64_t Wfc0[NRF_MAX_BLOCKS], Wfc1[NRF_MAX_BLOCKS], Wz[NRF_MAX_BLOCKS], bias1[NRF_MAX_BLOCKS];
  int64_t Wfc0T[NRF_MAX_BLOCKS], Wfc1T[NRF_MAX_BLOCKS], WzT[NRF_MAX_BLOCKS];
  int64_t total;
};

int x3_layout(const NrfMlpParams* p, X3Layout* L) {
  NRF_REQUIRE(p, NRF_EINVAL, "mlp(bf16x3): null params");
  NRF_REQUIRE(p->n_blocks >= 1 && p->n_blocks <= NRF_MAX_BLOCKS && p->n_lin_z >= 0 && p->n_lin_z <= p->n_blocks,
              NRF_EINVAL, "mlp(bf16x3): n_blocks=%d n_lin_z=%d", p->n_blocks, p->n_lin_z);
  NRF_REQUIRE(p->d_in > 0 && p->d_hidden > 0 && p->d_out > 0 && p->d_latent >= 0, NRF_EINVAL, "mlp(bf16x3): bad dims");
  L->H = p->d_hidden; L->C = p->d_latent; L->Din = p->d_in; L->Dout = p->d_out;
  L->nb = p->n_blocks; L->nz = p->d_latent > 0 ? p->n_lin_z : 0;
  NRF_REQUIRE(L->H % 128 == 0, NRF_ENOSUP, "mlp(bf16x3): d_hidden=%d must be a multiple of 128", L->H);
  NRF_REQUIRE(L->C % 64 == 0 && L->C > 0, NRF_ENOSUP, "mlp(bf16x3): d_latent=%d must be a multiple of 64", L->C);
  L->kin_pad = (int)rup(L->C + L->Din, 64);
  L->dout_pad = (int)rup(L->Dout, 64);
  L->nout_pad = (int)rup(L->Dout, 128);
  L->cpad = (int)rup(L->C, 128);
  int64_t off = 0;
  auto take = [&](int64_t bytes) { int64_t o = off; off = rup(off + bytes, 1024); return o; };
  L->W0 = take((int64_t)L->H * 3 * L->kin_pad * 2);
  L->bias0 = take((int64_t)L->H * 4);
  for (int b = 0; b < L->nb; ++b) {
    L->Wfc0[b] = take((int64_t)L->H * 3 * L->H * 2);
    L->Wfc1[b] = take((int64_t)L->H * 3 * L->H * 2);
    L->Wz[b] = take((int64_t)L->H * 3 * L->C * 2);              // lin_z[b] as a forward operand (b >= 1 is used)
    L->bias1[b] = take((int64_t)L->H * 4);
    L->Wfc0T[b] = take((int64_t)L->H * 3 * L->H * 2);
    L->Wfc1T[b] = take((int64_t)L->H * 3 * L->H * 2);
    L->WzT[b] = take((int64_t)L->cpad * 3 * L->H * 2);          // lin_z[b]^T (cpad rows, K = H) for dL/dz
  }
  L->Wout = take((int64_t)L->nout_pad * 3 * L->H * 2);
  L->bias_out = take((int64_t)L->nout_pad * 4);
  L->WoutT = take((int64_t)L->H * 3 * L->dout_pad * 2);
  L->total = off;
  return NRF_OK;
}

// ---------------------------------------------------------------------------------------- packing
struct X3Seg {
  const float* src; const float* src2;
  int64_t dst_off;
  int kpad, rows, cols, ld_src, mode, col0;   // mode 0 / 1: matrix (plain / transposed source), 2: fp32 bias sum
};
constexpr int kX3MaxSegs = 72;
struct X3Table { int n; X3Seg seg[kX3MaxSegs]; };

__device__ __forceinline__ void split_bf16(float v, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(v);
  lo = __float2bfloat16_rn(v - __bfloat162float(hi));      // exact difference (Sterbenz-like: same binade or below)
}

// weight rows [hi(kpad) | hi(kpad) | lo(kpad)], element (r, col0 + c) of the logical (rows, kpad) matrix
__global__ void __launch_bounds__(256) x3_pack_kernel(const __grid_constant__ X3Table tab, char* __restrict__ base) {
  const X3Seg& sg = tab.seg[blockIdx.y];
  const int64_t n = (int64_t)sg.rows * sg.cols;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(t / sg.cols), c = (int)(t % sg.cols);
    if (sg.mode == 2) {
      reinterpret_cast<float*>(base + sg.dst_off)[c] = (sg.src ? sg.src[c] : 0.0f) + (sg.src2 ? sg.src2[c] : 0.0f);
      continue;
    }
    const float v = sg.mode == 1 ? sg.src[(int64_t)c * sg.ld_src + r] : sg.src[(int64_t)r * sg.ld_src + c];
    __nv_bfloat16 hi, lo;
    split_bf16(v, hi, lo);
    __nv_bfloat16* row = reinterpret_cast<__nv_bfloat16*>(base + sg.dst_off) + (int64_t)r * 3 * sg.kpad + sg.col0 + c;
    row[0] = hi; row[sg.kpad] = hi; row[2 * sg.kpad] = lo;
  }
}

// ------------------------------------------------------------------------------- element-wise stages
// rows of `n4 * 4` fp32 columns -> [hi | lo] rows of 2 * ld_half bf16 (columns >= n4 * 4 of each half are left alone:
// the buffers are zero-initialised once where padding matters)
struct X3Epi {
  const float* acc; int ld_acc;       // GEMM result
  const float* acc2; int ld_acc2;     // second GEMM result added to it (lin_z tail), or NULL
  const __nv_bfloat16* mask; int ld_mask;   // hi half of the saved forward operand: v = 0 where it is <= 0
  float* stream; int has_resid, write_stream;   // fp32 residual stream (N, ld_acc): v += stream; stream = v
  __nv_bfloat16* out; int ld_half;    // [hi | lo] rows, pitch 2 * ld_half; NULL: nothing written
  int relu;
  int64_t rows; int cols;
};

__global__ void __launch_bounds__(256) x3_epi_kernel(const X3Epi e) {
  const int c4 = e.cols / 4;
  const int64_t n = e.rows * c4;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = t / c4;
    const int c = (int)(t % c4) * 4;
    float4 v = *reinterpret_cast<const float4*>(e.acc + r * e.ld_acc + c);
    if (e.acc2) {
      const float4 w = *reinterpret_cast<const float4*>(e.acc2 + r * e.ld_acc2 + c);
      v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
    }
    if (e.mask) {
      const uint2 m = *reinterpret_cast<const uint2*>(e.mask + r * e.ld_mask + c);
      if (!((int16_t)(m.x & 0xffffu) > 0)) v.x = 0.f;
      if (!((int16_t)(m.x >> 16) > 0)) v.y = 0.f;
      if (!((int16_t)(m.y & 0xffffu) > 0)) v.z = 0.f;
      if (!((int16_t)(m.y >> 16) > 0)) v.w = 0.f;
    }
    if (e.has_resid) {
      const float4 s = *reinterpret_cast<const float4*>(e.stream + r * e.ld_acc + c);
      v.x += s.x; v.y += s.y; v.z += s.z; v.w += s.w;
    }
    if (e.write_stream) *reinterpret_cast<float4*>(e.stream + r * e.ld_acc + c) = v;
    if (e.out) {
      if (e.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
      __nv_bfloat16 h[4], l[4];
      split_bf16(v.x, h[0], l[0]); split_bf16(v.y, h[1], l[1]); split_bf16(v.z, h[2], l[2]); split_bf16(v.w, h[3], l[3]);
      __nv_bfloat16* row = e.out + r * 2 * e.ld_half + c;
      *reinterpret_cast<uint2*>(row) = *reinterpret_cast<const uint2*>(h);
      *reinterpret_cast<uint2*>(row + e.ld_half) = *reinterpret_cast<const uint2*>(l);
    }
  }
}

// dst[r, c] += src[r, c]  (dL/dz contributions of the lin_z layers)
__global__ void __launch_bounds__(256) x3_add_kernel(float* __restrict__ dst, int ld_dst, const float* __restrict__ src,
                                                     int ld_src, int64_t rows, int cols) {
  const int c4 = cols / 4;
  const int64_t n = rows * c4;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = t / c4;
    const int c = (int)(t % c4) * 4;
    float4 a = *reinterpret_cast<float4*>(dst + r * ld_dst + c);
    const float4 b = *reinterpret_cast<const float4*>(src + r * ld_src + c);
    a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    *reinterpret_cast<float4*>(dst + r * ld_dst + c) = a;
  }
}

int x3_epi(const X3Epi& e, cudaStream_t s) {
  const int64_t n = e.rows * (e.cols / 4);
  int64_t blocks = (n + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  { LaunchScope ls_(NRF_CAT_MISC, s);
  x3_epi_kernel<<<(unsigned)blocks, 256, 0, s>>>(e);
  }
  NRF_LAUNCH_OK();
  return NRF_OK;
}

// [hi | lo] rows of an fp32 matrix (N, cols) with pitch ld_src -> (N, 2 * ld_half) bf16
int x3_split(const float* src, int ld_src, __nv_bfloat16* dst, int ld_half, int64_t rows, int cols, cudaStream_t s) {
  X3Epi e;
  memset(&e, 0, sizeof(e));
  e.acc = src; e.ld_acc = ld_src; e.out = dst; e.ld_half = ld_half; e.rows = rows; e.cols = cols;
  return x3_epi(e, s);
}

// acc (N, n_store) fp32 = [A(2K) | A_hi(K)] . Wx^T (+ bias), Wx rows [hi | hi | lo] of pitch 3K
int x3_gemm(const __nv_bfloat16* A, int K, const char* Wx, int n_rows_pad, int n_store, const float* bias,
            float* out, int ldo, int64_t N, cudaStream_t s) {
  NrfGemm g;
  memset(&g, 0, sizeof(g));
  g.M = (int)N; g.N = n_rows_pad; g.n_store = n_store;
  g.A[0] = A; g.K[0] = 2 * K; g.lda[0] = 2 * K;
  g.A[1] = A; g.K[1] = K; g.lda[1] = 2 * K;
  g.B = Wx; g.ldb = 3 * K;
  g.bias = bias;
  g.out_f32 = out; g.ldo = ldo;
  OpFmt f = {0, 0, 0};
  return gemm_tc_launch(g, f, s);
}

// the lin_z tail: acc2 = [z_hi | z_lo | z_hi] . [Wz_hi | Wz_hi | Wz_lo]^T with z = columns [0, C) of both halves of FIN
int x3_gemm_z(const __nv_bfloat16* FIN, int kin_pad, int C, const char* Wzx, int H, float* out, int64_t N,
              cudaStream_t s) {
  NrfGemm g;
  memset(&g, 0, sizeof(g));
  g.M = (int)N; g.N = H; g.n_store = H;
  g.A[0] = FIN; g.K[0] = C; g.lda[0] = 2 * kin_pad;
  g.A[1] = FIN + kin_pad; g.K[1] = C; g.lda[1] = 2 * kin_pad;
  g.A[2] = FIN; g.K[2] = C; g.lda[2] = 2 * kin_pad;
  g.B = Wzx; g.ldb = 3 * C;
  g.out_f32 = out; g.ldo = H;
  OpFmt f = {0, 0, 0};
  return gemm_tc_launch(g, f, s);
}

// dW (n_valid, k_valid) += G^T A for split operands: G rows [hi(Nn) | lo(Nn)] (pitch ldg), A_hi / A_lo given separately
int x3_wgrad(const __nv_bfloat16* G, int Nn, int ldg, const __nv_bfloat16* A_hi, const __nv_bfloat16* A_lo, int lda,
             int K, int n_valid, int k_valid, float* dW, int ldw, float* dbias, void* ws, int64_t M, cudaStream_t s) {
  if (!dW) return NRF_OK;
  OpFmt f = {0, 0, 0};
  int rc;
  if ((rc = wgrad_tc_launch(G, ldg, A_hi, lda, (int)M, Nn, K, n_valid, k_valid, dW, ldw, dbias, ws, f, s))) return rc;
  if ((rc = wgrad_tc_launch(G, ldg, A_lo, lda, (int)M, Nn, K, n_valid, k_valid, dW, ldw, nullptr, ws, f, s))) return rc;
  return wgrad_tc_launch(G + Nn, ldg, A_hi, lda, (int)M, Nn, K, n_valid, k_valid, dW, ldw, dbias, ws, f, s);
}

// buffer carving (both passes): byte offsets per sample-row buffer, all multiples of 16 B per row
struct X3Acts {
  int64_t fin, ax0, an0, stream, acc, acc2, layer;   // byte offsets; ax(b) = ax0 + b * layer, an(b) = an0 + b * layer
  int64_t total;
};
X3Acts x3_acts(const X3Layout& L, int64_t N) {
  X3Acts a;
  int64_t off = 0;
  auto take = [&](int64_t bytes) { int64_t o = off; off = rup(off + bytes, 1024); return o; };
  a.layer = rup(N * 2 * L.H * 2, 1024);
  a.fin = take(N * 2 * L.kin_pad * 2);
  a.ax0 = take(a.layer * (L.nb + 1));
  a.an0 = take(a.layer * L.nb);
  a.stream = take(N * L.H * 4);
  a.acc = take(N * L.H * 4);
  a.acc2 = take(N * L.H * 4);
  a.total = off;
  return a;
}

}  // namespace

int mlp_x3_sizes(const NrfMlpParams* p, NrfMlpSizes* out) {
  X3Layout L;
  int rc = x3_layout(p, &L);
  if (rc) return rc;
  out->kin_pad = L.kin_pad;
  out->dout_pad = L.dout_pad;
  out->packed_bytes = L.total;
  // per sample (+ 1 KB alignment slack per buffer, charged to the fixed part): FIN, 2 nb + 1 operand layers, the fp32
  // residual stream and two fp32 GEMM results
  out->fwd_bytes_per_sample = (int64_t)2 * L.kin_pad * 2 + (int64_t)(2 * L.nb + 1) * 2 * L.H * 2 + (int64_t)3 * L.H * 4 + 64;
  // backward: split d_field, the fp32 gradient stream, two [hi | lo] gradient operands, one fp32 GEMM result, dL/dz part
  out->bwd_bytes_per_sample = (int64_t)2 * L.dout_pad * 2 + (int64_t)L.H * 4 + (int64_t)2 * 2 * L.H * 2 +
                              (int64_t)L.H * 4 + (int64_t)L.cpad * 4 + 64;
  out->bwd_fixed_bytes = rup(nrf_wgrad_workspace_bytes(L.H, L.H), 1024) + 16 * 1024;
  return NRF_OK;
}

int mlp_x3_pack(const NrfMlpParams* p, void* packed, cudaStream_t s) {
  X3Layout L;
  int rc = x3_layout(p, &L);
  if (rc) return rc;
  NRF_CUDA_OK(cudaMemsetAsync(packed, 0, (size_t)L.total, s));
  X3Table tab;
  tab.n = 0;
  auto mat = [&](int64_t off, int kpad, int rows, int cols, const float* src, int ld_src, int transpose, int col0) {
    if (!src || tab.n >= kX3MaxSegs) return;
    X3Seg& g = tab.seg[tab.n++];
    g.src = src; g.src2 = nullptr; g.dst_off = off; g.kpad = kpad; g.rows = rows; g.cols = cols; g.ld_src = ld_src;
    g.mode = transpose ? 1 : 0; g.col0 = col0;
  };
  auto bias = [&](int64_t off, const float* a, const float* b, int n) {
    if (tab.n >= kX3MaxSegs) return;
    X3Seg& g = tab.seg[tab.n++];
    g.src = a; g.src2 = b; g.dst_off = off; g.kpad = n; g.rows = 1; g.cols = n; g.ld_src = n; g.mode = 2; g.col0 = 0;
  };
  if (L.nz > 0) mat(L.W0, L.kin_pad, L.H, L.C, p->lin_z_w[0], L.C, 0, 0);
  mat(L.W0, L.kin_pad, L.H, L.Din, p->lin_in_w, L.Din, 0, L.C);
  bias(L.bias0, p->lin_in_b, L.nz > 0 ? p->lin_z_b[0] : nullptr, L.H);
  for (int b = 0; b < L.nb; ++b) {
    mat(L.Wfc0[b], L.H, L.H, L.H, p->fc0_w[b], L.H, 0, 0);
    mat(L.Wfc1[b], L.H, L.H, L.H, p->fc1_w[b], L.H, 0, 0);
    const bool cat = b + 1 < L.nz;
    if (cat) mat(L.Wz[b + 1], L.C, L.H, L.C, p->lin_z_w[b + 1], L.C, 0, 0);
    bias(L.bias1[b], p->fc1_b[b], cat ? p->lin_z_b[b + 1] : nullptr, L.H);
    mat(L.Wfc0T[b], L.H, L.H, L.H, p->fc0_w[b], L.H, 1, 0);
    mat(L.Wfc1T[b], L.H, L.H, L.H, p->fc1_w[b], L.H, 1, 0);
    if (b < L.nz) mat(L.WzT[b], L.H, L.C, L.H, p->lin_z_w[b], L.C, 1, 0);
  }
  mat(L.Wout, L.H, L.Dout, L.H, p->lin_out_w, L.H, 0, 0);
  bias(L.bias_out, p->lin_out_b, nullptr, L.Dout);
  mat(L.WoutT, L.dout_pad, L.H, L.Dout, p->lin_out_w, L.H, 1, 0);
  NRF_REQUIRE(tab.n < kX3MaxSegs, NRF_ENOSUP, "nrf_mlp_pack(bf16x3): segment table overflow");
  { LaunchScope ls_(NRF_CAT_MISC, s);
  x3_pack_kernel<<<dim3(64, tab.n), 256, 0, s>>>(tab, reinterpret_cast<char*>(packed));
  }
  NRF_LAUNCH_OK();
  return NRF_OK;
}

int mlp_x3_fwd(const NrfMlpParams* p, const void* packed, const float* field_in, int64_t N, void* acts,
               float* field_out, cudaStream_t s) {
  NRF_REQUIRE(acts, NRF_EINVAL, "nrf_mlp_fwd(bf16x3): the activation buffer is required");
  X3Layout L;
  int rc = x3_layout(p, &L);
  if (rc) return rc;
  const X3Acts a = x3_acts(L, N);
  const char* W = reinterpret_cast<const char*>(packed);
  char* act = reinterpret_cast<char*>(acts);
  auto bf = [&](int64_t off) { return reinterpret_cast<__nv_bfloat16*>(act + off); };
  auto fl = [&](int64_t off) { return reinterpret_cast<float*>(act + off); };
  auto bias = [&](int64_t off) { return reinterpret_cast<const float*>(W + off); };
  __nv_bfloat16* FIN = bf(a.fin);
  float* X = fl(a.stream);
  float* acc = fl(a.acc);
  float* acc2 = fl(a.acc2);
#define TRY(x) do { rc = (x); if (rc) return rc; } while (0)
  TRY(x3_split(field_in, L.kin_pad, FIN, L.kin_pad, N, L.kin_pad, s));
  // x'_0 = [z | p] . [W_z0 | W_in]^T + (b_in + b_z0)
  TRY(x3_gemm(FIN, L.kin_pad, W + L.W0, L.H, L.H, bias(L.bias0), acc, L.H, N, s));
  X3Epi e;
  memset(&e, 0, sizeof(e));
  e.rows = N; e.cols = L.H; e.ld_acc = L.H; e.ld_acc2 = L.H; e.ld_half = L.H; e.relu = 1;
  e.acc = acc; e.stream = X; e.write_stream = 1; e.out = bf(a.ax0);
  TRY(x3_epi(e, s));
  for (int b = 0; b < L.nb; ++b) {
    // net_b = relu(x'_b) . W_fc0^T + b_fc0
    TRY(x3_gemm(bf(a.ax0 + b * a.layer), L.H, W + L.Wfc0[b], L.H, L.H, p->fc0_b[b], acc, L.H, N, s));
    X3Epi n = e;
    n.stream = nullptr; n.write_stream = 0; n.has_resid = 0; n.acc2 = nullptr; n.out = bf(a.an0 + b * a.layer);
    TRY(x3_epi(n, s));
    // x'_{b+1} = x'_b + relu(net_b) . W_fc1^T (+ z . W_z,b+1^T) + biases
    TRY(x3_gemm(bf(a.an0 + b * a.layer), L.H, W + L.Wfc1[b], L.H, L.H, bias(L.bias1[b]), acc, L.H, N, s));
    const bool cat = b + 1 < L.nz;
    if (cat) TRY(x3_gemm_z(FIN, L.kin_pad, L.C, W + L.Wz[b + 1], L.H, acc2, N, s));
    X3Epi x = e;
    x.acc2 = cat ? acc2 : nullptr; x.has_resid = 1; x.write_stream = b + 1 < L.nb; x.out = bf(a.ax0 + (b + 1) * a.layer);
    TRY(x3_epi(x, s));
  }
  TRY(x3_gemm(bf(a.ax0 + L.nb * a.layer), L.H, W + L.Wout, L.nout_pad, (int)rup(L.Dout, 4), bias(L.bias_out), field_out,
              (int)rup(L.Dout, 4), N, s));
#undef TRY
  return NRF_OK;
}

int mlp_x3_bwd(const NrfMlpParams* p, const void* packed, int64_t N, const void* acts, const float* d_field,
               const NrfMlpGrads* gr, float* dlatent, void* scratch, cudaStream_t s) {
  X3Layout L;
  int rc = x3_layout(p, &L);
  if (rc) return rc;
  NRF_REQUIRE(L.nz == 0 || dlatent, NRF_EINVAL, "nrf_mlp_bwd(bf16x3): dlatent is required when d_latent > 0");
  const X3Acts a = x3_acts(L, N);
  const char* W = reinterpret_cast<const char*>(packed);
  const char* act = reinterpret_cast<const char*>(acts);
  auto abf = [&](int64_t off) { return reinterpret_cast<const __nv_bfloat16*>(act + off); };
  const __nv_bfloat16* FIN = abf(a.fin);
  char* sc = reinterpret_cast<char*>(scratch);
  int64_t off = rup(nrf_wgrad_workspace_bytes(L.H, L.H), 1024);
  void* wws = gr->deterministic ? sc : nullptr;
  auto take = [&](int64_t bytes) { char* o = sc + off; off = rup(off + bytes, 1024); return o; };
  __nv_bfloat16* DF = reinterpret_cast<__nv_bfloat16*>(take(N * 2 * L.dout_pad * 2));
  float* Gf = reinterpret_cast<float*>(take(N * L.H * 4));
  __nv_bfloat16* GX = reinterpret_cast<__nv_bfloat16*>(take(N * 2 * L.H * 2));
  __nv_bfloat16* DN = reinterpret_cast<__nv_bfloat16*>(take(N * 2 * L.H * 2));
  float* acc = reinterpret_cast<float*>(take(N * L.H * 4));
  float* accz = reinterpret_cast<float*>(take(N * L.cpad * 4));
#define TRY(x) do { rc = (x); if (rc) return rc; } while (0)
  TRY(x3_split(d_field, L.dout_pad, DF, L.dout_pad, N, L.dout_pad, s));
  const __nv_bfloat16* AXn = abf(a.ax0 + L.nb * a.layer);
  TRY(x3_wgrad(DF, L.dout_pad, 2 * L.dout_pad, AXn, AXn + L.H, 2 * L.H, L.H, L.Dout, L.H, gr->lin_out_w, L.H,
               gr->lin_out_b, wws, N, s));
  // dL/dx_nb = (d_field . W_out) gated by relu(x_nb) > 0
  TRY(x3_gemm(DF, L.dout_pad, W + L.WoutT, L.H, L.H, nullptr, acc, L.H, N, s));
  X3Epi e;
  memset(&e, 0, sizeof(e));
  e.rows = N; e.cols = L.H; e.ld_acc = L.H; e.ld_half = L.H; e.relu = 0;
  e.acc = acc; e.mask = AXn; e.ld_mask = 2 * L.H; e.stream = Gf; e.write_stream = 1; e.out = GX;
  TRY(x3_epi(e, s));
  bool dz_first = true;
  for (int b = L.nb - 1; b >= 0; --b) {
    const __nv_bfloat16* AN = abf(a.an0 + b * a.layer);
    const __nv_bfloat16* AX = abf(a.ax0 + b * a.layer);
    TRY(x3_wgrad(GX, L.H, 2 * L.H, AN, AN + L.H, 2 * L.H, L.H, L.H, L.H, gr->fc1_w[b], L.H, gr->fc1_b[b], wws, N, s));
    if (b + 1 < L.nz)
      TRY(x3_wgrad(GX, L.H, 2 * L.H, FIN, FIN + L.kin_pad, 2 * L.kin_pad, (int)rup(L.C, 64), L.H, L.C,
                   gr->lin_z_w[b + 1], L.C, gr->lin_z_b[b + 1], wws, N, s));
    // dL/dnet_b = (dL/dx_{b+1} . W_fc1) gated by relu(net_b) > 0
    TRY(x3_gemm(GX, L.H, W + L.Wfc1T[b], L.H, L.H, nullptr, acc, L.H, N, s));
    X3Epi n = e;
    n.mask = AN; n.stream = nullptr; n.write_stream = 0; n.has_resid = 0; n.out = DN;
    TRY(x3_epi(n, s));
    TRY(x3_wgrad(DN, L.H, 2 * L.H, AX, AX + L.H, 2 * L.H, L.H, L.H, L.H, gr->fc0_w[b], L.H, gr->fc0_b[b], wws, N, s));
    // dL/dx'_b = dL/dx_{b+1} + (dL/dnet_b . W_fc0) gated by relu(x'_b) > 0
    TRY(x3_gemm(DN, L.H, W + L.Wfc0T[b], L.H, L.H, nullptr, acc, L.H, N, s));
    X3Epi x = e;
    x.mask = AX; x.has_resid = 1; x.write_stream = b > 0; x.out = GX;
    TRY(x3_epi(x, s));
    if (b < L.nz) {          // dL/dz += dL/dx'_b . W_z,b
      float* dst = dz_first ? dlatent : accz;
      TRY(x3_gemm(GX, L.H, W + L.WzT[b], L.cpad, L.C, nullptr, dst, L.C, N, s));
      if (!dz_first) {
        const int64_t n4 = N * (L.C / 4);
        int64_t blocks = (n4 + 255) / 256;
        const int64_t cap = (int64_t)sm_count() * 16;
        if (blocks > cap) blocks = cap;
        { LaunchScope ls_(NRF_CAT_MISC, s);
        x3_add_kernel<<<(unsigned)blocks, 256, 0, s>>>(dlatent, L.C, accz, L.C, N, L.C);
        }
        NRF_LAUNCH_OK();
      }
      dz_first = false;
    }
  }
  // first layer: x'_0 = [z | p] . [W_z0 | W_in]^T; GX = dL/dx'_0
  if (L.nz > 0)
    TRY(x3_wgrad(GX, L.H, 2 * L.H, FIN, FIN + L.kin_pad, 2 * L.kin_pad, (int)rup(L.C, 64), L.H, L.C, gr->lin_z_w[0],
                 L.C, gr->lin_z_b[0], wws, N, s));
  TRY(x3_wgrad(GX, L.H, 2 * L.H, FIN + L.C, FIN + L.kin_pad + L.C, 2 * L.kin_pad, L.kin_pad - L.C, L.H, L.Din,
               gr->lin_in_w, L.Din, gr->lin_in_b, wws, N, s));
#undef TRY
  return NRF_OK;
}

}  // namespace nrf
