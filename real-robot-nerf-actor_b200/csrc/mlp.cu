// The ResnetFC field MLP (resnetfc.py:55-64,146-195) as a chain of fused-epilogue GEMMs.
//
// Algebra (n_lin_z = min(combine_layer, n_blocks), z = latent, p = [PE | viewdir]):
//   x'_0     = [z | p] . [W_z0 | W_in]^T + (b_in + b_z0)                     one GEMM, K = kin_pad
//   net_b    = relu(x'_b) . W_fc0[b]^T + b_fc0[b]
//   x'_{b+1} = x'_b + [relu(net_b) | z] . [W_fc1[b] | W_z,b+1]^T + (b_fc1[b] + b_z,b+1)
//   out      = relu(x_nb) . W_out^T + b_out
// i.e. "x = x + lin_z[b](z)" (resnetfc.py:182-188) is folded into the preceding GEMM by
// concatenating along K; combine_interleaved (utils.py:509-519) is the identity here.
// Every GEMM writes the next layer's operand (ReLU'd, operand-typed) from its epilogue, so the
// forward leaves exactly the tensors the backward needs: relu(x'_b) and relu(net_b).  The residual
// stream x' itself lives in ONE operand-typed buffer updated in place (bf16 in the tensor-core mode:
// SURVEY.md section 10 measured no error on top of bf16 operands; fp32 in the parity mode).
// Backward, per block (g = dL/dx_{b+1}, one buffer updated in place):
//   dW_fc1 += g^T relu(net_b);  dnet = (g . W_fc1) gated by net_b > 0;  dW_fc0 += dnet^T relu(x'_b)
//   g <- g + (dnet . W_fc0) gated by x'_b > 0        (kept per block b < n_lin_z for dL/dz)
//   dL/dz = [g'_0 | g'_1 | g'_2] . [W_z0 ; W_z1 ; W_z2]                       one GEMM, K = 3 H
#include <stdlib.h>
#include <string.h>
#include <type_traits>
#include "gemm_common.cuh"

namespace nrf {

static inline int64_t round_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }
// tensor-core modes: 16-bit operands (bf16, or fp16 forward operands with bf16 gradients), fp32 accumulate
static inline bool is_tc(int precision) { return precision == NRF_PREC_BF16 || precision == NRF_PREC_FP16; }

struct MlpLayout {
  int H, C, Din, Dout, nb, nz;
  int kin_pad, dout_pad, nout_pad;
  size_t es;                 // operand element size
  // byte offsets into the packed buffer
  int64_t W0, bias0, Wout, bias_out, WoutT;
  int64_t Wfc0[NRF_MAX_BLOCKS], Wfc1[NRF_MAX_BLOCKS], bias1[NRF_MAX_BLOCKS];
  int64_t Wfc0T[NRF_MAX_BLOCKS], Wfc1T[NRF_MAX_BLOCKS], WzcatT;
  int k1cat[NRF_MAX_BLOCKS];  // K of the concatenated fc_1 weight
  int64_t total;
};

static int make_layout(const NrfMlpParams* p, int precision, MlpLayout* L) {
  NRF_REQUIRE(p, NRF_EINVAL, "mlp: null params");
  NRF_REQUIRE(p->n_blocks >= 1 && p->n_blocks <= NRF_MAX_BLOCKS && p->n_lin_z >= 0 &&
                  p->n_lin_z <= p->n_blocks, NRF_EINVAL, "mlp: n_blocks=%d n_lin_z=%d", p->n_blocks, p->n_lin_z);
  NRF_REQUIRE(p->d_in > 0 && p->d_hidden > 0 && p->d_out > 0 && p->d_latent >= 0, NRF_EINVAL, "mlp: bad dims");
  NRF_REQUIRE(is_tc(precision) || precision == NRF_PREC_FP32, NRF_EINVAL, "mlp: precision %d", precision);
  L->H = p->d_hidden; L->C = p->d_latent; L->Din = p->d_in; L->Dout = p->d_out;
  L->nb = p->n_blocks; L->nz = p->d_latent > 0 ? p->n_lin_z : 0;
  L->es = is_tc(precision) ? 2 : 4;
  L->kin_pad = (int)round_up(L->C + L->Din, 64);
  L->dout_pad = (int)round_up(L->Dout, 64);
  L->nout_pad = (int)round_up(L->Dout, 128);
  if (is_tc(precision)) {
    NRF_REQUIRE(L->H % 128 == 0, NRF_ENOSUP, "mlp(bf16): d_hidden=%d must be a multiple of 128", L->H);
    NRF_REQUIRE(L->C % 64 == 0 && L->C > 0, NRF_ENOSUP,
                "mlp(bf16): d_latent=%d must be a multiple of 64", L->C);
    NRF_REQUIRE(L->nz <= 3, NRF_ENOSUP, "mlp(bf16): n_lin_z=%d > 3", L->nz);
  }
  int64_t off = 0;
  auto take = [&](int64_t bytes) { int64_t o = off; off = round_up(off + bytes, 1024); return o; };
  L->W0 = take((int64_t)L->H * L->kin_pad * L->es);
  L->bias0 = take((int64_t)L->H * 4);
  for (int b = 0; b < L->nb; ++b) {
    L->k1cat[b] = L->H + ((b + 1 < L->nz) ? L->C : 0);
    L->Wfc0[b] = take((int64_t)L->H * L->H * L->es);
    L->Wfc1[b] = take((int64_t)L->H * L->k1cat[b] * L->es);
    L->bias1[b] = take((int64_t)L->H * 4);
    L->Wfc0T[b] = take((int64_t)L->H * L->H * L->es);
    L->Wfc1T[b] = take((int64_t)L->H * L->H * L->es);
  }
  L->WzcatT = take((int64_t)round_up(L->C > 0 ? L->C : 1, 128) * (L->nz > 0 ? L->nz : 1) * L->H * L->es);
  L->Wout = take((int64_t)L->nout_pad * L->H * L->es);
  L->bias_out = take((int64_t)L->nout_pad * 4);
  L->WoutT = take((int64_t)L->H * L->dout_pad * L->es);
  L->total = off;
  return NRF_OK;
}

// Packing = ONE kernel launch over a table of segments (it is re-run every training step: an optimizer writing through
// `p.data` leaves no trace a cached copy could be keyed on).
//   mode 0: dst[r*ld_dst + c] = T(src[r*ld_src + c])      mode 1: dst[r*ld_dst + c] = T(src[c*ld_src + r])
//   mode 2: fp32 dst[c] = src[c] + src2[c]  (merged biases; either source may be NULL)
struct PackSeg {
  const float* src; const float* src2;
  int64_t dst_off;            // bytes from the packed base
  int ld_dst, rows, cols, ld_src, mode, pad_;
};
constexpr int kMaxPackSegs = 72;
struct PackTable { int n; PackSeg seg[kMaxPackSegs]; };

template <typename T>
__global__ void __launch_bounds__(256) pack_table_kernel(const __grid_constant__ PackTable tab, char* __restrict__ base) {
  const PackSeg& sg = tab.seg[blockIdx.y];
  const int64_t n = (int64_t)sg.rows * sg.cols;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(t / sg.cols), c = (int)(t % sg.cols);
    if (sg.mode == 2) {
      reinterpret_cast<float*>(base + sg.dst_off)[c] = (sg.src ? sg.src[c] : 0.0f) + (sg.src2 ? sg.src2[c] : 0.0f);
      continue;
    }
    const float v = sg.mode == 1 ? sg.src[(int64_t)c * sg.ld_src + r] : sg.src[(int64_t)r * sg.ld_src + c];
    T* dst = reinterpret_cast<T*>(base + sg.dst_off) + (int64_t)r * sg.ld_dst + c;
    if constexpr (std::is_same<T, __nv_bfloat16>::value) *dst = __float2bfloat16_rn(v);
    else if constexpr (std::is_same<T, __half>::value) {
      // NRF_PREC_FP16: the forward's matrices are fp16; the transposed ones belong to the backward, which is bf16
      if (sg.mode == 1) *reinterpret_cast<__nv_bfloat16*>(dst) = __float2bfloat16_rn(v);
      else *dst = __float2half_rn(v);
    } else *dst = v;
  }
}

// fp16 -> bf16 copy of a forward operand the (bf16) backward multiplies with: (rows, cols) with cols % 8 == 0
__global__ void __launch_bounds__(256) half_to_bf16_kernel(const __half* __restrict__ src, int ld_src,
                                                           __nv_bfloat16* __restrict__ dst, int ld_dst, int64_t rows,
                                                           int cols) {
  const int c8 = cols / 8;
  const int64_t n = rows * c8;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = t / c8;
    const int c = (int)(t % c8) * 8;
    const uint4 u = *reinterpret_cast<const uint4*>(src + r * ld_src + c);
    const __half2* h = reinterpret_cast<const __half2*>(&u);
    uint4 o;
    __nv_bfloat162* q = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = __half22float2(h[i]);
      q[i] = __floats2bfloat162_rn(f.x, f.y);
    }
    *reinterpret_cast<uint4*>(dst + r * ld_dst + c) = o;
  }
}
static int half_to_bf16(const void* src, int ld_src, void* dst, int ld_dst, int64_t rows, int cols, cudaStream_t s) {
  const int64_t n = rows * (cols / 8);
  int64_t blocks = (n + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  { LaunchScope ls_(NRF_CAT_MISC, s);
  half_to_bf16_kernel<<<(unsigned)blocks, 256, 0, s>>>(reinterpret_cast<const __half*>(src), ld_src,
                                                       reinterpret_cast<__nv_bfloat16*>(dst), ld_dst, rows, cols);
  }
  NRF_LAUNCH_OK();
  return NRF_OK;
}

template <typename T>
static int pack_all(const NrfMlpParams* p, const MlpLayout& L, void* packed, cudaStream_t s) {
  NRF_CUDA_OK(cudaMemsetAsync(packed, 0, (size_t)L.total, s));        // the zero padding between / inside the matrices
  PackTable tab;
  tab.n = 0;
  auto mat = [&](int64_t off, int ld_dst, int rows, int cols, const float* src, int ld_src, int transpose, int col0) {
    if (!src || tab.n >= kMaxPackSegs) return;
    PackSeg& g = tab.seg[tab.n++];
    g.src = src; g.src2 = nullptr; g.dst_off = off + (int64_t)col0 * (int64_t)sizeof(T);
    g.ld_dst = ld_dst; g.rows = rows; g.cols = cols; g.ld_src = ld_src; g.mode = transpose ? 1 : 0; g.pad_ = 0;
  };
  auto bias = [&](int64_t off, const float* a, const float* b, int n) {
    if (tab.n >= kMaxPackSegs) return;
    PackSeg& g = tab.seg[tab.n++];
    g.src = a; g.src2 = b; g.dst_off = off; g.ld_dst = n; g.rows = 1; g.cols = n; g.ld_src = n; g.mode = 2; g.pad_ = 0;
  };
  // W0 = [W_z0 | W_in | 0]
  if (L.nz > 0) mat(L.W0, L.kin_pad, L.H, L.C, p->lin_z_w[0], L.C, 0, 0);
  mat(L.W0, L.kin_pad, L.H, L.Din, p->lin_in_w, L.Din, 0, L.C);
  bias(L.bias0, p->lin_in_b, L.nz > 0 ? p->lin_z_b[0] : nullptr, L.H);
  for (int b = 0; b < L.nb; ++b) {
    mat(L.Wfc0[b], L.H, L.H, L.H, p->fc0_w[b], L.H, 0, 0);
    mat(L.Wfc1[b], L.k1cat[b], L.H, L.H, p->fc1_w[b], L.H, 0, 0);
    const bool cat = b + 1 < L.nz;
    if (cat) mat(L.Wfc1[b], L.k1cat[b], L.H, L.C, p->lin_z_w[b + 1], L.C, 0, L.H);
    bias(L.bias1[b], p->fc1_b[b], cat ? p->lin_z_b[b + 1] : nullptr, L.H);
    mat(L.Wfc0T[b], L.H, L.H, L.H, p->fc0_w[b], L.H, 1, 0);
    mat(L.Wfc1T[b], L.H, L.H, L.H, p->fc1_w[b], L.H, 1, 0);
  }
  for (int b = 0; b < L.nz; ++b) mat(L.WzcatT, L.nz * L.H, L.C, L.H, p->lin_z_w[b], L.C, 1, b * L.H);
  mat(L.Wout, L.H, L.Dout, L.H, p->lin_out_w, L.H, 0, 0);
  bias(L.bias_out, p->lin_out_b, nullptr, L.Dout);
  mat(L.WoutT, L.dout_pad, L.H, L.Dout, p->lin_out_w, L.H, 1, 0);
  NRF_REQUIRE(tab.n < kMaxPackSegs, NRF_ENOSUP, "nrf_mlp_pack: segment table overflow");
  { LaunchScope ls_(NRF_CAT_MISC, s);
  pack_table_kernel<T><<<dim3(64, tab.n), 256, 0, s>>>(tab, reinterpret_cast<char*>(packed));
  }
  NRF_LAUNCH_OK();
  return NRF_OK;
}

static inline NrfGemm gemm_init(int64_t M, int N, int n_store) {
  NrfGemm g;
  memset(&g, 0, sizeof(g));
  g.M = (int)M; g.N = N; g.n_store = n_store;
  return g;
}
static inline void set_a(NrfGemm& g, int i, const void* A, int K, int lda) {
  g.A[i] = A; g.K[i] = K; g.lda[i] = lda;
}

// grad: a GEMM of the backward: bf16 operands in both tensor-core modes (see NRF_PREC_FP16)
static int run_gemm(const NrfGemm& g, int precision, bool grad, cudaStream_t s) {
  return is_tc(precision) ? gemm_tc_launch(g, grad ? kFmtBf16 : op_fmt(precision), s) : gemm_simt_launch(g, s);
}

// dW += G^T A and (fused, from the same G tiles) dbias += column sums of G.
// `ws`: split-reduction workspace (deterministic mode of the tensor-core kernel), may be NULL.
static int run_wgrad(const void* G, int ldg, const void* A, int lda, int64_t M, int N, int K, int n_valid,
                     int k_valid, float* dW, int ldw, float* dbias, void* ws, int precision, cudaStream_t s) {
  if (!dW) return NRF_OK;
  return is_tc(precision)
             ? wgrad_tc_launch(G, ldg, A, lda, (int)M, N, K, n_valid, k_valid, dW, ldw, dbias, ws, kFmtBf16, s)
             : wgrad_simt_launch(G, ldg, A, lda, (int)M, N, K, n_valid, k_valid, dW, ldw, dbias, s);
}

}  // namespace nrf

using namespace nrf;

extern "C" int nrf_mlp_sizes(const NrfMlpParams* p, int precision, NrfMlpSizes* out) {
  NRF_REQUIRE(out, NRF_EINVAL, "nrf_mlp_sizes: null out");
  if (precision == NRF_PREC_BF16X3) return mlp_x3_sizes(p, out);
  MlpLayout L;
  int rc = make_layout(p, precision, &L);
  if (rc) return rc;
  out->kin_pad = L.kin_pad;
  out->dout_pad = L.dout_pad;
  out->packed_bytes = L.total;
  out->fwd_bytes_per_sample = (int64_t)(2 * L.nb + 2) * L.H * L.es;
  // layer-by-layer chain: dL/dx, dL/dnet and the n_lin_z kept gradients; fused chain: every dL/dx'_b and dL/dnet_b
  out->bwd_bytes_per_sample = (int64_t)(2 * L.nb + 1 > 2 + L.nz ? 2 * L.nb + 1 : 2 + L.nz) * L.H * L.es;
  if (precision == NRF_PREC_FP16)       // bf16 copies for the backward: field_in, and (layer-by-layer chain) one layer
    out->bwd_bytes_per_sample += (int64_t)(L.kin_pad + L.H) * 2;
  out->bwd_fixed_bytes = (int64_t)round_up(nrf_wgrad_workspace_bytes(L.H, L.H), 1024) + 4096;   // + alignment slack
  return NRF_OK;
}

extern "C" int nrf_mlp_pack(const NrfMlpParams* p, int precision, void* packed, void* stream) {
  NRF_REQUIRE(packed, NRF_EINVAL, "nrf_mlp_pack: null buffer");
  if (precision == NRF_PREC_BF16X3) return mlp_x3_pack(p, packed, as_stream(stream));
  MlpLayout L;
  int rc = make_layout(p, precision, &L);
  if (rc) return rc;
  if (precision == NRF_PREC_BF16) return pack_all<__nv_bfloat16>(p, L, packed, as_stream(stream));
  if (precision == NRF_PREC_FP16) return pack_all<__half>(p, L, packed, as_stream(stream));
  return pack_all<float>(p, L, packed, as_stream(stream));
}

// The fused kernel is built for the reference's field MLP: 512 hidden units, a latent of 64 or 128 channels that
// fills whole k-blocks, PE + viewdir in one k-block.  Anything else runs the layer-by-layer chain below.
static bool fused_supported(const MlpLayout& L) {
  // (2 nb + 1) slots x 64 B of gate bits per sample must fit the spare activation layer (H * es bytes per sample)
  return (2 * L.nb + 1) * 64 <= L.H * (int)L.es && L.H == 512 && L.es == 2 && (L.C == 64 || L.C == 128) && L.kin_pad == L.C + 64 && L.nz >= 1 &&
         L.nout_pad <= 640 && 2 * L.nb + 2 <= kFusedMaxLayers;
}

// diagnostics: device buffer (148 x 32 int64) that the next fused launches fill with per-role cycle counters
static void* g_fused_prof = nullptr;
extern "C" void nrf_debug_set_fused_profile(void* device_buffer) { g_fused_prof = device_buffer; }

static int mlp_fwd_fused(const NrfMlpParams* p, const MlpLayout& L, const char* W, const void* field_in, int64_t N,
                         void* acts, float* field_out, int precision, const uint8_t* touch, cudaStream_t s) {
  FusedDesc d;
  memset(&d, 0, sizeof(d));
  d.half = precision == NRF_PREC_FP16;
  const int kbH = L.H / 64, kbC = L.C / 64;
  int l = 0;
  auto add = [&](const void* Wl, int ldw, int kb_main, int kb_z, int kind, int a_src, int first, int slot,
                 const float* bias) {
    FusedLayerDesc& f = d.L[l++];
    f.W = Wl; f.ldw = ldw; f.kb_main = kb_main; f.kb_z = kb_z; f.kind = kind; f.a_src = a_src; f.first = first;
    f.act_slot = slot; f.mask_slot = -1; f.bias = bias;
  };
  // slots of `acts`: relu(x'_b) at b, relu(net_b) at nb + 1 + b (the layout nrf_mlp_bwd reads)
  add(W + L.W0, L.kin_pad, kbC + 1, 0, 0, 0, 1, 0, reinterpret_cast<const float*>(W + L.bias0));
  d.L[0].skip_head = kbC;                          // [latent | PE + viewdir]: the latent k-panels come first
  for (int b = 0; b < L.nb; ++b) {
    add(W + L.Wfc0[b], L.H, kbH, 0, 1, 1, 0, L.nb + 1 + b, p->fc0_b[b]);
    add(W + L.Wfc1[b], L.k1cat[b], kbH, (b + 1 < L.nz) ? kbC : 0, 0, 2, 0, b + 1,
        reinterpret_cast<const float*>(W + L.bias1[b]));
    d.L[l - 1].skip_z = 1;                         // the lin_z[b + 1] tail of fc_1 multiplies the latent
  }
  add(W + L.Wout, L.H, kbH, 0, 2, 1, 0, -1, reinterpret_cast<const float*>(W + L.bias_out));
  d.L[l - 1].n_chunks = L.nout_pad / 128;          // 4 at d_embed = 384 (388 outputs), 5 at d_embed = 512 (516)
  d.n_layers = l;
  d.in = field_in; d.in_cols = L.kin_pad;
  d.N = N;
  d.saves = acts; d.n_slots = 2 * L.nb + 1;
  // the layer-by-layer chain keeps the residual stream in the last layer of `acts`; here it holds the gate bits
  d.gate_bits = acts ? reinterpret_cast<char*>(acts) + (int64_t)(2 * L.nb + 1) * N * L.H * (int64_t)L.es : nullptr;
  d.out = field_out; d.d_out = L.Dout; d.ldo = (int)round_up(L.Dout, 4);   // pad columns receive exact zeros
  d.prof = g_fused_prof;
  d.touch = touch;
  return mlp_fused_launch(d, s);
}

// Backward data-gradient chain in the fused kernel.  G (slots x (N,H) bf16) receives dL/dx'_b at slot b and
// dL/dnet_b at slot nb + 1 + b: the same slot numbering as the forward's `acts`, so that every weight gradient is
// G[slot]^T . acts[slot'] (see nrf_mlp_bwd).
static int mlp_bwd_fused(const MlpLayout& L, const char* W, const void* d_field, int64_t N, const void* acts,
                         void* G, int precision, cudaStream_t s) {
  FusedDesc d;
  memset(&d, 0, sizeof(d));
  (void)precision;                                 // the backward is bf16 in both tensor-core modes
  const int kbH = L.H / 64;
  int l = 0;
  auto add = [&](const void* Wl, int ldw, int kb_main, int kind, int a_src, int first, int slot, int mask_slot) {
    FusedLayerDesc& f = d.L[l++];
    f.W = Wl; f.ldw = ldw; f.kb_main = kb_main; f.kb_z = 0; f.kind = kind; f.a_src = a_src; f.first = first;
    f.act_slot = slot; f.mask_slot = mask_slot; f.bias = nullptr;
  };
  // dL/dx_nb = (d_field . W_out) gated by relu(x_nb) > 0
  // d_field has dout_pad / 64 k-panels: up to 7 of them are loaded into P; if there are more, an EVEN number stays
  // in P (producer and issuer walk the k-blocks in pairs that must not straddle the two sources) and the rest arrive
  // through the ring (d_embed = 512: 576 columns = 6 + 3 panels)
  const int kb_df = L.dout_pad / 64, kb_p = kb_df <= 7 ? kb_df : 6;
  add(W + L.WoutT, L.dout_pad, kb_p, 0, 0, 1, L.nb, L.nb);
  d.L[0].kb_z = kb_df - kb_p;
  d.L[0].ext_col = kb_p * 64;
  for (int b = L.nb - 1; b >= 0; --b) {
    // dL/dnet_b = (dL/dx_{b+1} . W_fc1[b]) gated by relu(net_b) > 0
    add(W + L.Wfc1T[b], L.H, kbH, 1, 1, 0, L.nb + 1 + b, L.nb + 1 + b);
    // dL/dx'_b = dL/dx_{b+1} + (dL/dnet_b . W_fc0[b]) gated by relu(x'_b) > 0
    add(W + L.Wfc0T[b], L.H, kbH, 0, 2, 0, b, b);
  }
  d.n_layers = l;
  d.backward = 1;
  d.in = d_field; d.in_cols = L.dout_pad;
  d.N = N;
  d.saves = G; d.n_slots = 2 * L.nb + 1;
  d.gate_bits = const_cast<char*>(reinterpret_cast<const char*>(acts)) + (int64_t)(2 * L.nb + 1) * N * L.H * (int64_t)L.es;
  d.prof = g_fused_prof;
  return mlp_fused_launch(d, s);
}

extern "C" int nrf_mlp_fused_supported(const NrfMlpParams* p, int precision) {
  MlpLayout L;
  if (make_layout(p, precision, &L)) return 0;
  return is_tc(precision) && fused_supported(L) ? 1 : 0;
}

static int mlp_fwd_impl(const NrfMlpParams* p, const void* packed, int precision, const void* field_in,
                        int64_t N, void* acts, float* field_out, void* stream, bool allow_fused,
                        const uint8_t* touch = nullptr) {
  NRF_REQUIRE(packed && field_in && field_out && N > 0, NRF_EINVAL, "nrf_mlp_fwd: bad arguments");
  NRF_REQUIRE(N < (int64_t)1 << 31, NRF_ENOSUP, "nrf_mlp_fwd: N too large for one call");
  if (precision == NRF_PREC_BF16X3)
    return mlp_x3_fwd(p, packed, reinterpret_cast<const float*>(field_in), N, acts, field_out, as_stream(stream));
  MlpLayout L;
  int rc = make_layout(p, precision, &L);
  if (rc) return rc;
  cudaStream_t s = as_stream(stream);
  const char* W = reinterpret_cast<const char*>(packed);
  static const bool layered = getenv("NRF_MLP_LAYERED") != nullptr;
  if (allow_fused && is_tc(precision) && fused_supported(L) && !layered)
    return mlp_fwd_fused(p, L, W, field_in, N, acts, field_out, precision, touch, s);
  NRF_REQUIRE(acts, NRF_EINVAL, "nrf_mlp_fwd: the layer-by-layer chain needs the activation buffer");
  char* act = reinterpret_cast<char*>(acts);
  const int64_t layer = N * L.H * (int64_t)L.es;
  auto ax = [&](int b) { return act + (int64_t)b * layer; };                 // relu(x'_b), b = 0..nb
  auto an = [&](int b) { return act + (int64_t)(L.nb + 1 + b) * layer; };    // relu(net_b), b = 0..nb-1
  char* xcur = act + (int64_t)(2 * L.nb + 1) * layer;                        // residual stream x'

  NrfGemm g = gemm_init(N, L.H, L.H);
  set_a(g, 0, field_in, L.kin_pad, L.kin_pad);
  g.B = W + L.W0; g.ldb = L.kin_pad;
  g.bias = reinterpret_cast<const float*>(W + L.bias0);
  g.out_act = xcur; g.ldact = L.H;
  g.out_act2 = ax(0); g.ldact2 = L.H; g.relu_act2 = 1;
  rc = run_gemm(g, precision, false, s);
  if (rc) return rc;
  for (int b = 0; b < L.nb; ++b) {
    g = gemm_init(N, L.H, L.H);
    set_a(g, 0, ax(b), L.H, L.H);
    g.B = W + L.Wfc0[b]; g.ldb = L.H;
    g.bias = p->fc0_b[b];
    g.out_act = an(b); g.ldact = L.H; g.relu_act = 1;
    rc = run_gemm(g, precision, false, s);
    if (rc) return rc;
    g = gemm_init(N, L.H, L.H);
    set_a(g, 0, an(b), L.H, L.H);
    if (b + 1 < L.nz) set_a(g, 1, field_in, L.C, L.kin_pad);
    g.B = W + L.Wfc1[b]; g.ldb = L.k1cat[b];
    g.bias = reinterpret_cast<const float*>(W + L.bias1[b]);
    g.resid = xcur; g.ldr = L.H;
    g.out_act = xcur; g.ldact = L.H;                         // x' updated in place; after the last block it is x_nb, the
    g.out_act2 = ax(b + 1); g.ldact2 = L.H; g.relu_act2 = 1; // MLP's second return value (resnetfc.py:192-195: last_feat)
    rc = run_gemm(g, precision, false, s);
    if (rc) return rc;
  }
  g = gemm_init(N, L.nout_pad, (int)round_up(L.Dout, 4));     // rows of field_out are round_up(d_out, 4) floats long;
  set_a(g, 0, ax(L.nb), L.H, L.H);                            // the pad columns come out as exact zeros (zero weight rows)
  g.B = W + L.Wout; g.ldb = L.H;
  g.bias = reinterpret_cast<const float*>(W + L.bias_out);
  g.out_f32 = field_out; g.ldo = (int)round_up(L.Dout, 4);
  return run_gemm(g, precision, false, s);
}

extern "C" int nrf_mlp_fwd(const NrfMlpParams* p, const void* packed, int precision, const void* field_in,
                           int64_t N, void* acts, float* field_out, void* stream) {
  return mlp_fwd_impl(p, packed, precision, field_in, N, acts, field_out, stream, true);
}

extern "C" int nrf_mlp_fwd_touch(const NrfMlpParams* p, const void* packed, int precision, const void* field_in,
                                 int64_t N, void* acts, float* field_out, const uint8_t* touch_flags, void* stream) {
  return mlp_fwd_impl(p, packed, precision, field_in, N, acts, field_out, stream, true, touch_flags);
}

extern "C" int nrf_mlp_fwd_layered(const NrfMlpParams* p, const void* packed, int precision, const void* field_in,
                                   int64_t N, void* acts, float* field_out, void* stream) {
  return mlp_fwd_impl(p, packed, precision, field_in, N, acts, field_out, stream, false);
}

static int mlp_bwd_impl(const NrfMlpParams* p, const void* packed, int precision, const void* field_in,
                        int64_t N, const void* acts, const void* d_field, const NrfMlpGrads* gr,
                        float* dlatent, void* scratch, void* stream, bool force_layered, bool* event_recorded) {
  NRF_REQUIRE(packed && field_in && acts && d_field && gr && scratch && N > 0, NRF_EINVAL,
              "nrf_mlp_bwd: bad arguments");
  NRF_REQUIRE(N < (int64_t)1 << 31, NRF_ENOSUP, "nrf_mlp_bwd: N too large for one call");
  NRF_REQUIRE(!(gr->d_last && precision == NRF_PREC_BF16X3), NRF_ENOSUP, "nrf_mlp_bwd: d_last in the bf16x3 mode");
  if (precision == NRF_PREC_BF16X3)
    return mlp_x3_bwd(p, packed, N, acts, reinterpret_cast<const float*>(d_field), gr, dlatent, scratch,
                      as_stream(stream));
  MlpLayout L;
  int rc = make_layout(p, precision, &L);
  if (rc) return rc;
  NRF_REQUIRE(L.nz == 0 || dlatent, NRF_EINVAL, "nrf_mlp_bwd: dlatent is required when d_latent > 0");
  NRF_REQUIRE(!gr->d_last || force_layered, NRF_ENOSUP,
              "nrf_mlp_bwd: d_last (gradient of the last residual stream) needs nrf_mlp_bwd_layered");
  cudaStream_t s = as_stream(stream);
  const char* W = reinterpret_cast<const char*>(packed);
  const char* act = reinterpret_cast<const char*>(acts);
  const int64_t layer = N * L.H * (int64_t)L.es;
  auto ax = [&](int b) { return act + (int64_t)b * layer; };
  auto an = [&](int b) { return act + (int64_t)(L.nb + 1 + b) * layer; };
  char* sc = reinterpret_cast<char*>(scratch);
  int64_t fixed = round_up(nrf_wgrad_workspace_bytes(L.H, L.H), 1024);
  void* wws = gr->deterministic ? sc : nullptr;   // per-split partial tiles of the weight gradients (ordered reduce)
#define TRY(x) do { rc = (x); if (rc) return rc; } while (0)
  static const bool layered = getenv("NRF_MLP_LAYERED") != nullptr;
  const bool fused = is_tc(precision) && fused_supported(L) && !layered && !force_layered;
  const int64_t n_grad_layers = fused ? 2 * L.nb + 1 : 2 + L.nz;
  // NRF_PREC_FP16: the backward is bf16 (A and B of an MMA must share a format, gradients are bf16), so every forward
  // operand a weight gradient multiplies with must be bf16: the fused forward saved its activations that way; field_in
  // (and, in the layer-by-layer chain, each saved fp16 activation layer right before its use) is converted here
  const bool half = precision == NRF_PREC_FP16;
  char* fin_bf = sc + fixed + n_grad_layers * layer;
  char* act_bf = fin_bf + round_up(N * L.kin_pad * 2, 1024);
  const char* fin = reinterpret_cast<const char*>(field_in);
  if (half) {
    TRY(half_to_bf16(field_in, L.kin_pad, fin_bf, L.kin_pad, N, L.kin_pad, s));
    fin = fin_bf;
  }
  auto act16 = [&](const char* a) -> const char* {   // a saved activation layer as the wgrad's bf16 A operand
    if (!half || fused) return a;
    rc = half_to_bf16(a, L.H, act_bf, L.H, N, L.H, s);
    return act_bf;
  };

  if (fused) {
    // one fused kernel for the whole data-gradient chain, then the weight gradients and dL/dz from its outputs
    char* G = sc + fixed;
    auto gx = [&](int b) { return G + (int64_t)b * layer; };                 // dL/dx'_b, b = 0..nb
    auto gn = [&](int b) { return G + (int64_t)(L.nb + 1 + b) * layer; };    // dL/dnet_b
    TRY(mlp_bwd_fused(L, W, d_field, N, acts, G, precision, s));
    // every weight gradient of the pass in one persistent launch (wgrad_multi.cu); the ordered-reduction mode and
    // NRF_WGRAD_MULTI=0 keep one launch per GEMM
    static const bool multi_off = getenv("NRF_WGRAD_MULTI") && atoi(getenv("NRF_WGRAD_MULTI")) == 0;
    if (!gr->deterministic && !multi_off && L.H == 512) {
      WgmHost h;
      memset(&h, 0, sizeof(h));
      h.M = (int)N;
      h.sync = reinterpret_cast<int*>(sc);        // the split-reduction workspace of the ordered mode is free here
      auto operand = [&](const void* base, int cols, int ld) {
        h.op[h.n_maps].base = base; h.op[h.n_maps].cols = cols; h.op[h.n_maps].ld = ld;
        return h.n_maps++;
      };
      // a problem over the 512 columns of a saved activation layer: two k tiles of four slabs
      auto square = [&](int g_map, int n_valid, const void* a, float* dW, float* db, float* db2) {
        if (!dW) return;                                       // as run_wgrad: no weight gradient wanted, none computed
        WgmProblem& q = h.prob[h.n_prob++];
        q.g_map = g_map; q.n_valid = n_valid; q.ns = 4; q.k_tiles = L.H / 256;
        q.dbias = db ? db : db2; q.dbias2 = db ? db2 : nullptr;
        const int am = operand(a, L.H, L.H);
        for (int j = 0; j < L.H / 64; ++j) {
          q.slab[j].dW = dW + 64 * j; q.slab[j].ldw = L.H; q.slab[j].k_valid = 64;
          q.slab[j].a_map = am; q.slab[j].a_col = 64 * j;
        }
      };
      const int fin_map = operand(fin, L.kin_pad, L.kin_pad);
      const int cs = L.C / 64;                                 // 64-wide slabs of the latent part of the field input
      // G = dL/dx'_b times the latent columns of the field input (lin_z[b]); b = 0 also takes the PE / viewdir columns
      // (lin_in) in the same tile.  Biases: lin_z_b[b] has the gradient of the layer that shares its G (added there).
      auto field_in_problem = [&](int g_map, int b) {
        float* wz = b < L.nz ? gr->lin_z_w[b] : nullptr;
        float* win = b == 0 ? gr->lin_in_w : nullptr;
        if (!wz && !win) return;
        WgmProblem& q = h.prob[h.n_prob++];
        const int pe = b == 0 ? (L.kin_pad - L.C) / 64 : 0;
        const int slabs = (b < L.nz ? cs : 0) + pe;
        q.g_map = g_map; q.n_valid = L.H; q.ns = slabs <= 2 ? 2 : 4; q.k_tiles = (slabs + q.ns - 1) / q.ns;
        int j = 0;
        for (int c = 0; b < L.nz && c < cs; ++c, ++j) {
          q.slab[j].dW = wz ? wz + 64 * c : nullptr; q.slab[j].ldw = L.C; q.slab[j].k_valid = wz ? 64 : 0;
          q.slab[j].a_map = fin_map; q.slab[j].a_col = 64 * c;
        }
        for (int c = 0; c < pe; ++c, ++j) {
          const int left = L.Din - 64 * c;
          q.slab[j].dW = win ? win + 64 * c : nullptr; q.slab[j].ldw = L.Din;
          q.slab[j].k_valid = !win || left < 0 ? 0 : (left > 64 ? 64 : left);
          q.slab[j].a_map = fin_map; q.slab[j].a_col = L.C + 64 * c;
        }
        for (; j < q.k_tiles * q.ns; ++j) {                    // padding slabs: any readable columns, results dropped
          q.slab[j].dW = nullptr; q.slab[j].ldw = 0; q.slab[j].k_valid = 0; q.slab[j].a_map = fin_map; q.slab[j].a_col = 0;
        }
        if (b == 0) {                                          // the per-GEMM path ties a bias to its weight's GEMM
          float* b1 = win ? gr->lin_in_b : nullptr;
          float* b2 = wz ? gr->lin_z_b[0] : nullptr;
          q.dbias = b1 ? b1 : b2; q.dbias2 = b1 ? b2 : nullptr;
        }
      };
      const bool fits = 2 * L.nb + 2 + L.nz <= kWgmMaxProblems && 4 * L.nb + 4 <= kWgmMaxMaps &&
                        (L.kin_pad - L.C) / 64 + cs <= kWgmMaxSlabs;
      if (fits) {
        int gx_map[NRF_MAX_BLOCKS + 1];
        for (int b = 0; b <= NRF_MAX_BLOCKS; ++b) gx_map[b] = -1;
        square(operand(d_field, L.dout_pad, L.dout_pad), L.Dout, ax(L.nb), gr->lin_out_w, gr->lin_out_b, nullptr);
        for (int b = L.nb - 1; b >= 0; --b) {
          const int gxm = operand(gx(b + 1), L.H, L.H);
          // lin_z_b[b + 1] has fc1_b[b]'s gradient: summed once, by fc_1's problem, or by lin_z's own when fc_1 has none
          const bool z_here = b + 1 < L.nz && gr->lin_z_w[b + 1];
          square(gxm, L.H, an(b), gr->fc1_w[b], gr->fc1_b[b], z_here ? gr->lin_z_b[b + 1] : nullptr);
          if (z_here) gx_map[b + 1] = gxm;
          square(operand(gn(b), L.H, L.H), L.H, ax(b), gr->fc0_w[b], gr->fc0_b[b], nullptr);
        }
        // the narrow problems after the 512 x 512 ones: problems of one shape in a row can share a wave (short passes).
        // lin_z[b >= 1] multiplies with the latent columns only, which are zero for every sample without a corner inside
        // the grid: with the encode's touch flags those problems visit the flagged 64-sample blocks only (one block in
        // five in the BASELINE camera set-up; they are bound by streaming G otherwise)
        const int64_t n_t128 = (N + 127) / 128, n_b64 = (N + 63) / 64;
        const int64_t blk_off = round_up(65536 + (n_t128 + 1) * 4, 256);
        int32_t* blocks = nullptr;
        int32_t* live = nullptr;                               // 128-sample tiles for the dL/dz GEMM below
        if (gr->touch_flags && blk_off + (n_b64 + 1) * 4 <= fixed) {
          if (L.nz > 0 && round_up(L.C, 128) == 128) live = reinterpret_cast<int32_t*>(sc + 65536);
          for (int b = 1; b < L.nz; ++b)
            if (gx_map[b] >= 0) blocks = reinterpret_cast<int32_t*>(sc + blk_off);
          if (live || blocks) TRY(live_tiles_launch(gr->touch_flags, N, live, blocks, s));
        }
        for (int b = L.nz - 1; b >= 1; --b)
          if (gx_map[b] >= 0) {
            field_in_problem(gx_map[b], b);
            if (!gr->fc1_w[b - 1]) h.prob[h.n_prob - 1].dbias = gr->lin_z_b[b];
            if (!h.prob[h.n_prob - 1].dbias) h.prob[h.n_prob - 1].blocks = blocks;
          }
        field_in_problem(operand(gx(0), L.H, L.H), 0);
        // With NrfMlpGrads.dlatent_ready_event the dL/dz GEMM goes FIRST and the event is recorded behind it: the
        // caller's volume scatter (on a stream of its own) then runs under the weight gradients.  Same kernels, same
        // operands: nothing changes numerically.
        const bool dz_first = gr->dlatent_ready_event != nullptr;
        if (!dz_first) TRY(wgrad_multi_launch(h, s));
        NrfGemm g = gemm_init(N, (int)round_up(L.C, 128), L.C);
        for (int b = 0; b < L.nz; ++b) set_a(g, b, gx(b), L.H, L.H);
        g.B = W + L.WzcatT; g.ldb = L.nz * L.H;
        g.out_f32 = dlatent; g.ldo = L.C;
        // dL/dz only for the 128-sample tiles in which some sample has a corner inside the grid (the others' rows are
        // never read: the scatter visits in-grid corners only); the list lives in the idle split-reduction workspace
        if (live && g.N == 128) TRY(gemm_tc_launch(g, kFmtBf16, s, live));
        else TRY(run_gemm(g, precision, true, s));
        if (dz_first) {
          NRF_CUDA_OK(cudaEventRecord(reinterpret_cast<cudaEvent_t>(gr->dlatent_ready_event), s));
          *event_recorded = true;
          TRY(wgrad_multi_launch(h, s));
        }
        return NRF_OK;
      }
    }
    TRY(run_wgrad(d_field, L.dout_pad, ax(L.nb), L.H, N, L.dout_pad, L.H, L.Dout, L.H, gr->lin_out_w, L.H,
                  gr->lin_out_b, wws, precision, s));
    for (int b = L.nb - 1; b >= 0; --b) {
      TRY(run_wgrad(gx(b + 1), L.H, an(b), L.H, N, L.H, L.H, L.H, L.H, gr->fc1_w[b], L.H, gr->fc1_b[b], wws,
                    precision, s));
      if (b + 1 < L.nz)
        TRY(run_wgrad(gx(b + 1), L.H, fin, L.kin_pad, N, L.H, (int)round_up(L.C, 64), L.H, L.C, gr->lin_z_w[b + 1],
                      L.C, gr->lin_z_b[b + 1], wws, precision, s));
      TRY(run_wgrad(gn(b), L.H, ax(b), L.H, N, L.H, L.H, L.H, L.H, gr->fc0_w[b], L.H, gr->fc0_b[b], wws, precision, s));
    }
    NrfGemm g = gemm_init(N, (int)round_up(L.C, 128), L.C);
    for (int b = 0; b < L.nz; ++b) set_a(g, b, gx(b), L.H, L.H);
    g.B = W + L.WzcatT; g.ldb = L.nz * L.H;
    g.out_f32 = dlatent; g.ldo = L.C;
    TRY(run_gemm(g, precision, true, s));
    TRY(run_wgrad(gx(0), L.H, fin, L.kin_pad, N, L.H, (int)round_up(L.C, 64), L.H, L.C, gr->lin_z_w[0], L.C,
                  gr->lin_z_b[0], wws, precision, s));
    TRY(run_wgrad(gx(0), L.H, fin + (int64_t)L.C * L.es, L.kin_pad, N, L.H, L.kin_pad - L.C, L.H, L.Din, gr->lin_in_w,
                  L.Din, gr->lin_in_b, wws, precision, s));
    return NRF_OK;
  }

  char* gbuf = sc + fixed;                    // dL/dx, updated in place while b >= n_lin_z
  char* dnet = gbuf + layer;
  auto gz = [&](int b) { return dnet + (int64_t)(1 + b) * layer; };   // dL/dx'_b kept for dL/dz, b < n_lin_z
  // lin_out: parameter gradients, then the gradient of x_nb (ReLU-gated by relu(x_nb) > 0)
  {
    const char* a16 = act16(ax(L.nb));
    if (rc) return rc;
    TRY(run_wgrad(d_field, L.dout_pad, a16, L.H, N, L.dout_pad, L.H, L.Dout, L.H, gr->lin_out_w, L.H,
                  gr->lin_out_b, wws, precision, s));
  }
  NrfGemm g = gemm_init(N, L.H, L.H);
  set_a(g, 0, d_field, L.dout_pad, L.dout_pad);
  g.B = W + L.WoutT; g.ldb = L.dout_pad;
  g.mask_src = ax(L.nb); g.ldmask = L.H;
  if (gr->d_last) { g.resid = gr->d_last; g.ldr = L.H; }      // + the gradient that reaches x_nb directly (ret_last_feat)
  g.out_act = gbuf; g.ldact = L.H;
  TRY(run_gemm(g, precision, true, s));

  const char* gcur = gbuf;                    // dL/dx_{b+1}
  for (int b = L.nb - 1; b >= 0; --b) {
    bool cat = b + 1 < L.nz;
    {
      const char* a16 = act16(an(b));
      if (rc) return rc;
      TRY(run_wgrad(gcur, L.H, a16, L.H, N, L.H, L.H, L.H, L.H, gr->fc1_w[b], L.H, gr->fc1_b[b], wws, precision, s));
    }
    if (cat)
      TRY(run_wgrad(gcur, L.H, fin, L.kin_pad, N, L.H, (int)round_up(L.C, 64), L.H, L.C, gr->lin_z_w[b + 1], L.C,
                    gr->lin_z_b[b + 1], wws, precision, s));
    // dnet_b = (g . W_fc1[b]) gated by relu(net_b) > 0
    g = gemm_init(N, L.H, L.H);
    set_a(g, 0, gcur, L.H, L.H);
    g.B = W + L.Wfc1T[b]; g.ldb = L.H;
    g.mask_src = an(b); g.ldmask = L.H;
    g.out_act = dnet; g.ldact = L.H;
    TRY(run_gemm(g, precision, true, s));
    {
      const char* a16 = act16(ax(b));
      if (rc) return rc;
      TRY(run_wgrad(dnet, L.H, a16, L.H, N, L.H, L.H, L.H, L.H, gr->fc0_w[b], L.H, gr->fc0_b[b], wws, precision, s));
    }
    // dL/dx'_b = dL/dx_{b+1} + (dnet_b . W_fc0[b]) gated by relu(x'_b) > 0
    char* target = b < L.nz ? gz(b) : gbuf;
    g = gemm_init(N, L.H, L.H);
    set_a(g, 0, dnet, L.H, L.H);
    g.B = W + L.Wfc0T[b]; g.ldb = L.H;
    g.mask_src = ax(b); g.ldmask = L.H;
    g.resid = gcur; g.ldr = L.H;
    g.out_act = target; g.ldact = L.H;
    TRY(run_gemm(g, precision, true, s));
    gcur = target;
  }
  // dL/dz = [g'_0 | g'_1 | ...] . [W_z0 ; W_z1 ; ...]
  if (L.nz > 0) {
    int cpad = (int)round_up(L.C, 128);
    if (is_tc(precision) || L.nz <= 3) {
      g = gemm_init(N, is_tc(precision) ? cpad : L.C, L.C);
      for (int b = 0; b < L.nz; ++b) set_a(g, b, gz(b), L.H, L.H);
      g.B = W + L.WzcatT; g.ldb = L.nz * L.H;
      g.out_f32 = dlatent; g.ldo = L.C;
      TRY(run_gemm(g, precision, true, s));
    } else {
      set_error("nrf_mlp_bwd: n_lin_z=%d > 3 is not supported", L.nz);
      return NRF_ENOSUP;
    }
  }
  // first layer: x'_0 = [z | p] . [W_z0 | W_in]^T ; gcur = dL/dx'_0
  if (L.nz > 0)
    TRY(run_wgrad(gcur, L.H, fin, L.kin_pad, N, L.H, (int)round_up(L.C, 64), L.H, L.C, gr->lin_z_w[0], L.C,
                  gr->lin_z_b[0], wws, precision, s));
  TRY(run_wgrad(gcur, L.H, fin + (int64_t)L.C * L.es, L.kin_pad, N, L.H, L.kin_pad - L.C, L.H, L.Din, gr->lin_in_w,
                L.Din, gr->lin_in_b, wws, precision, s));
#undef TRY
  return NRF_OK;
}

// NrfMlpGrads.dlatent_ready_event: the one-launch weight-gradient path records it right behind the dL/dz GEMM, before
// the weight gradients; every other path records it here, when the whole backward is enqueued - "dlatent is complete
// once the event has fired" holds either way.  (Never both: a stream that waits on an event waits for its LATEST record.)
static int mlp_bwd_entry(const NrfMlpParams* p, const void* packed, int precision, const void* field_in, int64_t N,
                         const void* acts, const void* d_field, const NrfMlpGrads* gr, float* dlatent, void* scratch,
                         void* stream, bool force_layered) {
  bool recorded = false;
  int rc = mlp_bwd_impl(p, packed, precision, field_in, N, acts, d_field, gr, dlatent, scratch, stream, force_layered,
                        &recorded);
  if (rc == NRF_OK && gr && gr->dlatent_ready_event && !recorded)
    NRF_CUDA_OK(cudaEventRecord(reinterpret_cast<cudaEvent_t>(gr->dlatent_ready_event), as_stream(stream)));
  return rc;
}

extern "C" int nrf_mlp_bwd(const NrfMlpParams* p, const void* packed, int precision, const void* field_in,
                           int64_t N, const void* acts, const void* d_field, const NrfMlpGrads* gr,
                           float* dlatent, void* scratch, void* stream) {
  return mlp_bwd_entry(p, packed, precision, field_in, N, acts, d_field, gr, dlatent, scratch, stream, false);
}

extern "C" int nrf_mlp_bwd_layered(const NrfMlpParams* p, const void* packed, int precision, const void* field_in,
                                   int64_t N, const void* acts, const void* d_field, const NrfMlpGrads* gr,
                                   float* dlatent, void* scratch, void* stream) {
  return mlp_bwd_entry(p, packed, precision, field_in, N, acts, d_field, gr, dlatent, scratch, stream, true);
}
