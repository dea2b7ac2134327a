// Epilogue shared by the tcgen05 (bf16) and SIMT (fp32) GEMMs of the field MLP.
//   v = acc + bias[n];  v = mask_src[m,n] > 0 ? v : 0;  v += resid[m,n];
//   out_act[m,n] = T(relu ? max(v,0) : v);  out_act2 likewise;  out_f32[m,n] = v
// Forward layers use bias/resid/relu (resnetfc.py:57-64,172,183-191); the dgrad chain uses
// mask_src (ReLU gate of the saved activation) and resid (gradient of the skip connection).
#pragma once
#include "common.cuh"

namespace nrf {

template <typename T>
struct Epilogue {
  const float* bias;
  const T* mask_src; int ldmask;
  const T* resid; int ldr;
  T* out_act; int ldact; int relu_act;
  T* out_act2; int ldact2; int relu_act2;
  float* out_f32; int ldo;
  int M, n_store;
};

template <typename T>
static inline Epilogue<T> make_epilogue(const NrfGemm& g) {
  Epilogue<T> e;
  e.bias = g.bias;
  e.mask_src = reinterpret_cast<const T*>(g.mask_src); e.ldmask = g.ldmask;
  e.resid = reinterpret_cast<const T*>(g.resid); e.ldr = g.ldr;
  e.out_act = reinterpret_cast<T*>(g.out_act); e.ldact = g.ldact; e.relu_act = g.relu_act;
  e.out_act2 = reinterpret_cast<T*>(g.out_act2); e.ldact2 = g.ldact2; e.relu_act2 = g.relu_act2;
  e.out_f32 = g.out_f32; e.ldo = g.ldo;
  e.M = g.M; e.n_store = g.n_store;
  return e;
}

__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }

// ---- whole-MLP fused kernel (mlp_fused.cu): layer programs built by mlp.cu, forward and backward
constexpr int kFusedMaxLayers = 2 * NRF_MAX_BLOCKS + 2;
struct FusedLayerDesc {
  const void* W;       // (512, ldw) bf16 row-major, K = 64 * (kb_main + kb_z) columns used
  int ldw;
  int kb_main, kb_z;   // 64-wide k-blocks of the main operand / of the trailing part read from `in`
  int kind;            // 0: residual-stream layer (fwd: lin_in, fc_1; bwd: lin_out^T, fc_0^T)  1: fc_0 / fc_1^T
                       // 2: lin_out (forward only, last)
  int a_src;           // 0: rows of `in`   1: previous output in TMEM   2: previous output in shared memory
  int first;           // residual-stream layer that starts the stream (nothing added)
  int act_slot;        // slot of `saves` this layer's output operand is written to, -1: none
  int mask_slot;       // backward: slot of the forward's saved operand whose ReLU gate applies to this layer's output
  const float* bias;   // forward
  int n_chunks;        // 128-wide output chunks of this layer; 0 = 4 (the 512 hidden units).  lin_out: nout_pad / 128
  int ext_col;         // first column of `in` the kb_z trailing k-panels are read from
  int skip_head;       // forward: the first skip_head main k-panels multiply latent columns of `in` ...
  int skip_z;          // ... and so do the kb_z trailing panels: both are all-zero for a tile without a sample in the grid
};
struct FusedDesc {
  int n_layers;
  FusedLayerDesc L[kFusedMaxLayers];
  int backward;
  const void* in; int in_cols;       // forward: field input (N, kin_pad); backward: d_field (N, dout_pad); bf16
  int64_t N;
  void* saves; int n_slots;          // n_slots x (N, 512) bf16 written by the epilogues (forward: NULL = inference)
  void* gate_bits;                   // slots x (N, 64 B) bit-packed ReLU gates: written by the saving forward, read
                                     // by the backward (same slot numbering as `saves` of the forward)
  float* out; int d_out; int ldo;    // forward: raw field outputs
  void* prof;                        // optional: 32 int64 cycle counters per CTA (diagnostics), NULL otherwise
  const uint8_t* touch;              // forward, optional: one flag per 32 rows (nrf_encode_points_touch); a 256-row tile
                                     // whose flags are all 0 has an all-zero latent: its latent k-panels are skipped
  int half;                          // forward only: fp16 operands (weights, field input, activations, residual stream;
                                     // NRF_PREC_FP16); what it SAVES for the backward is bf16 either way
};
int mlp_fused_launch(const FusedDesc& d, cudaStream_t stream);

// Operand formats of a tcgen05 GEMM (0 bf16, 1 fp16: the same kind::f16 rate).  A and B must agree: the instruction
// descriptor has a format field per operand, but fp16 x bf16 faults with "illegal instruction" on B200 (measured,
// scripts/fp16_probe.py) - which is why the NRF_PREC_FP16 mode keeps a bf16 copy of everything its backward multiplies.
//   gemm_tc : a = [A0|A1|A2], b = B, io = resid / out_act / out_act2 (mask_src is read by its bits: any 16-bit float)
//   wgrad_tc: a = G, b = A
struct OpFmt { int a_half, b_half, io_half; };
static inline OpFmt op_fmt(int precision) {
  const int h = precision == NRF_PREC_FP16;
  OpFmt f = {h, h, h};
  return f;
}
constexpr OpFmt kFmtBf16 = {0, 0, 0};
// live (optional, N = 128 GEMMs without epilogue inputs): live[0] row tiles of 128 rows to compute, indices live[1..]
int gemm_tc_launch(const NrfGemm& g, OpFmt fmt, cudaStream_t stream, const int32_t* live = nullptr);
// builds such a list from one flag byte per 32 rows (nrf_encode_points_touch)
// (one launch: 128-row tiles into live128 and / or 64-row blocks into live64)
int live_tiles_launch(const void* flags, int64_t n_rows, int32_t* live128, int32_t* live64, cudaStream_t stream);
int gemm_simt_launch(const NrfGemm& g, cudaStream_t stream);
int wgrad_tc_launch(const void* G, int ldg, const void* A, int lda, int M, int N, int K, int n_valid,
                    int k_valid, float* dW, int ldw, float* dbias, void* workspace, OpFmt fmt, cudaStream_t stream);
// NRF_PREC_BF16X3 (mlp_x3.cu): split-bf16 operands, three MMAs per product, fp32-grade accuracy on the tensor cores
int mlp_x3_sizes(const NrfMlpParams* p, NrfMlpSizes* out);
int mlp_x3_pack(const NrfMlpParams* p, void* packed, cudaStream_t s);
int mlp_x3_fwd(const NrfMlpParams* p, const void* packed, const float* field_in, int64_t N, void* acts,
               float* field_out, cudaStream_t s);
int mlp_x3_bwd(const NrfMlpParams* p, const void* packed, int64_t N, const void* acts, const float* d_field,
               const NrfMlpGrads* gr, float* dlatent, void* scratch, cudaStream_t s);
// All weight gradients of one backward pass in one persistent launch (wgrad_multi.cu).  Operands are (M, cols) bf16
// matrices with row pitch ld; a problem multiplies the columns of ONE G operand with k_tiles x ns slabs of 64 columns,
// each slab taken from any operand at any 64-aligned column and added into its own destination.
constexpr int kWgmMaxMaps = 40, kWgmMaxProblems = 28, kWgmMaxSlabs = 8;
struct WgmSlab {
  float* dW; int ldw;      // destination of this slab's 64 columns: dW[n * ldw + j], j < k_valid
  int k_valid;             // 0: padding slab, results dropped
  int a_map, a_col;        // source operand and first column
};
struct WgmProblem {
  int g_map;               // operand G
  int n_valid;             // columns of G that count = rows of the dW's
  int ns;                  // slabs per k tile: 2 or 4 (MMA N = 64 ns)
  int k_tiles;
  float* dbias;            // += column sums of G (NULL: none)
  float* dbias2;           // a second bias with the same gradient (lin_in / lin_z[0]), or NULL
  const int32_t* blocks;   // optional (dbias == NULL only): blocks[0] 64-sample blocks to visit, indices blocks[1..] ascending
                           // (live_tiles_launch): the others contribute nothing (their A rows are all zero)
  WgmSlab slab[kWgmMaxSlabs];
};
struct WgmOperand { const void* base; int cols, ld; };
struct WgmHost {
  int n_maps; WgmOperand op[kWgmMaxMaps];
  int n_prob; WgmProblem prob[kWgmMaxProblems];
  int M;
  int* sync;               // >= kWgmMaxProblems * 128 ints of device scratch (zeroed by the launcher), or NULL
};
int wgrad_multi_launch(const WgmHost& h, cudaStream_t stream);
int wgrad_simt_launch(const void* G, int ldg, const void* A, int lda, int M, int N, int K, int n_valid,
                      int k_valid, float* dW, int ldw, float* dbias, cudaStream_t stream);

}  // namespace nrf
