// The whole ResnetFC field MLP (resnetfc.py:55-64,146-195) as ONE persistent tcgen05 kernel per sample tile.
//
// A CTA pair (cluster of 2, one TPC) owns 256 samples (128 per CTA) and walks them through every layer without
// touching HBM in between: only the field-input rows come in, the raw field outputs go out and - in training -
// the bf16 operands the backward needs (relu(x'_b), relu(net_b)) are streamed out by TMA stores as a side effect.
//
// Where everything lives (per CTA; the MMAs are tcgen05.mma.cta_group::2, M = 256, N = 128, K = 16):
//   weights        stream from L2 through a TMA ring of 16 KB stages (this CTA's 64 of the 128 weight rows x 128 k);
//                  the pair shares every weight byte, so the L2 -> SM weight traffic is 32 B/clk/SM at full rate
//   field input    the first layer's A operand ([latent | PE | viewdir]; backward: d_field) is loaded once per tile
//                  into P, which is dead at that point; the latent k-panels of the lin_z[b+1] tails of fc_1 come
//                  through the weight ring, one 16 KB stage per 128-row k-panel
//   accumulators   TMEM columns [0,256): two 128-column buffers; the epilogue of chunk i overlaps the MMAs of i+1
//   relu(x')       TMEM columns [256,512) as packed bf16 pairs: the A operand of fc_0 and lin_out comes straight
//                  from tensor memory (tcgen05.mma with A in TMEM), written by the epilogue with tcgen05.st
//   relu(net)      shared memory "P": 8 k-panels of 128 rows x 64 bf16 (128B-swizzled UMMA layout), the A operand
//                  of fc_1, written by the epilogue with st.shared; in training the same panels are the source of
//                  the TMA stores of relu(net_b)
//   residual x'    REGISTERS of the epilogue threads: thread (row, column half g) keeps its 256 elements of the bf16
//                  residual stream as 128 packed registers for the whole tile (setmaxnreg gives the epilogue
//                  warpgroups 232 registers, the service warps 40)
// Ping-ponging the operand between TMEM and shared memory removes every write-after-read hazard between a layer's
// MMAs and its own epilogue, so the next layer's MMAs start on k-block kb as soon as the epilogue has published it
// (a_ready[kb]) and the tensor pipe never drains between layers.
//
// The same kernel runs the BACKWARD data-gradient chain (kBwd): lin_out^T, then per block fc_1^T and fc_0^T with the
// ReLU gates read from the operands the forward saved, the gradient of the residual stream in the epilogue
// registers, and every intermediate gradient (dL/dx'_b, dL/dnet_b: the G operands of the weight-gradient GEMMs)
// streamed out by TMA stores.  Forward and backward are two "layer programs" for one executor.
//
// Operand formats (kind::f16 takes bf16 or fp16 at the same tensor rate): NRF_PREC_BF16 runs everything in bf16;
// NRF_PREC_FP16 (kHalf) runs the FORWARD with fp16 weights / activations / residual stream (11 significant bits: 6-8 x
// smaller output error).  Its backward is the bf16 backward unchanged - bf16 gradients need no
// loss scaling, and gradient error is set by the forward's ReLU-gate flips, not by gradient rounding - so the kHalf
// forward SAVES its operands as bf16 (converted on their way to the TMA-store staging slot): A and B of one MMA must
// share a format (fp16 x bf16 is an illegal instruction on B200), and the weight-gradient GEMMs multiply the saved
// operands with bf16 gradients.
//
// Warp roles: 0 TMA producer | 1 MMA issuer (leader CTA only) | 2 TMEM allocator | 3 first-layer input loader |
// 4-11 epilogue: warp (g, q4) owns
// rows [32 q4, 32 q4 + 32) x columns [64g, 64g+64) of every 128-column chunk (= its rows of k-block 2c+g of the next
// layer); the eight epilogue warps never synchronise with each other.
#include "tc_ptx.cuh"

namespace nrf {

namespace {

constexpr int kFThreads = 384;
constexpr int kFPanel = 128 * 64 * 2;        // 16 KB: 128 rows x 64 bf16, 128B swizzle
constexpr int kFBoxB = 64 * 64 * 2;          // 8 KB: this CTA's 64 weight rows x 64 k
constexpr int kFStageB = 2 * kFBoxB;         // a ring stage carries two k-blocks (8 MMAs per barrier round trip)
constexpr int kFChunks = 4;                  // 512 / 128 accumulator chunks per layer
constexpr int kFSmemP = 8 * kFPanel;         // 128 KB
constexpr int kFSlot = 32 * 128;             // 4 KB: 32 rows x 64 bf16 (128B swizzle), one per epilogue warp
constexpr int kFSmemBar = 320;
constexpr int kFMaxIn = 7;                   // k-panels of the first layer's global A operand (they live in P)
constexpr int kFSmemBias = 8 * 256;          // per epilogue warp: the 64 bias values of its current chunk
constexpr int kFAlignSlack = 704;            // dynamic shared memory starts 1024-aligned in practice (checked)
#ifndef NRF_FUSED_INFER_STAGES
#define NRF_FUSED_INFER_STAGES 6
#endif
#ifdef NRF_FUSED_DIRECT_SAVE          // A/B experiment: the saved copies of relu(x') leave with st.global from registers
constexpr bool kDirectSave = true;
#else
constexpr bool kDirectSave = false;
#endif
template <bool kSave>
struct FCfg {
  static constexpr int kStages = kSave ? (kDirectSave ? 6 : 4) : NRF_FUSED_INFER_STAGES;   // saving gives 32 KB up for the staging slots
  static constexpr int kRing = kStages * kFStageB;
  static constexpr int kStaging = kSave && !kDirectSave ? 8 * kFSlot : 0;
  static constexpr int kSmem = kFSmemP + kRing + kStaging + kFSmemBar + kFSmemBias + kFAlignSlack;
  static_assert(kSmem <= 232448, "fused MLP kernel exceeds the 227 KB shared-memory limit");
};
constexpr uint32_t kFColQ = 256;             // first TMEM column of the packed relu(x') operand

enum { kLayerX = 0, kLayerNet = 1, kLayerOut = 2 };
enum { kSrcIn = 0, kSrcQ = 1, kSrcP = 2 };

struct FLayer {
  int kind, a_src, kb_main, kb_z, first, act_slot;
  int n_chunks;            // 128-wide accumulator chunks of this layer (4 for the hidden layers, up to 5 for lin_out)
  int ext_col;             // first column of the global input the kb_z trailing k-panels come from
  int mask_slot;           // backward: slot of the forward's saved operand whose sign gates this layer's output
  int publish;             // the epilogue hands its output to the next layer's MMAs (a_ready)
  int skip_head, skip_z;   // k-panels that multiply the latent (see FusedLayerDesc): no MMAs, no weight loads on a dead tile
  const float* bias;       // forward only
};
struct FArgs {
  int n_layers, n_tiles, n_prod;   // n_prod: layers per tile whose epilogue publishes an A operand
  int l_p_free;                    // last layer of a tile whose MMAs read P: after it the next tile's input may land
  int N, d_out, ldo;
  uint32_t idesc;   // tcgen05 instruction descriptor (M 256, N 128, K-major; operand formats per precision mode)
  int dbg;      // NRF_DBG timing experiments (wrong results!; only in the profiling instantiations, i.e. with a
                // profile buffer set): 1 no weight TMA, 2 epilogue protocol only, 4 no MMAs, 8 MMA issuer ignores
                // acc_empty / a_ready
  float* out;
  uint32_t* saves;                 // kDirectSave: base of the (slots, N, 512) saved-operand tensor
  uint2* gate_bits;                // slots x (N, 8) x 64 bits: bit-packed ReLU gates, written by the training forward
                                   // (one bit per saved operand element: non-zero) and read by the backward
  long long* prof;                 // NRF_FUSED_PROF=1: per-CTA cycle counters of the warp roles (see mlp_fused_launch)
  const uint8_t* touch;            // optional: one flag per 32 samples "some corner inside the grid"; NULL: every tile live
  int n_flags;                     // (N + 31) / 32
  FLayer L[kFusedMaxLayers];
};
struct FMaps {
  CUtensorMap w[kFusedMaxLayers];
  CUtensorMap in;        // forward: field input (N, kin_pad); backward: d_field (N, dout_pad)
  CUtensorMap acts;      // (512, N, slots) bf16, box 64 x 32 x 1, 128B swizzle: a warp's 32 rows of one k-panel, from
                         // the P panels or from the warp's staging slot
};

__device__ __forceinline__ void umma_bf16_pair_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                                  uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// packed fp32 pairs (sm_100 add.f32x2: two IEEE rn additions per instruction) and packed bf16 ReLU
__device__ __forceinline__ uint64_t pair_u32(uint32_t lo, uint32_t hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
  return r;
}
__device__ __forceinline__ uint64_t pair_f32(float lo, float hi) {
  return pair_u32(__float_as_uint(lo), __float_as_uint(hi));
}
__device__ __forceinline__ uint64_t add2(uint64_t x, uint64_t y) {
  uint64_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(x), "l"(y));
  return r;
}
template <bool kHalf>
__device__ __forceinline__ uint32_t cvt_op16x2(uint64_t v) {     // {lo, hi} fp32 -> packed bf16 / fp16 (lo in bits 0-15)
  uint32_t lo, hi, r;
  asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(v));
  if (kHalf) asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(__uint_as_float(hi)), "f"(__uint_as_float(lo)));
  else asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(__uint_as_float(hi)), "f"(__uint_as_float(lo)));
  return r;
}
template <bool kHalf>
__device__ __forceinline__ uint32_t relu_op16x2(uint32_t w) {    // op16(relu(x)) == relu(op16(x))
  uint32_t r;
  if (kHalf) asm("max.f16x2 %0, %1, %2;" : "=r"(r) : "r"(w), "r"(0u));
  else asm("max.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(w), "r"(0u));
  return r;
}
__device__ __forceinline__ uint32_t f16x2_to_bf16x2(uint32_t w) {   // the saved copy of an fp16 operand pair
  uint32_t r;
  asm("{\n"
      ".reg .f16 l, h;\n"
      ".reg .f32 a, b;\n"
      "mov.b32 {l, h}, %1;\n"
      "cvt.f32.f16 a, l;\n"
      "cvt.f32.f16 b, h;\n"
      "cvt.rn.bf16x2.f32 %0, b, a;\n"
      "}\n"
      : "=r"(r)
      : "r"(w));
  return r;
}
template <bool kHalf>
__device__ __forceinline__ uint64_t unpack_op16x2(uint32_t r) {  // packed bf16 / fp16 pair -> {lo, hi} fp32 (exact)
  if (!kHalf) return pair_u32(r << 16, r & 0xffff0000u);
  uint64_t o;
  asm("{\n"
      ".reg .f16 l, h;\n"
      ".reg .f32 a, b;\n"
      "mov.b32 {l, h}, %1;\n"
      "cvt.f32.f16 a, l;\n"
      "cvt.f32.f16 b, h;\n"
      "mov.b64 %0, {a, b};\n"
      "}\n"
      : "=l"(o)
      : "r"(r));
  return o;
}

// Bit-packed ReLU gates.  A gate word covers one 32-column sub-chunk of one row: bit j / bit 16+j = element 2j /
// 2j+1 of the saved (ReLU'd, packed bf16) operand is non-zero.
//   forward : G |= pos_mask(x_j) & (0x00010001 << j), the same mask that applies the ReLU (kReluMask); the first form,
//             16 x gate_push(G, w_j)   (w + 0x7fff7fff carries "non-zero" into bits 15 / 31 of a non-negative pair),
//             is kept for the A/B builds
//   backward: gate_mask(G, j) = per-half 0xffff / 0 for pair j  (shift the two flags to the byte MSBs, PRMT
//             replicates them over the halves)
#ifdef NRF_GATE_PUSH
__device__ __forceinline__ uint32_t gate_push(uint32_t G, uint32_t w) {
  const uint32_t t = w + 0x7fff7fffu;
  return ((G >> 1) & 0x7fff7fffu) | (t & 0x80008000u);
}
#endif
// the same gate word built in place: HSET2 gives 0xffff per non-zero half of the (non-negative) pair, one LOP3 drops
// its bits j and 16 + j into G - two instructions per pair where the shift-merge above needs three
// 0xffff per half that is > 0: ReLU is `x & mask` (what max(x, 0) gives, -0 and NaN included: both -> +0) and the gate
// bits come from the same mask
template <bool kHalf>
__device__ __forceinline__ uint32_t pos_mask(uint32_t x) {
  uint32_t m;
  if (kHalf) asm("set.gt.u32.f16x2 %0, %1, %2;" : "=r"(m) : "r"(x), "r"(0u));
  else asm("set.gt.u32.bf16x2 %0, %1, %2;" : "=r"(m) : "r"(x), "r"(0u));
  return m;
}
#ifdef NRF_RELU_MAX
constexpr bool kReluMask = false;      // A/B build: max.bf16x2 for the ReLU, the gate bits from the result
#else
constexpr bool kReluMask = true;
#endif
template <bool kHalf>
__device__ __forceinline__ uint32_t gate_set(uint32_t G, uint32_t w, int j) {
#ifdef NRF_GATE_PUSH                    // A/B build: the shift-merge form (the same bits)
  (void)j;
  return gate_push(G, w);
#endif
  uint32_t m;
  if (kHalf) asm("set.gt.u32.f16x2 %0, %1, %2;" : "=r"(m) : "r"(w), "r"(0u));
  else asm("set.gt.u32.bf16x2 %0, %1, %2;" : "=r"(m) : "r"(w), "r"(0u));
  return G | (m & (0x00010001u << j));
}
__device__ __forceinline__ uint32_t gate_mask(uint32_t G, int j) {
  uint32_t m;                                  // selector nibble 8+k: the MSB of byte k replicated over the byte
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(m) : "r"(G << (15 - j)), "r"(0u), "r"(0xBB99u));
  return m;
}

// explicit shared-space accesses with 32-bit addresses (the aligned dynamic-smem base is a generic pointer to the
// compiler, which would otherwise emit generic LD/ST with 64-bit address arithmetic in the epilogue's hot loop)
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ float4 lds128f(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts32f(uint32_t addr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ void mbar_wait_u32(uint32_t addr, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spin = 0; !done; ++spin) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
#ifdef NRF_DEBUG_SPIN
    if (spin > (1u << 18)) {
      printf("mbar_wait_u32 timeout: block %d warp %d barrier smem 0x%x parity %u\n", (int)blockIdx.x,
             (int)(threadIdx.x >> 5), addr, parity);
      __trap();
    }
#else
    if (spin > (1u << 26)) __trap();
#endif
  }
}
__device__ __forceinline__ void mbar_arrive_leader_u32(uint32_t addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(addr & kPeerBitMask) : "memory");
}
// L2 cache-policy descriptors (the values CUTLASS passes as TMA cache hints): the saved operands are written once and
// read much later (evict first), the weights are re-read by every CTA for every tile (evict last).
constexpr uint64_t kL2EvictFirst = 0x12F0000000000000ull;
constexpr uint64_t kL2EvictLast = 0x14F0000000000000ull;
__device__ __forceinline__ void tma_store_3d_u32(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3, %4}], [%1], %5;" ::"l"(
          reinterpret_cast<uint64_t>(map)),
      "r"(src), "r"(c0), "r"(c1), "r"(c2), "l"(kL2EvictFirst)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair_hint(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                                      uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
      "[%0], [%1, {%3, %4}], [%2], %5;" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1), "l"(policy)
      : "memory");
}

// A 256-sample tile none of whose eight 32-sample groups touches the grid: its latent columns are exact zeros
// (zeros padding of the gather), so the k-panels that multiply them add nothing to any accumulator.
__device__ __forceinline__ bool tile_dead(const uint8_t* touch, int n_flags, int tile) {
  if (touch == nullptr) return false;
  uint32_t any = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int i = tile * 8 + j;
    if (i < n_flags) any |= __ldg(touch + i);
  }
  return any == 0;
}

// cycle accounting of the waits: compiled in only for the kProf instantiations (even a dormant run-time branch
// per wait cost 5-7 % of the kernel)
#define FUSED_TIMED(acc, stmt)                         \
  do {                                                 \
    if (kProf) {                                       \
      const long long t0__ = clock64();                \
      stmt;                                            \
      acc += clock64() - t0__;                         \
    } else {                                           \
      stmt;                                            \
    }                                                  \
  } while (0)

struct EpiCtx {            // per-thread constants of the epilogue
  uint32_t tmem_base, lane_off;
  uint32_t sP, slot, wbias;            // shared-space addresses: P, this warp's staging slot, this warp's bias values
  uint32_t acc_full, acc_empty, a_ready;
  int lane, g, q4, row;
  uint32_t sw128;
  long long t_wait;                    // cycles spent waiting for accumulators (profiling)
};

// One 128-column accumulator chunk of one layer, this thread's 64 columns.  KIND is the layer kind; xr is the thread's
// slice of the residual stream for this chunk (kLayerX only).
//   forward : v = acc + bias (+ x' for kLayerX);  operand for the next layer = bf16(relu(v))
//   backward: v = gate(acc) (+ g for kLayerX), gate = sign of the forward's saved operand; next operand = bf16(v)
// Prefetched here and parked until needed: the next chunk's bias values (`nxt`, forward: into the warp's smem slot) or
// the gate row of the chunk after next (`nxt2`, backward: into mk, which the caller double-buffers by chunk parity).
template <int KIND, bool kSave, bool kBwd, bool kProf, bool kHalf>
__device__ __forceinline__ void epi_chunk(EpiCtx& e, const FMaps& maps, const FArgs& a, const FLayer& L, int c,
                                          uint32_t n, int row0, bool first, uint32_t (&xr)[32], uint2& gate,
                                          const void* nxt, const uint2* gate2) {
  const uint32_t buf = n & 1;
  const int col0 = c * 128 + e.g * 64;         // first feature this thread handles in this chunk
  const bool save = kSave && L.act_slot >= 0;
  float bn0 = 0.f, bn1 = 0.f;
  if (!kBwd) {                                 // next chunk's bias: 2 values per lane, parked in registers
    bn0 = __ldg(reinterpret_cast<const float*>(nxt) + e.lane);
    bn1 = __ldg(reinterpret_cast<const float*>(nxt) + e.lane + 32);
  }
  FUSED_TIMED(e.t_wait, mbar_wait_u32(e.acc_full + buf * 8, (n >> 1) & 1));
  tc_fence_after();
  const uint32_t taddr = e.tmem_base + buf * 128 + e.g * 64 + e.lane_off;
  if (kProf && (a.dbg & 2)) {
    tc_fence_before();
    __syncwarp();
    if (e.lane == 0) {
      mbar_arrive_leader_u32(e.acc_empty + buf * 8);
      if (L.publish) mbar_arrive_leader_u32(e.a_ready + (2 * c + e.g) * 8);
    }
    return;
  }
  uint32_t v[64];
  tmem_ld32_nowait(taddr, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
  tmem_ld32_nowait(taddr + 32, *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
  tmem_ld_wait();
  tc_fence_before();                           // the accumulator buffer is free again
  __syncwarp();
  if (e.lane == 0) mbar_arrive_leader_u32(e.acc_empty + buf * 8);
  const uint32_t prow = e.sP + (2 * c + e.g) * kFPanel + e.row * 128;
  uint32_t w[32];                              // this thread's 64 outputs of the chunk as packed bf16 operand values
  uint32_t gacc[2] = {0u, 0u};                 // training forward: the two gate words of these 64 outputs
  constexpr bool kMaskRelu = kReluMask && kSave && !kBwd;
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    const uint32_t gw = s == 0 ? gate.x : gate.y;    // backward: gate word of this sub-chunk
    uint64_t x2[16];                           // fp32 pairs
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (!kBwd) {                             // acc + bias
        const float4 b = lds128f(e.wbias + (s * 32 + 4 * j) * 4);
        x2[2 * j] = add2(pair_u32(v[s * 32 + 4 * j], v[s * 32 + 4 * j + 1]), pair_f32(b.x, b.y));
        x2[2 * j + 1] = add2(pair_u32(v[s * 32 + 4 * j + 2], v[s * 32 + 4 * j + 3]), pair_f32(b.z, b.w));
      } else {
        x2[2 * j] = pair_u32(v[s * 32 + 4 * j], v[s * 32 + 4 * j + 1]);
        x2[2 * j + 1] = pair_u32(v[s * 32 + 4 * j + 2], v[s * 32 + 4 * j + 3]);
      }
    }
    if (KIND == kLayerOut) {
      // raw field outputs, fp32, written row-wise (128 B per thread per sub-chunk)
      if (row0 + e.row < a.N) {
        float* dst = a.out + (int64_t)(row0 + e.row) * a.ldo + col0 + s * 32;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (col0 + s * 32 + 4 * j < a.d_out)
            *reinterpret_cast<uint4*>(dst + 4 * j) =
                make_uint4((uint32_t)x2[2 * j], (uint32_t)(x2[2 * j] >> 32), (uint32_t)x2[2 * j + 1],
                           (uint32_t)(x2[2 * j + 1] >> 32));
      }
    } else if (KIND == kLayerX) {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        uint64_t t = x2[j];
        const uint32_t r = xr[s * 16 + j];
        if (!first) t = add2(t, unpack_op16x2<kHalf>(r));
        uint32_t xb = cvt_op16x2<kHalf>(t);
        if (kBwd) {                            // ReLU gate: closed -> the residual gradient passes unchanged
          const uint32_t m = gate_mask(gw, j);
          xb = first ? (xb & m) : ((xb & m) | (r & ~m));
        }
        xr[s * 16 + j] = xb;
        if (kMaskRelu) {
          const uint32_t m = pos_mask<kHalf>(xb);
          w[s * 16 + j] = xb & m;
          gacc[s] |= m & (0x00010001u << j);
        } else {
          w[s * 16 + j] = kBwd ? xb : relu_op16x2<kHalf>(xb);
        }
      }
      if (L.publish)
        tmem_st16(e.tmem_base + kFColQ + e.lane_off + (uint32_t)(col0 / 2 + s * 16),
                  *reinterpret_cast<uint32_t(*)[16]>(&w[s * 16]));
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (kMaskRelu) {
          const uint32_t xb = cvt_op16x2<kHalf>(x2[j]);
          const uint32_t m = pos_mask<kHalf>(xb);
          w[s * 16 + j] = xb & m;
          gacc[s] |= m & (0x00010001u << j);
        } else {
          w[s * 16 + j] = kBwd ? (cvt_op16x2<kHalf>(x2[j]) & gate_mask(gw, j))
                               : relu_op16x2<kHalf>(cvt_op16x2<kHalf>(x2[j]));
        }
#pragma unroll
      for (int j = 0; j < 4; ++j)
        sts128(prow + (((uint32_t)(s * 4 + j) ^ e.sw128) << 4), w[s * 16 + 4 * j], w[s * 16 + 4 * j + 1],
               w[s * 16 + 4 * j + 2], w[s * 16 + 4 * j + 3]);
    }
  }
  if (KIND != kLayerOut) {
    // 1. publish this warp's rows of k-block 2c+g of the next layer's A operand to the MMA issuer: this is what the
    //    tensor pipe may be waiting for, so nothing that only serves the saved copies comes before it
    if (L.publish) {
      if (KIND == kLayerX) tmem_st_wait();
      else fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (e.lane == 0) mbar_arrive_leader_u32(e.a_ready + (2 * c + e.g) * 8);
    }
    // 2. the copies kept for the other pass: operand values by TMA store, ReLU gates bit-packed
    if (save && kDirectSave && (KIND == kLayerX || kHalf)) {
      if (row0 + e.row < a.N) {
        uint4* dst = reinterpret_cast<uint4*>(a.saves + (((int64_t)L.act_slot * a.N + row0 + e.row) * 512 + col0) / 2);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          dst[j] = kHalf ? make_uint4(f16x2_to_bf16x2(w[4 * j]), f16x2_to_bf16x2(w[4 * j + 1]),
                                      f16x2_to_bf16x2(w[4 * j + 2]), f16x2_to_bf16x2(w[4 * j + 3]))
                         : make_uint4(w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
      }
      if (!kBwd && row0 + e.row < a.N) {
        uint32_t g0 = gacc[0], g1 = gacc[1];
        if (!kMaskRelu) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            g0 = gate_set<kHalf>(g0, w[j], j);
            g1 = gate_set<kHalf>(g1, w[16 + j], j);
          }
        }
        a.gate_bits[((int64_t)L.act_slot * a.N + row0 + e.row) * 8 + c * 2 + e.g] = make_uint2(g0, g1);
      }
    } else if (save) {
      if (KIND == kLayerX || kHalf) {          // (kHalf: relu(net) is saved as bf16 too, so not from its fp16 P panel)
        if (e.lane == 0) bulk_wait_read0();    // the slot's previous TMA store (a whole chunk ago) has read it
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (kHalf)
            sts128(e.slot + e.lane * 128 + (((uint32_t)j ^ e.sw128) << 4), f16x2_to_bf16x2(w[4 * j]),
                   f16x2_to_bf16x2(w[4 * j + 1]), f16x2_to_bf16x2(w[4 * j + 2]), f16x2_to_bf16x2(w[4 * j + 3]));
          else
            sts128(e.slot + e.lane * 128 + (((uint32_t)j ^ e.sw128) << 4), w[4 * j], w[4 * j + 1], w[4 * j + 2],
                   w[4 * j + 3]);
        }
        fence_proxy_async();
        __syncwarp();
      } else if (!L.publish) {
        fence_proxy_async();
        __syncwarp();
      }
      if (e.lane == 0) {
        tma_store_3d_u32(&maps.acts,
                         KIND == kLayerNet && !kHalf ? e.sP + (2 * c + e.g) * kFPanel + e.q4 * 32 * 128 : e.slot,
                         col0, row0 + e.q4 * 32, L.act_slot);
        bulk_commit();
      }
      if (!kBwd && row0 + e.row < a.N) {
        uint32_t g0 = gacc[0], g1 = gacc[1];
        if (!kMaskRelu) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            g0 = gate_set<kHalf>(g0, w[j], j);
            g1 = gate_set<kHalf>(g1, w[16 + j], j);
          }
        }
        a.gate_bits[((int64_t)L.act_slot * a.N + row0 + e.row) * 8 + c * 2 + e.g] = make_uint2(g0, g1);
      }
    }
  }
  if (!kBwd) {
    __syncwarp();
    sts32f(e.wbias + e.lane * 4, bn0);         // every lane is past its reads of this chunk's bias
    sts32f(e.wbias + (e.lane + 32) * 4, bn1);
    __syncwarp();
  } else {
    gate = __ldg(gate2);                       // gate words of the chunk after next (the caller double-buffers)
  }
}

template <bool kSave, bool kBwd, bool kProf, bool kHalf>
__global__ void __launch_bounds__(kFThreads, 1)
mlp_fused_kernel(const __grid_constant__ FMaps maps, const __grid_constant__ FArgs a) {
  using Cfg = FCfg<kSave>;
  constexpr int kStages = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sP = smem;
  uint8_t* sRing = sP + kFSmemP;
  uint8_t* sStage = sRing + Cfg::kRing;
  float* sBias = reinterpret_cast<float*>(sStage + Cfg::kStaging);
  uint64_t* full = reinterpret_cast<uint64_t*>(sStage + Cfg::kStaging + kFSmemBias);
  if (smem - smem_raw > kFAlignSlack) __trap();
  uint64_t* empty = full + kStages;
  uint64_t* acc_full = empty + kStages;
  uint64_t* acc_empty = acc_full + 2;
  uint64_t* a_ready = acc_empty + 2;           // 8: k-block kb of the next layer's A operand is in place
  uint64_t* in_full = a_ready + 8;             // kFMaxIn: k-panel kb of this tile's first-layer A operand is in P
  uint64_t* in_free = in_full + kFMaxIn;       // P may receive the next tile's input
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(in_free + 1);

  const int warp = uniform_warp_idx();
  int lane;                                      // volatile: kept in a register, never re-read with S2R in the loops
  asm volatile("mov.u32 %0, %%laneid;" : "=r"(lane));
  uint32_t crank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
  const bool cta_leader = crank == 0;
  const int pair = blockIdx.x / 2, n_pairs = gridDim.x / 2;
  const int n_iter = pair < a.n_tiles ? (a.n_tiles - pair + n_pairs - 1) / n_pairs : 0;
  const int nl = a.n_layers;

  if (warp == 0 && lane == 0) {
    for (int l = 0; l < nl; ++l) tma_prefetch_desc(&maps.w[l]);
    tma_prefetch_desc(&maps.in);
    if (kSave) tma_prefetch_desc(&maps.acts);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(full + s, 2); mbar_init(empty + s, 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(acc_full + s, 1); mbar_init(acc_empty + s, 16); }
    for (int s = 0; s < 8; ++s) mbar_init(a_ready + s, 8);
    for (int s = 0; s < kFMaxIn; ++s) mbar_init(in_full + s, 2);
    mbar_init(in_free, 9);                     // the issuer's commit + the 8 epilogue warps (their P stores drained)
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc2(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    // A "unit" is a pair of k-blocks (the last one of a layer may be single).  Units of the latent tail of fc_1
    // (their A operand is the latent part of the field input) take one ring stage per A k-panel followed by the
    // weight stage; all other units take the weight stage only.  Producer and issuer walk the same sequence.
    if (warp == 0) {
      // ---- TMA producer: runs ahead of the MMAs by the depth of the ring, across layers and tiles
      PipeState st;
      long long t_empty = 0;
      const long long t_begin = clock64();
      bool dead_n = !kBwd && n_iter > 0 && tile_dead(a.touch, a.n_flags, pair);
      for (int it = 0; it < n_iter; ++it) {
        const int row0 = ((pair + it * n_pairs) * 2 + (int)crank) * 128;
        const bool dead = dead_n;                      // (the next tile's flags are fetched a tile ahead)
        dead_n = !kBwd && it + 1 < n_iter && tile_dead(a.touch, a.n_flags, pair + (it + 1) * n_pairs);
        for (int l = 0; l < nl; ++l) {
          const int kb_main = a.L[l].kb_main, kb_tot = kb_main + a.L[l].kb_z;
          const int nc = a.L[l].n_chunks, ext_col = a.L[l].ext_col;
          // a dead tile walks [skip_head, kb_main) only: the units of latent k-panels in front and behind are left out
          const int kb_begin = dead ? a.L[l].skip_head : 0;
          const int kb_end = dead && a.L[l].skip_z ? kb_main : kb_tot;
          for (int c = 0; c < nc; ++c)
            for (int kb = kb_begin; kb < kb_end; kb += 2) {
              if (kProf && (a.dbg & 1)) continue;
              const int nk = kb_end - kb < 2 ? 1 : 2;
              if (kb >= kb_main) {                     // this CTA's 128 rows of the trailing global A k-panels
                for (int h = 0; h < nk; ++h) {
                  FUSED_TIMED(t_empty, mbar_wait(empty + st.stage, st.phase ^ 1));
                  if (elect_one()) {
                    if (cta_leader) mbar_expect_tx(full + st.stage, 2 * kFPanel);
                    else mbar_arrive_leader(full + st.stage);
                    tma_load_2d_pair(sRing + st.stage * kFStageB, &maps.in, full + st.stage,
                                     ext_col + (kb + h - kb_main) * 64, row0);
                  }
                  __syncwarp();
                  st.advance(kStages);
                }
              }
              FUSED_TIMED(t_empty, mbar_wait(empty + st.stage, st.phase ^ 1));
              if (elect_one()) {
                if (cta_leader) mbar_expect_tx(full + st.stage, 2 * nk * kFBoxB);
                else mbar_arrive_leader(full + st.stage);
                uint8_t* dst = sRing + st.stage * kFStageB;
                const int wrow = c * 128 + (int)crank * 64;
                tma_load_2d_pair_hint(dst, &maps.w[l], full + st.stage, kb * 64, wrow, kL2EvictLast);
                if (nk == 2)
                  tma_load_2d_pair_hint(dst + kFBoxB, &maps.w[l], full + st.stage, kb * 64 + 64, wrow, kL2EvictLast);
              }
              __syncwarp();
              st.advance(kStages);
            }
        }
      }
      if (kProf && lane == 0) {
        a.prof[blockIdx.x * 32 + 0] = clock64() - t_begin;
        a.prof[blockIdx.x * 32 + 1] = t_empty;
      }
    } else if (warp == 3) {
      // ---- first-layer A operand (field input / d_field): k-panels straight into P, which is dead between the
      // last layer that reads it and the second layer's epilogue.  One barrier per panel: the MMAs start on panel 0.
      const int kb_in = a.L[0].kb_main;
      for (int it = 0; it < n_iter; ++it) {
        if (it > 0) mbar_wait(in_free, (it - 1) & 1);
        const int row0 = ((pair + it * n_pairs) * 2 + (int)crank) * 128;
        if (elect_one()) {
          for (int kb = 0; kb < kb_in; ++kb) {
            if (cta_leader) mbar_expect_tx(in_full + kb, 2 * kFPanel);
            else mbar_arrive_leader(in_full + kb);
            tma_load_2d_pair(sP + kb * kFPanel, &maps.in, in_full + kb, kb * 64, row0);
          }
        }
        __syncwarp();
      }
    } else if (warp == 1 && cta_leader) {
      // ---- MMA issuer for the pair: the whole warp runs the loop (uniform control flow), one elected lane issues
      const uint32_t idesc = a.idesc;
      const uint32_t sP_u = smem_u32(sP), sRing_u = smem_u32(sRing);
      PipeState st;
      uint32_t n = 0;                              // running chunk counter -> accumulator buffer / phase
      long long t_acc = 0, t_ready = 0, t_full = 0;
      const long long t_begin = clock64();
      bool dead_n = !kBwd && n_iter > 0 && tile_dead(a.touch, a.n_flags, pair);
      for (int it = 0; it < n_iter; ++it) {
        const bool dead = dead_n;
        dead_n = !kBwd && it + 1 < n_iter && tile_dead(a.touch, a.n_flags, pair + (it + 1) * n_pairs);
        for (int l = 0; l < nl; ++l) {
          const int kb_main = a.L[l].kb_main, kb_tot = kb_main + a.L[l].kb_z;
          const int a_src = a.L[l].a_src;
          const uint32_t a_par = (uint32_t)(it * a.n_prod + l - 1) & 1;   // phase of the epilogue that produced A
          const int nc = a.L[l].n_chunks;
          // a dead tile walks [kb_begin, kb_end) = [skip_head, kb_main) only (the producer leaves the same ring stages
          // out): the latent k-panels in front and behind would add exact zeros.  The chunk's first MMA - k-block
          // kb_begin, step 0 - overwrites the accumulator.
          const int kb_begin = dead ? a.L[l].skip_head : 0;
          const int kb_end = dead && a.L[l].skip_z ? kb_main : kb_tot;
          if (kb_begin > 0 && a_src == kSrcIn && !(kProf && (a.dbg & 8)))
            for (int kb = 0; kb < kb_begin; ++kb)        // the skipped first-layer panels still land in P: nothing may be
              mbar_wait(in_full + kb, it & 1);          // in flight towards P when the next layers' epilogues write it
          for (int c = 0; c < nc; ++c, ++n) {
            const uint32_t buf = n & 1;
            if (!(kProf && (a.dbg & 8))) FUSED_TIMED(t_acc, mbar_wait(acc_empty + buf, ((n >> 1) & 1) ^ 1));
            const uint32_t tmem_d = tmem_base + buf * 128;
            for (int kb0 = kb_begin; kb0 < kb_end; kb0 += 2) {
              const int nk = kb_end - kb0 < 2 ? 1 : 2;
              const bool ext = kb0 >= kb_main;                             // A k-panels arrive through the ring
              uint32_t pa0 = 0, pa1 = 0;                                    // smem A panels of the unit's k-blocks
              int sa0 = 0, sa1 = 0;
              if (ext) {
                sa0 = st.stage;
                pa0 = sRing_u + st.stage * kFStageB;
                if (!(kProf && (a.dbg & 1))) FUSED_TIMED(t_full, mbar_wait(full + st.stage, st.phase));
                st.advance(kStages);
                if (nk == 2) {
                  sa1 = st.stage;
                  pa1 = sRing_u + st.stage * kFStageB;
                  if (!(kProf && (a.dbg & 1))) FUSED_TIMED(t_full, mbar_wait(full + st.stage, st.phase));
                  st.advance(kStages);
                }
              } else if (a_src != kSrcQ) {
                pa0 = sP_u + kb0 * kFPanel;
                pa1 = pa0 + kFPanel;
              }
              if (!ext && c == 0 && !(kProf && (a.dbg & 8))) {
                if (a_src == kSrcIn) {                  // loaded by warp 3
                  FUSED_TIMED(t_ready, mbar_wait(in_full + kb0, it & 1));
                  if (nk == 2) FUSED_TIMED(t_ready, mbar_wait(in_full + kb0 + 1, it & 1));
                } else {                                // produced by the previous layer's epilogue
                  FUSED_TIMED(t_ready, mbar_wait(a_ready + kb0, a_par));
                  if (nk == 2) FUSED_TIMED(t_ready, mbar_wait(a_ready + kb0 + 1, a_par));
                }
              }
              if (!(kProf && (a.dbg & 1))) FUSED_TIMED(t_full, mbar_wait(full + st.stage, st.phase));
              tc_fence_after();
              const uint32_t sb = sRing_u + st.stage * kFStageB;
              if (elect_one()) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                  if (h < nk && !(kProf && (a.dbg & 4))) {
                    const int kb = kb0 + h;
                    const uint64_t bdesc = make_sdesc(sb + h * kFBoxB, 16, 1024);
                    if (!ext && a_src == kSrcQ) {
                      const uint32_t ta = tmem_base + kFColQ + kb * 32;
#pragma unroll
                      for (int k = 0; k < 4; ++k)
                        umma_bf16_pair_ts(tmem_d, ta + k * 8, bdesc + 2 * k, idesc, ((kb ^ kb_begin) | k) != 0);
                    } else {
                      const uint64_t adesc = make_sdesc(h == 0 ? pa0 : pa1, 16, 1024);
#pragma unroll
                      for (int k = 0; k < 4; ++k)
                        umma_bf16_pair(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, ((kb ^ kb_begin) | k) != 0);
                    }
                  }
                }
                if (!(kProf && (a.dbg & 1))) {
                  if (ext) {
                    umma_commit_pair(empty + sa0);
                    if (nk == 2) umma_commit_pair(empty + sa1);
                  }
                  umma_commit_pair(empty + st.stage);
                }
              }
              __syncwarp();
              st.advance(kStages);
            }
            if (elect_one()) {
              umma_commit_pair(acc_full + buf);
              if (l == a.l_p_free && c == nc - 1) umma_commit_pair(in_free);
            }
            __syncwarp();
          }
        }
      }
      if (kProf && lane == 0) {
        a.prof[blockIdx.x * 32 + 2] = clock64() - t_begin;
        a.prof[blockIdx.x * 32 + 3] = t_acc;
        a.prof[blockIdx.x * 32 + 4] = t_ready;
        a.prof[blockIdx.x * 32 + 5] = t_full;
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
    // ---- epilogue: 8 warps that never synchronise with each other.  Warp (g, q4) owns rows [32 q4, 32 q4 + 32)
    // x columns [64g, 64g+64) of every chunk; it stages and TMA-stores its own 32-row boxes, keeps its own copy of
    // the chunk's bias values in shared memory and signals the MMA issuer on its own.
    const int ew = warp - 4;
    EpiCtx e;
    e.tmem_base = tmem_base;
    e.g = ew >> 2;                                   // column half of every chunk
    e.q4 = warp & 3;                                 // TMEM lane quarter
    e.lane = lane;
    e.row = e.q4 * 32 + lane;
    e.lane_off = (uint32_t)(e.q4 * 32) << 16;
    e.sw128 = (uint32_t)(e.row & 7);
    e.sP = smem_u32(sP);
    e.slot = smem_u32(sStage) + ew * kFSlot;         // this warp's staging slot
    e.wbias = smem_u32(sBias) + ew * 256;            // this warp's bias values of the current chunk
    e.acc_full = smem_u32(acc_full);
    e.acc_empty = smem_u32(acc_empty);
    e.a_ready = smem_u32(a_ready);
    e.t_wait = 0;
    const long long t_begin = clock64();
    const int g = e.g;
    uint32_t xres[kFChunks][32];                     // this thread's slice of the bf16 residual stream (x' or dL/dx')
    uint2 gate[2] = {make_uint2(0u, 0u), make_uint2(0u, 0u)};   // backward: gate words of this chunk and the next
    uint32_t n = 0;
    // what the next chunk (c of layer l of this CTA's tile `it`) needs prefetched: its bias values / its gate row
    // what a later chunk (c of layer l of this CTA's tile `it`; c may run past the layer) needs prefetched
    auto norm = [&](int& it, int& l, int& c) {          // c runs at most two chunks past its layer
      if (c >= a.L[l].n_chunks) { c -= a.L[l].n_chunks; ++l; }
      if (l == nl) { l = 0; ++it; }
    };
    auto pre_bias = [&](int it, int l, int c) -> const void* {
      norm(it, l, c);
      return kBwd ? nullptr : a.L[l].bias + c * 128 + g * 64;
    };
    auto pre_gate = [&](int it, int l, int c) -> const uint2* {
      if (!kBwd) return nullptr;
      norm(it, l, c);
      int64_t r = (int64_t)((pair + it * n_pairs) * 2 + (int)crank) * 128 + e.row;
      if (r > a.N - 1) r = a.N - 1;                  // ragged tail / past the last tile: any valid row will do
      return a.gate_bits + ((int64_t)a.L[l].mask_slot * a.N + r) * 8 + c * 2 + g;
    };
    if (!kBwd) {
      const float* b0 = reinterpret_cast<const float*>(pre_bias(0, 0, 0));
      sts32f(e.wbias + lane * 4, __ldg(b0 + lane));
      sts32f(e.wbias + (lane + 32) * 4, __ldg(b0 + lane + 32));
      __syncwarp();
    } else {
      gate[0] = __ldg(pre_gate(0, 0, 0));
      gate[1] = __ldg(pre_gate(0, 0, 1));
    }
    for (int it = 0; it < n_iter; ++it) {
      const int row0 = ((pair + it * n_pairs) * 2 + (int)crank) * 128;
      for (int l = 0; l < nl; ++l) {
        const FLayer& L = a.L[l];
        const int kind = L.kind;
        const bool first = L.first != 0;
        if (l == nl - 1 && lane == 0) {              // this warp's TMA stores that read P panels have drained
          if (kSave) bulk_wait_read0();
          mbar_arrive(in_free);
        }
        if (kind == kLayerX) {
#pragma unroll
          for (int c = 0; c < kFChunks; ++c, ++n)
            epi_chunk<kLayerX, kSave, kBwd, kProf, kHalf>(e, maps, a, L, c, n, row0, first, xres[c], gate[c & 1],
                                            pre_bias(it, l, c + 1), pre_gate(it, l, c + 2));
        } else if (kind == kLayerNet) {
#pragma unroll 2
          for (int c = 0; c < kFChunks; ++c, ++n)
            epi_chunk<kLayerNet, kSave, kBwd, kProf, kHalf>(e, maps, a, L, c, n, row0, false, xres[0], gate[c & 1],
                                              pre_bias(it, l, c + 1), pre_gate(it, l, c + 2));
        } else if (!kBwd) {
#pragma unroll 1
          for (int c = 0; c < L.n_chunks; ++c, ++n)
            epi_chunk<kLayerOut, kSave, kBwd, kProf, kHalf>(e, maps, a, L, c, n, row0, false, xres[0], gate[0],
                                              pre_bias(it, l, c + 1), pre_gate(it, l, c + 2));
        }
      }
    }
    if (kProf && lane == 0) {
      a.prof[blockIdx.x * 32 + 8 + 2 * ew] = clock64() - t_begin;
      a.prof[blockIdx.x * 32 + 9 + 2 * ew] = e.t_wait;
    }
    if (kSave && lane == 0) bulk_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, 512);
  }
}

static int make_acts_map(CUtensorMap* map, void* acts, int64_t N, int n_slots) {
  EncodeTiledFn fn = encode_tiled_fn();
  NRF_REQUIRE(fn != nullptr, NRF_ECUDA, "cuTensorMapEncodeTiled entry point not found");
  cuuint64_t gdim[3] = {512, (cuuint64_t)N, (cuuint64_t)n_slots};
  cuuint64_t gstride[2] = {512 * 2, (cuuint64_t)N * 512 * 2};
  cuuint32_t box[3] = {64, 32, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, acts, gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  NRF_REQUIRE(r == CUDA_SUCCESS, NRF_ECUDA, "cuTensorMapEncodeTiled(acts) failed (%d)", (int)r);
  return NRF_OK;
}

}  // namespace

int mlp_fused_launch(const FusedDesc& d, cudaStream_t stream) {
  NRF_REQUIRE(d.n_layers >= 2 && d.n_layers <= kFusedMaxLayers, NRF_EINVAL, "mlp_fused: %d layers", d.n_layers);
  NRF_REQUIRE(d.in_cols % 64 == 0 && d.in_cols >= 64, NRF_ENOSUP, "mlp_fused: %d input columns", d.in_cols);
  NRF_REQUIRE(!d.backward || d.saves, NRF_EINVAL, "mlp_fused: the backward needs its output buffer");
  FMaps maps;
  FArgs a;
  memset(&a, 0, sizeof(a));
  int rc;
  int n_prod = 0;
  for (int l = 0; l < d.n_layers; ++l) {
    const FusedLayerDesc& L = d.L[l];
    const int ktot = (L.kb_main + L.kb_z) * 64;
    const int nchunks = L.n_chunks > 0 ? L.n_chunks : kFChunks;
    NRF_REQUIRE(nchunks == kFChunks || (L.kind == 2 && nchunks >= 1 && nchunks <= 8), NRF_ENOSUP,
                "mlp_fused: layer %d has %d output chunks (only lin_out may differ from %d)", l, nchunks, kFChunks);
    if ((rc = make_map(&maps.w[l], L.W, ktot, nchunks * 128, L.ldw, 64, 64))) return rc;   // box = one k-block of a CTA's 64 rows
    NRF_REQUIRE(d.backward || ((reinterpret_cast<uintptr_t>(L.bias) & 15) == 0 && L.bias), NRF_EINVAL,
                "mlp_fused: bias alignment");
    a.L[l].kind = L.kind; a.L[l].a_src = L.a_src; a.L[l].kb_main = L.kb_main; a.L[l].kb_z = L.kb_z;
    a.L[l].first = L.first; a.L[l].act_slot = d.saves ? L.act_slot : -1; a.L[l].bias = L.bias;
    a.L[l].mask_slot = L.mask_slot;
    a.L[l].n_chunks = nchunks; a.L[l].ext_col = L.ext_col;
    a.L[l].skip_head = d.touch && !d.backward && L.skip_head % 2 == 0 ? L.skip_head : 0;   // whole units of two k-panels
    a.L[l].skip_z = d.touch && !d.backward ? L.skip_z : 0;
    NRF_REQUIRE(a.L[l].skip_head < L.kb_main, NRF_EINVAL, "mlp_fused: layer %d would skip all of its main k-panels", l);
    a.L[l].publish = l + 1 < d.n_layers && L.kind != 2;
    n_prod += a.L[l].publish;
  }
  for (int l = d.n_layers; l < kFusedMaxLayers; ++l) maps.w[l] = maps.w[0];
  if ((rc = make_map(&maps.in, d.in, d.in_cols, d.N, d.in_cols, 64, 128))) return rc;
  const bool save = d.saves != nullptr;
  if (save) {
    if ((rc = make_acts_map(&maps.acts, d.saves, d.N, d.n_slots))) return rc;
  } else {
    maps.acts = maps.in;
  }
  a.n_layers = d.n_layers;
  a.n_tiles = (int)((d.N + 255) / 256);
  a.n_prod = n_prod;
  NRF_REQUIRE(d.L[0].a_src == 0 && d.L[0].kb_main <= kFMaxIn && (d.L[0].kb_main + d.L[0].kb_z) * 64 == d.in_cols &&
                  (d.L[0].kb_z == 0 || (d.L[0].ext_col == d.L[0].kb_main * 64 && d.L[0].kb_main % 2 == 0)),
              NRF_ENOSUP, "mlp_fused: the first layer reads %d input columns: at most %d k-panels in P, the rest "
              "through the ring", d.in_cols, kFMaxIn);
  a.l_p_free = 0;
  for (int l = 0; l < d.n_layers; ++l)
    if (d.L[l].a_src == 2 || (l == 0)) a.l_p_free = l;
  NRF_REQUIRE(a.l_p_free < d.n_layers - 1 || d.backward, NRF_EINVAL, "mlp_fused: forward programs end on lin_out");
  NRF_REQUIRE(n_prod == d.n_layers - 1, NRF_EINVAL, "mlp_fused: every layer but the last must feed the next");
  a.N = (int)d.N; a.d_out = d.d_out; a.ldo = d.ldo; a.out = d.out;
  a.touch = d.backward ? nullptr : d.touch;
  a.n_flags = (int)((d.N + 31) / 32);
  a.gate_bits = reinterpret_cast<uint2*>(d.gate_bits);
  a.saves = reinterpret_cast<uint32_t*>(d.saves);
  NRF_REQUIRE(!d.saves || d.gate_bits, NRF_EINVAL, "mlp_fused: saving needs the gate-bit buffer");
  { const char* e = getenv("NRF_DBG"); a.dbg = e ? atoi(e) : 0; }
  a.prof = reinterpret_cast<long long*>(d.prof);
  int grid = sm_count() / 2 * 2;
  if (grid > 2 * a.n_tiles) grid = 2 * a.n_tiles;
  const bool prof = a.prof != nullptr;
  // the backward is bf16 in every mode; the forward's operands are bf16 or fp16 (kHalf); the cycle-counter
  // instantiations exist for bf16 only
  NRF_REQUIRE(!(d.backward && d.half), NRF_EINVAL, "mlp_fused: the backward runs in bf16");
  const bool half = d.half != 0;
  NRF_REQUIRE(!(prof && half), NRF_ENOSUP, "mlp_fused: the profiling build covers the bf16 kernels only");
  a.idesc = make_idesc(256, 128, 0, 0, half, half);
  auto kern = d.backward ? (prof ? mlp_fused_kernel<true, true, true, false> : mlp_fused_kernel<true, true, false, false>)
              : save     ? (prof   ? mlp_fused_kernel<true, false, true, false>
                            : half ? mlp_fused_kernel<true, false, false, true>
                                   : mlp_fused_kernel<true, false, false, false>)
                         : (prof   ? mlp_fused_kernel<false, false, true, false>
                            : half ? mlp_fused_kernel<false, false, false, true>
                                   : mlp_fused_kernel<false, false, false, false>);
  const int smem_bytes = save ? FCfg<true>::kSmem : FCfg<false>::kSmem;
  // per launch: the attribute is per (function, device); a process-wide "already set" flag breaks the second GPU of a
  // process, and the call costs well under a microsecond
  NRF_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kFThreads); cfg.dynamicSmemBytes = smem_bytes; cfg.stream = stream;
  { LaunchScope ls_(d.backward ? NRF_CAT_FUSED_BWD : NRF_CAT_FUSED_FWD, stream);
  NRF_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, maps, a));
  }
  NRF_LAUNCH_OK();
  return NRF_OK;
}

}  // namespace nrf
