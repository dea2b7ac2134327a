// tcgen05 / TMEM / TMA GEMMs of the field MLP (NRF_PREC_BF16): bf16 operands, fp32 accumulate.
//
//   gemm_tc_kernel<BN>   out = epilogue([A1|A2] . B^T)        forward layers and the dgrad chain
//   wgrad_tc_kernel<BK_> dW += G^T . A  (split over samples)  weight gradients
//
// Structure (both): persistent CTAs, 8 warps.
//   warp 0    TMA producer      cp.async.bulk.tensor.2d -> 128B-swizzled smem ring, mbarrier tx counts
//   warp 1    MMA issuer        one elected lane issues tcgen05.mma.cta_group::1.kind::f16 (M=128),
//                               tcgen05.commit releases smem stages / publishes the accumulator
//   warp 2    TMEM allocator    2 accumulator stages so the epilogue of tile i overlaps the MMAs of i+1
//   warps 4-7 epilogue          tcgen05.ld 32x32b (one TMEM lane = one output row per thread), fused
//                               bias / ReLU-gate / residual / bf16 re-quantisation; ReLU-gate and residual
//                               tiles arrive by TMA (prefetched 3 chunks ahead), results leave by TMA store

#include "tc_ptx.cuh"

namespace nrf {

// ------------------------------------------------------------------------------- forward / dgrad
// Epilogue I/O goes through TMA as well: the ReLU-gate and residual tiles are prefetched three
// 32-column chunks ahead into 64B-swizzled shared-memory buffers, results are written to swizzled
// staging buffers and leave through cp.async.bulk.tensor stores (coalesced, asynchronous, clipped to
// the tensor bounds, so there are no row/column predicates anywhere in the kernel).
template <int BN, int CG>
struct FwdCfg {
  static constexpr int kStageA = kTileM * kTileK * 2;          // 16 KB: this CTA's 128 rows of A
  static constexpr int kStageB = BN * kTileK * 2 / CG;         // B tile; a CTA pair (CG = 2) holds half each
  static constexpr int kStage = kStageA + kStageB;
  static constexpr int kStages = (CG == 2 || BN == 128) ? 4 : 2;   // 32 KB stages x 4 (BN=256 single CTA: 48 KB x 2)
  static constexpr int kEpiBuf = kTileM * 64;                  // 128 rows x 64 B (32 bf16 columns)
  static constexpr int kEpiBufs = 12;                          // per epilogue group: 2 slots of outA | outB or mask | resid
  static constexpr int kInSlots = 2;
  static constexpr int kTmemCols = 2 * BN;                     // 2 accumulator stages
  static constexpr int kSmem = kStages * kStage + kEpiBufs * kEpiBuf + 1024 /*align*/ + 512 /*barriers*/ +
                               BN * 4 /*bias slice*/;
};

struct FwdArgs {
  int M, N, n_store;
  int kb[3];                 // 64-wide k-blocks per A source
  const float* bias;
  int has_mask, has_resid, has_outA, relu_a, has_outB, relu_b, f32_out;
  int dbg;                   // NRF_DBG experiments: 1 = no TMA stores, 2 = no epilogue at all (timing only!)
  uint32_t idesc;            // tcgen05 instruction descriptor (tile shape + the operand formats, OpFmt)
  int io_half;               // resid / out_act / out_act2 are fp16 (1) or bf16 (0)
  const int32_t* live;       // optional (single-CTA kernels, one n tile): live[0] = number of row tiles to compute,
                             // live[1 + i] = their indices, ascending; the other row tiles are left unwritten
};
// 16-bit operand <-> fp32 in either format (warp-uniform run-time choice: this is the layer-by-layer chain, not the
// fused kernel's hot loop)
__device__ __forceinline__ uint32_t pack_op16(float lo, float hi, bool half) {
  uint32_t r;
  if (half) asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  else asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ float op16_to_f32(uint16_t h, bool half) {
  return half ? __half2float(__ushort_as_half(h)) : __uint_as_float((uint32_t)h << 16);
}
// compile-time epilogue kinds (EPI < 0: decided at run time from FwdArgs)
constexpr int kEpiMask = 1, kEpiResid = 2, kEpiOutB = 4, kEpiF32 = 8;

// CG = 2: the two CTAs of a cluster (one TPC) run ONE tcgen05.mma.cta_group::2 of M = 256: each CTA holds its
// own 128 rows of A and HALF of the B tile, the leader CTA issues the MMAs for both, each CTA's TMEM receives
// the accumulator of its own rows.  Per MMA every SM ingests 32 KB per k-block instead of 48 KB, which is what
// bounded the single-CTA kernel (profiles/r01_gemm_v2_ncu.md).
template <int BN, int CG, int EPI>
__global__ void __launch_bounds__(kFwdThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
               const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmMask, const __grid_constant__ CUtensorMap tmResid,
               const __grid_constant__ CUtensorMap tmOutA, const __grid_constant__ CUtensorMap tmOutB,
               const __grid_constant__ CUtensorMap tmOutF, const FwdArgs a) {
  using Cfg = FwdCfg<BN, CG>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* epi = smem + Cfg::kStages * Cfg::kStage;
  uint64_t* full = reinterpret_cast<uint64_t*>(epi + Cfg::kEpiBufs * Cfg::kEpiBuf);
  uint64_t* empty = full + Cfg::kStages;
  uint64_t* acc_full = empty + Cfg::kStages;
  uint64_t* acc_empty = acc_full + 2;
  uint64_t* in_full = acc_empty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(in_full + 2 * Cfg::kInSlots);
  float* sbias = reinterpret_cast<float*>(epi + Cfg::kEpiBufs * Cfg::kEpiBuf + 512);

  const int warp = uniform_warp_idx(), lane = threadIdx.x % 32;
  const int m_tiles = (a.M + kTileM - 1) / kTileM, n_tiles = a.N / BN;
  const int kb01 = a.kb[0] + a.kb[1];
  const int num_kb = kb01 + a.kb[2];
  // tile sequence of this CTA: iteration `it` -> (m_blk, n_blk).  A pair walks (row pair, n block) tiles.
  uint32_t crank = 0;
  if (CG == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
  const bool cta_leader = crank == 0;
  const int unit = CG == 2 ? blockIdx.x / 2 : blockIdx.x, n_units = CG == 2 ? gridDim.x / 2 : gridDim.x;
  const int32_t* live = CG == 1 ? a.live : nullptr;
  const int work = live ? __ldg(live) : (CG == 2 ? (m_tiles + 1) / 2 : m_tiles) * n_tiles;
  const int n_iter = unit < work ? (work - unit + n_units - 1) / n_units : 0;
  auto tile_of = [&](int it, int& m_blk, int& n_blk) {
    int t = unit + it * n_units;
    if (live) { m_blk = __ldg(live + 1 + t); n_blk = 0; return; }
    m_blk = CG == 2 ? 2 * (t / n_tiles) + (int)crank : t / n_tiles;
    n_blk = t % n_tiles;
  };
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA0);
    tma_prefetch_desc(&tmB);
    if (a.kb[1] > 0) tma_prefetch_desc(&tmA1);
    if (a.kb[2] > 0) tma_prefetch_desc(&tmA2);
    if (a.has_mask) tma_prefetch_desc(&tmMask);
    if (a.has_resid) tma_prefetch_desc(&tmResid);
    if (a.has_outA) tma_prefetch_desc(&tmOutA);
    if (a.has_outB) tma_prefetch_desc(&tmOutB);
    if (a.f32_out) tma_prefetch_desc(&tmOutF);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < Cfg::kStages; ++s) { mbar_init(full + s, CG); mbar_init(empty + s, 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(acc_full + s, 1); mbar_init(acc_empty + s, 8 * CG); }
    for (int s = 0; s < 2 * Cfg::kInSlots; ++s) mbar_init(in_full + s, 1);
    fence_barrier_init();
  }
  if (warp == 2) { if (CG == 2) tmem_alloc2(tmem_slot, Cfg::kTmemCols); else tmem_alloc(tmem_slot, Cfg::kTmemCols); }
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync();   // both CTAs' barriers exist before the peer signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // whole warp runs the loop (uniform control flow: addresses stay in uniform registers), one elected lane issues
    PipeState st;
    for (int it = 0; it < n_iter; ++it) {
      int m_blk, n_blk;
      tile_of(it, m_blk, n_blk);
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(empty + st.stage, st.phase ^ 1);
        uint8_t* sa = smem + st.stage * Cfg::kStage;
        uint8_t* sb = sa + Cfg::kStageA;
        const CUtensorMap* mapA = kb < a.kb[0] ? &tmA0 : (kb < kb01 ? &tmA1 : &tmA2);
        const int kcol = (kb < a.kb[0] ? kb : (kb < kb01 ? kb - a.kb[0] : kb - kb01)) * kTileK;
        if (elect_one()) {
          if (CG == 2) {
            // both CTAs' loads complete on the LEADER's barrier (it issues the MMAs for the pair)
            if (cta_leader) mbar_expect_tx(full + st.stage, 2 * Cfg::kStage);
            else mbar_arrive_leader(full + st.stage);
            tma_load_2d_pair(sa, mapA, full + st.stage, kcol, m_blk * kTileM);
            tma_load_2d_pair(sb, &tmB, full + st.stage, kb * kTileK, n_blk * BN + (int)crank * (BN / 2));
          } else {
            mbar_expect_tx(full + st.stage, Cfg::kStage);
            tma_load_2d(sa, mapA, full + st.stage, kcol, m_blk * kTileM);
            tma_load_2d(sb, &tmB, full + st.stage, kb * kTileK, n_blk * BN);
          }
        }
        __syncwarp();
        st.advance(Cfg::kStages);
      }
    }
  } else if (warp == 1) {
    if (cta_leader) {
      const uint32_t idesc = a.idesc;
      PipeState st;
      for (int it = 0; it < n_iter; ++it) {
        int as = it & 1;
        uint32_t aphase = (it >> 1) & 1;
        mbar_wait(acc_empty + as, aphase ^ 1);
        uint32_t tmem_d = tmem_base + as * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(full + st.stage, st.phase);
          tc_fence_after();
          uint32_t sa = smem_u32(smem + st.stage * Cfg::kStage);
          uint32_t sb = sa + Cfg::kStageA;
          uint64_t adesc = make_sdesc(sa, 16, 1024);
          uint64_t bdesc = make_sdesc(sb, 16, 1024);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < kTileK / kUmmaK; ++k) {
              // +32 B per 16-element K step inside the 128 B swizzle row: +2 in the 16 B address field
              if (CG == 2) umma_bf16_pair(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
              else umma_bf16(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
            }
            if (CG == 2) umma_commit_pair(empty + st.stage);      // frees the stage in both CTAs
            else umma_commit(empty + st.stage);
          }
          __syncwarp();
          st.advance(Cfg::kStages);
        }
        if (elect_one()) {
          if (CG == 2) umma_commit_pair(acc_full + as);           // both CTAs' epilogues
          else umma_commit(acc_full + as);
        }
        __syncwarp();
      }
    }
  } else if (warp >= kEpiWarp0) {
    // ---- epilogue: two groups of 4 warps; group g owns the 32-column chunks c with c % 2 == g.  Every warp
    // reads its own TMEM lane quarter (warp % 4).  Each group has its own leader thread, named barrier,
    // staging slots and TMA prefetch cursor.
    const int eg = (warp - kEpiWarp0) >> 2;          // group 0 / 1
    const int q4 = warp & 3;                         // TMEM lane quarter this warp may read
    const int row = q4 * 32 + lane;                  // row of the tile == TMEM lane
    const bool leader = (row == 0);
    const bool has_mask = EPI < 0 ? a.has_mask != 0 : (EPI & kEpiMask) != 0;
    const bool has_resid = EPI < 0 ? a.has_resid != 0 : (EPI & kEpiResid) != 0;
    const bool has_outB = EPI < 0 ? a.has_outB != 0 : (EPI & kEpiOutB) != 0;
    const bool f32_out = EPI < 0 ? a.f32_out != 0 : (EPI & kEpiF32) != 0;
    const bool has_in = has_mask || has_resid;
    const bool io_half = a.io_half != 0;
    const uint32_t in_bytes = (uint32_t)((has_mask ? 1 : 0) + (has_resid ? 1 : 0)) * Cfg::kEpiBuf;
    // per group: [outA x2 | outB-or-mask x2 | resid x2] 8 KB buffers; an fp32 output chunk (16 KB) uses the
    // first four.  A second output and a ReLU-gate input never occur in the same GEMM.
    uint8_t* gbase = epi + eg * (Cfg::kEpiBufs / 2) * Cfg::kEpiBuf;
    uint8_t* bufOutA = gbase;
    uint8_t* bufOutB = gbase + 2 * Cfg::kEpiBuf;
    uint8_t* bufMask = gbase + 2 * Cfg::kEpiBuf;
    uint8_t* bufResid = gbase + 4 * Cfg::kEpiBuf;
    uint64_t* my_in_full = in_full + eg * Cfg::kInSlots;
    auto group_barrier = [&]() {
      if (eg == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
      else asm volatile("bar.sync 2, 128;" ::: "memory");
    };
    auto chunks_of = [&](int it) {                   // chunks of tile `it` (both groups together)
      int m_blk, n_blk;
      tile_of(it, m_blk, n_blk);
      int left = a.n_store - n_blk * BN;
      int c = (left + 31) / 32;
      return c < 0 ? 0 : (c > BN / 32 ? BN / 32 : c);
    };
    // prefetch cursor (leader only): runs kInSlots of this group's chunks ahead of the consumer
    int pf_it = 0, pf_c = eg;
    uint32_t pf_q = 0;
    auto prefetch_one = [&]() {
      while (pf_it < n_iter && pf_c >= chunks_of(pf_it)) { ++pf_it; pf_c = eg; }
      if (pf_it >= n_iter) return;
      int slot = pf_q % Cfg::kInSlots;
      int pm, pn;
      tile_of(pf_it, pm, pn);
      int m0 = pm * kTileM, n0 = pn * BN + pf_c * 32;
      mbar_expect_tx(my_in_full + slot, in_bytes);
      if (has_mask) tma_load_2d(bufMask + slot * Cfg::kEpiBuf, &tmMask, my_in_full + slot, n0, m0);
      if (has_resid) tma_load_2d(bufResid + slot * Cfg::kEpiBuf, &tmResid, my_in_full + slot, n0, m0);
      ++pf_q;
      pf_c += 2;
    };
    if (leader && has_in)
      for (int i = 0; i < Cfg::kInSlots; ++i) prefetch_one();

    uint32_t q = 0;                                  // running chunk counter of this group
    const uint32_t sw64 = (uint32_t)((row >> 1) & 3);   // 64B swizzle: 16 B chunk index ^= (row/2)%4
    const uint32_t sw128 = (uint32_t)(row & 7);         // 128B swizzle: 16 B chunk index ^= row%8
    int bias_blk = -1;                               // n block whose bias slice sits in sbias
    for (int it = 0; it < n_iter; ++it) {
      int tm, tn;
      tile_of(it, tm, tn);
      const int m0 = tm * kTileM, n_base = tn * BN;
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const int nch = chunks_of(it);
      if (a.bias && bias_blk != tn) {                // once per CTA when the CTA keeps its n block
        asm volatile("bar.sync 3, 256;" ::: "memory");            // nobody still reads the previous slice
        for (int i = (eg * 128 + row); i < BN; i += 256) sbias[i] = __ldg(a.bias + n_base + i);
        asm volatile("bar.sync 3, 256;" ::: "memory");
        bias_blk = tn;
      }
      mbar_wait(acc_full + as, aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + as * BN + ((uint32_t)(q4 * 32) << 16);
      const int last_c = nch - 1 - (((nch - 1) & 1) != eg ? 1 : 0);   // this group's last chunk (may be < eg: none)
      if ((a.dbg & 2) || last_c < eg) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { if (CG == 2) mbar_arrive_leader(acc_empty + as); else mbar_arrive(acc_empty + as); }
        continue;
      }
#pragma unroll 1
      for (int c = eg; c < nch; c += 2, ++q) {
        const int n0 = n_base + c * 32;
        uint32_t v[32];
        tmem_ld32(taddr + c * 32, v);
        if (c == last_c) {                           // this warp is done with the accumulator stage
          tc_fence_before();
          __syncwarp();
          if (lane == 0) { if (CG == 2) mbar_arrive_leader(acc_empty + as); else mbar_arrive(acc_empty + as); }
        }
        float x[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) x[j] = __uint_as_float(v[j]);
        if (a.bias) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float4 b = *reinterpret_cast<const float4*>(sbias + c * 32 + j);
            x[j] += b.x; x[j + 1] += b.y; x[j + 2] += b.z; x[j + 3] += b.w;
          }
        }
        if (has_in) {
          const int slot = q % Cfg::kInSlots;
          mbar_wait(my_in_full + slot, (q / Cfg::kInSlots) & 1);
          if (has_mask) {
            const uint8_t* mrow = bufMask + slot * Cfg::kEpiBuf + row * 64;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4 mk = *reinterpret_cast<const uint4*>(mrow + ((j ^ sw64) << 4));
              const uint16_t* h = reinterpret_cast<const uint16_t*>(&mk);
#pragma unroll
              for (int t = 0; t < 8; ++t)      // "> 0" read off the bits: sign clear and non-zero, in bf16 and fp16 alike
                if (!((int16_t)h[t] > 0)) x[j * 8 + t] = 0.0f;
            }
          }
          if (has_resid) {
            const uint8_t* rrow = bufResid + slot * Cfg::kEpiBuf + row * 64;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4 rk = *reinterpret_cast<const uint4*>(rrow + ((j ^ sw64) << 4));
              const uint16_t* h = reinterpret_cast<const uint16_t*>(&rk);
#pragma unroll
              for (int t = 0; t < 8; ++t) x[j * 8 + t] += op16_to_f32(h[t], io_half);
            }
          }
        }
        const int ob = q & 1;
        if (f32_out) {
          uint8_t* orow = gbase + ob * (2 * Cfg::kEpiBuf) + row * 128;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<float4*>(orow + ((j ^ sw128) << 4)) =
                make_float4(x[j * 4], x[j * 4 + 1], x[j * 4 + 2], x[j * 4 + 3]);
        } else {
          {
            uint8_t* orow = bufOutA + ob * Cfg::kEpiBuf + row * 64;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint32_t w[4];
#pragma unroll
              for (int t = 0; t < 4; ++t) {
                float p = x[j * 8 + 2 * t], r = x[j * 8 + 2 * t + 1];
                if (a.relu_a) { p = fmaxf(p, 0.0f); r = fmaxf(r, 0.0f); }
                w[t] = pack_op16(p, r, io_half);
              }
              *reinterpret_cast<uint4*>(orow + ((j ^ sw64) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
            }
          }
          if (has_outB) {
            uint8_t* orow = bufOutB + ob * Cfg::kEpiBuf + row * 64;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint32_t w[4];
#pragma unroll
              for (int t = 0; t < 4; ++t) {
                float p = x[j * 8 + 2 * t], r = x[j * 8 + 2 * t + 1];
                if (a.relu_b) { p = fmaxf(p, 0.0f); r = fmaxf(r, 0.0f); }
                w[t] = pack_op16(p, r, io_half);
              }
              *reinterpret_cast<uint4*>(orow + ((j ^ sw64) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
            }
          }
        }
        fence_proxy_async();                 // staging writes -> visible to the TMA (async proxy)
        if (leader) bulk_wait_read0();       // the other slot (chunk q-1 of this group) has been drained
        group_barrier();
        if (leader) {
          if (a.dbg & 1) {
          } else if (f32_out) {
            tma_store_2d(&tmOutF, gbase + ob * (2 * Cfg::kEpiBuf), n0, m0);
          } else {
            tma_store_2d(&tmOutA, bufOutA + ob * Cfg::kEpiBuf, n0, m0);
            if (has_outB) tma_store_2d(&tmOutB, bufOutB + ob * Cfg::kEpiBuf, n0, m0);
          }
          bulk_commit();
          if (has_in) prefetch_one();        // every thread of the group is past its reads of slot q % kInSlots
        }
      }
    }
    if (leader) bulk_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync();   // neither CTA leaves while the pair's MMAs / barrier traffic may still touch it
  if (warp == 2) {
    tc_fence_after();
    if (CG == 2) tmem_dealloc2(tmem_base, Cfg::kTmemCols); else tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}


// ------------------------------------------------------------------------------------- wgrad
// dW[n,k] += sum_m G[m,n] A[m,k].  UMMA A-operand = G^T tile (128 n x 64 m), B-operand = A^T tile
// (BK_ k x 64 m); both MN-major (the sample index m is the slow axis of G and A in memory).
// grid = (n tiles * k tiles, splits); each CTA reduces its sample range and adds its partial tile
// into dW with fp32 reductions (red.global.add.f32).
template <int BK_, int CG>
struct WgCfg {
  static constexpr int kStageG = kTileM * kTileK * 2;          // 2 slabs of [64 m][64 n]: this CTA's 128 n
  static constexpr int kStageA = BK_ * kTileK * 2 / CG;        // slabs of [64 m][64 k]; a CTA pair holds half each
  static constexpr int kStage = kStageG + kStageA;
  static constexpr int kStages = CG == 2 ? 6 : (BK_ == 256 ? 4 : 6);
  static constexpr int kTmemCols = BK_ == 256 ? 512 : 2 * BK_;  // accumulator + 32 columns for the bias sums
  static constexpr int kOnes = 2048;                           // 16 x 64 bf16 ones: B operand of the bias MMA
  static constexpr int kSmem = kStages * kStage + kOnes + 1024 + 256;
};

// CG = 2: the two CTAs of a cluster run ONE tcgen05.mma.cta_group::2 of M = 256: a pair covers 256 n x BK_ k of dW;
// each CTA loads the G^T slabs of its own 128 n and HALF of the A^T slabs (the tensor cores of both SMs read both
// halves), so every SM ingests 32 KB instead of 48 KB per 64 samples at BK_ = 256: the single-CTA kernel was bound
// by that L2 -> SM traffic (96 B/clk at full MMA rate).
template <int BK_, int CG>
__global__ void __launch_bounds__(kThreads, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmG, const __grid_constant__ CUtensorMap tmA, int M,
                int k_tiles, int m_per_split, int n_valid, int k_valid, float* __restrict__ dW, int ldw,
                float* __restrict__ dbias, float* __restrict__ ws, int n_pad, int k_pad, uint32_t idesc,
                uint32_t idesc_bias, uint32_t ones_word) {
  using Cfg = WgCfg<BK_, CG>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* ones = smem + Cfg::kStages * Cfg::kStage;
  uint64_t* full = reinterpret_cast<uint64_t*>(ones + Cfg::kOnes);
  uint64_t* empty = full + Cfg::kStages;
  uint64_t* acc_full = empty + Cfg::kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int warp = uniform_warp_idx(), lane = threadIdx.x % 32;
  uint32_t crank = 0;
  if (CG == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
  const bool cta_leader = crank == 0;
  const int unit = blockIdx.x / CG;                     // (n tile [pair], k tile)
  const int n_blk = (unit / k_tiles) * CG + (int)crank, k_blk = unit % k_tiles;
  const int m_begin = blockIdx.y * m_per_split;
  const int m_end = min(M, m_begin + m_per_split);
  const int num_mb = (m_end - m_begin + kTileK - 1) / kTileK;
  // bias gradient = column sums of G = G^T . 1: one extra N=16 MMA per k-step against a constant tile of ones (no
  // extra memory traffic, no extra barriers).  It re-reads the 4 KB G^T operand from shared memory, a quarter of a
  // main MMA's time, so the k-steps are dealt round-robin over the CTAs that share this G tile (the k tiles): every
  // CTA sums its share of the samples and the launch finishes together.
  const bool do_bias = dbias != nullptr;
  // k-step i belongs to k tile (i & bias_mask); a k-tile count that is not a power of two leaves all of it to tile 0
  const int bias_mask = (k_tiles & (k_tiles - 1)) == 0 ? k_tiles - 1 : 0;
  if (warp >= kEpiWarp0) {
    for (int i = threadIdx.x - kEpiWarp0 * 32; i < Cfg::kOnes / 4; i += 128)
      reinterpret_cast<uint32_t*>(ones)[i] = ones_word;        // two 1.0 in the B operand's format
    fence_proxy_async();
  }

  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmG); tma_prefetch_desc(&tmA); }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < Cfg::kStages; ++s) { mbar_init(full + s, CG); mbar_init(empty + s, 1); }
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 2) { if (CG == 2) tmem_alloc2(tmem_slot, Cfg::kTmemCols); else tmem_alloc(tmem_slot, Cfg::kTmemCols); }
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync();   // both CTAs' barriers and ones tiles exist before the peer signals / reads them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    PipeState st;
    for (int mb = 0; mb < num_mb; ++mb) {
      mbar_wait(empty + st.stage, st.phase ^ 1);
      uint8_t* sg = smem + st.stage * Cfg::kStage;
      uint8_t* sa = sg + Cfg::kStageG;
      int m0 = m_begin + mb * kTileK;
      if (elect_one()) {
        // rows >= M are zero-filled by TMA; rows in [m_end, M) of the last chunk belong to the next
        // split, so chunks are aligned: m_per_split is a multiple of 64.
        if (CG == 2) {
          if (cta_leader) mbar_expect_tx(full + st.stage, 2 * Cfg::kStage);
          else mbar_arrive_leader(full + st.stage);
#pragma unroll
          for (int s = 0; s < kTileM / 64; ++s)
            tma_load_2d_pair(sg + s * (kTileK * 128), &tmG, full + st.stage, n_blk * kTileM + s * 64, m0);
#pragma unroll
          for (int s = 0; s < BK_ / 128; ++s)
            tma_load_2d_pair(sa + s * (kTileK * 128), &tmA, full + st.stage,
                             k_blk * BK_ + (int)crank * (BK_ / 2) + s * 64, m0);
        } else {
          mbar_expect_tx(full + st.stage, Cfg::kStage);
#pragma unroll
          for (int s = 0; s < kTileM / 64; ++s)
            tma_load_2d(sg + s * (kTileK * 128), &tmG, full + st.stage, n_blk * kTileM + s * 64, m0);
#pragma unroll
          for (int s = 0; s < BK_ / 64; ++s)
            tma_load_2d(sa + s * (kTileK * 128), &tmA, full + st.stage, k_blk * BK_ + s * 64, m0);
        }
      }
      __syncwarp();
      st.advance(Cfg::kStages);
    }
  } else if (warp == 1) {
    if (cta_leader) {
      const uint64_t odesc = make_sdesc(smem_u32(ones), kTileK * 128, 1024);
      PipeState st;
      uint32_t bias_acc = 0;                       // the first bias MMA of this CTA overwrites its accumulator columns
      for (int mb = 0; mb < num_mb; ++mb) {
        mbar_wait(full + st.stage, st.phase);
        tc_fence_after();
        uint32_t sg = smem_u32(smem + st.stage * Cfg::kStage);
        uint32_t sa = sg + Cfg::kStageG;
        uint64_t gdesc = make_sdesc(sg, kTileK * 128, 1024);
        uint64_t adesc = make_sdesc(sa, kTileK * 128, 1024);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < kTileK / kUmmaK; ++k) {
            // 16 samples = two 8-row atoms = 2048 B: +128 in the 16 B address field
            const bool bias_step = do_bias && ((mb * (kTileK / kUmmaK) + k) & bias_mask) == k_blk;
            if (CG == 2) {
              umma_bf16_pair(tmem_base, gdesc + 128 * k, adesc + 128 * k, idesc, (mb | k) != 0);
              if (bias_step) umma_bf16_pair(tmem_base + BK_, gdesc + 128 * k, odesc, idesc_bias, bias_acc);
            } else {
              umma_bf16(tmem_base, gdesc + 128 * k, adesc + 128 * k, idesc, (mb | k) != 0);
              if (bias_step) umma_bf16(tmem_base + BK_, gdesc + 128 * k, odesc, idesc_bias, bias_acc);
            }
            if (bias_step) bias_acc = 1;
          }
          if (CG == 2) umma_commit_pair(empty + st.stage); else umma_commit(empty + st.stage);
        }
        __syncwarp();
        st.advance(Cfg::kStages);
      }
      if (elect_one()) { if (CG == 2) umma_commit_pair(acc_full); else umma_commit(acc_full); }
      __syncwarp();
    }
  } else if (warp >= kEpiWarp0) {
    const int q = warp - kEpiWarp0;
    if (num_mb > 0) {
      mbar_wait(acc_full, 0);
      tc_fence_after();
      const int n = n_blk * kTileM + q * 32 + lane;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
      for (int c = 0; c < BK_ / 32; ++c) {
        int k0 = k_blk * BK_ + c * 32;
        if (k0 >= k_valid) break;
        uint32_t v[32];
        tmem_ld32(taddr + c * 32, v);
        if (ws) {
          // deterministic mode: this split's partial tile goes to the workspace; wgrad_reduce_kernel adds the
          // splits up in a fixed order
          float4* dst = reinterpret_cast<float4*>(ws + ((int64_t)blockIdx.y * n_pad + n) * k_pad + k0);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            dst[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                 __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
        } else if (n < n_valid) {
          float* dst = dW + (int64_t)n * ldw + k0;
          if ((ldw & 3) == 0 && (reinterpret_cast<uintptr_t>(dW) & 15) == 0 && k0 + 32 <= k_valid) {
            // 16 B aligned rows: vector reductions (red.global.add.v4.f32)
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              atomicAdd(reinterpret_cast<float4*>(dst + j),
                        make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                                    __uint_as_float(v[j + 3])));
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (k0 + j < k_valid) atomicAdd(dst + j, __uint_as_float(v[j]));
          }
        }
      }
      // column BK_ of the accumulator: this CTA's share of the sum over the samples of G[:, n] (nothing if its share
      // of the k-steps was empty: a split shorter than k_tiles k-steps)
      if (do_bias && (bias_mask ? num_mb * (kTileK / kUmmaK) > k_blk : k_blk == 0)) {
        uint32_t v[32];
        tmem_ld32(taddr + BK_, v);
        if (ws) ws[(int64_t)gridDim.y * n_pad * k_pad + ((int64_t)k_blk * gridDim.y + blockIdx.y) * n_pad + n] = __uint_as_float(v[0]);
        else if (n < n_valid) atomicAdd(dbias + n, __uint_as_float(v[0]));
      } else if (do_bias && ws) {
        ws[(int64_t)gridDim.y * n_pad * k_pad + ((int64_t)k_blk * gridDim.y + blockIdx.y) * n_pad + n] = 0.0f;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync();   // neither CTA leaves while the pair's MMAs / barrier traffic may still touch it
  if (warp == 2) {
    tc_fence_after();
    if (CG == 2) tmem_dealloc2(tmem_base, Cfg::kTmemCols); else tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// dW[n,k] += sum over splits (ascending) of the partial tiles; dbias likewise.  One thread per 4 columns.
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* __restrict__ ws, int splits, int n_pad,
                                                           int k_pad, int n_valid, int k_valid,
                                                           float* __restrict__ dW, int ldw, float* __restrict__ dbias,
                                                           int bias_parts) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int kq = (k_valid + 3) / 4;
  if (t < (int64_t)n_valid * kq) {
    int n = (int)(t / kq), k = (int)(t % kq) * 4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s = 0; s < splits; ++s) {
      float4 p = *reinterpret_cast<const float4*>(ws + ((int64_t)s * n_pad + n) * k_pad + k);
      acc.x += p.x; acc.y += p.y; acc.z += p.z; acc.w += p.w;
    }
    float* dst = dW + (int64_t)n * ldw + k;
    dst[0] += acc.x;
    if (k + 1 < k_valid) dst[1] += acc.y;
    if (k + 2 < k_valid) dst[2] += acc.z;
    if (k + 3 < k_valid) dst[3] += acc.w;
  } else if (dbias) {
    int64_t n = t - (int64_t)n_valid * kq;
    if (n < n_valid) {
      const float* wb = ws + (int64_t)splits * n_pad * k_pad;
      float acc = 0.f;
      for (int s = 0; s < splits * bias_parts; ++s) acc += wb[(int64_t)s * n_pad + n];
      dbias[n] += acc;
    }
  }
}

// ----------------------------------------------------------------------------------- host side
// Row tiles (64 or 128 rows) that have work, from one flag byte per 32 rows: live[0] = count, live[1..] = tile indices.
// Block 0 builds the list of 128-row tiles into live128, block 1 (if launched) the list of 64-row blocks into live64.
__global__ void __launch_bounds__(1024) live_tiles_kernel(const uint8_t* __restrict__ flags, int64_t n_groups, int64_t n_rows,
                                                          int32_t* __restrict__ live128, int32_t* __restrict__ live64) {
  const int gpt = blockIdx.x == 0 && live128 ? 4 : 2;
  int32_t* __restrict__ live = gpt == 4 ? live128 : live64;
  const int n_tiles = (int)((n_rows + gpt * 32 - 1) / (gpt * 32));
  __shared__ int warp_tot[32];
  __shared__ int base_s;
  const int lane = threadIdx.x % 32, wid = threadIdx.x / 32;
  if (threadIdx.x == 0) base_s = 0;
  __syncthreads();
  for (int t0 = 0; t0 < n_tiles; t0 += 1024) {
    const int t = t0 + threadIdx.x;
    bool on = false;
    if (t < n_tiles)
      for (int j = 0; j < gpt; ++j) {
        const int64_t gidx = (int64_t)t * gpt + j;
        on = on || (gidx < n_groups && flags[gidx] != 0);
      }
    const unsigned m = __ballot_sync(0xffffffffu, on);
    if (lane == 0) warp_tot[wid] = __popc(m);
    __syncthreads();
    int before = base_s;
    for (int w = 0; w < wid; ++w) before += warp_tot[w];
    if (on) live[1 + before + __popc(m & ((1u << lane) - 1u))] = t;
    __syncthreads();
    if (threadIdx.x == 0) { int tot = 0; for (int w = 0; w < 32; ++w) tot += warp_tot[w]; base_s += tot; }
    __syncthreads();
  }
  if (threadIdx.x == 0) live[0] = base_s;
}

int live_tiles_launch(const void* flags, int64_t n_rows, int32_t* live128, int32_t* live64, cudaStream_t stream) {
  NRF_REQUIRE(flags && (live128 || live64), NRF_EINVAL, "live_tiles: null pointer");
  LaunchScope ls_(NRF_CAT_MISC, stream);
  live_tiles_kernel<<<(live128 && live64) ? 2 : 1, 1024, 0, stream>>>(reinterpret_cast<const uint8_t*>(flags),
                                                                      (n_rows + 31) / 32, n_rows, live128, live64);
  NRF_LAUNCH_OK();
  return NRF_OK;
}

template <int BN, int CG, int EPI>
static int launch_fwd(const NrfGemm& g, OpFmt fmt, cudaStream_t stream, const int32_t* live = nullptr) {
  using Cfg = FwdCfg<BN, CG>;
  const int a_rows = kTileM, b_rows = BN / CG;   // TMA box heights
  CUtensorMap tmA[3], tmB, tmMask, tmResid, tmOutA, tmOutB, tmOutF;
  int rc = 0;
  int ktot = 0;
  for (int i = 0; i < 3; ++i) {
    if (g.K[i] > 0) {
      rc = make_map(&tmA[i], g.A[i], g.K[i], g.M, g.lda[i], kTileK, a_rows);
      if (rc) return rc;
    } else {
      tmA[i] = tmA[0];
    }
    ktot += g.K[i];
  }
  rc = make_map(&tmB, g.B, ktot, g.N, g.ldb, kTileK, b_rows);
  if (rc) return rc;
  // epilogue tiles: 32 bf16 columns (64 B rows, 64 B swizzle) or 32 fp32 columns (128 B rows, 128 B swizzle)
  auto epi_map = [&](CUtensorMap* m, const void* ptr, int ld) {
    return make_map_ex(m, ptr, g.n_store, g.M, ld, 32, kTileM, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                       CU_TENSOR_MAP_SWIZZLE_64B);
  };
  tmMask = tmResid = tmOutA = tmOutB = tmOutF = tmB;
  if (g.mask_src && (rc = epi_map(&tmMask, g.mask_src, g.ldmask))) return rc;
  if (g.resid && (rc = epi_map(&tmResid, g.resid, g.ldr))) return rc;
  if (g.out_act && (rc = epi_map(&tmOutA, g.out_act, g.ldact))) return rc;
  if (g.out_act2 && (rc = epi_map(&tmOutB, g.out_act2, g.ldact2))) return rc;
  if (g.out_f32 && (rc = make_map_ex(&tmOutF, g.out_f32, g.n_store, g.M, g.ldo, 32, kTileM,
                                     CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, CU_TENSOR_MAP_SWIZZLE_128B)))
    return rc;
  FwdArgs a;
  a.M = g.M; a.N = g.N; a.n_store = g.n_store;
  for (int i = 0; i < 3; ++i) a.kb[i] = g.K[i] / kTileK;
  a.bias = g.bias;
  a.has_mask = g.mask_src != nullptr; a.has_resid = g.resid != nullptr;
  a.has_outA = g.out_act != nullptr; a.relu_a = g.relu_act;
  a.has_outB = g.out_act2 != nullptr; a.relu_b = g.relu_act2;
  a.f32_out = g.out_f32 != nullptr;
  { const char* e = getenv("NRF_DBG"); a.dbg = e ? atoi(e) : 0; }
  a.idesc = make_idesc(kTileM * CG, BN, 0, 0, fmt.a_half, fmt.b_half);
  a.io_half = fmt.io_half;
  a.live = live;
  // per launch: the attribute is per (function, device), see mlp_fused_launch
  NRF_CUDA_OK(cudaFuncSetAttribute(gemm_tc_kernel<BN, CG, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmem));
  int m_tiles = (g.M + kTileM - 1) / kTileM, n_tiles = g.N / BN;
  int tiles = m_tiles * n_tiles;
  int grid = tiles < sm_count() ? tiles : sm_count();
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cudaLaunchAttribute attr[1];
  if (CG == 2) {
    grid = sm_count() / 2 * 2;                       // whole CTA pairs; a pair walks (row pair, n block) tiles
    int work = (m_tiles + 1) / 2 * n_tiles;
    if (grid > work * 2) grid = work * 2;
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
  }
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kFwdThreads); cfg.dynamicSmemBytes = Cfg::kSmem; cfg.stream = stream;
  { LaunchScope ls_(NRF_CAT_GEMM, stream);
  NRF_CUDA_OK(cudaLaunchKernelEx(&cfg, gemm_tc_kernel<BN, CG, EPI>, tmA[0], tmA[1], tmA[2], tmB, tmMask, tmResid, tmOutA,
                                 tmOutB, tmOutF, a));
  }
  NRF_LAUNCH_OK();
  return NRF_OK;
}

int gemm_tc_launch(const NrfGemm& g, OpFmt fmt, cudaStream_t stream, const int32_t* live) {
  NRF_REQUIRE(fmt.a_half == fmt.b_half, NRF_ENOSUP, "gemm_tc: mixed fp16 / bf16 operands are not executable");
  NRF_REQUIRE(g.K[0] > 0 && g.K[0] % kTileK == 0 && g.K[1] % kTileK == 0 && g.K[2] % kTileK == 0, NRF_ENOSUP,
              "gemm_tc: K = %d,%d,%d must be multiples of 64", g.K[0], g.K[1], g.K[2]);
  NRF_REQUIRE(g.K[2] == 0 || g.K[1] > 0, NRF_EINVAL, "gemm_tc: A[2] needs A[1]");
  NRF_REQUIRE(g.N % 128 == 0, NRF_ENOSUP, "gemm_tc: N=%d must be a multiple of 128", g.N);
  NRF_REQUIRE(!g.out_f32 || (!g.mask_src && !g.resid && !g.out_act && !g.out_act2), NRF_ENOSUP,
              "gemm_tc: out_f32 excludes mask_src / resid / out_act* in bf16 mode");
  NRF_REQUIRE(g.out_act || !g.out_act2, NRF_EINVAL, "gemm_tc: out_act2 without out_act");
  NRF_REQUIRE(!(g.out_act2 && g.mask_src), NRF_ENOSUP, "gemm_tc: out_act2 and mask_src share staging buffers");
  NRF_REQUIRE((reinterpret_cast<uintptr_t>(g.bias) & 15) == 0, NRF_EINVAL, "gemm_tc: bias must be 16 B aligned");
  const int BN = g.N % 256 == 0 ? 256 : 128;
  NRF_REQUIRE(g.N - g.n_store < BN, NRF_EINVAL, "gemm_tc: N=%d over-padded for n_store=%d", g.N, g.n_store);
  // 256-wide tiles: CTA pairs with tcgen05.mma.cta_group::2 (NRF_GEMM_1CTA=1 keeps the single-CTA kernel);
  // the pair kernel is specialised on the epilogue kind (less code in the hot loop: it is i-cache sensitive)
  static const bool one_cta = getenv("NRF_GEMM_1CTA") != nullptr;
  if (BN == 256 && g.M > kTileM && !one_cta) {
    int epi = (g.mask_src ? kEpiMask : 0) | (g.resid ? kEpiResid : 0) | (g.out_act2 ? kEpiOutB : 0) |
              (g.out_f32 ? kEpiF32 : 0);
    switch (epi) {
      case 0: return launch_fwd<256, 2, 0>(g, fmt, stream);
      case kEpiOutB: return launch_fwd<256, 2, kEpiOutB>(g, fmt, stream);
      case kEpiResid: return launch_fwd<256, 2, kEpiResid>(g, fmt, stream);
      case kEpiResid | kEpiOutB: return launch_fwd<256, 2, kEpiResid | kEpiOutB>(g, fmt, stream);
      case kEpiMask: return launch_fwd<256, 2, kEpiMask>(g, fmt, stream);
      case kEpiMask | kEpiResid: return launch_fwd<256, 2, kEpiMask | kEpiResid>(g, fmt, stream);
      case kEpiF32: return launch_fwd<256, 2, kEpiF32>(g, fmt, stream);
      default: return launch_fwd<256, 2, -1>(g, fmt, stream);
    }
  }
  NRF_REQUIRE(!live || (BN == 128 && g.N == 128 && !g.mask_src && !g.resid), NRF_ENOSUP,
              "gemm_tc: a live-tile list needs the single-CTA kernel with one n tile and no epilogue inputs");
  if (BN == 256) return launch_fwd<256, 1, -1>(g, fmt, stream);
  return launch_fwd<128, 1, -1>(g, fmt, stream, live);
}

template <int BK_, int CG>
static int launch_wgrad(const void* G, int ldg, const void* A, int lda, int M, int N, int K, int n_valid,
                        int k_valid, float* dW, int ldw, float* dbias, void* workspace, OpFmt fmt,
                        cudaStream_t stream) {
  using Cfg = WgCfg<BK_, CG>;
  CUtensorMap tmG, tmA;
  int rc = make_map(&tmG, G, N, M, ldg, 64, kTileK);
  if (rc) return rc;
  rc = make_map(&tmA, A, K, M, lda, 64, kTileK);
  if (rc) return rc;
  NRF_CUDA_OK(cudaFuncSetAttribute(wgrad_tc_kernel<BK_, CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmem));
  // UMMA A operand = G^T (fmt.a_half), B operand = A^T and the ones tile of the bias sums (fmt.b_half)
  const uint32_t idesc = make_idesc(kTileM * CG, BK_, 1, 1, fmt.a_half, fmt.b_half);
  const uint32_t idesc_bias = make_idesc(kTileM * CG, 16, 1, 1, fmt.a_half, fmt.b_half);
  const uint32_t ones_word = fmt.b_half ? 0x3C003C00u : 0x3F803F80u;
  int n_tiles = (n_valid + kTileM - 1) / kTileM;
  if (CG == 2) n_tiles = (n_tiles + 1) / 2 * 2;      // whole CTA pairs
  int k_tiles = (k_valid + BK_ - 1) / BK_;
  int out_tiles = n_tiles * k_tiles;
  int splits = sm_count() / out_tiles;      // one wave: (output tiles) x (sample splits) <= #SMs
  int max_splits = (M + 4 * kTileK - 1) / (4 * kTileK);
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  int m_per = ((M + splits - 1) / splits + kTileK - 1) / kTileK * kTileK;
  splits = (M + m_per - 1) / m_per;
  const int n_pad = n_tiles * kTileM, k_pad = k_tiles * BK_;
  float* ws = reinterpret_cast<float*>(workspace);
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cudaLaunchAttribute attr[1];
  if (CG == 2) {
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
  }
  cfg.gridDim = dim3(out_tiles, splits); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = Cfg::kSmem;
  cfg.stream = stream;
  { LaunchScope ls_(NRF_CAT_WGRAD, stream);
  NRF_CUDA_OK(cudaLaunchKernelEx(&cfg, wgrad_tc_kernel<BK_, CG>, tmG, tmA, M, k_tiles, m_per, n_valid, k_valid, dW, ldw,
                                 dbias, ws, n_pad, k_pad, idesc, idesc_bias, ones_word));
  }
  NRF_LAUNCH_OK();
  if (ws) {
    int64_t work = (int64_t)n_valid * ((k_valid + 3) / 4) + (dbias ? n_valid : 0);
    { LaunchScope ls_(NRF_CAT_WGRAD, stream);
    wgrad_reduce_kernel<<<(unsigned)((work + 255) / 256), 256, 0, stream>>>(ws, splits, n_pad, k_pad, n_valid, k_valid,
                                                                           dW, ldw, dbias, k_tiles);
    }
    NRF_LAUNCH_OK();
  }
  return NRF_OK;
}

int wgrad_tc_launch(const void* G, int ldg, const void* A, int lda, int M, int N, int K, int n_valid,
                    int k_valid, float* dW, int ldw, float* dbias, void* workspace, OpFmt fmt, cudaStream_t stream) {
  NRF_REQUIRE(N % 64 == 0 && K % 64 == 0, NRF_ENOSUP, "wgrad_tc: N=%d, K=%d must be multiples of 64", N, K);
  int rc;
  // measured on B200: kind::f16 with DIFFERENT A / B formats (fp16 x bf16) faults with "illegal instruction"
  NRF_REQUIRE(fmt.a_half == fmt.b_half, NRF_ENOSUP, "wgrad_tc: mixed fp16 / bf16 operands are not executable");
  static const bool one_cta = getenv("NRF_WGRAD_1CTA") != nullptr;
  if (k_valid > 128 && n_valid > kTileM && !one_cta)
    rc = launch_wgrad<256, 2>(G, ldg, A, lda, M, N, K, n_valid, k_valid, dW, ldw, dbias, workspace, fmt, stream);
  else if (k_valid > 128) rc = launch_wgrad<256, 1>(G, ldg, A, lda, M, N, K, n_valid, k_valid, dW, ldw, dbias, workspace, fmt, stream);
  else if (k_valid > 64) rc = launch_wgrad<128, 1>(G, ldg, A, lda, M, N, K, n_valid, k_valid, dW, ldw, dbias, workspace, fmt, stream);
  else rc = launch_wgrad<64, 1>(G, ldg, A, lda, M, N, K, n_valid, k_valid, dW, ldw, dbias, workspace, fmt, stream);
  return rc;
}

}  // namespace nrf
