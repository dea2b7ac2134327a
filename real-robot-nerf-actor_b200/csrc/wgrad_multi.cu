// ALL weight gradients of one MLP backward pass in ONE persistent launch (NRF_PREC_BF16 / NRF_PREC_FP16):
//   for every problem p:  dW_p[n, k] += sum_m G_p[m, n] A_p[m, k],   db_p[n] += sum_m G_p[m, n]
// (resnetfc.py:55-64,171-195: the autograd of every nn.Linear of ResnetFC; 13 + n_lin_z problems per pass).
//
// Why one launch: the per-GEMM kernel (wgrad_tc_kernel, gemm_tc.cu) pays a prologue (barrier init, TMEM allocation,
// cluster sync), a cold TMA ring, a drained tensor pipe during the fp32 reductions of its accumulator and a launch gap
// for each of its 30 launches per step; its bias sums are extra N = 16 MMAs that re-read the G^T operand (12 % of the
// tensor-pipe time); it fills 144 of the 148 SMs.  Here 74 CTA pairs walk the problem list:
//   * a pair owns ONE (256 n x 64 ns k) tile of dW over ONE sample split of the problem, then moves to the next problem;
//     the tiles of a split sit on neighbouring pairs and run at the same time, so G and A come from DRAM once and from
//     L2 for the other tiles (as the grid of the per-GEMM kernel arranged it);
//   * the TMA ring never drains: the producer prefetches the next problem's stages while the last MMAs of this one run;
//   * TWO 256-column fp32 accumulators in TMEM: the reductions (red.global.add.v4.f32) of problem i overlap the MMAs of
//     problem i + 1;
//   * the bias sums leave the tensor pipe: four otherwise idle warps add up the G^T slab of a stage from shared memory
//     after the stage's MMAs have retired (tcgen05.commit -> mma_done) and only then hand the slot back to the producer;
//     the k tiles that share a G tile take turns by sample block, as before;
//   * problems that share G and have few k columns ride in one tile: lin_z[0] and lin_in both multiply dL/dx'_0 with
//     columns of the field input - one N = 256 tile [latent 0:64 | latent 64:128 | PE+viewdir | -] instead of two launches
//     that each stream G again.
// Warp roles (384 threads): 0 TMA producer, 1 MMA issuer (leader CTA), 2 TMEM allocator, 4-7 accumulator flush
// (TMEM lane quarter = warp % 4), 8-11 bias sums.
#include "tc_ptx.cuh"

namespace nrf {

constexpr int kWgmThreads = 384;
constexpr int kWgmStages = 6;
constexpr int kWgmStageG = kTileM * kTileK * 2;          // 16 KB: this CTA's 128 n x 64 samples of G^T (two 64-wide slabs)
constexpr int kWgmSlab = kTileK * kTileK * 2;            // 8 KB: one 64 k x 64 samples slab of A^T
constexpr int kWgmStage = kWgmStageG + 2 * kWgmSlab;     // this CTA's half of the (up to four) A^T slabs
constexpr int kWgmSmem = kWgmStages * kWgmStage + 1024 /*align*/ + 512 /*barriers*/;
constexpr int kWgmAccCols = 256;

struct WgmDevProblem {
  int g_map, n_valid, ns, k_tiles;
  int tiles, splits, m_per, bias_mask;       // bias_mask < 0: k tile 0 sums every sample block
  uint32_t idesc;
  int pair0;                                 // first CTA pair of this problem (several short problems share a wave)
  float* dbias;
  float* dbias2;
  const int32_t* blocks;                     // optional list of the 64-sample blocks to visit ([0] = count)
  WgmSlab slab[kWgmMaxSlabs];
};
struct WgmArgs {
  CUtensorMap maps[kWgmMaxMaps];
  WgmDevProblem prob[kWgmMaxProblems];
  int n_prob, M;
  int* sync;   // kWgmSyncStride zeroed counters per problem: the pairs of a sample split start the problem together
  int dbg;     // NRF_WGM_DBG timing experiments (wrong results!): 1 no bias reads, 2 slots recycled on mma_done (no bias
               // warps at all), 4 no reductions of the accumulator
};

struct WgmWork { int np, kt, split, m_begin, num_mb, first; };   // first: position of the split's first block in P.blocks
constexpr int kWgmSyncStride = 128;
// This pair's share of problem p (the same arithmetic in every warp role).
__device__ __forceinline__ bool wgm_work(const WgmArgs& a, int p, int pair, WgmWork& w) {
  const WgmDevProblem& P = a.prob[p];
  const int rel = pair - P.pair0;
  if (rel < 0 || rel >= P.tiles * P.splits) return false;
  const int tile = rel % P.tiles, split = rel / P.tiles;
  w.np = tile / P.k_tiles;
  w.kt = tile - w.np * P.k_tiles;
  w.split = split;
  if (P.blocks) {                              // the splits share the LISTED blocks evenly (every tile of a split alike)
    const int n = __ldg(P.blocks);
    const int per = (n + P.splits - 1) / P.splits;
    w.first = split * per;
    w.m_begin = 0;
    w.num_mb = max(0, min(per, n - w.first));
    return w.num_mb > 0;
  }
  w.first = 0;
  w.m_begin = split * P.m_per;
  const int m_end = min(a.M, w.m_begin + P.m_per);
  w.num_mb = (m_end - w.m_begin + kTileK - 1) / kTileK;
  return w.num_mb > 0;
}

__global__ void __launch_bounds__(kWgmThreads, 1) wgrad_multi_kernel(const __grid_constant__ WgmArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + kWgmStages * kWgmStage);   // leader CTA: both CTAs' loads landed
  uint64_t* mma_done = full + kWgmStages;      // the stage's MMAs have retired (multicast commit: both CTAs)
  uint64_t* slot_free = mma_done + kWgmStages; // ... and the bias warps are done with its G^T slab
  uint64_t* acc_full = slot_free + kWgmStages; // [2]
  uint64_t* acc_empty = acc_full + 2;          // [2] leader CTA: 4 flush warps x 2 CTAs
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = uniform_warp_idx(), lane = threadIdx.x % 32;
  uint32_t crank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
  const bool cta_leader = crank == 0;
  const int pair = blockIdx.x >> 1;

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kWgmStages; ++s) { mbar_init(full + s, 2); mbar_init(mma_done + s, 1); mbar_init(slot_free + s, 4); }
    for (int s = 0; s < 2; ++s) { mbar_init(acc_full + s, 1); mbar_init(acc_empty + s, 8); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc2(tmem_slot, 2 * kWgmAccCols);
  tc_fence_before();
  __syncthreads();
  cluster_sync();   // both CTAs' barriers exist before the peer signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ---------------------------------------------------------------- TMA producer (both CTAs)
    PipeState st;
    for (int p = 0; p < a.n_prob; ++p) {
      WgmWork w;
      if (!wgm_work(a, p, pair, w)) continue;
      const WgmDevProblem& P = a.prob[p];
      const CUtensorMap* mg = &a.maps[P.g_map];
      const int half = P.ns >> 1;                                  // A^T slabs this CTA loads (its half of the MMA's N)
      const WgmSlab* sl = &P.slab[w.kt * P.ns + (int)crank * half];
      const CUtensorMap* ma0 = &a.maps[sl[0].a_map];
      const CUtensorMap* ma1 = &a.maps[sl[half - 1].a_map];
      const int c0 = sl[0].a_col, c1 = sl[half - 1].a_col;
      const uint32_t bytes = 2u * (uint32_t)(kWgmStageG + half * kWgmSlab);
      const int n0 = w.np * 256 + (int)crank * kTileM;
      if (lane == 0) {
        tma_prefetch_desc(mg); tma_prefetch_desc(ma0); tma_prefetch_desc(ma1);
        if (p + 1 < a.n_prob) {                                    // and the next problem's, a whole problem ahead
          const WgmDevProblem& Q = a.prob[p + 1];
          tma_prefetch_desc(&a.maps[Q.g_map]);
          tma_prefetch_desc(&a.maps[Q.slab[0].a_map]);
        }
      }
      if (a.sync && P.tiles > 1 && !(a.dbg & 8)) {
        // The tiles of a sample split read the same G / A rows: started together they take them from DRAM once and
        // from L2 otherwise.  Without this the pairs drift apart over the launch (measured: +16 % DRAM reads).  A
        // performance hint only - the wait is bounded and nothing depends on it.
        int* ctr = a.sync + p * kWgmSyncStride + w.split;
        if (lane == 0) {
          if (cta_leader) atomicAdd(ctr, 1);
          for (int spin = 0; spin < 4096; ++spin) {
            if (*reinterpret_cast<volatile int*>(ctr) >= P.tiles) break;
            __nanosleep(64);
          }
        }
        __syncwarp();
      }
      int blk32 = 0;                                   // listed problems: 32 block indices at a time, one per lane
      for (int mb = 0; mb < w.num_mb; ++mb) {
        if (P.blocks && (mb & 31) == 0) blk32 = mb + lane < w.num_mb ? __ldg(P.blocks + 1 + w.first + mb + lane) : 0;
        mbar_wait(((a.dbg & 2) ? mma_done : slot_free) + st.stage, st.phase ^ 1);
        uint8_t* sg = smem + st.stage * kWgmStage;
        uint8_t* sa = sg + kWgmStageG;
        // rows >= M are zero-filled by TMA; splits are 64-aligned
        const int m0 = P.blocks ? __shfl_sync(0xffffffffu, blk32, mb & 31) * kTileK : w.m_begin + mb * kTileK;
        if (elect_one()) {
          if (cta_leader) mbar_expect_tx(full + st.stage, bytes);
          else mbar_arrive_leader(full + st.stage);
          tma_load_2d_pair(sg, mg, full + st.stage, n0, m0);
          tma_load_2d_pair(sg + kWgmSlab, mg, full + st.stage, n0 + 64, m0);
          tma_load_2d_pair(sa, ma0, full + st.stage, c0, m0);
          if (half == 2) tma_load_2d_pair(sa + kWgmSlab, ma1, full + st.stage, c1, m0);
        }
        __syncwarp();
        st.advance(kWgmStages);
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer (leader CTA)
    if (cta_leader) {
      PipeState st;
      int it = 0;
      for (int p = 0; p < a.n_prob; ++p) {
        WgmWork w;
        if (!wgm_work(a, p, pair, w)) continue;
        const uint32_t idesc = a.prob[p].idesc;
        const int b = it & 1;
        mbar_wait(acc_empty + b, ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + b * kWgmAccCols;
        for (int mb = 0; mb < w.num_mb; ++mb) {
          mbar_wait(full + st.stage, st.phase);
          tc_fence_after();
          const uint32_t sg = smem_u32(smem + st.stage * kWgmStage);
          const uint64_t gdesc = make_sdesc(sg, kWgmSlab, 1024);
          const uint64_t adesc = make_sdesc(sg + kWgmStageG, kWgmSlab, 1024);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < kTileK / kUmmaK; ++k)   // 16 samples = two 8-row atoms = 2048 B: +128 in the address field
              umma_bf16_pair(tmem_d, gdesc + 128 * k, adesc + 128 * k, idesc, (mb | k) != 0);
            umma_commit_pair(mma_done + st.stage);
          }
          __syncwarp();
          st.advance(kWgmStages);
        }
        if (elect_one()) umma_commit_pair(acc_full + b);
        __syncwarp();
        ++it;
      }
    }
  } else if (warp >= 4 && warp < 8) {
    // ---------------------------------------------------------------- accumulator flush (both CTAs)
    const int q = warp & 3;
    int it = 0;
    for (int p = 0; p < a.n_prob; ++p) {
      WgmWork w;
      if (!wgm_work(a, p, pair, w)) continue;
      const WgmDevProblem& P = a.prob[p];
      const int b = it & 1;
      mbar_wait(acc_full + b, (it >> 1) & 1);
      tc_fence_after();
      const int n = w.np * 256 + (int)crank * kTileM + q * 32 + lane;
      const uint32_t taddr = tmem_base + b * kWgmAccCols + ((uint32_t)(q * 32) << 16);
      const WgmSlab* sl = &P.slab[w.kt * P.ns];
      const int n_chunks = 2 * P.ns;
#pragma unroll 1
      for (int c = 0; c < n_chunks; ++c) {
        uint32_t v[32];
        tmem_ld32(taddr + c * 32, v);
        if (c == n_chunks - 1) {                       // this warp is done with the accumulator
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_leader(acc_empty + b);
        }
        const WgmSlab& s = sl[c >> 1];
        const int kv = s.k_valid - (c & 1) * 32;       // valid columns of this 32-column chunk
        if (n < P.n_valid && kv > 0 && !(a.dbg & 4)) {
          float* dst = s.dW + (int64_t)n * s.ldw + (c & 1) * 32;
          if (kv >= 32 && (s.ldw & 3) == 0 && (reinterpret_cast<uintptr_t>(s.dW) & 15) == 0) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)              // 16 B aligned rows: red.global.add.v4.f32
              atomicAdd(reinterpret_cast<float4*>(dst + j),
                        make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                                    __uint_as_float(v[j + 3])));
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < kv) atomicAdd(dst + j, __uint_as_float(v[j]));
          }
        }
      }
      ++it;
    }
  } else if (warp >= 8 && !(a.dbg & 2)) {
    // ---------------------------------------------------------------- bias sums + slot release (both CTAs)
    // thread t reads the 16 B chunk at physical position pc of rows r0, r0 + 8, .. of slab sg (128B swizzle: that is
    // the logical chunk pc ^ r0 of each of those rows, i.e. the same 8 n for all of them)
    const int t = threadIdx.x - 256;
    const int gslab = t >> 6, pc = t & 7, r0 = (t >> 3) & 7;
    const uint32_t off = (uint32_t)(gslab * kWgmSlab + r0 * 128 + pc * 16);
    PipeState st;
    for (int p = 0; p < a.n_prob; ++p) {
      WgmWork w;
      if (!wgm_work(a, p, pair, w)) continue;
      const WgmDevProblem& P = a.prob[p];
      const bool has_bias = P.dbias != nullptr;
      float acc[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = 0.f;
      const int mb0 = w.m_begin / kTileK;
      for (int mb = 0; mb < w.num_mb; ++mb) {
        mbar_wait(mma_done + st.stage, st.phase);
        const bool mine = has_bias && !(a.dbg & 1) && (P.bias_mask < 0 ? w.kt == 0 : ((mb0 + mb) & P.bias_mask) == w.kt);
        if (mine) {
          const uint8_t* src = smem + st.stage * kWgmStage + off;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint4 x = *reinterpret_cast<const uint4*>(src + j * 1024);
            acc[0] += __uint_as_float(x.x << 16); acc[1] += __uint_as_float(x.x & 0xffff0000u);
            acc[2] += __uint_as_float(x.y << 16); acc[3] += __uint_as_float(x.y & 0xffff0000u);
            acc[4] += __uint_as_float(x.z << 16); acc[5] += __uint_as_float(x.z & 0xffff0000u);
            acc[6] += __uint_as_float(x.w << 16); acc[7] += __uint_as_float(x.w & 0xffff0000u);
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(slot_free + st.stage);
        st.advance(kWgmStages);
      }
      if (has_bias) {
        const int n = w.np * 256 + (int)crank * kTileM + gslab * 64 + ((pc ^ r0) << 3);
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (n + i < P.n_valid && acc[i] != 0.f) {
            atomicAdd(P.dbias + n + i, acc[i]);
            if (P.dbias2) atomicAdd(P.dbias2 + n + i, acc[i]);
          }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync();   // neither CTA leaves while the pair's MMAs / barrier traffic may still touch it
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, 2 * kWgmAccCols);
  }
}

// ----------------------------------------------------------------------------------- host side
int wgrad_multi_launch(const WgmHost& h, cudaStream_t stream) {
  NRF_REQUIRE(h.n_maps > 0 && h.n_maps <= kWgmMaxMaps && h.n_prob > 0 && h.n_prob <= kWgmMaxProblems && h.M > 0,
              NRF_EINVAL, "wgrad_multi: %d operands, %d problems, M=%d", h.n_maps, h.n_prob, h.M);
  const int pairs = sm_count() / 2;
  WgmArgs args;
  memset(&args, 0, sizeof(args));
  for (int i = 0; i < h.n_maps; ++i) {
    int rc = make_map(&args.maps[i], h.op[i].base, h.op[i].cols, h.M, h.op[i].ld, 64, kTileK);
    if (rc) return rc;
  }
  for (int p = 0; p < h.n_prob; ++p) {
    const WgmProblem& s = h.prob[p];
    WgmDevProblem& d = args.prob[p];
    NRF_REQUIRE((s.ns == 2 || s.ns == 4) && s.k_tiles >= 1 && s.k_tiles * s.ns <= kWgmMaxSlabs && s.n_valid > 0 &&
                    s.g_map >= 0 && s.g_map < h.n_maps, NRF_EINVAL, "wgrad_multi: problem %d malformed", p);
    d.g_map = s.g_map; d.n_valid = s.n_valid; d.ns = s.ns; d.k_tiles = s.k_tiles; d.dbias = s.dbias; d.dbias2 = s.dbias2;
    NRF_REQUIRE(!s.blocks || !s.dbias, NRF_EINVAL, "wgrad_multi: a block list excludes bias sums (they need every sample)");
    d.blocks = s.blocks;
    for (int j = 0; j < s.k_tiles * s.ns; ++j) {
      NRF_REQUIRE(s.slab[j].a_map >= 0 && s.slab[j].a_map < h.n_maps, NRF_EINVAL, "wgrad_multi: slab operand");
      d.slab[j] = s.slab[j];
    }
    const int n_pairs = (s.n_valid + 255) / 256;
    d.tiles = n_pairs * s.k_tiles;
    NRF_REQUIRE(d.tiles <= pairs, NRF_ENOSUP, "wgrad_multi: %d output tiles > %d CTA pairs", d.tiles, pairs);
    d.bias_mask = (s.k_tiles & (s.k_tiles - 1)) == 0 ? s.k_tiles - 1 : -1;
    d.idesc = make_idesc(2 * kTileM, 64 * s.ns, 1, 1, 0, 0);
  }
  // Sample splits.  Long passes: one problem at a time, its tiles x splits fill the CTA pairs.  Short passes (the
  // reference's own 512-ray chunks): a split of fewer than ~8192 samples costs more in fp32 reductions of its
  // accumulator (256 KB per pair and problem, whatever M is) and in pipeline restarts than it computes, so up to `g`
  // consecutive problems of the same shape run side by side on disjoint pairs, each with 1/g of the splits.
  const int want = (h.M + 8191) / 8192;                        // splits a problem can keep busy
  for (int p = 0; p < h.n_prob;) {
    const int T = args.prob[p].tiles;
    int gmax = 1;
    while (p + gmax < h.n_prob && args.prob[p + gmax].tiles == T && args.prob[p + gmax].ns == args.prob[p].ns) ++gmax;
    int g = 1;
    for (int c = 1; c <= gmax; ++c) {                          // the most problems per wave that still leave >= `want`
      const int sp = pairs / (c * T);                          // splits each and idle <= 5 % of the pairs
      if (sp >= 1 && sp >= (want < pairs / T ? want : pairs / T) && c * T * sp * 20 >= pairs * 19) g = c;
    }
    int splits = pairs / (g * T);
    const int max_splits = (h.M + 4 * kTileK - 1) / (4 * kTileK);
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    const int m_per = ((h.M + splits - 1) / splits + kTileK - 1) / kTileK * kTileK;
    splits = (h.M + m_per - 1) / m_per;
    for (int j = 0; j < g; ++j) {
      WgmDevProblem& d = args.prob[p + j];
      d.m_per = m_per; d.splits = splits; d.pair0 = j * T * splits;
    }
    p += g;
  }
  args.n_prob = h.n_prob; args.M = h.M;
  args.sync = h.sync;
  if (h.sync) NRF_CUDA_OK(cudaMemsetAsync(h.sync, 0, (size_t)h.n_prob * kWgmSyncStride * sizeof(int), stream));
  { const char* e = getenv("NRF_WGM_DBG"); args.dbg = e ? atoi(e) : 0; }
  NRF_CUDA_OK(cudaFuncSetAttribute(wgrad_multi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgmSmem));
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  cfg.gridDim = dim3(pairs * 2); cfg.blockDim = dim3(kWgmThreads); cfg.dynamicSmemBytes = kWgmSmem;
  cfg.stream = stream;
  { LaunchScope ls_(NRF_CAT_WGRAD, stream);
  NRF_CUDA_OK(cudaLaunchKernelEx(&cfg, wgrad_multi_kernel, args));
  }
  NRF_LAUNCH_OK();
  return NRF_OK;
}

}  // namespace nrf
