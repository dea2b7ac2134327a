// Voxel rows of a volume gradient <-> a compact (n, C) list: the two device-side steps of the sparse exchange of
// dL/dvoxel_feat between ranks that render different rays of ONE scene (BASELINE config 5; the reference
// has no multi-GPU path - precedent: featurenerf_robo/featurenerf/src/render/nerf_embed.py:412-429 `bind_parallel`).
//
//   nrf_rows_gather : rows[i, :] = grad[voxel idx[i], :]
//   nrf_rows_update : grad[voxel idx[i], :] (+)= rows[i, :]      (rows == NULL: the rows are cleared)
//
// idx: ascending, unique flat indices `scene * V + voxel` (what `counts > 0` of the scatter's per-voxel entry counts
// enumerates), so an update has no collisions and needs no atomics; the caller applies the ranks' lists one launch
// after the other in rank order, which makes the sum bit-identical on every rank.
// grad is (SB, C, V) (channels_first) or (SB, V, C).  Channel-first is the hard case: a voxel's C values lie V floats
// apart, so every element costs a 32 B sector.  A block stages 32 list entries x C channels in shared memory and walks
// the volume with LANES OVER ENTRIES for one channel at a time: neighbouring voxels of a ray's footprint (x-adjacent
// corner pairs are consecutive indices) fall into the same sectors and are merged by the coalescer, while the row list
// is read / written as whole 512 B rows.
#include <string.h>
#include "common.cuh"

namespace nrf {

constexpr int kRowsTile = 32;

template <bool kGather>
__global__ void __launch_bounds__(128) rows_cf_kernel(float* __restrict__ grad, const int64_t* __restrict__ idx,
                                                      int64_t n, float* __restrict__ rows, int C, int64_t V,
                                                      int add) {
  extern __shared__ float tile[];                       // [kRowsTile][C + 1]
  const int ldt = C + 1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int64_t i0 = (int64_t)blockIdx.x * kRowsTile; i0 < n; i0 += (int64_t)gridDim.x * kRowsTile) {
    const int cnt = (int)min((int64_t)kRowsTile, n - i0);
    int64_t base = -1;                                  // element offset of (scene, channel 0, voxel) of this lane's entry
    if (lane < cnt) {
      const int64_t f = idx[i0 + lane];
      const int64_t scene = f / V;
      base = scene * C * V + (f - scene * V);
    }
    if (!kGather) {
      if (rows) {
        for (int e = warp; e < cnt; e += 4)             // whole rows in, 16 B per lane
          for (int c = lane * 4; c < C; c += 128) {
            const float4 v = *reinterpret_cast<const float4*>(rows + (i0 + e) * C + c);
            float* t = tile + e * ldt + c;
            t[0] = v.x; t[1] = v.y; t[2] = v.z; t[3] = v.w;
          }
        __syncthreads();
      }
      if (base >= 0)
        for (int c = warp; c < C; c += 4) {
          float* g = grad + base + (int64_t)c * V;
          if (!rows) *g = 0.0f;
          else if (add) *g = *g + tile[lane * ldt + c];
          else *g = tile[lane * ldt + c];
        }
      __syncthreads();
    } else {
      if (base >= 0)
        for (int c = warp; c < C; c += 4) tile[lane * ldt + c] = grad[base + (int64_t)c * V];
      __syncthreads();
      for (int e = warp; e < cnt; e += 4)
        for (int c = lane * 4; c < C; c += 128) {
          const float* t = tile + e * ldt + c;
          *reinterpret_cast<float4*>(rows + (i0 + e) * C + c) = make_float4(t[0], t[1], t[2], t[3]);
        }
      __syncthreads();
    }
  }
}

// channels-last: a row is C contiguous floats; one warp per entry
template <bool kGather>
__global__ void __launch_bounds__(256) rows_cl_kernel(float* __restrict__ grad, const int64_t* __restrict__ idx,
                                                      int64_t n, float* __restrict__ rows, int C, int add) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < n; i += warps) {
    float* g = grad + idx[i] * C;
    float* r = rows ? rows + i * C : nullptr;
    for (int c = lane * 4; c < C; c += 128) {
      float4* gp = reinterpret_cast<float4*>(g + c);
      if (kGather) { *reinterpret_cast<float4*>(r + c) = *gp; continue; }
      if (!r) { *gp = make_float4(0.f, 0.f, 0.f, 0.f); continue; }
      float4 v = *reinterpret_cast<const float4*>(r + c);
      if (add) { const float4 o = *gp; v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w; }
      *gp = v;
    }
  }
}

// ---------------------------------------------------------------------------------------------- merge
// grad[voxel] = sum over ranks r = 0, 1, .. (in that order) of rank r's row for that voxel, for every voxel that any
// rank lists; voxels nobody lists keep their value (zero: the local scatter wrote them).  Lists: rows (world, cap, C),
// idx (world, cap) ascending within a rank, cnt[r] valid entries.  A block owns a CONTIGUOUS range of 32-voxel tiles, so
// its cursor into every rank's list only moves forward (one binary search per rank and block, then one coalesced
// 32-index probe per rank and tile); a tile is summed in shared memory and written once, channel rows of 128 B - no
// read-modify-write of the volume and none of the 8x sector amplification of per-rank updates in the (C, V) layout.
constexpr int kMergeMaxWorld = 64;
struct MergeArgs {
  const float* rows; const int64_t* idx; int64_t cap;
  int64_t cnt[kMergeMaxWorld];
  int world, C, channels_first;
  int whole_tiles;           // unlisted voxels are known to be zero: write whole 32-voxel tiles (full 32 B sectors)
  int64_t V, total;          // voxels per scene, SB * V
};

__global__ void __launch_bounds__(128) rows_merge_kernel(float* __restrict__ grad, const __grid_constant__ MergeArgs a) {
  extern __shared__ float acc[];                        // [32][C + 1]
  __shared__ int64_t cur[kMergeMaxWorld];               // this block's cursor into rank r's list
  __shared__ unsigned listed;                           // voxels of the current tile that some rank lists
  __shared__ int n_in[kMergeMaxWorld];                  // entries of rank r inside the current tile
  const int C = a.C, ldt = C + 1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t n_tiles = (a.total + 31) / 32;
  const int64_t per = (n_tiles + gridDim.x - 1) / gridDim.x;
  const int64_t t_begin = (int64_t)blockIdx.x * per, t_end = min(n_tiles, t_begin + per);
  if (t_begin >= t_end) return;
  for (int r = threadIdx.x; r < a.world; r += blockDim.x) {      // lower_bound(idx_r, first voxel of the range)
    const int64_t* ix = a.idx + (int64_t)r * a.cap;
    int64_t lo = 0, hi = a.cnt[r];
    const int64_t key = t_begin * 32;
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      if (ix[mid] < key) lo = mid + 1; else hi = mid;
    }
    cur[r] = lo;
  }
  __syncthreads();
  for (int64_t t = t_begin; t < t_end; ++t) {
    const int64_t f0 = t * 32;
    int mine = 0;
    for (int r = warp; r < a.world; r += 4) {            // how many of rank r's next entries fall into this tile (<= 32)
      const int64_t p = cur[r] + lane;
      const bool in = p < a.cnt[r] && a.idx[(int64_t)r * a.cap + p] < f0 + 32;
      const int k = __popc(__ballot_sync(0xffffffffu, in));
      if (lane == 0) n_in[r] = k;
      mine |= k;
    }
    if (!__syncthreads_or(mine)) continue;                // nobody lists a voxel of this tile (n_in is all zero: the
                                                          // next tile's counts may overwrite it right away)
    for (int i = threadIdx.x; i < 32 * ldt; i += blockDim.x) acc[i] = 0.0f;
    if (threadIdx.x == 0) listed = 0u;
    __syncthreads();
    for (int r = 0; r < a.world; ++r) {                   // rank order: the same bits on every rank
      const int k = n_in[r];
      if (k == 0) continue;
      const int64_t p0 = cur[r];
      for (int e = warp; e < k; e += 4) {                 // distinct voxels within a rank: no two warps share a row
        const int64_t p = (int64_t)r * a.cap + p0 + e;
        const int v = (int)(a.idx[p] - f0);
        if (lane == 0) atomicOr(&listed, 1u << v);
        for (int c = lane * 4; c < C; c += 128) {
          const float4 x = *reinterpret_cast<const float4*>(a.rows + p * C + c);
          float* d = acc + v * ldt + c;
          d[0] += x.x; d[1] += x.y; d[2] += x.z; d[3] += x.w;
        }
      }
      __syncthreads();
    }
    if (a.channels_first) {
      const int64_t f = f0 + lane;
      if (f < a.total && (a.whole_tiles || ((listed >> lane) & 1u))) {   // voxels nobody lists keep their value
        const int64_t scene = f / a.V;
        float* g = grad + scene * C * a.V + (f - scene * a.V);
        for (int c = warp; c < C; c += 4) g[(int64_t)c * a.V] = acc[lane * ldt + c];
      }
    } else {
      for (int v = warp; v < 32 && f0 + v < a.total; v += 4)
        if (a.whole_tiles || ((listed >> v) & 1u))
          for (int c = lane; c < C; c += 32) grad[(f0 + v) * C + c] = acc[v * ldt + c];
    }
    __syncthreads();
    if (threadIdx.x < a.world) cur[threadIdx.x] += n_in[threadIdx.x];
    for (int r = threadIdx.x + blockDim.x; r < a.world; r += blockDim.x) cur[r] += n_in[r];
    __syncthreads();
  }
}

template <bool kGather>
static int rows_launch(float* grad, int channels_first, int C, int64_t V, const int64_t* idx, int64_t n, float* rows,
                       int add, cudaStream_t s) {
  if (n == 0) return NRF_OK;
  NRF_REQUIRE(grad && idx && n > 0 && C > 0 && C % 4 == 0 && V > 0, NRF_EINVAL, "nrf_rows_*: bad arguments (C=%d)", C);
  NRF_REQUIRE((reinterpret_cast<uintptr_t>(rows) & 15) == 0 && (reinterpret_cast<uintptr_t>(grad) & 15) == 0, NRF_EINVAL,
              "nrf_rows_*: buffers must be 16 B aligned");
  LaunchScope ls_(NRF_CAT_SCATTER, s);
  if (channels_first) {
    const size_t smem = (size_t)kRowsTile * (C + 1) * sizeof(float);
    NRF_REQUIRE(smem <= 48 * 1024, NRF_ENOSUP, "nrf_rows_*: C=%d too wide for the channel-first tile", C);
    const int64_t tiles = (n + kRowsTile - 1) / kRowsTile;
    const int grid = (int)min(tiles, (int64_t)sm_count() * 16);
    rows_cf_kernel<kGather><<<grid, 128, smem, s>>>(grad, idx, n, rows, C, V, add);
  } else {
    const int grid = (int)min((n + 7) / 8, (int64_t)sm_count() * 8);
    rows_cl_kernel<kGather><<<grid, 256, 0, s>>>(grad, idx, n, rows, C, add);
  }
  NRF_LAUNCH_OK();
  return NRF_OK;
}

}  // namespace nrf

using namespace nrf;

extern "C" int nrf_rows_gather(const float* grad, int channels_first, int C, int64_t V, const int64_t* idx, int64_t n,
                               float* rows, void* stream) {
  NRF_REQUIRE(rows || n == 0, NRF_EINVAL, "nrf_rows_gather: rows is NULL");
  return rows_launch<true>(const_cast<float*>(grad), channels_first, C, V, idx, n, rows, 0, as_stream(stream));
}

extern "C" int nrf_rows_update(float* grad, int channels_first, int C, int64_t V, const int64_t* idx, int64_t n,
                               const float* rows, int add, void* stream) {
  return rows_launch<false>(grad, channels_first, C, V, idx, n, const_cast<float*>(rows), add, as_stream(stream));
}

extern "C" int nrf_rows_merge(float* grad, int channels_first, int C, int64_t V, int SB, const float* rows,
                              const int64_t* idx, int64_t cap, const int64_t* counts_host, int world,
                              int unlisted_are_zero, void* stream) {
  NRF_REQUIRE(grad && rows && idx && counts_host && C > 0 && C % 4 == 0 && V > 0 && SB > 0 && cap > 0, NRF_EINVAL,
              "nrf_rows_merge: bad arguments");
  NRF_REQUIRE(world >= 1 && world <= kMergeMaxWorld, NRF_ENOSUP, "nrf_rows_merge: world=%d > %d", world, kMergeMaxWorld);
  NRF_REQUIRE((reinterpret_cast<uintptr_t>(rows) & 15) == 0, NRF_EINVAL, "nrf_rows_merge: rows must be 16 B aligned");
  MergeArgs a;
  memset(&a, 0, sizeof(a));
  a.rows = rows; a.idx = idx; a.cap = cap; a.world = world; a.C = C; a.channels_first = channels_first;
  a.V = V; a.total = (int64_t)SB * V; a.whole_tiles = unlisted_are_zero != 0;
  int64_t n = 0;
  for (int r = 0; r < world; ++r) {
    NRF_REQUIRE(counts_host[r] >= 0 && counts_host[r] <= cap, NRF_EINVAL, "nrf_rows_merge: counts[%d] out of range", r);
    a.cnt[r] = counts_host[r];
    n += counts_host[r];
  }
  if (n == 0) return NRF_OK;
  const size_t smem = (size_t)32 * (C + 1) * sizeof(float);
  NRF_REQUIRE(smem <= 40 * 1024, NRF_ENOSUP, "nrf_rows_merge: C=%d too wide", C);
  const int64_t n_tiles = (a.total + 31) / 32;
  const int grid = (int)min(n_tiles, (int64_t)sm_count() * 16);
  cudaStream_t s = as_stream(stream);
  LaunchScope ls_(NRF_CAT_SCATTER, s);
  rows_merge_kernel<<<grid, 128, smem, s>>>(grad, a);
  NRF_LAUNCH_OK();
  return NRF_OK;
}
