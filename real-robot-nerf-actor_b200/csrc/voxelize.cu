// Point cloud -> bounded voxel grid (the producer of the volume the PerAct encoder turns into voxel_feat).
//   nrf_voxelize <- voxel_grid_real.py:175-233 VoxelGrid.coords_to_bounding_voxel_grid:
//       index  = clamp(floor((p - (bb_min - res)) / (res + 1e-12)), 0, S+1)     (:187-191; a one-voxel border)
//       values = scatter-mean of [xyz, features, 1] over the points of each voxel (:117-134, :136-155)
//       crop the border (:212), occupancy = (count > 0) (:220-222), voxel index / S inserted before it (:224-226)
// The reference builds a (B*(S+2)^3*(4+F)) index tensor and runs two scatter_add_ (float atomics on the GPU: the sum
// order, hence the bits, change run to run).  Here: counting sort of the points by voxel (integer atomics only), then
// one thread per voxel adds its points in ascending point index -- the order the reference's CPU scatter_add_ uses --
// and writes the 3+F+3+1 output channels once.  Bit-reproducible, no float atomics, no (S+2)^3 intermediate.
#include "common.cuh"
#include "sortscan.cuh"

namespace nrf {

constexpr int kVoxMaxCh = 16;   // 3 + F <= 16

// geom: per scene [shifted_min(3) | denominator(3)], computed by the caller with the reference's own fp32 ops
__global__ void __launch_bounds__(256) voxel_key_kernel(const float* __restrict__ coords, const float* __restrict__ geom,
                                                        int B, int N, int S, int32_t* __restrict__ key,
                                                        int32_t* __restrict__ count) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (int64_t)B * N) return;
  const int b = (int)(e / N);
  const float* g = geom + b * 6;
  int idx[3];
  bool inside = true;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    float f = floorf(__fdiv_rn(__fsub_rn(coords[e * 3 + a], g[a]), g[3 + a]));
    // .int() then min(dims-1) / max(0) (:189-191); clamping the float first avoids the integer overflow
    f = fminf(fmaxf(f, -1.0f), (float)(S + 2));
    int i = (f == f) ? (int)f : 0;
    i = min(max(i, 0), S + 1);
    idx[a] = i;
    inside = inside && i >= 1 && i <= S;          // border voxels are cropped away (:212)
  }
  int32_t k = -1;
  if (inside) {
    k = (int32_t)((((int64_t)b * S + (idx[0] - 1)) * S + (idx[1] - 1)) * S + (idx[2] - 1));
    atomicAdd(count + k, 1);
  }
  key[e] = k;
}

__global__ void __launch_bounds__(256) voxel_reduce_kernel(const float* __restrict__ coords,
                                                           const float* __restrict__ feats, int N, int F, int S,
                                                           const int32_t* __restrict__ offset,
                                                           const int32_t* __restrict__ count,
                                                           const int32_t* __restrict__ list, float* __restrict__ out,
                                                           int64_t T) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  const int nch = 3 + F;
  const int cnt = count[t];
  float acc[kVoxMaxCh];
#pragma unroll
  for (int c = 0; c < kVoxMaxCh; ++c) acc[c] = 0.f;
  if (cnt > 0) {
    int32_t* lst = const_cast<int32_t*>(list) + offset[t];      // this voxel's own segment of the list
    if (cnt > 16) {
      // long lists (a cloud collapsed onto few voxels): heap sort of the segment in place, O(cnt log cnt), instead of
      // the cnt^2 selection loop below (1e4 points in one voxel were 1e8 iterations of one thread)
      auto sift = [&](int root, int end) {
        const int32_t v = lst[root];
        for (int child = 2 * root + 1; child < end; child = 2 * root + 1) {
          if (child + 1 < end && lst[child + 1] > lst[child]) ++child;
          if (lst[child] <= v) break;
          lst[root] = lst[child];
          root = child;
        }
        lst[root] = v;
      };
      for (int i = cnt / 2 - 1; i >= 0; --i) sift(i, cnt);
      for (int end = cnt - 1; end > 0; --end) {
        const int32_t top = lst[0];
        lst[0] = lst[end];
        lst[end] = top;
        sift(0, end);
      }
    }
    int32_t last = -1;
    for (int step = 0; step < cnt; ++step) {          // ascending point index
      int32_t best = 0x7fffffff;
      if (cnt > 16) {
        best = lst[step];
      } else {
        for (int i = 0; i < cnt; ++i) {
          const int32_t e = lst[i];
          if (e > last && e < best) best = e;
        }
      }
      last = best;
      const float* p = coords + (int64_t)best * 3;
      acc[0] = __fadd_rn(acc[0], p[0]);
      acc[1] = __fadd_rn(acc[1], p[1]);
      acc[2] = __fadd_rn(acc[2], p[2]);
      if (feats) {
        const float* q = feats + (int64_t)best * F;
#pragma unroll
        for (int c = 0; c < kVoxMaxCh - 3; ++c)
          if (c < F) acc[3 + c] = __fadd_rn(acc[3 + c], q[c]);
      }
    }
    const float d = (float)cnt;
#pragma unroll
    for (int c = 0; c < kVoxMaxCh; ++c) acc[c] = __fdiv_rn(acc[c], d);   // out.true_divide_(count) (:130)
  }
  float* o = out + t * (nch + 4);
#pragma unroll
  for (int c = 0; c < kVoxMaxCh; ++c)
    if (c < nch) o[c] = acc[c];
  const int k = (int)(t % S), j = (int)((t / S) % S), i = (int)((t / ((int64_t)S * S)) % S);
  const float sd = (float)S;
  o[nch + 0] = __fdiv_rn((float)i, sd);               // index_grid / voxel_d (:224-225)
  o[nch + 1] = __fdiv_rn((float)j, sd);
  o[nch + 2] = __fdiv_rn((float)k, sd);
  o[nch + 3] = cnt > 0 ? 1.0f : 0.0f;                 // occupied (:220)
}

}  // namespace nrf

using namespace nrf;

extern "C" int64_t nrf_voxelize_workspace_bytes(int B, int N, int S) {
  const int64_t T = (int64_t)B * S * S * S, E = (int64_t)B * N;
  const int64_t nb = (T + 1023) / 1024;
  // count[T] cursor[T] offset[T] block_sums[nb+1] key[E] list[E]
  return (3 * T + nb + 1 + 2 * E) * 4 + 1024;
}

extern "C" int nrf_voxelize(const float* coords, const float* feats, int B, int N, int F, const float* geom, int S,
                            float* out, void* workspace, void* stream) {
  NRF_REQUIRE(coords && geom && out && workspace, NRF_EINVAL, "nrf_voxelize: null pointer");
  NRF_REQUIRE(B > 0 && N > 0 && S > 0 && F >= 0, NRF_EINVAL, "nrf_voxelize: bad sizes");
  NRF_REQUIRE(3 + F <= kVoxMaxCh, NRF_ENOSUP, "nrf_voxelize: at most %d feature channels", kVoxMaxCh - 3);
  NRF_REQUIRE(F == 0 || feats, NRF_EINVAL, "nrf_voxelize: features missing");
  const int64_t T = (int64_t)B * S * S * S, E = (int64_t)B * N;
  NRF_REQUIRE(T < ((int64_t)1 << 31) && E < ((int64_t)1 << 31), NRF_ENOSUP, "nrf_voxelize: more than 2^31 voxels or points");
  cudaStream_t s = as_stream(stream);
  const int64_t nb = (T + 1023) / 1024;
  int32_t* count = reinterpret_cast<int32_t*>(workspace);
  int32_t* cursor = count + T;
  int32_t* offset = cursor + T;
  int32_t* block_sums = offset + T;
  int32_t* key = block_sums + nb + 1;
  int32_t* list = key + E;
  NRF_CUDA_OK(cudaMemsetAsync(count, 0, (size_t)(2 * T) * 4, s));       // count and cursor
  { LaunchScope ls_(NRF_CAT_MISC, s);
    voxel_key_kernel<<<(unsigned)((E + 255) / 256), 256, 0, s>>>(coords, geom, B, N, S, key, count); }
  NRF_LAUNCH_OK();
  int rc = exclusive_scan_i32(count, offset, block_sums, T, NRF_CAT_MISC, s);
  if (rc) return rc;
  { LaunchScope ls_(NRF_CAT_MISC, s);
    scatter_fill_kernel<<<(unsigned)((E + 255) / 256), 256, 0, s>>>(key, E, offset, cursor, list); }
  NRF_LAUNCH_OK();
  { LaunchScope ls_(NRF_CAT_MISC, s);
    voxel_reduce_kernel<<<(unsigned)((T + 255) / 256), 256, 0, s>>>(coords, F > 0 ? feats : nullptr, N, F, S, offset,
                                                                    count, list, out, T); }
  NRF_LAUNCH_OK();
  return NRF_OK;
}
