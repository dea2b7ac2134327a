// Field-input producer and its transpose (volume-gradient scatter), plus the volume re-layout.
//   nrf_encode_points        <- neural_rendering.py:246-283, models_embed.py:185-203,259-277,366,405,
//                               utils.py:545-557
//   nrf_scatter_volume_grad  <- autograd of F.grid_sample at models_embed.py:275
//   nrf_volume_to_channels_* <- layout change so one trilinear corner is one contiguous C-vector
#include "common.cuh"

namespace nrf {

// (SB, C, V) -> (SB, V, C) and back: 32x32 tiles through padded shared memory; both sides coalesced.
template <bool kToLast>
__global__ void volume_transpose_kernel(const float* __restrict__ src, float* __restrict__ dst, int C,
                                        int64_t V) {
  __shared__ float tile[32][33];
  int b = blockIdx.z;
  int64_t v0 = (int64_t)blockIdx.x * 32;
  int c0 = blockIdx.y * 32;
  const float* s = src + (int64_t)b * C * V;
  float* d = dst + (int64_t)b * C * V;
  int tx = threadIdx.x, ty = threadIdx.y;   // 32 x 8
  if (kToLast) {
#pragma unroll
    for (int i = 0; i < 32; i += 8) {
      int c = c0 + ty + i;
      int64_t v = v0 + tx;
      if (c < C && v < V) tile[ty + i][tx] = s[(int64_t)c * V + v];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 32; i += 8) {
      int64_t v = v0 + ty + i;
      int c = c0 + tx;
      if (c < C && v < V) d[v * C + c] = tile[tx][ty + i];
    }
  } else {
#pragma unroll
    for (int i = 0; i < 32; i += 8) {
      int64_t v = v0 + ty + i;
      int c = c0 + tx;
      if (c < C && v < V) tile[ty + i][tx] = s[v * C + c];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 32; i += 8) {
      int c = c0 + ty + i;
      int64_t v = v0 + tx;
      if (c < C && v < V) d[(int64_t)c * V + v] = tile[tx][ty + i];
    }
  }
}

struct EncodeArgs {
  const float* rays;
  const float* z;
  int R, K, rays_per_scene;
  const float* vol;
  int SB, C, S0, S1, S2;
  float bmin[3], bext[3];
  int num_freqs;
  float freq_factor;
  void* out;
  int ld_out;
  float* points;
};

template <typename T> __device__ __forceinline__ T to_out(float v);
template <> __device__ __forceinline__ float to_out<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 to_out<__nv_bfloat16>(float v) {
  return __float2bfloat16_rn(v);
}

template <typename T>
__device__ __forceinline__ void store4(T* p, float4 v);
template <> __device__ __forceinline__ void store4<float>(float* p, float4 v) {
  *reinterpret_cast<float4*>(p) = v;
}
template <> __device__ __forceinline__ void store4<__nv_bfloat16>(__nv_bfloat16* p, float4 v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
  uint2 u;
  u.x = *reinterpret_cast<uint32_t*>(&a);
  u.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = u;
}

// One warp per sample, grid-stride.  A lane owns 4 consecutive channels of each 128-channel slab, so
// every trilinear corner is one coalesced 512 B read.  Accumulation order and rounding follow ATen's
// CPU grid_sampler_3d (out += v*w, product and sum rounded separately) -> bit-identical latents.
template <typename T>
__global__ void __launch_bounds__(256) encode_points_kernel(EncodeArgs a) {
  int lane = threadIdx.x % kWarp;
  int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / kWarp;
  int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) / kWarp;
  int64_t N = (int64_t)a.R * a.K;
  T* out = reinterpret_cast<T*>(a.out);
  const int C = a.C;
  const int n_pe = 3 + 6 * a.num_freqs;
  for (int64_t n = warp; n < N; n += nwarps) {
    int r = (int)(n / a.K);
    int scene = r / a.rays_per_scene;
    const float* ray = a.rays + (int64_t)r * 8;
    float zv = a.z[n];
    SampleGeom g = sample_geometry(ray, zv, a.bmin, a.bext);
    Corner8 c8;
    trilinear_corners(g.cx, g.cy, g.cz, a.S0, a.S1, a.S2, C, c8);
    const float* vol = a.vol + (int64_t)scene * a.S0 * a.S1 * a.S2 * C;
    T* row = out + n * a.ld_out;
    for (int c0 = lane * 4; c0 < C; c0 += kWarp * 4) {
      float4 v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k)
        v[k] = c8.off[k] >= 0 ? __ldg(reinterpret_cast<const float4*>(vol + c8.off[k] + c0))
                              : make_float4(0.f, 0.f, 0.f, 0.f);
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (c8.off[k] >= 0) {
          acc.x = __fadd_rn(acc.x, __fmul_rn(v[k].x, c8.w[k]));
          acc.y = __fadd_rn(acc.y, __fmul_rn(v[k].y, c8.w[k]));
          acc.z = __fadd_rn(acc.z, __fmul_rn(v[k].z, c8.w[k]));
          acc.w = __fadd_rn(acc.w, __fmul_rn(v[k].w, c8.w[k]));
        }
      }
      store4<T>(row + c0, acc);
    }
    // positional encoding [x y z | per frequency: sin(xyz), cos(xyz)] then the view direction
    float cxyz[3] = {g.cx, g.cy, g.cz};
    for (int e = lane; e < a.ld_out - C; e += kWarp) {
      float val = 0.f;
      if (e < 3) {
        val = cxyz[e];
      } else if (e < n_pe) {
        int q = e - 3;
        int f = q / 6, w = q % 6;
        float freq = a.freq_factor * (float)(1 << f);
        float phase = (w >= 3) ? 1.57079637050628662109375f : 0.0f;   // fp32(pi/2), utils.py:542
        val = sinf(__fmaf_rn(cxyz[w % 3], freq, phase));   // addcmul contracts to an FMA in ATen
      } else if (e < n_pe + 3) {
        val = ray[3 + (e - n_pe)];
      }
      row[C + e] = to_out<T>(val);
    }
    if (a.points && lane < 3) a.points[n * 3 + lane] = lane == 0 ? g.px : (lane == 1 ? g.py : g.pz);
  }
}

struct ScatterArgs {
  const float* rays;
  const float* z;
  int R, K, rays_per_scene;
  const float* dlatent;
  int ld;
  float* grad;
  int SB, C, S0, S1, S2;
  float bmin[3], bext[3];
};

// Transpose of the gather: grad[corner] += w * dlatent.  One warp per sample; vector fp32 reductions
// into the channels-last gradient volume (red.global.add.v4.f32, sm_90+).
__global__ void __launch_bounds__(256) scatter_volume_grad_kernel(ScatterArgs a) {
  int lane = threadIdx.x % kWarp;
  int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / kWarp;
  int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) / kWarp;
  int64_t N = (int64_t)a.R * a.K;
  const int C = a.C;
  for (int64_t n = warp; n < N; n += nwarps) {
    int r = (int)(n / a.K);
    int scene = r / a.rays_per_scene;
    SampleGeom g = sample_geometry(a.rays + (int64_t)r * 8, a.z[n], a.bmin, a.bext);
    Corner8 c8;
    trilinear_corners(g.cx, g.cy, g.cz, a.S0, a.S1, a.S2, C, c8);
    float* grad = a.grad + (int64_t)scene * a.S0 * a.S1 * a.S2 * C;
    const float* dl = a.dlatent + n * a.ld;
    for (int c0 = lane * 4; c0 < C; c0 += kWarp * 4) {
      float4 d = *reinterpret_cast<const float4*>(dl + c0);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (c8.off[k] >= 0) {
          float w = c8.w[k];
          float4 v = make_float4(d.x * w, d.y * w, d.z * w, d.w * w);
          atomicAdd(reinterpret_cast<float4*>(grad + c8.off[k] + c0), v);
        }
      }
    }
  }
}

static int fill_bounds(const float* bounds_host, float bmin[3], float bext[3]) {
  for (int i = 0; i < 3; ++i) {
    bmin[i] = bounds_host[i];
    // bb_max - bb_min is an fp32 tensor subtraction in the reference (models_embed.py:201)
    volatile float e = bounds_host[3 + i] - bounds_host[i];
    bext[i] = e;
  }
  return 0;
}

}  // namespace nrf

using namespace nrf;

static int volume_transpose(const float* src, float* dst, int SB, int C, int64_t V, bool to_last,
                            void* stream) {
  NRF_REQUIRE(src && dst && SB > 0 && C > 0 && V > 0, NRF_EINVAL, "volume transpose: bad args");
  NRF_REQUIRE(SB <= 65535 && (C + 31) / 32 <= 65535, NRF_ENOSUP, "volume transpose: grid too large");
  dim3 block(32, 8);
  dim3 grid((unsigned)((V + 31) / 32), (unsigned)((C + 31) / 32), (unsigned)SB);
  if (to_last)
    { LaunchScope ls_(NRF_CAT_TRANSPOSE, as_stream(stream));
    volume_transpose_kernel<true><<<grid, block, 0, as_stream(stream)>>>(src, dst, C, V);
    }
  else
    { LaunchScope ls_(NRF_CAT_TRANSPOSE, as_stream(stream));
    volume_transpose_kernel<false><<<grid, block, 0, as_stream(stream)>>>(src, dst, C, V);
    }
  NRF_LAUNCH_OK();
  return NRF_OK;
}

extern "C" int nrf_volume_to_channels_last(const float* src, float* dst, int SB, int C, int64_t V,
                                           void* stream) {
  return volume_transpose(src, dst, SB, C, V, true, stream);
}
extern "C" int nrf_volume_to_channels_first(const float* src, float* dst, int SB, int C, int64_t V,
                                            void* stream) {
  return volume_transpose(src, dst, SB, C, V, false, stream);
}

extern "C" int nrf_encode_points(const float* rays, const float* z, int R, int K, int rays_per_scene,
                                 const float* vol_cl, int SB, int C, int S0, int S1, int S2,
                                 const float* bounds_host, int num_freqs, float freq_factor, void* out,
                                 int ld_out, int out_bf16, float* points_out, void* stream) {
  NRF_REQUIRE(rays && z && vol_cl && bounds_host && out, NRF_EINVAL, "nrf_encode_points: null pointer");
  NRF_REQUIRE(R > 0 && K > 0 && rays_per_scene > 0 && R == SB * rays_per_scene, NRF_EINVAL,
              "nrf_encode_points: R=%d must equal SB*rays_per_scene=%d*%d", R, SB, rays_per_scene);
  NRF_REQUIRE(C % 4 == 0 && C > 0, NRF_ENOSUP, "nrf_encode_points: C=%d must be a multiple of 4", C);
  NRF_REQUIRE(ld_out >= C + 6 + 6 * num_freqs && ld_out % 4 == 0, NRF_EINVAL,
              "nrf_encode_points: ld_out=%d too small / unaligned", ld_out);
  EncodeArgs a;
  a.rays = rays; a.z = z; a.R = R; a.K = K; a.rays_per_scene = rays_per_scene;
  a.vol = vol_cl; a.SB = SB; a.C = C; a.S0 = S0; a.S1 = S1; a.S2 = S2;
  fill_bounds(bounds_host, a.bmin, a.bext);
  a.num_freqs = num_freqs; a.freq_factor = freq_factor;
  a.out = out; a.ld_out = ld_out; a.points = points_out;
  int64_t N = (int64_t)R * K;
  int threads = 256;
  int64_t want = (N + 7) / 8;
  int max_blocks = sm_count() * 16;
  int blocks = (int)(want < max_blocks ? want : max_blocks);
  if (out_bf16)
    { LaunchScope ls_(NRF_CAT_ENCODE, as_stream(stream));
    encode_points_kernel<__nv_bfloat16><<<blocks, threads, 0, as_stream(stream)>>>(a);
    }
  else
    { LaunchScope ls_(NRF_CAT_ENCODE, as_stream(stream));
    encode_points_kernel<float><<<blocks, threads, 0, as_stream(stream)>>>(a);
    }
  NRF_LAUNCH_OK();
  return NRF_OK;
}

extern "C" int nrf_scatter_volume_grad(const float* rays, const float* z, int R, int K,
                                       int rays_per_scene, const float* dlatent, int ld, float* grad_cl,
                                       int SB, int C, int S0, int S1, int S2, const float* bounds_host,
                                       void* stream) {
  NRF_REQUIRE(rays && z && dlatent && grad_cl && bounds_host, NRF_EINVAL,
              "nrf_scatter_volume_grad: null pointer");
  NRF_REQUIRE(R > 0 && K > 0 && R == SB * rays_per_scene, NRF_EINVAL,
              "nrf_scatter_volume_grad: R != SB*rays_per_scene");
  NRF_REQUIRE(C % 4 == 0 && ld % 4 == 0 && ld >= C, NRF_ENOSUP, "nrf_scatter_volume_grad: C/ld alignment");
  ScatterArgs a;
  a.rays = rays; a.z = z; a.R = R; a.K = K; a.rays_per_scene = rays_per_scene;
  a.dlatent = dlatent; a.ld = ld; a.grad = grad_cl;
  a.SB = SB; a.C = C; a.S0 = S0; a.S1 = S1; a.S2 = S2;
  fill_bounds(bounds_host, a.bmin, a.bext);
  int64_t N = (int64_t)R * K;
  int threads = 256;
  int64_t want = (N + 7) / 8;
  int max_blocks = sm_count() * 16;
  int blocks = (int)(want < max_blocks ? want : max_blocks);
  { LaunchScope ls_(NRF_CAT_SCATTER, as_stream(stream));
  scatter_volume_grad_kernel<<<blocks, threads, 0, as_stream(stream)>>>(a);
  }
  NRF_LAUNCH_OK();
  return NRF_OK;
}

// ------------------------------------------------------------------------------------------------
// Atomics-free, deterministic volume-gradient scatter (counting sort by voxel + segmented reduce).
//   1. scatter_count : per (sample, corner) entry e = 8n + c: key[e] = scene*V + voxel (or -1), w[e] = corner
//                      weight, count[key]++            (integer atomics only: their result is order-free)
//   2. exclusive scan of count -> offset               (three small kernels)
//   3. scatter_fill  : list[offset[key] + cursor[key]++] = e
//   4. scatter_reduce: one warp per voxel: its entries are visited in ascending e (selection by warp-min, the
//                      lists are short), grad[voxel,:] (+)= sum_e w[e] * dlatent[e/8,:] -- every voxel row is
//                      written by exactly one warp, in a fixed order: no float atomics, bit-reproducible.
namespace nrf {

__global__ void __launch_bounds__(256) scatter_count_kernel(ScatterArgs a, int64_t V, int32_t* __restrict__ key,
                                                            float* __restrict__ wts, int32_t* __restrict__ count) {
  int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t N = (int64_t)a.R * a.K;
  if (n >= N) return;
  int r = (int)(n / a.K);
  int scene = r / a.rays_per_scene;
  SampleGeom g = sample_geometry(a.rays + (int64_t)r * 8, a.z[n], a.bmin, a.bext);
  Corner8 c8;
  trilinear_corners(g.cx, g.cy, g.cz, a.S0, a.S1, a.S2, 1, c8);      // C = 1: off = voxel index
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    int32_t kk = -1;
    if (c8.off[k] >= 0) {
      kk = (int32_t)((int64_t)scene * V + c8.off[k]);
      atomicAdd(count + kk, 1);
    }
    key[n * 8 + k] = kk;
    wts[n * 8 + k] = c8.w[k];
  }
}

// exclusive scan, 1024 elements per block
__global__ void __launch_bounds__(256) scan_block_kernel(const int32_t* __restrict__ in, int32_t* __restrict__ out,
                                                         int32_t* __restrict__ block_sums, int64_t T) {
  __shared__ int32_t warp_tot[8];
  int64_t base = (int64_t)blockIdx.x * 1024 + threadIdx.x * 4;
  int32_t v[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] = base + i < T ? in[base + i] : 0;
  int32_t tsum = v[0] + v[1] + v[2] + v[3];
  int lane = threadIdx.x % 32, wid = threadIdx.x / 32;
  int32_t incl = tsum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int32_t t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) warp_tot[wid] = incl;
  __syncthreads();
  int32_t woff = 0;
  for (int w = 0; w < wid; ++w) woff += warp_tot[w];
  int32_t excl = woff + incl - tsum;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (base + i < T) out[base + i] = excl;
    excl += v[i];
  }
  if (threadIdx.x == 255) block_sums[blockIdx.x] = woff + incl;
}

__global__ void __launch_bounds__(1024) scan_sums_kernel(int32_t* __restrict__ block_sums, int nb,
                                                         int32_t* __restrict__ total) {
  // single block: serial over chunks of 1024 with a running carry
  __shared__ int32_t s[1024];
  __shared__ int32_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int b0 = 0; b0 < nb; b0 += 1024) {
    int i = b0 + threadIdx.x;
    int32_t v = i < nb ? block_sums[i] : 0;
    s[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      int32_t t = threadIdx.x >= o ? s[threadIdx.x - o] : 0;
      __syncthreads();
      s[threadIdx.x] += t;
      __syncthreads();
    }
    int32_t c = carry;
    if (i < nb) block_sums[i] = c + s[threadIdx.x] - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry = c + s[1023];
    __syncthreads();
  }
  if (threadIdx.x == 0 && total) *total = carry;
}

__global__ void __launch_bounds__(256) scan_add_kernel(int32_t* __restrict__ out, const int32_t* __restrict__ block_sums,
                                                       int64_t T) {
  int64_t base = (int64_t)blockIdx.x * 1024 + threadIdx.x * 4;
  int32_t add = block_sums[blockIdx.x];
#pragma unroll
  for (int i = 0; i < 4; ++i)
    if (base + i < T) out[base + i] += add;
}

__global__ void __launch_bounds__(256) scatter_fill_kernel(const int32_t* __restrict__ key, int64_t E,
                                                           const int32_t* __restrict__ offset,
                                                           int32_t* __restrict__ cursor, int32_t* __restrict__ list) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  int32_t k = key[e];
  if (k < 0) return;
  int32_t pos = offset[k] + atomicAdd(cursor + k, 1);
  list[pos] = (int32_t)e;
}

__global__ void __launch_bounds__(256) scatter_reduce_kernel(const int32_t* __restrict__ offset,
                                                             const int32_t* __restrict__ count,
                                                             const int32_t* __restrict__ list,
                                                             const float* __restrict__ wts,
                                                             const float* __restrict__ dlatent, int ld,
                                                             float* __restrict__ grad, int C, int64_t T, int accumulate) {
  int lane = threadIdx.x % kWarp;
  int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / kWarp;
  int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) / kWarp;
  for (int64_t v = warp; v < T; v += nwarps) {
    int cnt = count[v];
    float* grow = grad + v * C;
    if (cnt == 0) {
      if (!accumulate)
        for (int c0 = lane * 4; c0 < C; c0 += kWarp * 4) *reinterpret_cast<float4*>(grow + c0) = make_float4(0.f, 0.f, 0.f, 0.f);
      continue;
    }
    const int32_t* lst = list + offset[v];
    // 128 channels (one float4 per lane) per walk of the list; wider volumes walk it again per slab
    for (int cb = 0; cb < C; cb += kWarp * 4) {
      const int c0 = cb + lane * 4;
      const bool active = c0 < C;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      int32_t last = -1;
      for (int step = 0; step < cnt; ++step) {
        // next entry in ascending order: smallest id > last (warp-wide selection; the lists are short)
        int32_t best = 0x7fffffff;
        for (int i = lane; i < cnt; i += kWarp) {
          int32_t e = lst[i];
          if (e > last && e < best) best = e;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
        last = best;
        if (active) {
          float w = wts[best];
          float4 d = *reinterpret_cast<const float4*>(dlatent + (int64_t)(best >> 3) * ld + c0);
          acc.x = fmaf(w, d.x, acc.x); acc.y = fmaf(w, d.y, acc.y);
          acc.z = fmaf(w, d.z, acc.z); acc.w = fmaf(w, d.w, acc.w);
        }
      }
      if (active) {
        float4* dst = reinterpret_cast<float4*>(grow + c0);
        if (accumulate) {
          float4 o = *dst;
          acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w;
        }
        *dst = acc;
      }
    }
  }
}

}  // namespace nrf

extern "C" int64_t nrf_scatter_sorted_workspace_bytes(int64_t N, int SB, int64_t V) {
  int64_t T = (int64_t)SB * V, E = N * 8;
  int64_t nb = (T + 1023) / 1024;
  // count[T] cursor[T] offset[T] block_sums[nb+1] key[E] list[E] wts[E]
  return (3 * T + nb + 1 + 2 * E) * 4 + E * 4 + 1024;
}

extern "C" int nrf_scatter_volume_grad_sorted(const float* rays, const float* z, int R, int K, int rays_per_scene,
                                              const float* dlatent, int ld, float* grad_cl, int SB, int C, int S0,
                                              int S1, int S2, const float* bounds_host, int accumulate,
                                              void* workspace, void* stream) {
  NRF_REQUIRE(rays && z && dlatent && grad_cl && bounds_host && workspace, NRF_EINVAL,
              "nrf_scatter_volume_grad_sorted: null pointer");
  NRF_REQUIRE(R > 0 && K > 0 && R == SB * rays_per_scene, NRF_EINVAL,
              "nrf_scatter_volume_grad_sorted: R != SB*rays_per_scene");
  NRF_REQUIRE(C % 4 == 0 && ld % 4 == 0 && ld >= C, NRF_ENOSUP, "nrf_scatter_volume_grad_sorted: C/ld alignment");
  int64_t V = (int64_t)S0 * S1 * S2, T = (int64_t)SB * V, N = (int64_t)R * K, E = N * 8;
  NRF_REQUIRE(T < ((int64_t)1 << 31) && E < ((int64_t)1 << 31), NRF_ENOSUP,
              "nrf_scatter_volume_grad_sorted: more than 2^31 voxels or entries");
  cudaStream_t s = as_stream(stream);
  int64_t nb = (T + 1023) / 1024;
  int32_t* count = reinterpret_cast<int32_t*>(workspace);
  int32_t* cursor = count + T;
  int32_t* offset = cursor + T;
  int32_t* block_sums = offset + T;
  int32_t* key = block_sums + nb + 1;
  int32_t* list = key + E;
  float* wts = reinterpret_cast<float*>(list + E);
  NRF_CUDA_OK(cudaMemsetAsync(count, 0, (size_t)(2 * T) * 4, s));       // count and cursor
  ScatterArgs a;
  a.rays = rays; a.z = z; a.R = R; a.K = K; a.rays_per_scene = rays_per_scene;
  a.dlatent = dlatent; a.ld = ld; a.grad = grad_cl;
  a.SB = SB; a.C = C; a.S0 = S0; a.S1 = S1; a.S2 = S2;
  fill_bounds(bounds_host, a.bmin, a.bext);
  { LaunchScope ls_(NRF_CAT_SCATTER, s);
    scatter_count_kernel<<<(unsigned)((N + 255) / 256), 256, 0, s>>>(a, V, key, wts, count); }
  NRF_LAUNCH_OK();
  { LaunchScope ls_(NRF_CAT_SCATTER, s);
    scan_block_kernel<<<(unsigned)nb, 256, 0, s>>>(count, offset, block_sums, T); }
  NRF_LAUNCH_OK();
  { LaunchScope ls_(NRF_CAT_SCATTER, s);
    scan_sums_kernel<<<1, 1024, 0, s>>>(block_sums, (int)nb, nullptr); }
  NRF_LAUNCH_OK();
  { LaunchScope ls_(NRF_CAT_SCATTER, s);
    scan_add_kernel<<<(unsigned)nb, 256, 0, s>>>(offset, block_sums, T); }
  NRF_LAUNCH_OK();
  { LaunchScope ls_(NRF_CAT_SCATTER, s);
    scatter_fill_kernel<<<(unsigned)((E + 255) / 256), 256, 0, s>>>(key, E, offset, cursor, list); }
  NRF_LAUNCH_OK();
  int64_t want = (T + 7) / 8;
  int max_blocks = sm_count() * 32;
  int blocks = (int)(want < max_blocks ? want : max_blocks);
  { LaunchScope ls_(NRF_CAT_SCATTER, s);
    scatter_reduce_kernel<<<blocks, 256, 0, s>>>(offset, count, list, wts, dlatent, ld, grad_cl, C, T, accumulate); }
  NRF_LAUNCH_OK();
  return NRF_OK;
}
