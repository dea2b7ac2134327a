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
