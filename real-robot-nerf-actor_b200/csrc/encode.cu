// Field-input producer and its transpose (volume-gradient scatter), plus the volume re-layout.
//   nrf_encode_points        <- neural_rendering.py:246-283, models_embed.py:185-203,259-277,366,405,
//                               utils.py:545-557
//   nrf_scatter_volume_grad  <- autograd of F.grid_sample at models_embed.py:275
//   nrf_volume_to_channels_* <- layout change so one trilinear corner is one contiguous C-vector
#include "common.cuh"
#include "sortscan.cuh"
#include "tc_ptx.cuh"

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

// (SB, C, V) -> (SB, V, C) for C % 128 == 0: a CTA moves 128 channels x 32 voxels.  Every thread has 16 independent
// 128 B-coalesced row loads in flight before the first shared-memory store (the 32x32 kernel above has 4), tile row
// stride 129 floats: conflict-free in both directions; the voxel rows leave as four 128 B segments per warp store.
__global__ void __launch_bounds__(256) volume_to_last_c128_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                                  int C, int64_t V) {
  __shared__ float tile[32][129];
  const int lane = threadIdx.x % kWarp, wid = threadIdx.x / kWarp;
  const int b = blockIdx.z;
  const int64_t v0 = (int64_t)blockIdx.x * 32;
  const int c0 = blockIdx.y * 128;
  const float* s = src + ((int64_t)b * C + c0) * V + v0;
  float* d = dst + ((int64_t)b * V + v0) * C + c0;
  const bool vok = v0 + lane < V;
  float x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = vok ? __ldg(s + (int64_t)(wid + 8 * i) * V + lane) : 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) tile[lane][wid + 8 * i] = x[i];
  __syncthreads();
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int v = wid * 4 + q;
    if (v0 + v < V) {
#pragma unroll
      for (int i = 0; i < 4; ++i) d[(int64_t)v * C + lane + 32 * i] = tile[v][lane + 32 * i];
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Re-layout of the voxels the rays actually touch.  A training step of 2 x 2048 rays reads a few per cent of a 100^3
// volume (one scene's 2048 rays of 16 384 on a 200^3 grid: 1.5 %), yet the dense (C, V) -> (V, C) pass moves all of it
// - 8.2 GB for the 200^3 volume, every step, on every rank.  mark_voxels_kernel flags the in-grid trilinear corners of
// a pass's samples (the same geometry code as the gather); volume_to_last_marked_kernel then moves only the 32-voxel
// tiles that hold a newly flagged voxel and marks the whole tile as done, so that the second pass of a step adds what
// it touches beyond the first.  Tiles without a flag stay unwritten in the channels-last buffer: the gather never
// reads them (it reads in-grid corners of its samples only - exactly the flagged voxels).
constexpr uint8_t kVoxWanted = 2, kVoxDone = 1;

__global__ void __launch_bounds__(256) mark_voxels_kernel(const float* __restrict__ rays, const float* __restrict__ z,
                                                          int R, int K, int rays_per_scene, int S0, int S1, int S2,
                                                          float b0, float b1, float b2, float e0, float e1, float e2,
                                                          uint8_t* __restrict__ flags) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= (int64_t)R * K) return;
  const int r = (int)(n / K);
  const float bmin[3] = {b0, b1, b2}, bext[3] = {e0, e1, e2};
  const SampleGeom g = sample_geometry(rays + (int64_t)r * 8, z[n], bmin, bext);
  Corner8 c8;
  trilinear_corners(g.cx, g.cy, g.cz, S0, S1, S2, 1, c8);               // C = 1: off = voxel index
  uint8_t* f = flags + (int64_t)(r / rays_per_scene) * S0 * S1 * S2;
#pragma unroll
  for (int k = 0; k < 8; ++k)
    if (c8.off[k] >= 0 && f[c8.off[k]] == 0) f[c8.off[k]] = kVoxWanted;  // (racing writers all store the same value)
}

// grid (tiles of 32 voxels, 1, SB), 256 threads; 64 channels x 32 voxels per round through shared memory
__global__ void __launch_bounds__(256) volume_to_last_marked_kernel(const float* __restrict__ src,
                                                                    float* __restrict__ dst, int C, int64_t V,
                                                                    uint8_t* __restrict__ flags) {
  __shared__ float tile[32][65];
  const int lane = threadIdx.x % kWarp, wid = threadIdx.x / kWarp;
  const int b = blockIdx.z;
  const int64_t v0 = (int64_t)blockIdx.x * 32;
  const bool vok = v0 + lane < V;
  uint8_t* fl = flags + (int64_t)b * V + v0;
  const uint8_t mine = vok ? fl[lane] : (uint8_t)0;
  if (__ballot_sync(0xffffffffu, mine == kVoxWanted) == 0u) return;      // same flags in every warp: uniform exit
  for (int c0 = 0; c0 < C; c0 += 64) {
    const float* s = src + ((int64_t)b * C + c0) * V + v0;
    float* d = dst + ((int64_t)b * V + v0) * C + c0;
    float x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = (vok && c0 + wid + 8 * i < C) ? __ldg(s + (int64_t)(wid + 8 * i) * V + lane) : 0.f;
    __syncthreads();                                                     // the previous round's reads of the tile
#pragma unroll
    for (int i = 0; i < 8; ++i) tile[lane][wid + 8 * i] = x[i];
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int v = wid * 4 + q;
      if (v0 + v < V) {
#pragma unroll
        for (int i = 0; i < 2; ++i)
          if (c0 + lane + 32 * i < C) d[(int64_t)v * C + lane + 32 * i] = tile[v][lane + 32 * i];
      }
    }
  }
  __syncthreads();                                                       // every warp has read the flags
  if (wid == 0 && vok) fl[lane] = kVoxDone;
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
  uint8_t* touch;   // optional: one byte per group of 32 consecutive samples, 1 if any of them has a corner inside the grid
  int fma;          // corner accumulation out = fma(v, w, out) (ATen's CUDA grid_sampler_3d) instead of the separately
                    // rounded multiply and add of ATen's CPU kernel (the default)
  int pe_dirs;      // use_code_viewdirs (models_embed.py:370-372): the positional encoding runs over [xyz | viewdir] (6
                    // inputs, tail 6 + 12 F) instead of over xyz with the raw view direction behind it (tail 6 + 6 F)
};

// One trilinear corner: out += v * w in ATen's rounding.  The CPU kernel rounds the product and the sum separately;
// nvcc contracts the same source line of the CUDA kernel into an FMA.  Measured on B200 (tests/test_gpu_bench_sizes.py):
// the default is bit-identical to CPU ATen, the FMA form to CUDA-eager ATen.
__device__ __forceinline__ float4 corner_acc(float4 acc, float4 v, float w, int fma) {
  if (fma) {
    acc.x = __fmaf_rn(v.x, w, acc.x); acc.y = __fmaf_rn(v.y, w, acc.y);
    acc.z = __fmaf_rn(v.z, w, acc.z); acc.w = __fmaf_rn(v.w, w, acc.w);
  } else {
    acc.x = __fadd_rn(acc.x, __fmul_rn(v.x, w)); acc.y = __fadd_rn(acc.y, __fmul_rn(v.y, w));
    acc.z = __fadd_rn(acc.z, __fmul_rn(v.z, w)); acc.w = __fadd_rn(acc.w, __fmul_rn(v.w, w));
  }
  return acc;
}

template <typename T> __device__ __forceinline__ T to_out(float v);
template <> __device__ __forceinline__ float to_out<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 to_out<__nv_bfloat16>(float v) {
  return __float2bfloat16_rn(v);
}
template <> __device__ __forceinline__ __half to_out<__half>(float v) { return __float2half_rn(v); }

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
template <> __device__ __forceinline__ void store4<__half>(__half* p, float4 v) {
  __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
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
  const int d_pe = a.pe_dirs ? 6 : 3;
  const int n_pe = d_pe + 2 * d_pe * a.num_freqs;
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
          acc = corner_acc(acc, v[k], c8.w[k], a.fma);
        }
      }
      store4<T>(row + c0, acc);
    }
    // positional encoding [x y z | per frequency: sin(xyz), cos(xyz)] then the view direction; with pe_dirs the six
    // inputs [x y z dx dy dz] are encoded together: [in(6) | per frequency: sin(in), cos(in)] (utils.py:545-557)
    const float cin[6] = {g.cx, g.cy, g.cz, ray[3], ray[4], ray[5]};
    for (int e = lane; e < a.ld_out - C; e += kWarp) {
      float val = 0.f;
      if (e < d_pe) {
        val = cin[e];
      } else if (e < n_pe) {
        int q = e - d_pe;
        int f = q / (2 * d_pe), w = q % (2 * d_pe);
        float freq = a.freq_factor * (float)(1 << f);
        float phase = (w >= d_pe) ? 1.57079637050628662109375f : 0.0f;   // fp32(pi/2), utils.py:542
        val = sinf(__fmaf_rn(cin[w % d_pe], freq, phase));   // addcmul contracts to an FMA in ATen
      } else if (!a.pe_dirs && e < n_pe + 3) {
        val = ray[3 + (e - n_pe)];
      }
      row[C + e] = to_out<T>(val);
    }
    if (a.points && lane < 3) a.points[n * 3 + lane] = lane == 0 ? g.px : (lane == 1 ? g.py : g.pz);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// 32 samples per warp iteration (num_freqs == 6, ld_out == C + 64: the BASELINE / nerfact.conf shapes).
// The one-warp-per-sample kernel above is issue-bound (ncu: 643 warp instructions per sample, 72 % issue slots, 6 %
// DRAM): every lane repeats the sample's geometry and the 42 tail values are produced 32 at a time behind integer
// div/mod and a local-memory array.  Here a LANE owns a sample for everything scalar (geometry, trilinear setup, the
// 36 sines with compile-time element indices, its 128 B / 256 B tail written as 16 B vectors), and the warp then walks
// its 32 samples for the channel vectors: an out-of-grid sample costs one 256 B zero store, a sample that touches the
// grid gets its setup by shuffle and the same 8 coalesced corner reads and rounding order as above (bit-identical).
constexpr int kTailW = 64;
constexpr int kTailFreqs = 6;

__device__ __forceinline__ float tail_value(int e, const float c[3], const float d[3], float freq_factor) {
  constexpr int n_pe = 3 + 6 * kTailFreqs;
  if (e < 3) return c[e];
  if (e < n_pe) {
    const int q = e - 3, f = q / 6, w = q % 6;
    const float freq = freq_factor * (float)(1 << f);
    const float phase = (w >= 3) ? 1.57079637050628662109375f : 0.0f;   // fp32(pi/2), utils.py:542
    return sinf(__fmaf_rn(c[w % 3], freq, phase));
  }
  if (e < n_pe + 3) return d[e - n_pe];
  return 0.0f;
}

template <typename T> struct TailVec;
template <> struct TailVec<float> {
  static constexpr int kPer = 4;
  static __device__ __forceinline__ void store(float* p, const float* v) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};
template <> struct TailVec<__nv_bfloat16> {
  static constexpr int kPer = 8;
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float* v) {
    __nv_bfloat162 q0 = __floats2bfloat162_rn(v[0], v[1]), q1 = __floats2bfloat162_rn(v[2], v[3]);
    __nv_bfloat162 q2 = __floats2bfloat162_rn(v[4], v[5]), q3 = __floats2bfloat162_rn(v[6], v[7]);
    uint4 u;
    u.x = *reinterpret_cast<uint32_t*>(&q0); u.y = *reinterpret_cast<uint32_t*>(&q1);
    u.z = *reinterpret_cast<uint32_t*>(&q2); u.w = *reinterpret_cast<uint32_t*>(&q3);
    *reinterpret_cast<uint4*>(p) = u;
  }
};
template <> struct TailVec<__half> {
  static constexpr int kPer = 8;
  static __device__ __forceinline__ void store(__half* p, const float* v) {
    __half2 q0 = __floats2half2_rn(v[0], v[1]), q1 = __floats2half2_rn(v[2], v[3]);
    __half2 q2 = __floats2half2_rn(v[4], v[5]), q3 = __floats2half2_rn(v[6], v[7]);
    uint4 u;
    u.x = *reinterpret_cast<uint32_t*>(&q0); u.y = *reinterpret_cast<uint32_t*>(&q1);
    u.z = *reinterpret_cast<uint32_t*>(&q2); u.w = *reinterpret_cast<uint32_t*>(&q3);
    *reinterpret_cast<uint4*>(p) = u;
  }
};

template <typename T>
__global__ void __launch_bounds__(256) encode_points_w32_kernel(EncodeArgs a) {
  const int lane = threadIdx.x % kWarp;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / kWarp;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) / kWarp;
  const int64_t N = (int64_t)a.R * a.K;
  const int64_t groups = (N + kWarp - 1) / kWarp;
  T* out = reinterpret_cast<T*>(a.out);
  const int C = a.C;
  const int64_t scene_stride = (int64_t)a.S0 * a.S1 * a.S2 * C;
  for (int64_t grp = warp; grp < groups; grp += nwarps) {
    const int64_t n = grp * kWarp + lane;
    const bool valid = n < N;
    const int64_t nn = valid ? n : N - 1;
    const int r = (int)(nn / a.K);
    const float4 ra = __ldg(reinterpret_cast<const float4*>(a.rays + (int64_t)r * 8));
    const float4 rb = __ldg(reinterpret_cast<const float4*>(a.rays + (int64_t)r * 8) + 1);
    const float ray[6] = {ra.x, ra.y, ra.z, ra.w, rb.x, rb.y};
    const SampleGeom g = sample_geometry(ray, __ldg(a.z + nn), a.bmin, a.bext);
    const TriSetup ts = trilinear_setup(g.cx, g.cy, g.cz, a.S0, a.S1, a.S2);
    if (valid) {
      const float c[3] = {g.cx, g.cy, g.cz}, d[3] = {ra.w, rb.x, rb.y};
      T* tail = out + n * a.ld_out + C;
      constexpr int kPer = TailVec<T>::kPer;
#pragma unroll
      for (int j = 0; j < kTailW / kPer; ++j) {
        float v[kPer];
#pragma unroll
        for (int i = 0; i < kPer; ++i) v[i] = tail_value(j * kPer + i, c, d, a.freq_factor);
        TailVec<T>::store(tail + j * kPer, v);
      }
      if (a.points) {
        a.points[n * 3 + 0] = g.px; a.points[n * 3 + 1] = g.py; a.points[n * 3 + 2] = g.pz;
      }
    }
    const unsigned touch = __ballot_sync(0xffffffffu, valid && trilinear_touches(ts, a.S0, a.S1, a.S2));
    if (a.touch && lane == 0) a.touch[grp] = touch != 0u;
    const int scene_l = r / a.rays_per_scene;
    const int cnt = (int)((N - grp * kWarp) < kWarp ? (N - grp * kWarp) : kWarp);
    // The eight corners (voxel index within the scene or -1, weight) are worked out ONCE, by the lane that owns the
    // sample, and handed to the warp by 16 shuffles per sample: the warp used to redo the whole corner set-up (bounds
    // tests, 64-bit offsets, weight products: ~120 instructions) for each of its 32 samples.
    int cvox[8];
    float cw[8];
    {
      Corner8 c8;
      corners_from_setup(ts, a.S0, a.S1, a.S2, 1, c8);                  // C = 1: off = voxel index
#pragma unroll
      for (int k = 0; k < 8; ++k) { cvox[k] = (int)c8.off[k]; cw[k] = c8.w[k]; }
    }
    for (int s = 0; s < cnt; ++s) {
      T* row = out + (grp * kWarp + s) * a.ld_out;
      if (!((touch >> s) & 1u)) {
        for (int c0 = lane * 4; c0 < C; c0 += kWarp * 4) store4<T>(row + c0, make_float4(0.f, 0.f, 0.f, 0.f));
        continue;
      }
      const int scene = __shfl_sync(0xffffffffu, scene_l, s);
      int vx[8];
      float w[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        vx[k] = __shfl_sync(0xffffffffu, cvox[k], s);
        w[k] = __shfl_sync(0xffffffffu, cw[k], s);
      }
      const float* vol = a.vol + (int64_t)scene * scene_stride;
      for (int c0 = lane * 4; c0 < C; c0 += kWarp * 4) {
        float4 v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k)
          v[k] = vx[k] >= 0 ? __ldg(reinterpret_cast<const float4*>(vol + (int64_t)vx[k] * C + c0))
                            : make_float4(0.f, 0.f, 0.f, 0.f);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          if (vx[k] >= 0) {
            acc = corner_acc(acc, v[k], w[k], a.fma);
          }
        }
        store4<T>(row + c0, acc);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// The gather staged through TMA and shared memory (BASELINE north_star, subsystem 2).  The channels-last volume is a
// 5-D tensor (C, S2, S1, S0, SB) to the TMA unit; the 2 x 2 x 2 voxel neighbourhood of a sample is ONE tile of that
// tensor - box (C, 2, 2, 2, 1) = 8 rows of 512 B - so the lane that owns a sample fetches all eight corners with a
// single cp.async.bulk.tensor.5d at coordinates (0, x0, y0, z0, scene), straight from its own registers: no corner
// offsets, no bounds tests, no shuffles of addresses.  Corners outside the grid are the TMA unit's out-of-bounds
// zero fill, which IS F.grid_sample's zeros padding (models_embed.py:275); a zero row adds +-0 to the sum, so the
// latents are bit-identical to the kernels that skip those corners.  The box arrives in ATen's corner order
// (x fastest, then y, then z).  Per warp: kTmaStages boxes of 4 KB in flight (mbarrier per stage), the warp consumes
// box i (8 x LDS.128 per lane, the eight weights by shuffle from the owner) while boxes i+1 .. are on their way.
// 64 or 128 channels, the w32 tail layout; NRF_ENCODE_TMA=0 falls back to encode_points_w32_kernel (A/B, tests).
constexpr int kTmaWarps = 4;

__device__ __forceinline__ void tma_load_box5(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c1, int c2,
                                              int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}

template <typename T, int kTmaStages, int C>
__global__ void __launch_bounds__(kTmaWarps * 32) encode_points_tma_kernel(const __grid_constant__ CUtensorMap vmap,
                                                                          EncodeArgs a) {
  extern __shared__ uint8_t tma_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(tma_smem_raw) + 127) & ~uintptr_t(127));
  __shared__ __align__(8) uint64_t bars[kTmaWarps][kTmaStages];
  int lane;
  asm volatile("mov.u32 %0, %%laneid;" : "=r"(lane));
  const int wid = uniform_warp_idx();
  const int64_t warp = (int64_t)blockIdx.x * kTmaWarps + wid;
  const int64_t nwarps = (int64_t)gridDim.x * kTmaWarps;
  const int64_t N = (int64_t)a.R * a.K;
  const int64_t groups = (N + kWarp - 1) / kWarp;
  T* out = reinterpret_cast<T*>(a.out);
  constexpr int kBox = 8 * C * 4;                  // bytes: 8 corners x C channels fp32 (a lane owns 4 channels: C = 64
                                                   // leaves the upper half-warp idle in the sums, as in the LDG kernel)
  if (lane == 0) {
    for (int s = 0; s < kTmaStages; ++s) mbar_init(&bars[wid][s], 1);
    fence_barrier_init();
  }
  if (threadIdx.x == 0) tma_prefetch_desc(&vmap);
  __syncthreads();
  const uint32_t slot0 = smem_u32(smem) + wid * kTmaStages * kBox;
  const uint32_t bar0 = smem_u32(&bars[wid][0]);
  uint32_t n_issued = 0, n_done = 0;             // boxes of this warp so far: stage = n % stages, parity = (n / stages) & 1
  for (int64_t grp = warp; grp < groups; grp += nwarps) {
    const int64_t n = grp * kWarp + lane;
    const bool valid = n < N;
    const int64_t nn = valid ? n : N - 1;
    const int r = (int)(nn / a.K);
    const float4 ra = __ldg(reinterpret_cast<const float4*>(a.rays + (int64_t)r * 8));
    const float4 rb = __ldg(reinterpret_cast<const float4*>(a.rays + (int64_t)r * 8) + 1);
    const float ray[6] = {ra.x, ra.y, ra.z, ra.w, rb.x, rb.y};
    const SampleGeom g = sample_geometry(ray, __ldg(a.z + nn), a.bmin, a.bext);
    const TriSetup ts = trilinear_setup(g.cx, g.cy, g.cz, a.S0, a.S1, a.S2);
    if (valid) {
      const float c[3] = {g.cx, g.cy, g.cz}, d[3] = {ra.w, rb.x, rb.y};
      T* tail = out + n * a.ld_out + C;
      constexpr int kPer = TailVec<T>::kPer;
#pragma unroll
      for (int j = 0; j < kTailW / kPer; ++j) {
        float v[kPer];
#pragma unroll
        for (int i = 0; i < kPer; ++i) v[i] = tail_value(j * kPer + i, c, d, a.freq_factor);
        TailVec<T>::store(tail + j * kPer, v);
      }
      if (a.points) {
        a.points[n * 3 + 0] = g.px; a.points[n * 3 + 1] = g.py; a.points[n * 3 + 2] = g.pz;
      }
    }
    const unsigned touch = __ballot_sync(0xffffffffu, valid && trilinear_touches(ts, a.S0, a.S1, a.S2));
    if (a.touch && lane == 0) a.touch[grp] = touch != 0u;
    const int scene = r / a.rays_per_scene;
    float cw[8];                                  // the corner weights, (wx * wy) * wz in ATen's rounding order
    {
      Corner8 c8;
      corners_from_setup(ts, a.S0, a.S1, a.S2, 1, c8);
#pragma unroll
      for (int k = 0; k < 8; ++k) cw[k] = c8.w[k];
    }
    auto issue = [&](int s) {                     // the owner of sample s fetches its 2 x 2 x 2 x C box
      if (lane == s) {
        const uint32_t st = n_issued % kTmaStages;
        const uint32_t bar = bar0 + st * 8;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)kBox) : "memory");
        tma_load_box5(slot0 + st * kBox, &vmap, bar, ts.x0, ts.y0, ts.z0, scene);
      }
      ++n_issued;
    };
    // samples that touch the grid, in order; up to kTmaStages - 1 boxes ahead of the one being summed
    unsigned ahead = touch;
    for (int i = 0; i < kTmaStages - 1 && ahead; ++i) {
      const int s = __ffs(ahead) - 1;
      ahead &= ahead - 1;
      issue(s);
    }
    unsigned todo = touch;
    for (int s = 0; s < kWarp; ++s) {
      if (grp * kWarp + s >= N) break;
      T* row = out + (grp * kWarp + s) * a.ld_out;
      if (!((todo >> s) & 1u)) {
        if (lane * 4 < C) store4<T>(row + lane * 4, make_float4(0.f, 0.f, 0.f, 0.f));
        continue;
      }
      __syncwarp();                               // every lane is done with the slot the next box lands in
      if (ahead) {
        const int sn = __ffs(ahead) - 1;
        ahead &= ahead - 1;
        issue(sn);
      }
      float w[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) w[k] = __shfl_sync(0xffffffffu, cw[k], s);
      const uint32_t st = n_done % kTmaStages;
      mbar_wait(&bars[wid][st], (n_done / kTmaStages) & 1);
      ++n_done;
      const uint32_t src = slot0 + st * kBox + lane * 16;
      if (lane * 4 < C) {
        float4 v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k)
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                       : "=f"(v[k].x), "=f"(v[k].y), "=f"(v[k].z), "=f"(v[k].w) : "r"(src + k * (C * 4)));
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int k = 0; k < 8; ++k) acc = corner_acc(acc, v[k], w[k], a.fma);
        store4<T>(row + lane * 4, acc);
      }
    }
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
  if (to_last && C % 128 == 0) {
    dim3 grid128((unsigned)((V + 31) / 32), (unsigned)(C / 128), (unsigned)SB);
    LaunchScope ls_(NRF_CAT_TRANSPOSE, as_stream(stream));
    volume_to_last_c128_kernel<<<grid128, 256, 0, as_stream(stream)>>>(src, dst, C, V);
    NRF_LAUNCH_OK();
    return NRF_OK;
  }
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

namespace nrf {
// The same at voxel granularity: a warp scans 32 flags and moves only the flagged voxels - C strided 4 B reads (a 32 B
// sector each) and one coalesced row per voxel.  8 x the bytes of the tile kernel per voxel moved, so it pays when the
// touched voxels are sparse WITHIN the tiles too: a ray crosses a 32-voxel run along x in one or two voxels, so 2048
// rays on a 200^3 grid touch 1.5 % of the voxels but half of the tiles.
__global__ void __launch_bounds__(256) volume_to_last_marked_voxels_kernel(const float* __restrict__ src,
                                                                           float* __restrict__ dst, int C, int64_t V,
                                                                           uint8_t* __restrict__ flags) {
  const int lane = threadIdx.x % kWarp, wid = threadIdx.x / kWarp;
  const int b = blockIdx.z;
  const int64_t v0 = ((int64_t)blockIdx.x * 8 + wid) * 32;
  if (v0 >= V) return;
  const bool vok = v0 + lane < V;
  uint8_t* fl = flags + (int64_t)b * V + v0;
  const uint8_t mine = vok ? fl[lane] : (uint8_t)0;
  unsigned m = __ballot_sync(0xffffffffu, mine == kVoxWanted);
  const float* s = src + (int64_t)b * C * V + v0;
  float* d = dst + ((int64_t)b * V + v0) * C;
  while (m) {
    const int vi = __ffs(m) - 1;
    m &= m - 1;
    for (int c = lane; c < C; c += 128) {
      float x[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) x[i] = c + 32 * i < C ? __ldg(s + (int64_t)(c + 32 * i) * V + vi) : 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (c + 32 * i < C) d[(int64_t)vi * C + c + 32 * i] = x[i];
    }
  }
  if (mine == kVoxWanted) fl[lane] = kVoxDone;
}
}  // namespace nrf

extern "C" int nrf_mark_voxels(const float* rays, const float* z, int R, int K, int rays_per_scene, int SB, int S0,
                               int S1, int S2, const float* bounds_host, uint8_t* flags, void* stream) {
  NRF_REQUIRE(rays && z && bounds_host && flags && R > 0 && K > 0 && rays_per_scene > 0 && R == SB * rays_per_scene,
              NRF_EINVAL, "nrf_mark_voxels: bad arguments");
  float bmin[3], bext[3];
  fill_bounds(bounds_host, bmin, bext);
  const int64_t N = (int64_t)R * K;
  { LaunchScope ls_(NRF_CAT_TRANSPOSE, as_stream(stream));
    mark_voxels_kernel<<<(unsigned)((N + 255) / 256), 256, 0, as_stream(stream)>>>(
        rays, z, R, K, rays_per_scene, S0, S1, S2, bmin[0], bmin[1], bmin[2], bext[0], bext[1], bext[2], flags); }
  NRF_LAUNCH_OK();
  return NRF_OK;
}

extern "C" int nrf_volume_to_channels_last_marked(const float* src, float* dst, int SB, int C, int64_t V,
                                                  uint8_t* flags, int per_voxel, void* stream) {
  NRF_REQUIRE(src && dst && flags && SB > 0 && C > 0 && V > 0, NRF_EINVAL, "nrf_volume_to_channels_last_marked: bad args");
  NRF_REQUIRE(SB <= 65535 && (V + 31) / 32 < ((int64_t)1 << 31), NRF_ENOSUP,
              "nrf_volume_to_channels_last_marked: grid too large");
  { LaunchScope ls_(NRF_CAT_TRANSPOSE, as_stream(stream));
    if (per_voxel) {
      dim3 grid((unsigned)((V + 255) / 256), 1, (unsigned)SB);
      volume_to_last_marked_voxels_kernel<<<grid, 256, 0, as_stream(stream)>>>(src, dst, C, V, flags);
    } else {
      dim3 grid((unsigned)((V + 31) / 32), 1, (unsigned)SB);
      volume_to_last_marked_kernel<<<grid, 256, 0, as_stream(stream)>>>(src, dst, C, V, flags);
    } }
  NRF_LAUNCH_OK();
  return NRF_OK;
}

static int encode_points_impl(const float* rays, const float* z, int R, int K, int rays_per_scene,
                              const float* vol_cl, int SB, int C, int S0, int S1, int S2,
                              const float* bounds_host, int num_freqs, float freq_factor, void* out,
                              int ld_out, int out_bf16, float* points_out, uint8_t* touch_flags, void* stream) {
  NRF_REQUIRE(rays && z && vol_cl && bounds_host && out, NRF_EINVAL, "nrf_encode_points: null pointer");
  NRF_REQUIRE(R > 0 && K > 0 && rays_per_scene > 0 && R == SB * rays_per_scene, NRF_EINVAL,
              "nrf_encode_points: R=%d must equal SB*rays_per_scene=%d*%d", R, SB, rays_per_scene);
  NRF_REQUIRE(C % 4 == 0 && C > 0, NRF_ENOSUP, "nrf_encode_points: C=%d must be a multiple of 4", C);
  const int pe_dirs = (out_bf16 >> 9) & 1;
  NRF_REQUIRE(num_freqs >= 0 && num_freqs <= 24, NRF_EINVAL, "nrf_encode_points: num_freqs=%d", num_freqs);
  NRF_REQUIRE(ld_out >= C + 6 + (pe_dirs ? 12 : 6) * num_freqs && ld_out % 4 == 0, NRF_EINVAL,
              "nrf_encode_points: ld_out=%d too small / unaligned", ld_out);
  EncodeArgs a;
  a.rays = rays; a.z = z; a.R = R; a.K = K; a.rays_per_scene = rays_per_scene;
  a.vol = vol_cl; a.SB = SB; a.C = C; a.S0 = S0; a.S1 = S1; a.S2 = S2;
  fill_bounds(bounds_host, a.bmin, a.bext);
  a.num_freqs = num_freqs; a.freq_factor = freq_factor;
  a.out = out; a.ld_out = ld_out; a.points = points_out;
  a.fma = (out_bf16 >> 8) & 1;
  a.pe_dirs = pe_dirs;
  a.touch = touch_flags;
  out_bf16 &= 0xff;
  int64_t N = (int64_t)R * K;
  int threads = 256;
  const bool w32 = !pe_dirs && num_freqs == kTailFreqs && ld_out == C + kTailW &&
                   (reinterpret_cast<uintptr_t>(rays) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0;
  const char* tma_env = getenv("NRF_ENCODE_TMA");             // "0": encode_points_w32_kernel (A/B, tests)
  if (w32 && (C == 128 || C == 64) && S0 >= 2 && S1 >= 2 && S2 >= 2 && !(tma_env && atoi(tma_env) == 0) &&
      (reinterpret_cast<uintptr_t>(vol_cl) & 15) == 0) {
    // the corners through TMA boxes and shared memory (encode_points_tma_kernel)
    EncodeTiledFn fn = encode_tiled_fn();
    NRF_REQUIRE(fn != nullptr, NRF_ECUDA, "cuTensorMapEncodeTiled entry point not found");
    CUtensorMap vmap;
    cuuint64_t gdim[5] = {(cuuint64_t)C, (cuuint64_t)S2, (cuuint64_t)S1, (cuuint64_t)S0, (cuuint64_t)SB};
    cuuint64_t gstr[4] = {(cuuint64_t)C * 4, (cuuint64_t)S2 * C * 4, (cuuint64_t)S1 * S2 * C * 4,
                          (cuuint64_t)S0 * S1 * S2 * C * 4};
    cuuint32_t box[5] = {(cuuint32_t)C, 2, 2, 2, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult cr = fn(&vmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<float*>(vol_cl), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    NRF_REQUIRE(cr == CUDA_SUCCESS, NRF_ECUDA, "cuTensorMapEncodeTiled(volume) failed (%d)", (int)cr);
    static const int stages_env = getenv("NRF_ENCODE_TMA_STAGES") ? atoi(getenv("NRF_ENCODE_TMA_STAGES")) : 3;
    const int stages = stages_env == 2 || stages_env == 4 ? stages_env : 3;
    const int smem_bytes = kTmaWarps * stages * (8 * C * 4) + 128;
    int64_t want = ((N + 31) / 32 + kTmaWarps - 1) / kTmaWarps;
    int per_sm = (227 * 1024) / (smem_bytes + 1024);
    if (per_sm > 8) per_sm = 8;
    int max_blocks = sm_count() * per_sm;
    int blocks = (int)(want < max_blocks ? want : max_blocks);
    LaunchScope ls_(NRF_CAT_ENCODE, as_stream(stream));
#define NRF_ENC_TMA3(TT, NS, CC)                                                                                     \
    do {                                                                                                             \
      NRF_CUDA_OK(cudaFuncSetAttribute(encode_points_tma_kernel<TT, NS, CC>,                                         \
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));                    \
      encode_points_tma_kernel<TT, NS, CC><<<blocks, kTmaWarps * 32, smem_bytes, as_stream(stream)>>>(vmap, a);      \
    } while (0)
#define NRF_ENC_TMA2(TT, NS)                                                                                         \
    do {                                                                                                             \
      if (C == 128) NRF_ENC_TMA3(TT, NS, 128);                                                                       \
      else NRF_ENC_TMA3(TT, NS, 64);                                                                                 \
    } while (0)
#define NRF_ENC_TMA(TT)                                                                                              \
    do {                                                                                                             \
      if (stages == 2) NRF_ENC_TMA2(TT, 2);                                                                          \
      else if (stages == 4) NRF_ENC_TMA2(TT, 4);                                                                     \
      else NRF_ENC_TMA2(TT, 3);                                                                                      \
    } while (0)
    if (out_bf16 == 2) NRF_ENC_TMA(__half);
    else if (out_bf16) NRF_ENC_TMA(__nv_bfloat16);
    else NRF_ENC_TMA(float);
#undef NRF_ENC_TMA
#undef NRF_ENC_TMA2
#undef NRF_ENC_TMA3
  } else if (w32) {                            // 32 samples per warp iteration
    int64_t want = ((N + 31) / 32 + 7) / 8;
    int max_blocks = sm_count() * 8;
    int blocks = (int)(want < max_blocks ? want : max_blocks);
    LaunchScope ls_(NRF_CAT_ENCODE, as_stream(stream));
    if (out_bf16 == 2) encode_points_w32_kernel<__half><<<blocks, threads, 0, as_stream(stream)>>>(a);
    else if (out_bf16) encode_points_w32_kernel<__nv_bfloat16><<<blocks, threads, 0, as_stream(stream)>>>(a);
    else encode_points_w32_kernel<float><<<blocks, threads, 0, as_stream(stream)>>>(a);
  } else {                                     // any other shape: one warp per sample
    if (touch_flags)                           // (this kernel does not report them: every group counts as touching)
      NRF_CUDA_OK(cudaMemsetAsync(touch_flags, 1, (size_t)((N + 31) / 32), as_stream(stream)));
    int64_t want = (N + 7) / 8;
    int max_blocks = sm_count() * 16;
    int blocks = (int)(want < max_blocks ? want : max_blocks);
    LaunchScope ls_(NRF_CAT_ENCODE, as_stream(stream));
    if (out_bf16 == 2) encode_points_kernel<__half><<<blocks, threads, 0, as_stream(stream)>>>(a);
    else if (out_bf16) encode_points_kernel<__nv_bfloat16><<<blocks, threads, 0, as_stream(stream)>>>(a);
    else encode_points_kernel<float><<<blocks, threads, 0, as_stream(stream)>>>(a);
  }
  NRF_LAUNCH_OK();
  return NRF_OK;
}

extern "C" int nrf_encode_points(const float* rays, const float* z, int R, int K, int rays_per_scene,
                                 const float* vol_cl, int SB, int C, int S0, int S1, int S2,
                                 const float* bounds_host, int num_freqs, float freq_factor, void* out,
                                 int ld_out, int out_bf16, float* points_out, void* stream) {
  return encode_points_impl(rays, z, R, K, rays_per_scene, vol_cl, SB, C, S0, S1, S2, bounds_host, num_freqs, freq_factor,
                            out, ld_out, out_bf16, points_out, nullptr, stream);
}

extern "C" int nrf_encode_points_touch(const float* rays, const float* z, int R, int K, int rays_per_scene,
                                       const float* vol_cl, int SB, int C, int S0, int S1, int S2,
                                       const float* bounds_host, int num_freqs, float freq_factor, void* out,
                                       int ld_out, int out_bf16, float* points_out, uint8_t* touch_flags,
                                       void* stream) {
  return encode_points_impl(rays, z, R, K, rays_per_scene, vol_cl, SB, C, S0, S1, S2, bounds_host, num_freqs, freq_factor,
                            out, ld_out, out_bf16, points_out, touch_flags, stream);
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

// ---------------------------------------------------------------------------------------------------------------
// Both render passes in ONE counting sort, and (optionally) the gradient written straight in the caller's
// channel-first layout.  Entry e = 8 n + corner with n < N_a: sample n of pass a, else sample n - N_a of pass b;
// a voxel's entries are summed in ascending e (all of pass a, then pass b): fixed order, no float atomics.
// Replaces {sorted scatter a, sorted scatter b with read-modify-write, (V,C) -> (C,V) transpose of the whole
// gradient volume}: the dense gradient is written exactly once.
namespace nrf {

struct ScatterPass {
  const float* z;
  const float* dlat;
  int K, ld;
};

__global__ void __launch_bounds__(256) scatter_count2_kernel(const float* __restrict__ rays, int R, int rays_per_scene,
                                                             ScatterPass pa, ScatterPass pb, int64_t Na, int64_t Nb,
                                                             ScatterArgs g0, int64_t V, int32_t* __restrict__ key,
                                                             float* __restrict__ wts, int32_t* __restrict__ count) {
  int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= Na + Nb) return;
  const bool second = n >= Na;
  const int64_t nl = second ? n - Na : n;
  const int K = second ? pb.K : pa.K;
  const float* z = second ? pb.z : pa.z;
  int r = (int)(nl / K);
  int scene = r / rays_per_scene;
  SampleGeom g = sample_geometry(rays + (int64_t)r * 8, z[nl], g0.bmin, g0.bext);
  Corner8 c8;
  trilinear_corners(g.cx, g.cy, g.cz, g0.S0, g0.S1, g0.S2, 1, c8);      // C = 1: off = voxel index
  int4 kk[2];
  int32_t* kp = reinterpret_cast<int32_t*>(kk);
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    int32_t q = -1;
    if (c8.off[k] >= 0) {
      q = (int32_t)((int64_t)scene * V + c8.off[k]);
      atomicAdd(count + q, 1);
    }
    kp[k] = q;
  }
  reinterpret_cast<int4*>(key + n * 8)[0] = kk[0];
  reinterpret_cast<int4*>(key + n * 8)[1] = kk[1];
  reinterpret_cast<float4*>(wts + n * 8)[0] = make_float4(c8.w[0], c8.w[1], c8.w[2], c8.w[3]);
  reinterpret_cast<float4*>(wts + n * 8)[1] = make_float4(c8.w[4], c8.w[5], c8.w[6], c8.w[7]);
}

// list + its weights side by side (the reduce reads both with independent loads)
__global__ void __launch_bounds__(256) scatter_fill2_kernel(const int32_t* __restrict__ key, const float* __restrict__ wts,
                                                            int64_t E, const int32_t* __restrict__ offset,
                                                            int32_t* __restrict__ cursor, int32_t* __restrict__ list,
                                                            float* __restrict__ wlist) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  int32_t k = key[e];
  if (k < 0) return;
  int32_t pos = offset[k] + atomicAdd(cursor + k, 1);
  list[pos] = (int32_t)e;
  wlist[pos] = wts[e];
}

// Sum of one voxel's entries for the channels lane, lane+32, ... (CJ of them), in ascending entry order.
// Up to 32 entries (the usual case): one entry per lane, ranked with shuffles, then the dlatent rows are read four
// at a time -- three dependent memory round trips per VOXEL; the selection loop below pays them per ENTRY.
template <int CJ>
__device__ __forceinline__ void voxel_sum(const int32_t* __restrict__ lst, const float* __restrict__ wl, int cnt,
                                          ScatterPass pa, ScatterPass pb, int64_t Ea, int lane, float acc[CJ]) {
#pragma unroll
  for (int j = 0; j < CJ; ++j) acc[j] = 0.f;
  auto row_of = [&](int32_t e) {
    return e < Ea ? pa.dlat + (int64_t)(e >> 3) * pa.ld : pb.dlat + (int64_t)((e - Ea) >> 3) * pb.ld;
  };
  if (cnt == 1) {
    const float w = wl[0];
    const float* r = row_of(lst[0]);
#pragma unroll
    for (int j = 0; j < CJ; ++j) acc[j] = fmaf(w, r[lane + j * kWarp], acc[j]);
    return;
  }
  if (cnt <= kWarp) {
    const int32_t e_l = lane < cnt ? lst[lane] : 0x7fffffff;
    const float w_l = lane < cnt ? wl[lane] : 0.f;
    int rank = 0;                                             // entries are distinct: ranks are a permutation
    for (int j = 0; j < cnt; ++j) rank += (__shfl_sync(0xffffffffu, e_l, j) < e_l) ? 1 : 0;
    int k = 0;
    for (; k + 4 <= cnt; k += 4) {
      int32_t e[4]; float w[4]; float d[4][CJ];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int src = __ffs(__ballot_sync(0xffffffffu, rank == k + q)) - 1;
        e[q] = __shfl_sync(0xffffffffu, e_l, src);
        w[q] = __shfl_sync(0xffffffffu, w_l, src);
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float* r = row_of(e[q]);
#pragma unroll
        for (int j = 0; j < CJ; ++j) d[q][j] = r[lane + j * kWarp];
      }
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int j = 0; j < CJ; ++j) acc[j] = fmaf(w[q], d[q][j], acc[j]);
    }
    for (; k < cnt; ++k) {
      const int src = __ffs(__ballot_sync(0xffffffffu, rank == k)) - 1;
      const int32_t e = __shfl_sync(0xffffffffu, e_l, src);
      const float w = __shfl_sync(0xffffffffu, w_l, src);
      const float* r = row_of(e);
#pragma unroll
      for (int j = 0; j < CJ; ++j) acc[j] = fmaf(w, r[lane + j * kWarp], acc[j]);
    }
    return;
  }
  // more than a warp's worth of entries in one voxel (coarse grids under many samples; degenerate inputs): the voxel's
  // segment of the entry / weight lists belongs to this warp alone, so it is sorted IN PLACE by a bitonic network whose
  // comparators all point the same way (any length: the virtual padding behind `cnt` never moves) -
  // O(cnt log^2 cnt / 32) steps instead of the cnt^2 / 32 of a selection loop - and then summed in order
  int32_t* se = const_cast<int32_t*>(lst);
  float* sw = const_cast<float*>(wl);
  auto cmpxchg = [&](int i, int p) {
    const int32_t a = se[i], b = se[p];
    if (a > b) {
      se[i] = b; se[p] = a;
      const float wa = sw[i], wb = sw[p];
      sw[i] = wb; sw[p] = wa;
    }
  };
  for (int k = 2; (k >> 1) < cnt; k <<= 1) {
    for (int i = lane; i < cnt; i += kWarp) {
      const int p = i ^ (k - 1);
      if (p > i && p < cnt) cmpxchg(i, p);
    }
    __syncwarp();
    for (int j = k >> 2; j > 0; j >>= 1) {
      for (int i = lane; i < cnt; i += kWarp) {
        const int p = i ^ j;
        if (p > i && p < cnt) cmpxchg(i, p);
      }
      __syncwarp();
    }
  }
  for (int k = 0; k < cnt; ++k) {
    const float w = sw[k];
    const float* d = row_of(se[k]);
#pragma unroll
    for (int j = 0; j < CJ; ++j) acc[j] = fmaf(w, d[lane + j * kWarp], acc[j]);
  }
}

// Channel-first output: a CTA owns 32 consecutive voxels; its warps share the voxels that have entries (round-robin
// over the set bits), sum each into a (C x 32) shared tile, then every channel row leaves coalesced.  kVec (V % 32
// == 0: tiles never straddle scenes, rows are 16 B aligned): row stride 36, LDS.128 + STG.128, a warp writes four
// channel rows per instruction.  A tile without any entry is zero-filled without touching shared memory.
// The kernel is latency-bound (count/offset -> list -> dlatent rows -> barrier -> stores, one tile per CTA at a time):
// small CTAs (4 warps, up to 10 resident per SM) and the next tile's count / offset prefetched during the current one.
constexpr int kRcfWarps = 4;
template <int CJ, bool kVec>
__global__ void __launch_bounds__(kRcfWarps * 32, 10) scatter_reduce_cf_kernel(const int32_t* __restrict__ offset,
                                                                const int32_t* __restrict__ count,
                                                                const int32_t* __restrict__ list,
                                                                const float* __restrict__ wlist, ScatterPass pa,
                                                                ScatterPass pb, int64_t Ea, float* __restrict__ grad,
                                                                int64_t V, int64_t T) {
  constexpr int C = CJ * kWarp;
  constexpr int LD = kVec ? 36 : 33;
  constexpr int NW = kRcfWarps;
  __shared__ __align__(16) float tile[C][LD];
  const int lane = threadIdx.x % kWarp, wid = threadIdx.x / kWarp;
  const int r4 = lane >> 3, l8 = lane & 7;
  const int64_t ntiles = (T + 31) / 32;
  int cnt_n = 0, off_n = 0;
  if ((int64_t)blockIdx.x < ntiles) {
    const int64_t t = (int64_t)blockIdx.x * 32 + lane;
    if (t < T) { cnt_n = count[t]; off_n = offset[t]; }
  }
  for (int64_t tl = blockIdx.x; tl < ntiles; tl += gridDim.x) {
    const int64_t t = tl * 32 + lane;
    const bool valid = t < T;
    const int cnt_l = cnt_n, off_l = off_n;
    {
      const int64_t tn = (tl + gridDim.x) * 32 + lane;
      cnt_n = 0; off_n = 0;
      if (tn < T) { cnt_n = count[tn]; off_n = offset[tn]; }
    }
    const unsigned any = __ballot_sync(0xffffffffu, cnt_l > 0);   // uniform over the CTA (same counts in every warp)
    float* gcol;                                                  // scalar path: this lane's voxel, + c * V per channel
    float* gvec;                                                  // vector path: this lane's voxel quad
    {
      const int64_t scene = valid ? t / V : 0, v = valid ? t - scene * V : 0;
      gcol = grad + (scene * C) * V + v;
      const int64_t t0 = tl * 32, scene0 = t0 / V;
      gvec = grad + (scene0 * C) * V + (t0 - scene0 * V) + 4 * l8;
    }
    if (any == 0u) {
      if (kVec) {
#pragma unroll
        for (int i = 0; i < C / (4 * NW); ++i)
          *reinterpret_cast<float4*>(gvec + (int64_t)(wid * 4 + r4 + 4 * NW * i) * V) = make_float4(0.f, 0.f, 0.f, 0.f);
      } else if (valid) {
        for (int c = wid; c < C; c += NW) gcol[(int64_t)c * V] = 0.f;
      }
      continue;
    }
    unsigned m = any;
    int idx = 0;
    while (m) {
      const int vi = __ffs(m) - 1;
      m &= m - 1;
      if ((idx++ % NW) != wid) continue;
      const int cnt = __shfl_sync(0xffffffffu, cnt_l, vi);
      const int off = __shfl_sync(0xffffffffu, off_l, vi);
      float acc[CJ];
      voxel_sum<CJ>(list + off, wlist + off, cnt, pa, pb, Ea, lane, acc);
#pragma unroll
      for (int j = 0; j < CJ; ++j) tile[lane + j * kWarp][vi] = acc[j];
    }
    __syncthreads();
    if (kVec) {
      const unsigned sel = (any >> (4 * l8)) & 0xFu;
#pragma unroll
      for (int i = 0; i < C / (4 * NW); ++i) {
        const int c = wid * 4 + r4 + 4 * NW * i;
        float4 x = *reinterpret_cast<const float4*>(&tile[c][4 * l8]);
        x.x = (sel & 1u) ? x.x : 0.f; x.y = (sel & 2u) ? x.y : 0.f;
        x.z = (sel & 4u) ? x.z : 0.f; x.w = (sel & 8u) ? x.w : 0.f;
        *reinterpret_cast<float4*>(gvec + (int64_t)c * V) = x;
      }
    } else if (valid) {
      const bool has = cnt_l > 0;
      for (int c = wid; c < C; c += NW) gcol[(int64_t)c * V] = has ? tile[c][lane] : 0.f;
    }
    __syncthreads();
  }
}

// The same reduction for the kVec case, restructured for the regime where most voxels of a tile have entries (rays
// that stay inside the volume: 6 M entries over ~1 M voxels at config 2).  scatter_reduce_cf_kernel pays a dependent
// chain (list -> rows) and a shuffle-ranked selection PER VOXEL and reads every 512 B row as four 128 B pieces
// (ncu, in-box: 71 warp instructions per entry, 28 % DRAM).  Here a warp owns 8 voxels of the tile (every fourth) and
//   1. loads up to 32 entries of whole voxels with one read of the entry and the weight list,
//   2. ranks every entry among the entries of its own voxel (shuffles over the longest voxel of the batch only) and
//      drops {entry, weight} at its sorted position in shared memory,
//   3. walks the sorted batch four entries at a time - one broadcast LDS.64 and ONE vector load of the lane's CJ
//      consecutive channels per entry, the four row loads of a group in flight together - and adds in ascending entry
//      order (the same fmaf chain per channel as voxel_sum: bit-identical results),
//   4. puts a finished voxel's row into the (32 voxel x C) tile with one vector STS; 16 B slots are XOR-swizzled by the
//      voxel quad so that the channel-first read-out below (4 scalar LDS -> STG.128 of four consecutive voxels, a
//      warp instruction = four full 128 B lines) is conflict-free too.
// Voxels with more than 32 entries take voxel_sum's in-place sort.
template <int CJ> struct LaneVec;
template <> struct LaneVec<4> {
  float4 v;
  __device__ __forceinline__ void zero() { v = make_float4(0.f, 0.f, 0.f, 0.f); }
  __device__ __forceinline__ void load(const float* p) { v = __ldg(reinterpret_cast<const float4*>(p)); }
  __device__ __forceinline__ void fma(float w, const LaneVec& d) {
    v.x = fmaf(w, d.v.x, v.x); v.y = fmaf(w, d.v.y, v.y); v.z = fmaf(w, d.v.z, v.z); v.w = fmaf(w, d.v.w, v.w);
  }
  __device__ __forceinline__ void store(float* p) const { *reinterpret_cast<float4*>(p) = v; }
};
template <> struct LaneVec<2> {
  float2 v;
  __device__ __forceinline__ void zero() { v = make_float2(0.f, 0.f); }
  __device__ __forceinline__ void load(const float* p) { v = __ldg(reinterpret_cast<const float2*>(p)); }
  __device__ __forceinline__ void fma(float w, const LaneVec& d) {
    v.x = fmaf(w, d.v.x, v.x); v.y = fmaf(w, d.v.y, v.y);
  }
  __device__ __forceinline__ void store(float* p) const { *reinterpret_cast<float2*>(p) = v; }
};

template <int CJ>
__global__ void __launch_bounds__(kRcfWarps * 32, 8) scatter_reduce_cf2_kernel(const int32_t* __restrict__ offset,
                                                                 const int32_t* __restrict__ count,
                                                                 const int32_t* __restrict__ list,
                                                                 const float* __restrict__ wlist, ScatterPass pa,
                                                                 ScatterPass pb, int64_t Ea, float* __restrict__ grad,
                                                                 int64_t V, int64_t T) {
  constexpr int C = CJ * kWarp;
  constexpr int NS = C / 4;                       // 16 B slots per tile row
  constexpr int NW = kRcfWarps;
  static_assert(NW == 4, "a warp owns 8 of the tile's 32 voxels");
  __shared__ __align__(16) float tile[32 * C];
  __shared__ __align__(8) const float* ent_row[NW][kWarp];    // sorted batch: the entry's gradient row ...
  __shared__ float ent_w[NW][kWarp];                          // ... and its trilinear weight
  int lane;                                       // volatile: kept in a register, never re-read with S2R in the loops
  asm volatile("mov.u32 %0, %%laneid;" : "=r"(lane));
  const int wid = __shfl_sync(0xffffffffu, (int)(threadIdx.x / kWarp), 0);      // warp-uniform to the compiler too
  const int32_t Ea32 = (int32_t)Ea;               // E < 2^31 (checked by the caller)
  const int r4 = lane >> 3, l8 = lane & 7;
  const int64_t ntiles = T / 32;                  // kVec: V % 32 == 0
  // element (voxel vi, channel c) of the tile: row vi, slot (c / 4) ^ (vi / 4), word c % 4
  auto tile_at = [&](int vi, int c) -> float* {
    return tile + vi * C + ((((c >> 2) ^ (vi >> 2)) & (NS - 1)) << 2) + (c & 3);
  };
  auto row_of = [&](int32_t e) {                  // worked out by the lane that owns the entry, once
    return e < Ea32 ? pa.dlat + (int64_t)(e >> 3) * pa.ld : pb.dlat + (int64_t)((e - Ea32) >> 3) * pb.ld;
  };
  int cnt_n = 0, off_n = 0;
  if ((int64_t)blockIdx.x < ntiles) {
    const int64_t t = (int64_t)blockIdx.x * 32 + lane;
    cnt_n = count[t]; off_n = offset[t];
  }
  for (int64_t tl = blockIdx.x; tl < ntiles; tl += gridDim.x) {
    const int cnt_l = cnt_n, off_l = off_n;
    {
      const int64_t tn = (tl + gridDim.x) * 32 + lane;
      cnt_n = 0; off_n = 0;
      if (tn < T) { cnt_n = count[tn]; off_n = offset[tn]; }
    }
    const unsigned any = __ballot_sync(0xffffffffu, cnt_l > 0);   // uniform over the CTA (same counts in every warp)
    const uint32_t t0 = (uint32_t)tl * 32u, scene0 = t0 / (uint32_t)V;        // T < 2^31 (checked by the caller)
    // this lane's voxel quad, channel wid * 4 + r4; the write-out walks the channels in steps of 16
    float* gvec = grad + ((int64_t)scene0 * C + wid * 4 + r4) * V + (t0 - scene0 * (uint32_t)V) + 4 * l8;
    if (any == 0u) {
#pragma unroll
      for (int i = 0; i < C / (4 * NW); ++i)
        *reinterpret_cast<float4*>(gvec + (int64_t)(4 * NW * i) * V) = make_float4(0.f, 0.f, 0.f, 0.f);
      continue;
    }
    // lanes 0..7: count, list offset and inclusive prefix of this warp's voxels 4 j + wid (interleaved: the voxels a
    // ray crosses are runs along x, so consecutive ownership would leave one warp with the whole tile's work)
    const int cj = __shfl_sync(0xffffffffu, cnt_l, 4 * l8 + wid);
    const int oj = __shfl_sync(0xffffffffu, off_l, 4 * l8 + wid);
    int pin = cj;
#pragma unroll
    for (int d = 1; d < 8; d <<= 1) {
      const int o = __shfl_up_sync(0xffffffffu, pin, d, 8);
      if (l8 >= d) pin += o;
    }
    int vs = 0, pstart = 0;                       // first voxel not done yet, entries before it
    while (vs < 8) {
      const unsigned fit = __ballot_sync(0xffffffffu, lane < 8 && lane >= vs && pin - pstart <= kWarp);
      const int nv = __popc(fit);
      if (nv == 0) {                              // a voxel with more than a warp's worth of entries, on its own
        const int cnt = __shfl_sync(0xffffffffu, cj, vs);
        const int off = __shfl_sync(0xffffffffu, oj, vs);
        float acc[CJ];
        voxel_sum<CJ>(list + off, wlist + off, cnt, pa, pb, Ea, lane, acc);
#pragma unroll
        for (int j = 0; j < CJ; ++j) *tile_at(4 * vs + wid, lane + j * kWarp) = acc[j];
        pstart += cnt;
        vs += 1;
        continue;
      }
      const int ve = vs + nv;
      const int n = __shfl_sync(0xffffffffu, pin, ve - 1) - pstart;           // entries of the batch (may be 0)
      LaneVec<CJ> acc;
      acc.zero();
      int v = vs;
      if (n > 0) {
        int vi_l = vs;                            // this entry's voxel: the first whose inclusive prefix exceeds `lane`
        int maxc = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int pj = __shfl_sync(0xffffffffu, pin, j) - pstart;
          const int cc = __shfl_sync(0xffffffffu, cj, j);
          if (j >= vs && j < ve) {
            vi_l += (pj <= lane) ? 1 : 0;
            maxc = cc > maxc ? cc : maxc;
          }
        }
        if (vi_l > 7) vi_l = 7;                    // lanes past the batch: any valid source lane
        const int seg_end = __shfl_sync(0xffffffffu, pin, vi_l) - pstart;
        const int seg_beg = seg_end - __shfl_sync(0xffffffffu, cj, vi_l);
        const int src_l = __shfl_sync(0xffffffffu, oj, vi_l) + lane - seg_beg;   // a voxel's entries are contiguous
        int32_t e_l = 0x7fffffff;
        float w_l = 0.f;
        if (lane < n) { e_l = list[src_l]; w_l = wlist[src_l]; }
        int rank = 0;                              // entries are distinct: ranks within a voxel are a permutation
        for (int d = 0; d < maxc; ++d) {
          const int src = seg_beg + d;
          const int32_t ej = __shfl_sync(0xffffffffu, e_l, src & 31);
          rank += (src < seg_end && ej < e_l) ? 1 : 0;
        }
        __syncwarp();                              // the previous batch's reads of ent_*[] are done
        if (lane < n) {
          ent_row[wid][seg_beg + rank] = row_of(e_l);
          ent_w[wid][seg_beg + rank] = w_l;
        }
        __syncwarp();
        int vend = __shfl_sync(0xffffffffu, pin, v) - pstart;
        for (int k0 = 0; k0 < n; k0 += 4) {
          const float* row[4];
          float w[4];
          LaneVec<CJ> d[4];
          const float* const* er = &ent_row[wid][k0];      // k0 + q <= 31: n <= 32, k0 a multiple of 4
          const float* ewt = &ent_w[wid][k0];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            row[q] = er[q];
            w[q] = ewt[q];
          }
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (k0 + q < n) d[q].load(row[q] + lane * CJ);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            if (k0 + q < n) {
              while (k0 + q >= vend) {             // voxel v is complete (or empty): its row goes into the tile
                acc.store(tile_at(4 * v + wid, lane * CJ));
                acc.zero();
                ++v;
                vend = __shfl_sync(0xffffffffu, pin, v & 7) - pstart;
              }
              acc.fma(w[q], d[q]);
            }
          }
        }
      }
      for (; v < ve; ++v) {                        // the last voxel with entries, then trailing empty ones
        acc.store(tile_at(4 * v + wid, lane * CJ));
        acc.zero();
      }
      pstart += n;
      vs = ve;
    }
    __syncthreads();
    // channel c = wid * 4 + r4 + 16 i of voxels 4 l8 + t: slot (c / 4) ^ l8 = ((wid ^ (l8 & 3)) | ((i ^ (l8 >> 2)) << 2)),
    // so the i-dependence is +-16 words on one of two per-thread bases (even / odd i) and every LDS has an immediate offset
    {
      const int h = l8 >> 2;
      const float* rd_even = tile + (4 * l8) * C + 4 * (wid ^ (l8 & 3)) + r4 + 16 * h;
      const float* rd_odd = rd_even - 32 * h;
#pragma unroll
      for (int i = 0; i < C / (4 * NW); ++i) {
        const float* rd = ((i & 1) ? rd_odd : rd_even) + 16 * i;
        float4 x;
        x.x = rd[0]; x.y = rd[C]; x.z = rd[2 * C]; x.w = rd[3 * C];
        *reinterpret_cast<float4*>(gvec + (int64_t)(4 * NW * i) * V) = x;
      }
    }
    __syncthreads();
  }
}

// Channels-last output (the producer works in torch.channels_last_3d): one warp per voxel row.
template <int CJ>
__global__ void __launch_bounds__(256) scatter_reduce_cl_kernel(const int32_t* __restrict__ offset,
                                                                const int32_t* __restrict__ count,
                                                                const int32_t* __restrict__ list,
                                                                const float* __restrict__ wlist, ScatterPass pa,
                                                                ScatterPass pb, int64_t Ea, float* __restrict__ grad,
                                                                int64_t T) {
  constexpr int C = CJ * kWarp;
  const int lane = threadIdx.x % kWarp;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / kWarp;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) / kWarp;
  for (int64_t t = warp; t < T; t += nwarps) {
    const int cnt = count[t];
    float acc[CJ];
    if (cnt > 0) {
      { const int off = offset[t]; voxel_sum<CJ>(list + off, wlist + off, cnt, pa, pb, Ea, lane, acc); }
    } else {
#pragma unroll
      for (int j = 0; j < CJ; ++j) acc[j] = 0.f;
    }
#pragma unroll
    for (int j = 0; j < CJ; ++j) grad[t * C + lane + j * kWarp] = acc[j];
  }
}

}  // namespace nrf

extern "C" int nrf_scatter_volume_grad_merged(const float* rays, int R, int rays_per_scene, const float* z_a, int K_a,
                                              const float* dlat_a, int ld_a, const float* z_b, int K_b,
                                              const float* dlat_b, int ld_b, float* grad, int channels_first, int SB,
                                              int C, int S0, int S1, int S2, const float* bounds_host,
                                              void* workspace, void* stream) {
  NRF_REQUIRE(rays && z_a && dlat_a && grad && bounds_host && workspace, NRF_EINVAL,
              "nrf_scatter_volume_grad_merged: null pointer");
  NRF_REQUIRE(R > 0 && K_a > 0 && R == SB * rays_per_scene, NRF_EINVAL,
              "nrf_scatter_volume_grad_merged: R != SB*rays_per_scene");
  const bool two = z_b != nullptr;
  NRF_REQUIRE(!two || (dlat_b && K_b > 0 && ld_b >= C), NRF_EINVAL, "nrf_scatter_volume_grad_merged: pass b");
  NRF_REQUIRE((C == 64 || C == 128) && ld_a >= C, NRF_ENOSUP,
              "nrf_scatter_volume_grad_merged: C=%d (64 or 128 channels; use nrf_scatter_volume_grad_sorted)", C);
  int64_t V = (int64_t)S0 * S1 * S2, T = (int64_t)SB * V;
  int64_t Na = (int64_t)R * K_a, Nb = two ? (int64_t)R * K_b : 0, N = Na + Nb, E = N * 8;
  NRF_REQUIRE(T < ((int64_t)1 << 31) && E < ((int64_t)1 << 31), NRF_ENOSUP,
              "nrf_scatter_volume_grad_merged: more than 2^31 voxels or entries");
  NRF_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 15) == 0, NRF_EINVAL,
              "nrf_scatter_volume_grad_merged: workspace must be 16 B aligned");
  cudaStream_t s = as_stream(stream);
  int64_t nb = (T + 1023) / 1024;
  // count[T] cursor[T] offset[T] block_sums[nb+1] | 16 B aligned: key[E] wts[E] list[E] wlist[E]
  int32_t* count = reinterpret_cast<int32_t*>(workspace);
  int32_t* cursor = count + T;
  int32_t* offset = cursor + T;
  int32_t* block_sums = offset + T;
  int64_t head = (3 * T + nb + 1 + 3) & ~(int64_t)3;
  int32_t* key = count + head;
  float* wts = reinterpret_cast<float*>(key + E);
  int32_t* list = reinterpret_cast<int32_t*>(wts + E);
  float* wlist = reinterpret_cast<float*>(list + E);
  NRF_CUDA_OK(cudaMemsetAsync(count, 0, (size_t)(2 * T) * 4, s));       // count and cursor
  ScatterArgs g0;
  g0.S0 = S0; g0.S1 = S1; g0.S2 = S2; g0.C = C; g0.SB = SB;
  fill_bounds(bounds_host, g0.bmin, g0.bext);
  ScatterPass pa{z_a, dlat_a, K_a, ld_a};
  ScatterPass pb{two ? z_b : z_a, two ? dlat_b : dlat_a, two ? K_b : K_a, two ? ld_b : ld_a};
  { LaunchScope ls_(NRF_CAT_SCATTER, s);
    scatter_count2_kernel<<<(unsigned)((N + 255) / 256), 256, 0, s>>>(rays, R, rays_per_scene, pa, pb, Na, Nb, g0, V,
                                                                      key, wts, count); }
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
    scatter_fill2_kernel<<<(unsigned)((E + 255) / 256), 256, 0, s>>>(key, wts, E, offset, cursor, list, wlist); }
  NRF_LAUNCH_OK();
  const int64_t Ea = Na * 8;
  { LaunchScope ls_(NRF_CAT_SCATTER, s);
    if (channels_first) {
      int64_t ntiles = (T + 31) / 32;
      int max_blocks = sm_count() * 10;
      int blocks = (int)(ntiles < max_blocks ? ntiles : max_blocks);
      const bool vec = V % 32 == 0 && (reinterpret_cast<uintptr_t>(grad) & 15) == 0;
      // rows read as one vector per lane: 16 B aligned gradient rows in both passes
      const bool rowvec = vec && ld_a % 4 == 0 && (reinterpret_cast<uintptr_t>(dlat_a) & 15) == 0 &&
                          (!two || (ld_b % 4 == 0 && (reinterpret_cast<uintptr_t>(dlat_b) & 15) == 0));
      const char* env = getenv("NRF_SCATTER_RCF");            // "1": the one-voxel-per-warp kernel (A/B, tests)
      const bool batched = rowvec && !(env && atoi(env) == 1);
#define NRF_RCF(CJ, VEC) scatter_reduce_cf_kernel<CJ, VEC><<<blocks, kRcfWarps * 32, 0, s>>>(offset, count, list, wlist, pa, pb, Ea, grad, V, T)
#define NRF_RCF2(CJ) scatter_reduce_cf2_kernel<CJ><<<blocks8, kRcfWarps * 32, 0, s>>>(offset, count, list, wlist, pa, pb, Ea, grad, V, T)
      const int max8 = sm_count() * 8;
      const int blocks8 = (int)(ntiles < max8 ? ntiles : max8);
      if (batched) { if (C == 128) NRF_RCF2(4); else NRF_RCF2(2); }
      else if (C == 128) { if (vec) NRF_RCF(4, true); else NRF_RCF(4, false); }
      else { if (vec) NRF_RCF(2, true); else NRF_RCF(2, false); }
#undef NRF_RCF
#undef NRF_RCF2
    } else {
      int64_t want = (T + 7) / 8;
      int max_blocks = sm_count() * 32;
      int blocks = (int)(want < max_blocks ? want : max_blocks);
      if (C == 128) scatter_reduce_cl_kernel<4><<<blocks, 256, 0, s>>>(offset, count, list, wlist, pa, pb, Ea, grad, T);
      else scatter_reduce_cl_kernel<2><<<blocks, 256, 0, s>>>(offset, count, list, wlist, pa, pb, Ea, grad, T);
    }
  }
  NRF_LAUNCH_OK();
  return NRF_OK;
}

extern "C" int64_t nrf_scatter_sorted_workspace_bytes(int64_t N, int SB, int64_t V) {
  int64_t T = (int64_t)SB * V, E = N * 8;
  int64_t nb = (T + 1023) / 1024;
  // count[T] cursor[T] offset[T] block_sums[nb+1] key[E] list[E] wts[E] (+ wlist[E] for the merged variant)
  return (3 * T + nb + 1 + 2 * E) * 4 + 2 * E * 4 + 1024;
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
