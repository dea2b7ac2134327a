// Shared helpers for the sm_100a kernels behind include/nrf_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/nrf_b200.h"

namespace nrf {

void set_error(const char* fmt, ...);

#define NRF_REQUIRE(cond, code, ...)                 \
  do {                                               \
    if (!(cond)) {                                   \
      nrf::set_error(__VA_ARGS__);                   \
      return (code);                                 \
    }                                                \
  } while (0)

#define NRF_CUDA_OK(expr)                                                        \
  do {                                                                           \
    cudaError_t e__ = (expr);                                                    \
    if (e__ != cudaSuccess) {                                                    \
      nrf::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__),    \
                     __FILE__, __LINE__);                                        \
      return NRF_ECUDA;                                                          \
    }                                                                            \
  } while (0)

#define NRF_LAUNCH_OK() NRF_CUDA_OK(cudaGetLastError())

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

int sm_count();

// Counts every kernel launch and, while nrf_timing_begin() is active, brackets it with an event pair.
struct LaunchScope {
  LaunchScope(int category, cudaStream_t stream);
  ~LaunchScope();
  int idx;
  cudaStream_t stream;
};

constexpr int kWarp = 32;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Geometry shared by the gather and the scatter: everything eager PyTorch rounds separately is
// rounded separately here (__f*_rn intrinsics are never contracted into FMAs).
struct Corner8 {
  int64_t off[8];   // element offset of the corner's channel vector, -1 if outside the grid
  float w[8];
};

struct SampleGeom {
  float px, py, pz;   // world point o + z*d                (neural_rendering.py:246)
  float cx, cy, cz;   // canonical coords in [0,1]          (models_embed.py:200-201)
};

__device__ __forceinline__ SampleGeom sample_geometry(const float* __restrict__ ray, float z,
                                                      const float bmin[3], const float bext[3]) {
  SampleGeom g;
  g.px = __fadd_rn(ray[0], __fmul_rn(z, ray[3]));
  g.py = __fadd_rn(ray[1], __fmul_rn(z, ray[4]));
  g.pz = __fadd_rn(ray[2], __fmul_rn(z, ray[5]));
  g.cx = __fdiv_rn(__fsub_rn(g.px, bmin[0]), bext[0]);
  g.cy = __fdiv_rn(__fsub_rn(g.py, bmin[1]), bext[1]);
  g.cz = __fdiv_rn(__fsub_rn(g.pz, bmin[2]), bext[2]);
  return g;
}

// ATen grid_sampler_3d, bilinear, align_corners=True, zeros padding.
// Corner order tnw, tne, tsw, tse, bnw, bne, bsw, bse; x -> S2 (fastest), y -> S1, z -> S0.
// TriSetup is the compact per-sample form (what one lane hands to the warp): low-corner indices and
// the six one-dimensional weights; a corner's weight is (wx*wy)*wz, rounded in that order.
struct TriSetup {
  int x0, y0, z0;     // floor indices, clamped to >= -2 (both corners of that axis out of range)
  float wx0, wx1, wy0, wy1, wz0, wz1;
  bool finite;
};

__device__ __forceinline__ TriSetup trilinear_setup(float cx, float cy, float cz, int S0, int S1, int S2) {
  float gx = __fsub_rn(__fmul_rn(cx, 2.0f), 1.0f);
  float gy = __fsub_rn(__fmul_rn(cy, 2.0f), 1.0f);
  float gz = __fsub_rn(__fmul_rn(cz, 2.0f), 1.0f);
  float ix = __fmul_rn(__fdiv_rn(__fadd_rn(gx, 1.0f), 2.0f), (float)(S2 - 1));
  float iy = __fmul_rn(__fdiv_rn(__fadd_rn(gy, 1.0f), 2.0f), (float)(S1 - 1));
  float iz = __fmul_rn(__fdiv_rn(__fadd_rn(gz, 1.0f), 2.0f), (float)(S0 - 1));
  float x0 = floorf(ix), y0 = floorf(iy), z0 = floorf(iz);
  float x1 = x0 + 1.0f, y1 = y0 + 1.0f, z1 = z0 + 1.0f;
  TriSetup t;
  t.wx0 = __fsub_rn(x1, ix); t.wx1 = __fsub_rn(ix, x0);
  t.wy0 = __fsub_rn(y1, iy); t.wy1 = __fsub_rn(iy, y0);
  t.wz0 = __fsub_rn(z1, iz); t.wz1 = __fsub_rn(iz, z0);
  // Points far outside the box can overflow int conversion; clamp the float first.
  auto toi = [](float v) { return (int)fminf(fmaxf(v, -2.0f), 1.0e9f); };
  t.x0 = toi(x0); t.y0 = toi(y0); t.z0 = toi(z0);
  t.finite = (ix == ix) && (iy == iy) && (iz == iz);
  return t;
}

// true if at least one of the 8 corners lies inside the grid
__device__ __forceinline__ bool trilinear_touches(const TriSetup& t, int S0, int S1, int S2) {
  return t.finite && t.x0 >= -1 && t.x0 < S2 && t.y0 >= -1 && t.y0 < S1 && t.z0 >= -1 && t.z0 < S0;
}

__device__ __forceinline__ void corners_from_setup(const TriSetup& t, int S0, int S1, int S2, int C, Corner8& c8) {
  const float wx[2] = {t.wx0, t.wx1}, wy[2] = {t.wy0, t.wy1}, wz[2] = {t.wz0, t.wz1};
  const int xi[2] = {t.x0, t.x0 + 1}, yi[2] = {t.y0, t.y0 + 1}, zi[2] = {t.z0, t.z0 + 1};
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    int a = k & 1, b = (k >> 1) & 1, c = (k >> 2) & 1;
    bool ok = t.finite && xi[a] >= 0 && xi[a] < S2 && yi[b] >= 0 && yi[b] < S1 && zi[c] >= 0 && zi[c] < S0;
    c8.off[k] = ok ? (((int64_t)zi[c] * S1 + yi[b]) * S2 + xi[a]) * (int64_t)C : (int64_t)-1;
    c8.w[k] = __fmul_rn(__fmul_rn(wx[a], wy[b]), wz[c]);
  }
}

__device__ __forceinline__ void trilinear_corners(float cx, float cy, float cz, int S0, int S1, int S2,
                                                  int C, Corner8& c8) {
  corners_from_setup(trilinear_setup(cx, cy, cz, S0, S1, S2), S0, S1, S2, C, c8);
}

}  // namespace nrf
