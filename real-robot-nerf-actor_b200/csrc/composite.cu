// Alpha compositing of RGB / depth / D-channel feature along each ray, forward and closed-form
// backward.  One warp per ray; transmittance is a warp-shuffle product scan.
//   forward  <- neural_rendering.py:239-243 (deltas), :339-359 (alphas, cumprod, weighted sums),
//               models_embed.py:444-466 (sigmoid / relu heads, applied here to the raw MLP outputs)
//   backward <- autograd of the same lines, closed form in SURVEY.md 9.2
#include "common.cuh"

namespace nrf {

constexpr int kMaxPerLane = 32;   // K <= 1024

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + __expf(-x)); }

// Where the field-output row (and the gradient row) of sorted sample k of ray r lives.  Plain pass: row r*K + k of one
// buffer.  Fine pass that reuses the coarse pass's evaluations (perm != NULL): perm[r][k] = p indexes the ray's
// [Ka coarse | Kb new] samples before the sort: p < Ka -> row r*Ka + p of buffer a (the coarse pass's outputs),
// else row r*Kb + (p - Ka) of buffer b (the newly evaluated samples).
struct Rows {
  const float* fa;
  const float* fb;
  const int* perm;     // this ray's row of the permutation, or nullptr
  int Ka, Kb, ld;
  int64_t r;
  __device__ __forceinline__ int64_t row(int k, bool& first) const {
    if (!perm) { first = true; return r * Ka + k; }
    const int p = perm[k];
    first = p < Ka;
    return first ? r * Ka + p : r * Kb + (p - Ka);
  }
  __device__ __forceinline__ const float* at(int k) const {
    bool first;
    const int64_t i = row(k, first);
    return (first ? fa : fb) + i * ld;
  }
};

struct NrfReuseDev {       // device copy of NrfCompositeReuse (+ the accumulate flag of nrf_composite_bwd)
  const float* field_new;
  const int* perm;
  int n_first;
  void* d_field_new;
  int accumulate;
};

__device__ __forceinline__ Rows make_rows(const float* field, int ldo, int r, int K, const NrfReuseDev& ru) {
  Rows q;
  q.fa = field; q.fb = ru.field_new; q.ld = ldo; q.r = r;
  if (ru.perm) { q.perm = ru.perm + (int64_t)r * K; q.Ka = ru.n_first; q.Kb = K - ru.n_first; }
  else { q.perm = nullptr; q.Ka = K; q.Kb = 0; }
  return q;
}

// Computes, for the warp's ray, alpha_k / T_k / w_k for the lane's contiguous chunk of samples.
// chunk = ceil(K/32); lane owns k in [lane*chunk, min(K,(lane+1)*chunk)).
struct RayScan {
  int k0, k1;
};

// Exclusive product scan across lanes of `local` (the product of a lane's chunk).
__device__ __forceinline__ float warp_exclusive_prod(float local, int lane) {
  float incl = local;
#pragma unroll
  for (int o = 1; o < kWarp; o <<= 1) {
    float v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl *= v;
  }
  float excl = __shfl_up_sync(0xffffffffu, incl, 1);
  return lane == 0 ? 1.0f : excl;
}

// smem per warp: w[K], aux[K]
template <bool kBackward>
__device__ __forceinline__ void ray_weights(const Rows& rows,
                                            const float* __restrict__ zrow, float far, int K, int lane,
                                            float* __restrict__ s_w, float* __restrict__ s_alpha,
                                            float* __restrict__ s_T, float* __restrict__ s_delta,
                                            const float* __restrict__ nz) {
  int chunk = (K + kWarp - 1) / kWarp;
  int k0 = lane * chunk, k1 = min(K, k0 + chunk);
  float local = 1.0f;
  for (int k = k0; k < k1; ++k) {
    float zk = zrow[k];
    float delta = (k + 1 < K ? zrow[k + 1] : far) - zk;
    float sigma = fmaxf(rows.at(k)[3], 0.0f);                          // models_embed.py:464
    if (nz) sigma = fmaxf(sigma + nz[k], 0.0f);                        // neural_rendering.py:336-339 (training noise)
    float alpha = 1.0f - expf(-delta * sigma);
    s_alpha[k] = alpha;
    if (kBackward) s_delta[k] = delta;
    local *= (1.0f - alpha) + 1e-10f;
  }
  float T = warp_exclusive_prod(local, lane);
  for (int k = k0; k < k1; ++k) {
    float alpha = s_alpha[k];
    s_w[k] = alpha * T;
    if (kBackward) s_T[k] = T;
    T *= (1.0f - alpha) + 1e-10f;
  }
  __syncwarp();
}

// field row = [r g b sigma | embed(D)], D % 4 == 0 and ldo % 4 == 0 -> float4 everywhere.
// One CTA of 4 warps per ray: warp 0 runs the transmittance scan, then warp w accumulates the samples
// k = w, w+4, ... and the four partial sums meet in shared memory.
constexpr int kRayWarps = 4;
__global__ void __launch_bounds__(128) composite_fwd_kernel(
    const float* __restrict__ field, int ldo, const float* __restrict__ z, const float* __restrict__ rays,
    int R, int K, int D, int white_bkgd, float* __restrict__ weights, float* __restrict__ rgb,
    float* __restrict__ embed, float* __restrict__ depth, const float* __restrict__ sig_noise, NrfReuseDev ru) {
  extern __shared__ float smem[];
  const int wid = threadIdx.x / kWarp, lane = threadIdx.x % kWarp;
  const int r = blockIdx.x;
  const float* nz = sig_noise ? sig_noise + (int64_t)r * K : nullptr;
  float* s_w = smem;
  float* s_alpha = s_w + K;
  float* s_part = smem + ((2 * K + 3) & ~3);   // kRayWarps x (4 + D + 4) partial sums, 16 B aligned
  const Rows rows = make_rows(field, ldo, r, K, ru);
  const float* zrow = z + (int64_t)r * K;
  float far = rays[(int64_t)r * 8 + 7];
  if (wid == 0) ray_weights<false>(rows, zrow, far, K, lane, s_w, s_alpha, nullptr, nullptr, nz);
  __syncthreads();

  const int nvec = (4 + D) / 4;             // float4 per row
  constexpr int kMaxVec = 8;                // up to (4+D) <= 1024 channels
  float4 acc[kMaxVec];
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  float dsum = 0.f, wsum = 0.f;
  for (int k = wid; k < K; k += kRayWarps) {
    float w = s_w[k];
    const float4* row = reinterpret_cast<const float4*>(rows.at(k));
#pragma unroll
    for (int i = 0; i < kMaxVec; ++i) {
      int v = lane + i * kWarp;
      if (v < nvec) {
        float4 x = __ldg(row + v);
        if (v == 0) {
          x.x = sigmoidf_(x.x); x.y = sigmoidf_(x.y); x.z = sigmoidf_(x.z); x.w = 0.f;
        }
        acc[i].x = fmaf(w, x.x, acc[i].x);
        acc[i].y = fmaf(w, x.y, acc[i].y);
        acc[i].z = fmaf(w, x.z, acc[i].z);
        acc[i].w = fmaf(w, x.w, acc[i].w);
      }
    }
    if (lane == 0) {
      dsum = fmaf(w, zrow[k], dsum);
      wsum += w;
    }
  }
  const int pstride = 4 + D + 4;
  float* mine = s_part + wid * pstride;
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i) {
    int v = lane + i * kWarp;
    if (v < nvec) *reinterpret_cast<float4*>(mine + v * 4) = acc[i];
  }
  if (lane == 0) { mine[4 + D] = dsum; mine[4 + D + 1] = wsum; }
  __syncthreads();
  for (int k = threadIdx.x; k < K; k += blockDim.x) weights[(int64_t)r * K + k] = s_w[k];
  for (int c = threadIdx.x; c < 4 + D + 2; c += blockDim.x) {
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < kRayWarps; ++w) v += s_part[w * pstride + c];
    if (c < 3) {
      float ws = 0.f;
#pragma unroll
      for (int w = 0; w < kRayWarps; ++w) ws += s_part[w * pstride + 4 + D + 1];
      rgb[(int64_t)r * 3 + c] = v + (white_bkgd ? 1.0f - ws : 0.0f);
    } else if (c == 3) {
    } else if (c < 4 + D) {
      embed[(int64_t)r * D + (c - 4)] = v;
    } else if (c == 4 + D) {
      depth[r] = v;
    }
  }
}

template <typename T> __device__ __forceinline__ float4 load_grad4(const T* p);
template <> __device__ __forceinline__ float4 load_grad4<float>(const float* p) {
  return *reinterpret_cast<const float4*>(p);
}
template <> __device__ __forceinline__ float4 load_grad4<__nv_bfloat16>(const __nv_bfloat16* p) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&u.x), b = *reinterpret_cast<const __nv_bfloat162*>(&u.y);
  const float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
  return make_float4(fa.x, fa.y, fb.x, fb.y);
}
template <typename T> __device__ __forceinline__ void store_grad4(T* p, float4 v);
template <> __device__ __forceinline__ void store_grad4<float>(float* p, float4 v) {
  *reinterpret_cast<float4*>(p) = v;
}
template <> __device__ __forceinline__ void store_grad4<__nv_bfloat16>(__nv_bfloat16* p, float4 v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
  uint2 u;
  u.x = *reinterpret_cast<uint32_t*>(&a);
  u.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = u;
}

// acc: the row already holds this sample's gradient from the fine pass that reused it: add to it.
template <typename T> __device__ __forceinline__ void put_grad4(T* p, float4 v, bool acc) {
  if (acc) {
    const float4 o = load_grad4<T>(p);
    v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
  }
  store_grad4<T>(p, v);
}

// The gradient row of sorted sample k (see Rows).
template <typename T>
__device__ __forceinline__ T* grad_row(const Rows& rows, T* d_a, void* d_b, int ldg, int k) {
  bool first;
  const int64_t i = rows.row(k, first);
  return (first ? d_a : reinterpret_cast<T*>(d_b)) + i * ldg;
}

// Backward.  With g_k = <d_rgb, rgb_k> + <d_embed, e_k> + d_depth z_k (+ d_w_k):
//   dL/dsigma_k = 1[raw>0] delta_k (1-alpha_k) ( T_k g_k - S_k / (1-alpha_k+eps) ),  S_k = sum_{j>k} w_j g_j
//   dL/draw_rgb = w_k d_rgb s(1-s);  dL/de_k = w_k d_embed
//   white_bkgd adds (-sum_c d_rgb_c) to every g_k.
//   dL/dz_k (optional) = d_depth w_k + dL/ddelta_{k-1} - dL/ddelta_k,
//   dL/ddelta_k = relu(sigma_k)(1-alpha_k)(same bracket).
template <typename T>
__global__ void __launch_bounds__(128) composite_bwd_kernel(
    const float* __restrict__ field, int ldo, const float* __restrict__ z, const float* __restrict__ rays,
    int R, int K, int D, int white_bkgd, const float* __restrict__ d_rgb,
    const float* __restrict__ d_embed, const float* __restrict__ d_depth,
    const float* __restrict__ d_weights, T* __restrict__ d_field, int ldg, float* __restrict__ d_z,
    const float* __restrict__ sig_noise, NrfReuseDev ru) {
  // One CTA of 4 warps per ray: scans on warp 0, the per-sample row work split over the 4 warps.
  extern __shared__ float smem[];
  const int wid = threadIdx.x / kWarp, lane = threadIdx.x % kWarp;
  const int r = blockIdx.x;
  const float* nz = sig_noise ? sig_noise + (int64_t)r * K : nullptr;
  float* s_w = smem;
  float* s_alpha = s_w + K;
  float* s_T = s_alpha + K;
  float* s_delta = s_T + K;
  float* s_g = s_delta + K;
  float* s_ds = s_g + K;   // dL/dsigma (pre relu-gate) then reused
  const Rows rows = make_rows(field, ldo, r, K, ru);
  const float* zrow = z + (int64_t)r * K;
  float far = rays[(int64_t)r * 8 + 7];
  if (wid == 0) ray_weights<true>(rows, zrow, far, K, lane, s_w, s_alpha, s_T, s_delta, nz);
  __syncthreads();

  const int nvec = (4 + D) / 4;
  constexpr int kMaxVec = 8;
  // per-lane slice of the output gradients
  float4 dout[kMaxVec];
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i) {
    int v = lane + i * kWarp;
    dout[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (v < nvec) {
      if (v == 0) {
        dout[i] = make_float4(d_rgb[(int64_t)r * 3 + 0], d_rgb[(int64_t)r * 3 + 1],
                              d_rgb[(int64_t)r * 3 + 2], 0.f);
      } else {
        dout[i] = *reinterpret_cast<const float4*>(d_embed + (int64_t)r * D + (v - 1) * 4);
      }
    }
  }
  float dd = d_depth ? d_depth[r] : 0.0f;
  float bk = 0.f;
  if (white_bkgd) bk = -(d_rgb[(int64_t)r * 3 + 0] + d_rgb[(int64_t)r * 3 + 1] + d_rgb[(int64_t)r * 3 + 2]);

  // pass A: g_k
  for (int k = wid; k < K; k += kRayWarps) {
    const float4* row = reinterpret_cast<const float4*>(rows.at(k));
    float part = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxVec; ++i) {
      int v = lane + i * kWarp;
      if (v < nvec) {
        float4 x = __ldg(row + v);
        if (v == 0) {
          x.x = sigmoidf_(x.x); x.y = sigmoidf_(x.y); x.z = sigmoidf_(x.z); x.w = 0.f;
        }
        part += x.x * dout[i].x + x.y * dout[i].y + x.z * dout[i].z + x.w * dout[i].w;
      }
    }
    part = warp_sum(part);
    if (lane == 0) {
      float g = part + dd * zrow[k] + bk;
      if (d_weights) g += d_weights[(int64_t)r * K + k];
      s_g[k] = g;
    }
  }
  __syncthreads();
  // suffix sums S_k = sum_{j>k} w_j g_j : lane-chunked reverse scan (warp 0)
  if (wid == 0) {
    int chunk = (K + kWarp - 1) / kWarp;
    int k0 = lane * chunk, k1 = min(K, k0 + chunk);
    float local = 0.f;
    for (int k = k0; k < k1; ++k) local += s_w[k] * s_g[k];
    // inclusive suffix over lanes
    float incl = local;
#pragma unroll
    for (int o = 1; o < kWarp; o <<= 1) {
      float v = __shfl_down_sync(0xffffffffu, incl, o);
      if (lane + o < kWarp) incl += v;
    }
    float run = incl - local;   // sum over later lanes
    for (int k = k1 - 1; k >= k0; --k) {
      float alpha = s_alpha[k];
      float one_m = 1.0f - alpha;
      float bracket = s_T[k] * s_g[k] - run / (one_m + 1e-10f);
      run += s_w[k] * s_g[k];
      s_ds[k] = one_m * bracket;     // multiply by delta (for dsigma) or relu(sigma) (for ddelta) later
    }
  }
  __syncthreads();
  // pass B: write d_field rows
  for (int k = wid; k < K; k += kRayWarps) {
    float w = s_w[k];
    const float4* row = reinterpret_cast<const float4*>(rows.at(k));
    const bool acc = ru.accumulate != 0;
    T* grow = grad_row<T>(rows, d_field, ru.d_field_new, ldg, k);
#pragma unroll
    for (int i = 0; i < kMaxVec; ++i) {
      int v = lane + i * kWarp;
      if (v < nvec) {
        float4 gq;
        if (v == 0) {
          float4 x = __ldg(row);
          float sr = sigmoidf_(x.x), sg = sigmoidf_(x.y), sb = sigmoidf_(x.z);
          gq.x = w * dout[i].x * sr * (1.0f - sr);
          gq.y = w * dout[i].y * sg * (1.0f - sg);
          gq.z = w * dout[i].z * sb * (1.0f - sb);
          const bool open = x.w > 0.0f && (!nz || x.w + nz[k] > 0.0f);   // both ReLUs pass
          gq.w = open ? s_delta[k] * s_ds[k] : 0.0f;
        } else {
          gq = make_float4(w * dout[i].x, w * dout[i].y, w * dout[i].z, w * dout[i].w);
        }
        put_grad4<T>(grow + v * 4, gq, acc);
      }
    }
    if (!acc)
      for (int c = 4 + D + lane; c < ldg; c += kWarp) grow[c] = T(0.0f);
  }
  if (d_z) {
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
      auto sig_at = [&](int q) {
        float sg = fmaxf(rows.at(q)[3], 0.0f);
        return nz ? fmaxf(sg + nz[q], 0.0f) : sg;
      };
      float dl_k = sig_at(k) * s_ds[k];
      float dl_km1 = 0.f;
      if (k > 0) dl_km1 = sig_at(k - 1) * s_ds[k - 1];
      d_z[(int64_t)r * K + k] = dd * s_w[k] + dl_km1 - dl_k;
    }
  }
}


// ---------------------------------------------------------------------------------------------------------------
// Fast paths for D = 128 * DV (384 at the BASELINE dims, 512 in nerfact.conf): the D feature channels are exactly DV
// float4 per lane, the [r g b sigma] head is handled by lane 0 outside the vector loop, nothing is predicated.
// The generic kernels above spend ~380 warp instructions per sample on predicates for the ragged 97-float4 row;
// these need ~70 and read every field row exactly once.
template <int DV>
__global__ void __launch_bounds__(128) composite_fwd_fast_kernel(
    const float* __restrict__ field, int ldo, const float* __restrict__ z, const float* __restrict__ rays,
    int K, int white_bkgd, float* __restrict__ weights, float* __restrict__ rgb, float* __restrict__ embed,
    float* __restrict__ depth, const float* __restrict__ sig_noise, NrfReuseDev ru) {
  constexpr int D = 128 * DV;
  extern __shared__ float smem[];
  const int wid = threadIdx.x / kWarp, lane = threadIdx.x % kWarp;
  const int r = blockIdx.x;
  const float* nz = sig_noise ? sig_noise + (int64_t)r * K : nullptr;
  float* s_w = smem;
  float* s_alpha = s_w + K;
  float* s_part = smem + ((2 * K + 3) & ~3);   // kRayWarps x (D + 8): embed sums, then rgb(3), depth, wsum
  const Rows rows = make_rows(field, ldo, r, K, ru);
  const float* zrow = z + (int64_t)r * K;
  const float far = rays[(int64_t)r * 8 + 7];
  if (wid == 0) ray_weights<false>(rows, zrow, far, K, lane, s_w, s_alpha, nullptr, nullptr, nz);
  __syncthreads();
  float4 acc[DV];
#pragma unroll
  for (int i = 0; i < DV; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  float hr = 0.f, hg = 0.f, hb = 0.f, dsum = 0.f, wsum = 0.f;
#pragma unroll 2
  for (int k = wid; k < K; k += kRayWarps) {
    const float w = s_w[k];
    const float4* row = reinterpret_cast<const float4*>(rows.at(k));
    float4 x[DV];
#pragma unroll
    for (int i = 0; i < DV; ++i) x[i] = __ldg(row + 1 + lane + i * kWarp);
    if (lane == 0) {
      const float4 h = __ldg(row);
      hr = fmaf(w, sigmoidf_(h.x), hr);
      hg = fmaf(w, sigmoidf_(h.y), hg);
      hb = fmaf(w, sigmoidf_(h.z), hb);
      dsum = fmaf(w, zrow[k], dsum);
      wsum += w;
    }
#pragma unroll
    for (int i = 0; i < DV; ++i) {
      acc[i].x = fmaf(w, x[i].x, acc[i].x);
      acc[i].y = fmaf(w, x[i].y, acc[i].y);
      acc[i].z = fmaf(w, x[i].z, acc[i].z);
      acc[i].w = fmaf(w, x[i].w, acc[i].w);
    }
  }
  constexpr int pstride = D + 8;
  float* mine = s_part + wid * pstride;
#pragma unroll
  for (int i = 0; i < DV; ++i) *reinterpret_cast<float4*>(mine + (lane + i * kWarp) * 4) = acc[i];
  if (lane == 0) { mine[D] = hr; mine[D + 1] = hg; mine[D + 2] = hb; mine[D + 3] = dsum; mine[D + 4] = wsum; }
  __syncthreads();
  for (int k = threadIdx.x; k < K; k += blockDim.x) weights[(int64_t)r * K + k] = s_w[k];
  for (int c = threadIdx.x; c < D + 4; c += blockDim.x) {
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < kRayWarps; ++w) v += s_part[w * pstride + c];
    if (c < D) {
      embed[(int64_t)r * D + c] = v;
    } else if (c < D + 3) {
      float ws = 0.f;
#pragma unroll
      for (int w = 0; w < kRayWarps; ++w) ws += s_part[w * pstride + D + 4];
      rgb[(int64_t)r * 3 + (c - D)] = v + (white_bkgd ? 1.0f - ws : 0.0f);
    } else {
      depth[r] = v;
    }
  }
}

template <int DV, typename T>
__global__ void __launch_bounds__(128) composite_bwd_fast_kernel(
    const float* __restrict__ field, int ldo, const float* __restrict__ z, const float* __restrict__ rays,
    int K, int white_bkgd, const float* __restrict__ d_rgb, const float* __restrict__ d_embed,
    const float* __restrict__ d_depth, const float* __restrict__ d_weights, T* __restrict__ d_field, int ldg,
    float* __restrict__ d_z, const float* __restrict__ sig_noise, NrfReuseDev ru) {
  constexpr int D = 128 * DV;
  extern __shared__ float smem[];
  const int wid = threadIdx.x / kWarp, lane = threadIdx.x % kWarp;
  const int r = blockIdx.x;
  const float* nz = sig_noise ? sig_noise + (int64_t)r * K : nullptr;
  float* s_w = smem;
  float* s_alpha = s_w + K;
  float* s_T = s_alpha + K;
  float* s_delta = s_T + K;
  float* s_g = s_delta + K;
  float* s_ds = s_g + K;       // (1 - alpha) * bracket
  float* s_sig = s_ds + K;     // the density that went into alpha_k: relu(raw) (+ noise, relu again)
  float* s_h = s_sig + K;      // 3 per sample: d_rgb_c * s(1-s), the sigmoid-head factors
  float* s_open = s_h + 3 * K; // 1 where the gradient reaches the raw density output (every ReLU on the way passes)
  const Rows rows = make_rows(field, ldo, r, K, ru);
  const float* zrow = z + (int64_t)r * K;
  const float far = rays[(int64_t)r * 8 + 7];
  if (wid == 0) ray_weights<true>(rows, zrow, far, K, lane, s_w, s_alpha, s_T, s_delta, nz);
  // this lane's slice of dL/d(embed of the ray) and the ray's scalar upstream gradients
  float4 de[DV];
#pragma unroll
  for (int i = 0; i < DV; ++i) de[i] = __ldg(reinterpret_cast<const float4*>(d_embed + (int64_t)r * D) + lane + i * kWarp);
  const float dr = d_rgb[(int64_t)r * 3 + 0], dg = d_rgb[(int64_t)r * 3 + 1], db = d_rgb[(int64_t)r * 3 + 2];
  const float dd = d_depth ? d_depth[r] : 0.0f;
  const float bk = white_bkgd ? -(dr + dg + db) : 0.0f;
  __syncthreads();

  // pass A: g_k = <d_rgb, rgb_k> + <d_embed, e_k> + d_depth z_k (+ d_w_k); every field row is read here, once
#pragma unroll 2
  for (int k = wid; k < K; k += kRayWarps) {
    const float4* row = reinterpret_cast<const float4*>(rows.at(k));
    float4 x[DV];
#pragma unroll
    for (int i = 0; i < DV; ++i) x[i] = __ldg(row + 1 + lane + i * kWarp);
    float part = 0.f;
#pragma unroll
    for (int i = 0; i < DV; ++i) part += x[i].x * de[i].x + x[i].y * de[i].y + x[i].z * de[i].z + x[i].w * de[i].w;
    part = warp_sum(part);
    if (lane == 0) {
      const float4 h = __ldg(row);
      const float sr = sigmoidf_(h.x), sg = sigmoidf_(h.y), sb = sigmoidf_(h.z);
      float g = part + sr * dr + sg * dg + sb * db + dd * zrow[k] + bk;
      if (d_weights) g += d_weights[(int64_t)r * K + k];
      s_g[k] = g;
      const float sden = nz ? fmaxf(fmaxf(h.w, 0.0f) + nz[k], 0.0f) : fmaxf(h.w, 0.0f);
      s_sig[k] = sden;
      s_open[k] = (h.w > 0.0f && sden > 0.0f) ? 1.0f : 0.0f;
      s_h[3 * k + 0] = dr * sr * (1.0f - sr);
      s_h[3 * k + 1] = dg * sg * (1.0f - sg);
      s_h[3 * k + 2] = db * sb * (1.0f - sb);
    }
  }
  __syncthreads();
  // suffix sums S_k = sum_{j>k} w_j g_j : lane-chunked reverse scan (warp 0)
  if (wid == 0) {
    int chunk = (K + kWarp - 1) / kWarp;
    int k0 = lane * chunk, k1 = min(K, k0 + chunk);
    float local = 0.f;
    for (int k = k0; k < k1; ++k) local += s_w[k] * s_g[k];
    float incl = local;
#pragma unroll
    for (int o = 1; o < kWarp; o <<= 1) {
      float v = __shfl_down_sync(0xffffffffu, incl, o);
      if (lane + o < kWarp) incl += v;
    }
    float run = incl - local;   // sum over later lanes
    for (int k = k1 - 1; k >= k0; --k) {
      float alpha = s_alpha[k];
      float one_m = 1.0f - alpha;
      float bracket = s_T[k] * s_g[k] - run / (one_m + 1e-10f);
      run += s_w[k] * s_g[k];
      s_ds[k] = one_m * bracket;
    }
  }
  __syncthreads();
  // pass B: the gradient rows, from shared memory and registers only
  const int pad4 = (ldg - (4 + D)) / 4;            // zero-filled float4 groups behind the last channel
#pragma unroll 2
  for (int k = wid; k < K; k += kRayWarps) {
    const float w = s_w[k];
    const bool acc = ru.accumulate != 0;
    T* grow = grad_row<T>(rows, d_field, ru.d_field_new, ldg, k);
#pragma unroll
    for (int i = 0; i < DV; ++i)
      put_grad4<T>(grow + 4 + (lane + i * kWarp) * 4, make_float4(w * de[i].x, w * de[i].y, w * de[i].z, w * de[i].w), acc);
    if (lane == 0) {
      put_grad4<T>(grow, make_float4(w * s_h[3 * k], w * s_h[3 * k + 1], w * s_h[3 * k + 2],
                                     s_open[k] != 0.0f ? s_delta[k] * s_ds[k] : 0.0f), acc);
    } else if (lane <= pad4 && !acc) {
      store_grad4<T>(grow + 4 + D + (lane - 1) * 4, make_float4(0.f, 0.f, 0.f, 0.f));
    }
  }
  if (d_z) {
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
      float dl_k = s_sig[k] * s_ds[k];
      float dl_km1 = k > 0 ? s_sig[k - 1] * s_ds[k - 1] : 0.f;
      d_z[(int64_t)r * K + k] = dd * s_w[k] + dl_km1 - dl_k;
    }
  }
}

}  // namespace nrf

using namespace nrf;

static int check_composite(const char* who, int R, int K, int D, int ldo) {
  NRF_REQUIRE(R > 0 && K > 0 && D >= 0, NRF_EINVAL, "%s: bad sizes", who);
  NRF_REQUIRE(K <= kWarp * kMaxPerLane, NRF_ENOSUP, "%s: K=%d > 1024", who, K);
  NRF_REQUIRE(D % 4 == 0 && ldo % 4 == 0 && ldo >= 4 + D, NRF_ENOSUP,
              "%s: D=%d, ldo=%d must be multiples of 4 with ldo >= 4+D", who, D, ldo);
  NRF_REQUIRE((4 + D) / 4 <= 8 * kWarp, NRF_ENOSUP, "%s: 4+D=%d > 1024", who, 4 + D);
  return NRF_OK;
}

extern "C" int nrf_composite_fwd(const float* field_out, int ldo, const float* z, const float* rays,
                                 int R, int K, int D, int white_bkgd, float* weights, float* rgb,
                                 float* embed, float* depth, const float* sigma_noise,
                                 const NrfCompositeReuse* reuse, void* stream) {
  NRF_REQUIRE(field_out && z && rays && weights && rgb && embed && depth, NRF_EINVAL,
              "nrf_composite_fwd: null pointer");
  int rc = check_composite("nrf_composite_fwd", R, K, D, ldo);
  if (rc) return rc;
  NrfReuseDev ru{nullptr, nullptr, 0, nullptr, 0};
  if (reuse) {
    NRF_REQUIRE(reuse->field_new && reuse->perm && reuse->n_first > 0 && reuse->n_first < K, NRF_EINVAL,
                "nrf_composite_fwd: reuse needs field_new, perm and 0 < n_first < K");
    ru = NrfReuseDev{reuse->field_new, reuse->perm, reuse->n_first, nullptr, 0};
  }
  size_t smem = ((size_t)((2 * K + 3) & ~3) + (size_t)kRayWarps * (4 + D + 4)) * sizeof(float);
  const bool aligned = (reinterpret_cast<uintptr_t>(field_out) & 15) == 0 && (reinterpret_cast<uintptr_t>(embed) & 15) == 0;
  { LaunchScope ls_(NRF_CAT_COMPOSITE_FWD, as_stream(stream));
  if (D == 384 && aligned && smem <= 48 * 1024)
    composite_fwd_fast_kernel<3><<<R, kRayWarps * kWarp, smem, as_stream(stream)>>>(
        field_out, ldo, z, rays, K, white_bkgd, weights, rgb, embed, depth, sigma_noise, ru);
  else if (D == 512 && aligned && smem <= 48 * 1024)
    composite_fwd_fast_kernel<4><<<R, kRayWarps * kWarp, smem, as_stream(stream)>>>(
        field_out, ldo, z, rays, K, white_bkgd, weights, rgb, embed, depth, sigma_noise, ru);
  else
    composite_fwd_kernel<<<R, kRayWarps * kWarp, smem, as_stream(stream)>>>(
        field_out, ldo, z, rays, R, K, D, white_bkgd, weights, rgb, embed, depth, sigma_noise, ru);
  }
  NRF_LAUNCH_OK();
  return NRF_OK;
}

extern "C" int nrf_composite_bwd(const float* field_out, int ldo, const float* z, const float* rays,
                                 int R, int K, int D, int white_bkgd, const float* d_rgb,
                                 const float* d_embed, const float* d_depth, const float* d_weights,
                                 void* d_field, int ldg, int out_bf16, float* d_z, const float* sigma_noise,
                                 const NrfCompositeReuse* reuse, int accumulate, void* stream) {
  NRF_REQUIRE(field_out && z && rays && d_rgb && d_embed && d_field, NRF_EINVAL,
              "nrf_composite_bwd: null pointer");
  int rc = check_composite("nrf_composite_bwd", R, K, D, ldo);
  if (rc) return rc;
  NrfReuseDev ru{nullptr, nullptr, 0, nullptr, 0};
  if (reuse) {
    NRF_REQUIRE(reuse->field_new && reuse->perm && reuse->d_field_new && reuse->n_first > 0 && reuse->n_first < K,
                NRF_EINVAL, "nrf_composite_bwd: reuse needs field_new, perm, d_field_new and 0 < n_first < K");
    NRF_REQUIRE(!accumulate, NRF_EINVAL, "nrf_composite_bwd: accumulate and reuse are exclusive");
    ru = NrfReuseDev{reuse->field_new, reuse->perm, reuse->n_first, reuse->d_field_new, 0};
  }
  ru.accumulate = accumulate ? 1 : 0;
  NRF_REQUIRE(ldg >= 4 + D && ldg % 4 == 0, NRF_EINVAL, "nrf_composite_bwd: ldg=%d", ldg);
  size_t smem = (size_t)6 * K * sizeof(float);
  dim3 grid(R), block(kRayWarps * kWarp);
  // fast path (D = 384 / 512): every field row read once, nothing predicated
  const size_t smem_fast = (size_t)11 * K * sizeof(float);
  const bool aligned = (reinterpret_cast<uintptr_t>(field_out) & 15) == 0 &&
                       (reinterpret_cast<uintptr_t>(d_embed) & 15) == 0 && (reinterpret_cast<uintptr_t>(d_field) & 15) == 0;
  const int pad4 = (ldg - (4 + D)) / 4;
  if ((D == 384 || D == 512) && aligned && smem_fast <= 48 * 1024 && pad4 < kWarp) {
    LaunchScope ls_(NRF_CAT_COMPOSITE_BWD, as_stream(stream));
#define NRF_CBWD_FAST(DV, T)                                                                                   \
    composite_bwd_fast_kernel<DV, T><<<grid, block, smem_fast, as_stream(stream)>>>(                          \
        field_out, ldo, z, rays, K, white_bkgd, d_rgb, d_embed, d_depth, d_weights, reinterpret_cast<T*>(d_field), \
        ldg, d_z, sigma_noise, ru)
    if (D == 384) { if (out_bf16) NRF_CBWD_FAST(3, __nv_bfloat16); else NRF_CBWD_FAST(3, float); }
    else { if (out_bf16) NRF_CBWD_FAST(4, __nv_bfloat16); else NRF_CBWD_FAST(4, float); }
#undef NRF_CBWD_FAST
  } else if (out_bf16) {
    if (smem > 48 * 1024)
      NRF_CUDA_OK(cudaFuncSetAttribute(composite_bwd_kernel<__nv_bfloat16>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    { LaunchScope ls_(NRF_CAT_COMPOSITE_BWD, as_stream(stream));
    composite_bwd_kernel<__nv_bfloat16><<<grid, block, smem, as_stream(stream)>>>(
        field_out, ldo, z, rays, R, K, D, white_bkgd, d_rgb, d_embed, d_depth, d_weights,
        reinterpret_cast<__nv_bfloat16*>(d_field), ldg, d_z, sigma_noise, ru);
    }
  } else {
    if (smem > 48 * 1024)
      NRF_CUDA_OK(cudaFuncSetAttribute(composite_bwd_kernel<float>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    { LaunchScope ls_(NRF_CAT_COMPOSITE_BWD, as_stream(stream));
    composite_bwd_kernel<float><<<grid, block, smem, as_stream(stream)>>>(
        field_out, ldo, z, rays, R, K, D, white_bkgd, d_rgb, d_embed, d_depth, d_weights,
        reinterpret_cast<float*>(d_field), ldg, d_z, sigma_noise, ru);
    }
  }
  NRF_LAUNCH_OK();
  return NRF_OK;
}
