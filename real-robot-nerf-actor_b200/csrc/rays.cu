// Ray generation and per-ray sampling kernels (bit-exact restatements of eager-PyTorch op chains).
//   nrf_raygen        <- utils.py:444-506   unproj_map + gen_rays
//   nrf_sample_coarse <- neural_rendering.py:159-176
//   nrf_sample_fine   <- neural_rendering.py:179-207
//   nrf_sort_rows     <- neural_rendering.py:463 (torch.sort along the sample axis)
// Every operation the reference rounds separately is rounded separately here (__f*_rn).
#include "common.cuh"

namespace nrf {

// One thread per pixel.  Default = the CPU-ATen bit pattern the golden fixtures were produced with:
// norm = sqrt(fma(z,z,fma(y,y,x*x))), direction = (r0*x + r1*y) + r2*z with separately rounded products and sums.
// `flags` selects the rounding pattern of the SAME formulas as other ATen back ends execute them (NRF_RAYGEN_*):
//   pixel coordinate  true division | product with the fp32 reciprocal (ATen's CUDA div by a host scalar)
//   norm              the four association orders of x*x + y*y + z*z a reduction kernel can produce
//   direction         separately rounded | FMA chain over k ascending | descending | two accumulators {0, 1} + {2}
// intr (device, [fx, fy, cx, cy]) overrides the by-value intrinsics: a focal length that lives on the GPU (as in the
// reference's callers) is then never read back by the host -- no stream sync at the start of every step.
__global__ void raygen_kernel(const float* __restrict__ poses, int n_img, int W, int H, float fx,
                              float fy, float cx, float cy, float z_near, float z_far,
                              float* __restrict__ rays, const float* __restrict__ intr, int flags) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t total = (int64_t)n_img * H * W;
  if (t >= total) return;
  if (intr) { fx = intr[0]; fy = intr[1]; cx = intr[2]; cy = intr[3]; }
  int j = (int)(t % W);
  int i = (int)((t / W) % H);
  int b = (int)(t / ((int64_t)W * H));
  const float* P = poses + (int64_t)b * 16;
  const int dir_mode = flags & 3, recip = (flags >> 2) & 1, norm_mode = (flags >> 3) & 3;
  float x, y;
  if (recip) {
    x = __fmul_rn(__fsub_rn((float)j, cx), __fdiv_rn(1.0f, fx));
    y = -__fmul_rn(__fsub_rn((float)i, cy), __fdiv_rn(1.0f, fy));
  } else {
    x = __fdiv_rn(__fsub_rn((float)j, cx), fx);
    y = -__fdiv_rn(__fsub_rn((float)i, cy), fy);
  }
  float zc = -1.0f;
  float n2;
  {
    const float xx = __fmul_rn(x, x), yy = __fmul_rn(y, y), zz = __fmul_rn(zc, zc);
    if (norm_mode == 0) n2 = __fmaf_rn(zc, zc, __fmaf_rn(y, y, xx));
    else if (norm_mode == 1) n2 = __fadd_rn(__fadd_rn(xx, yy), zz);
    else if (norm_mode == 2) n2 = __fadd_rn(__fadd_rn(xx, zz), yy);
    else n2 = __fadd_rn(xx, __fadd_rn(yy, zz));
  }
  float nrm = __fsqrt_rn(n2);
  x = __fdiv_rn(x, nrm);
  y = __fdiv_rn(y, nrm);
  zc = __fdiv_rn(zc, nrm);
  float4 lo, hi;
  lo.x = P[3];
  lo.y = P[7];
  lo.z = P[11];
  float d[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const float r0 = P[r * 4 + 0], r1 = P[r * 4 + 1], r2 = P[r * 4 + 2];
    if (dir_mode == 0) {
      d[r] = __fadd_rn(__fadd_rn(__fmul_rn(r0, x), __fmul_rn(r1, y)), __fmul_rn(r2, zc));
    } else if (dir_mode == 1) {
      d[r] = __fmaf_rn(r2, zc, __fmaf_rn(r1, y, __fmul_rn(r0, x)));
    } else if (dir_mode == 2) {
      d[r] = __fmaf_rn(r0, x, __fmaf_rn(r1, y, __fmul_rn(r2, zc)));
    } else {
      // cuBLAS' batched (3x3) x (3x1) product as torch.matmul reaches it on B200 (scripts/raygen_probe.py): two
      // accumulators that start at +0 - k = 0, 1 chained by FMA and k = 2 on its own - then added.  The +0 start only
      // shows in the sign of an exact zero (fma(0, x, +0) = +0 for negative x where 0 * x = -0).
      d[r] = __fadd_rn(__fmaf_rn(r1, y, __fmaf_rn(r0, x, 0.0f)), __fmaf_rn(r2, zc, 0.0f));
    }
  }
  lo.w = d[0];
  hi.x = d[1];
  hi.y = d[2];
  hi.z = z_near;
  hi.w = z_far;
  float4* o = reinterpret_cast<float4*>(rays + t * 8);
  o[0] = lo;
  o[1] = hi;
}

__device__ __forceinline__ float lerp_depth(float near, float far, float zs, int lindisp) {
  if (!lindisp) return __fadd_rn(__fmul_rn(near, __fsub_rn(1.0f, zs)), __fmul_rn(far, zs));
  float a = __fmul_rn(__fdiv_rn(1.0f, near), __fsub_rn(1.0f, zs));
  float b = __fmul_rn(__fdiv_rn(1.0f, far), zs);
  return __fdiv_rn(1.0f, __fadd_rn(a, b));
}

__global__ void sample_coarse_kernel(const float* __restrict__ rays, int R, int Kc,
                                     const float* __restrict__ base, const float* __restrict__ jitter,
                                     int lindisp, float* __restrict__ z) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (int64_t)R * Kc) return;
  int r = (int)(t / Kc), k = (int)(t % Kc);
  float step = (float)(1.0 / (double)Kc);
  float zs = base[k];
  if (jitter) zs = __fadd_rn(zs, __fmul_rn(jitter[t], step));
  z[t] = lerp_depth(rays[(int64_t)r * 8 + 6], rays[(int64_t)r * 8 + 7], zs, lindisp);
}

// One warp per ray.  cdf is either given (bit-exact path: identical (cdf,u,jitter) -> identical
// (ind,z)) or built here from the weights: pdf = (w+1e-5)/sum, cdf = cumsum(pdf) accumulated in
// double and rounded to fp32 per entry (the CPU-ATen cumsum bit pattern), or - flags & NRF_FINE_CUDA_EAGER,
// Kc = 64 / 128 - in the association order ATen's CUDA kernels use (scripts/cdf_probe.py, bit for bit on B200):
//   torch.sum     Kc = 64: lane t adds elements t and t + 32; Kc = 128: lane t adds its float4 ((e0+e1)+e2)+e3;
//                 then a shuffle-down tree with offsets 16, 8, 4, 2, 1
//   the division  IEEE
//   torch.cumsum  chunks of 32 elements, the running total added to a chunk's first element, then a 16-thread
//                 Sklansky network (5 steps: thread t adds element a - 1 to element a + t % s, a = (t / s) 2 s + s)
__global__ void sample_fine_kernel(const float* __restrict__ rays, const float* __restrict__ weights,
                                   const float* __restrict__ cdf_in, int R, int Kc,
                                   const float* __restrict__ u, const float* __restrict__ jitter,
                                   int Kf, int flags, float* __restrict__ z_out, int ldz,
                                   float* __restrict__ ind_out) {
  const int lindisp = flags & 1;
  extern __shared__ float smem[];
  int warps = blockDim.x / kWarp;
  int wid = threadIdx.x / kWarp, lane = threadIdx.x % kWarp;
  int r = blockIdx.x * warps + wid;
  float* cdf = smem + (size_t)wid * (Kc + 1);
  if (r >= R) return;
  if (cdf_in) {
    for (int k = lane; k <= Kc; k += kWarp) cdf[k] = cdf_in[(int64_t)r * (Kc + 1) + k];
  } else if (flags & NRF_FINE_CUDA_EAGER) {
    const float* w = weights + (int64_t)r * Kc;
    float a;
    if (Kc == 64) {
      a = __fadd_rn(__fadd_rn(w[lane], 1e-5f), __fadd_rn(w[lane + 32], 1e-5f));
    } else {                                       // Kc == 128
      const float4 v = *reinterpret_cast<const float4*>(w + 4 * lane);
      a = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(v.x, 1e-5f), __fadd_rn(v.y, 1e-5f)), __fadd_rn(v.z, 1e-5f)),
                    __fadd_rn(v.w, 1e-5f));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a = __fadd_rn(a, __shfl_down_sync(0xffffffffu, a, o));
    const float total = __shfl_sync(0xffffffffu, a, 0);
    if (lane == 0) cdf[0] = 0.0f;
    for (int k = lane; k < Kc; k += kWarp) cdf[k + 1] = __fdiv_rn(__fadd_rn(w[k], 1e-5f), total);
    __syncwarp();
    for (int c0 = 0; c0 < Kc; c0 += 32) {
      float* buf = cdf + 1 + c0;
      if (lane == 0 && c0 > 0) buf[0] = __fadd_rn(buf[0], cdf[c0]);     // the total of the chunks before
      __syncwarp();
#pragma unroll
      for (int sft = 0; sft < 5; ++sft) {
        const int st = 1 << sft;
        if (lane < 16) {
          const int base = (lane >> sft) * (2 * st) + st;
          const int ti = base + (lane & (st - 1)), si = base - 1;
          buf[ti] = __fadd_rn(buf[ti], buf[si]);
        }
        __syncwarp();
      }
    }
  } else {
    const float* w = weights + (int64_t)r * Kc;
    // sum of (w + 1e-5) in fp32 element order-independent double accumulation
    double s = 0.0;
    for (int k = lane; k < Kc; k += kWarp) s += (double)__fadd_rn(w[k], 1e-5f);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    float total = (float)s;
    // inclusive double prefix sums of pdf, chunked per lane
    int per = (Kc + kWarp - 1) / kWarp;
    int k0 = lane * per;
    double local = 0.0;
    for (int k = k0; k < min(k0 + per, Kc); ++k)
      local += (double)__fdiv_rn(__fadd_rn(w[k], 1e-5f), total);
    double pre = local;
#pragma unroll
    for (int o = 1; o < kWarp; o <<= 1) {
      double v = __shfl_up_sync(0xffffffffu, pre, o);
      if (lane >= o) pre += v;
    }
    double run = pre - local;
    if (lane == 0) cdf[0] = 0.0f;
    for (int k = k0; k < min(k0 + per, Kc); ++k) {
      run += (double)__fdiv_rn(__fadd_rn(w[k], 1e-5f), total);
      cdf[k + 1] = (float)run;
    }
  }
  __syncwarp();
  float near = rays[(int64_t)r * 8 + 6], far = rays[(int64_t)r * 8 + 7];
  for (int k = lane; k < Kf; k += kWarp) {
    float uu = u[(int64_t)r * Kf + k];
    // searchsorted(cdf, u, right=True): number of entries <= u  (cdf is non-decreasing)
    int lo = 0, hi = Kc + 1;
    while (lo < hi) {
      int mid = (lo + hi) >> 1;
      if (cdf[mid] <= uu) lo = mid + 1; else hi = mid;
    }
    float ind = fmaxf((float)lo - 1.0f, 0.0f);
    float jit = jitter ? jitter[(int64_t)r * Kf + k] : 0.0f;
    float zs = __fdiv_rn(__fadd_rn(ind, jit), (float)Kc);
    z_out[(int64_t)r * ldz + k] = lerp_depth(near, far, zs, lindisp);
    if (ind_out) ind_out[(int64_t)r * Kf + k] = ind;
  }
}

// One CTA per ray: bitonic sort of up to 1024 keys in shared memory, stable w.r.t. the original
// position (ties broken by index, as torch.sort(stable=False) happens to do for small rows is not
// guaranteed -- only the sorted VALUES are compared bit-for-bit; perm is used for the backward).
__global__ void sort_rows_kernel(float* __restrict__ z, int R, int K, int P, int32_t* __restrict__ perm) {
  extern __shared__ unsigned char raw[];
  float* key = reinterpret_cast<float*>(raw);
  int* idx = reinterpret_cast<int*>(key + P);
  int r = blockIdx.x;
  for (int i = threadIdx.x; i < P; i += blockDim.x) {
    key[i] = i < K ? z[(int64_t)r * K + i] : __int_as_float(0x7f800000);
    idx[i] = i;
  }
  __syncthreads();
  for (int size = 2; size <= P; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = threadIdx.x; t < P / 2; t += blockDim.x) {
        int lo = 2 * t - (t & (stride - 1));
        int hi = lo + stride;
        bool up = ((lo & size) == 0);
        float a = key[lo], b = key[hi];
        int ia = idx[lo], ib = idx[hi];
        // a TOTAL order: NaN sorts last among the real entries (as torch.sort does) and the padding behind K stays
        // behind every real entry - a NaN depth can no longer pull a padding slot (index >= K) into the first K
        auto ord = [](float x) { int i = __float_as_int(x); return x != x ? 0x7fffffff : i ^ ((i >> 31) & 0x7fffffff); };
        const int ka = ord(a), kb = ord(b);
        const bool pa = ia >= K, pb = ib >= K;
        bool gt = pa != pb ? pa : (ka > kb) || (ka == kb && ia > ib);
        if (gt == up) {
          key[lo] = b; key[hi] = a;
          idx[lo] = ib; idx[hi] = ia;
        }
      }
      __syncthreads();
    }
  }
  for (int i = threadIdx.x; i < K; i += blockDim.x) {
    z[(int64_t)r * K + i] = key[i];
    if (perm) perm[(int64_t)r * K + i] = idx[i];
  }
}

}  // namespace nrf

using namespace nrf;

extern "C" int nrf_raygen_ex(const float* poses, int n_img, int W, int H, float fx, float fy, float cx,
                             float cy, float z_near, float z_far, float* rays_out, const float* intrinsics_dev,
                             int flags, void* stream) {
  NRF_REQUIRE(poses && rays_out && n_img > 0 && W > 0 && H > 0, NRF_EINVAL, "nrf_raygen: bad args");
  NRF_REQUIRE(flags >= 0 && flags < 32, NRF_EINVAL, "nrf_raygen: unknown flags %d", flags);
  int64_t total = (int64_t)n_img * H * W;
  int threads = 256;
  { LaunchScope ls_(NRF_CAT_SAMPLING, as_stream(stream));
  raygen_kernel<<<(unsigned)((total + threads - 1) / threads), threads, 0, as_stream(stream)>>>(
      poses, n_img, W, H, fx, fy, cx, cy, z_near, z_far, rays_out, intrinsics_dev, flags);
  }
  NRF_LAUNCH_OK();
  return NRF_OK;
}

extern "C" int nrf_raygen(const float* poses, int n_img, int W, int H, float fx, float fy, float cx,
                          float cy, float z_near, float z_far, float* rays_out, const float* intrinsics_dev,
                          void* stream) {
  return nrf_raygen_ex(poses, n_img, W, H, fx, fy, cx, cy, z_near, z_far, rays_out, intrinsics_dev, 0, stream);
}

extern "C" int nrf_sample_coarse(const float* rays, int R, int Kc, const float* base,
                                 const float* jitter, int lindisp, float* z_out, void* stream) {
  NRF_REQUIRE(rays && base && z_out && R > 0 && Kc > 0, NRF_EINVAL, "nrf_sample_coarse: bad args");
  int64_t total = (int64_t)R * Kc;
  int threads = 256;
  { LaunchScope ls_(NRF_CAT_SAMPLING, as_stream(stream));
  sample_coarse_kernel<<<(unsigned)((total + threads - 1) / threads), threads, 0, as_stream(stream)>>>(
      rays, R, Kc, base, jitter, lindisp, z_out);
  }
  NRF_LAUNCH_OK();
  return NRF_OK;
}

extern "C" int nrf_sample_fine(const float* rays, const float* weights, const float* cdf, int R,
                               int Kc, const float* u, const float* jitter, int Kf, int lindisp,
                               float* z_out, int ldz, float* ind_out, void* stream) {
  NRF_REQUIRE(rays && (weights || cdf) && u && z_out && R > 0 && Kc > 0 && Kf > 0 && ldz >= Kf,
              NRF_EINVAL, "nrf_sample_fine: bad args");
  NRF_REQUIRE(Kc <= 4096, NRF_ENOSUP, "nrf_sample_fine: Kc > 4096");
  NRF_REQUIRE(!(lindisp & NRF_FINE_CUDA_EAGER) || cdf || ((Kc == 64 || Kc == 128) &&
              (reinterpret_cast<uintptr_t>(weights) & 15) == 0), NRF_ENOSUP,
              "nrf_sample_fine: the CUDA-eager cdf pattern is known for Kc = 64 / 128 (Kc = %d)", Kc);
  int warps = 4;
  size_t smem = (size_t)warps * (Kc + 1) * sizeof(float);
  { LaunchScope ls_(NRF_CAT_SAMPLING, as_stream(stream));
  sample_fine_kernel<<<(R + warps - 1) / warps, warps * kWarp, smem, as_stream(stream)>>>(
      rays, weights, cdf, R, Kc, u, jitter, Kf, lindisp, z_out, ldz, ind_out);
  }
  NRF_LAUNCH_OK();
  return NRF_OK;
}

extern "C" int nrf_sort_rows(float* z, int R, int K, int32_t* perm_out, void* stream) {
  NRF_REQUIRE(z && R > 0 && K > 0, NRF_EINVAL, "nrf_sort_rows: bad args");
  NRF_REQUIRE(K <= 1024, NRF_ENOSUP, "nrf_sort_rows: K > 1024");
  int P = 2;
  while (P < K) P <<= 1;
  int threads = P / 2 < 32 ? 32 : (P / 2 > 512 ? 512 : P / 2);
  { LaunchScope ls_(NRF_CAT_SAMPLING, as_stream(stream));
  sort_rows_kernel<<<R, threads, (size_t)P * 8, as_stream(stream)>>>(z, R, K, P, perm_out);
  }
  NRF_LAUNCH_OK();
  return NRF_OK;
}
