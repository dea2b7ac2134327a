// Counting-sort building blocks shared by the volume-gradient scatter (encode.cu) and the voxelizer (voxelize.cu):
// an exclusive scan of per-bin counts (three small kernels) and the fill of per-bin entry lists.  `static`: every
// translation unit that includes this gets its own copy (the library is built without relocatable device code).
#pragma once
#include "common.cuh"

namespace nrf {

// exclusive scan, 1024 elements per block
static __global__ void __launch_bounds__(256) scan_block_kernel(const int32_t* __restrict__ in, int32_t* __restrict__ out,
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

static __global__ void __launch_bounds__(1024) scan_sums_kernel(int32_t* __restrict__ block_sums, int nb,
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

static __global__ void __launch_bounds__(256) scan_add_kernel(int32_t* __restrict__ out, const int32_t* __restrict__ block_sums,
                                                       int64_t T) {
  int64_t base = (int64_t)blockIdx.x * 1024 + threadIdx.x * 4;
  int32_t add = block_sums[blockIdx.x];
#pragma unroll
  for (int i = 0; i < 4; ++i)
    if (base + i < T) out[base + i] += add;
}

static __global__ void __launch_bounds__(256) scatter_fill_kernel(const int32_t* __restrict__ key, int64_t E,
                                                           const int32_t* __restrict__ offset,
                                                           int32_t* __restrict__ cursor, int32_t* __restrict__ list) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  int32_t k = key[e];
  if (k < 0) return;
  int32_t pos = offset[k] + atomicAdd(cursor + k, 1);
  list[pos] = (int32_t)e;
}

// offset[t] = sum_{u<t} count[u]; block_sums: (T+1023)/1024 + 1 ints of scratch
static inline int exclusive_scan_i32(const int32_t* count, int32_t* offset, int32_t* block_sums, int64_t T, int category,
                                     cudaStream_t s) {
  const int64_t nb = (T + 1023) / 1024;
  { LaunchScope ls_(category, s);
    scan_block_kernel<<<(unsigned)nb, 256, 0, s>>>(count, offset, block_sums, T); }
  NRF_LAUNCH_OK();
  { LaunchScope ls_(category, s);
    scan_sums_kernel<<<1, 1024, 0, s>>>(block_sums, (int)nb, nullptr); }
  NRF_LAUNCH_OK();
  { LaunchScope ls_(category, s);
    scan_add_kernel<<<(unsigned)nb, 256, 0, s>>>(offset, block_sums, T); }
  NRF_LAUNCH_OK();
  return NRF_OK;
}

}  // namespace nrf
