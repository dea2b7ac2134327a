// Rendering losses of one training step in one pass over the rendered rays.
//   nrf_render_loss <- neural_rendering.py:653-677 (target gather `[:, idx]`, the four F.mse_loss terms) and the
//                      autograd of those terms (mse_loss backward): the gradients w.r.t. the rendered rgb / embed of
//                      both passes are produced here, in the same read of the data.
// The reference runs ~10 small ATen kernels forward (2 index gathers, 4 x (square-difference + mean)) and 8 backward;
// here: one CTA per ray (sums of squares of its 3 + D channels against the gathered target pixel, gradients written
// on the way), then one CTA that adds the per-ray partials in a fixed order (bit-reproducible, no float atomics).
#include "common.cuh"

namespace nrf {

constexpr int kLossThreads = 128;

__device__ __forceinline__ float block_sum_128(float v, float* s_red) {
  v = warp_sum(v);
  const int wid = threadIdx.x / kWarp, lane = threadIdx.x % kWarp;
  __syncthreads();                       // s_red may still be read from the previous call
  if (lane == 0) s_red[wid] = v;
  __syncthreads();
  return s_red[0] + s_red[1] + s_red[2] + s_red[3];
}

__global__ void __launch_bounds__(kLossThreads) render_loss_kernel(
    const float* __restrict__ rgb_c, const float* __restrict__ rgb_f, const float* __restrict__ emb_c,
    const float* __restrict__ emb_f, int D, int rays_per_scene, const float* __restrict__ gt_rgb,
    const float* __restrict__ gt_embed, int64_t n_pix, const int64_t* __restrict__ idx, float* __restrict__ partial,
    float* __restrict__ d_rgb_c, float* __restrict__ d_rgb_f, float* __restrict__ d_emb_c,
    float* __restrict__ d_emb_f, float k_rgb, float k_emb) {
  __shared__ float s_red[4];
  const int r = blockIdx.x;
  const int scene = r / rays_per_scene, j = r - scene * rays_per_scene;
  const int64_t pix = idx ? (int64_t)scene * n_pix + idx[j] : (int64_t)r;     // idx == NULL: targets already per ray
  float sc = 0.f, sf = 0.f;
  if (threadIdx.x < 3) {
    const float t = gt_rgb[pix * 3 + threadIdx.x];
    const float a = rgb_c[(int64_t)r * 3 + threadIdx.x] - t, b = rgb_f[(int64_t)r * 3 + threadIdx.x] - t;
    sc = a * a;
    sf = b * b;
    if (d_rgb_c) d_rgb_c[(int64_t)r * 3 + threadIdx.x] = k_rgb * a;
    if (d_rgb_f) d_rgb_f[(int64_t)r * 3 + threadIdx.x] = k_rgb * b;
  }
  const float rc = block_sum_128(sc, s_red);
  const float rf = block_sum_128(sf, s_red);
  float ec = 0.f, ef = 0.f;
  const float* tg = gt_embed + pix * D;
  const float* xc = emb_c + (int64_t)r * D;
  const float* xf = emb_f + (int64_t)r * D;
  for (int c = threadIdx.x; c < D; c += kLossThreads) {
    const float t = tg[c];
    const float a = xc[c] - t, b = xf[c] - t;
    ec = fmaf(a, a, ec);
    ef = fmaf(b, b, ef);
    if (d_emb_c) d_emb_c[(int64_t)r * D + c] = k_emb * a;
    if (d_emb_f) d_emb_f[(int64_t)r * D + c] = k_emb * b;
  }
  const float tc = block_sum_128(ec, s_red);
  const float tf = block_sum_128(ef, s_red);
  if (threadIdx.x == 0)
    *reinterpret_cast<float4*>(partial + (int64_t)r * 4) = make_float4(rc, rf, tc, tf);
}

// terms[i] = (sum over rays of partial[r][i]) / n_i, rays added in ascending order within 256 fixed strides
__global__ void __launch_bounds__(256) render_loss_finish_kernel(const float* __restrict__ partial, int R,
                                                                 float inv_rgb, float inv_emb,
                                                                 float* __restrict__ terms) {
  __shared__ float4 s[256];
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int r = threadIdx.x; r < R; r += 256) {
    const float4 p = *reinterpret_cast<const float4*>(partial + (int64_t)r * 4);
    a.x += p.x; a.y += p.y; a.z += p.z; a.w += p.w;
  }
  s[threadIdx.x] = a;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      float4 b = s[threadIdx.x + o];
      s[threadIdx.x].x += b.x; s[threadIdx.x].y += b.y; s[threadIdx.x].z += b.z; s[threadIdx.x].w += b.w;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    terms[0] = s[0].x * inv_rgb;
    terms[1] = s[0].y * inv_rgb;
    terms[2] = s[0].z * inv_emb;
    terms[3] = s[0].w * inv_emb;
  }
}

}  // namespace nrf

using namespace nrf;

extern "C" int nrf_render_loss(const float* rgb_c, const float* rgb_f, const float* emb_c, const float* emb_f, int R,
                               int D, int rays_per_scene, const float* gt_rgb, const float* gt_embed, int64_t n_pix,
                               const int64_t* idx, float* partial, float* terms, float* d_rgb_c, float* d_rgb_f,
                               float* d_emb_c, float* d_emb_f, void* stream) {
  NRF_REQUIRE(rgb_c && rgb_f && emb_c && emb_f && gt_rgb && gt_embed && partial && terms, NRF_EINVAL,
              "nrf_render_loss: null pointer");
  NRF_REQUIRE(R > 0 && D > 0 && rays_per_scene > 0 && R % rays_per_scene == 0, NRF_EINVAL,
              "nrf_render_loss: R=%d must be a multiple of rays_per_scene=%d", R, rays_per_scene);
  NRF_REQUIRE((reinterpret_cast<uintptr_t>(partial) & 15) == 0, NRF_EINVAL, "nrf_render_loss: partial alignment");
  cudaStream_t s = as_stream(stream);
  const float n_rgb = (float)R * 3.0f, n_emb = (float)R * (float)D;
  { LaunchScope ls_(NRF_CAT_MISC, s);
    render_loss_kernel<<<R, kLossThreads, 0, s>>>(rgb_c, rgb_f, emb_c, emb_f, D, rays_per_scene, gt_rgb, gt_embed,
                                                  n_pix, idx, partial, d_rgb_c, d_rgb_f, d_emb_c, d_emb_f,
                                                  2.0f / n_rgb, 2.0f / n_emb); }
  NRF_LAUNCH_OK();
  { LaunchScope ls_(NRF_CAT_MISC, s);
    render_loss_finish_kernel<<<1, 256, 0, s>>>(partial, R, 1.0f / n_rgb, 1.0f / n_emb, terms); }
  NRF_LAUNCH_OK();
  return NRF_OK;
}
