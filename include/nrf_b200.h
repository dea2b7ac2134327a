/* nrf_b200.h -- C ABI of the B200-native feature-NeRF render path.
 *
 * One shared library (real-robot-nerf-actor_b200/libnrf_b200.so), plain pointers and sizes,
 * no torch types.  Every entry point
 *   - takes DEVICE pointers (unless the name says otherwise) and a cudaStream_t passed as void*,
 *   - enqueues work on that stream and returns immediately,
 *   - returns 0 on success or a negative NRF_E* code (never throws, keeps no global mutable state
 *     apart from a read-only driver entry point looked up once),
 *   - does not allocate: the caller (PyTorch's caching allocator on the Python side) owns all memory.
 *
 * The reference has no FFI: its boundary is the Python class NeuralRenderer
 * (/root/reference/neural_rendering.py:86-711).  Each function below names the reference
 * lines whose work it replaces; INTEGRATION.md shows the ctypes binding and how
 * NeuralRenderer.forward_nerf is re-expressed on top of them.
 *
 * Layouts
 *   rays      (R,8) fp32 row-major  [origin xyz, dir xyz, near, far]      (utils.py:504-506)
 *   z         (R,K) fp32            sample depths along each ray
 *   volume    channel-first  (SB,C,S0,S1,S2) fp32 as the caller holds it   (models_embed.py:147)
 *             channels-last  (SB,S0,S1,S2,C) fp32 as the kernels read it
 *             grid x indexes S2, y indexes S1, z indexes S0 (models_embed.py:275 quirk)
 *   field in  (N, kin_pad) rows [latent C | PE(xyz) 39 | viewdir 3 | zeros], N = R*K,
 *             sample n = ray (n / K), depth index (n % K); bf16 (tensor-core mode) or fp32
 *   field out (N, 4+D) fp32 RAW MLP outputs [rgb(3) | sigma | embed(D)] before sigmoid / relu
 */
#ifndef NRF_B200_H
#define NRF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NRF_OK 0
#define NRF_EINVAL (-1)   /* bad argument (null pointer, size, alignment)   */
#define NRF_ECUDA (-2)    /* a CUDA runtime / driver call failed            */
#define NRF_ENOSUP (-3)   /* shape outside what the kernels are built for   */

/* precision of the field MLP */
#define NRF_PREC_BF16 0   /* bf16 operands, fp32 accumulate, tcgen05 tensor cores */
#define NRF_PREC_FP32 1   /* fp32 SIMT FFMA: parity-grade mode                    */
#define NRF_PREC_FP16 2   /* fp16 FORWARD operands (weights, field input, activations, residual stream: 11 significant
                             bits, same tcgen05 kind::f16 rate as bf16, 6-8 x smaller output error: SURVEY.md section 10)
                             with bf16 GRADIENTS (d_field, dL/dx', dL/dnet: their range needs no loss scaling; the
                             gradient error of either mode is set by ReLU-gate flips of the forward, not by gradient
                             rounding).  "Operand-typed" below then means fp16 for forward tensors (field_in, acts, packed
                             weights) and bf16 for gradient tensors (d_field, the backward's scratch). */

#define NRF_PREC_BF16X3 3 /* split-bf16 operands (a = hi + lo, 16 significant bits), every product as hi.hi + lo.hi + hi.lo
                             on the tensor cores with fp32 accumulation: fp32-grade results (1.5e-5 on
                             rgb) at a third of the bf16 rate, layer by layer (csrc/mlp_x3.cu).  The caller-facing tensors
                             are those of NRF_PREC_FP32: fp32 field input, fp32 d_field.  nrf_mlp_* only (nrf_gemm /
                             nrf_wgrad take the three 16-bit / fp32 modes). */

const char* nrf_version(void);
const char* nrf_last_error(void);   /* text of the last failure on this thread */

/* ---- rays (utils.py:444-506 unproj_map + gen_rays) ------------------------------------------
 * poses (n_img,4,4) cam->world fp32; rays_out (n_img,H,W,8).
 * intrinsics_dev: NULL, or 4 floats [fx, fy, cx, cy] in device memory that replace the by-value arguments (the
 * reference's callers keep `focal` on the GPU, utils.py:485: it is then never read back by the host). */
int nrf_raygen(const float* poses, int n_img, int W, int H, float fx, float fy, float cx, float cy,
               float z_near, float z_far, float* rays_out, const float* intrinsics_dev, void* stream);

/* The same formulas under another ATen back end's rounding pattern (utils.py:444-506 run on a GPU executes ATen's CUDA
 * kernels and a cuBLAS batched K = 3 product, the golden fixtures were written by the CPU kernels; nrf_raygen = flags 0).
 * flags = direction | pixel | norm:
 *   NRF_RAYGEN_DIR_SEPARATE (r0*x + r1*y) + r2*z, every product and sum rounded | _DIR_FMA_ASC fma(r2,z,fma(r1,y,r0*x))
 *   | _DIR_FMA_DESC fma(r0,x,fma(r1,y,r2*z)) | _DIR_SPLIT fma(r1,y,fma(r0,x,+0)) + fma(r2,z,+0);
 *   NRF_RAYGEN_PIXEL_RECIP: (j - cx) * (1/fx) instead of (j - cx) / fx (ATen CUDA divides by a host scalar that way);
 *   NRF_RAYGEN_NORM_FMA sqrt(fma(z,z,fma(y,y,x*x))) | _NORM_XY_Z (xx + yy) + zz | _NORM_XZ_Y (xx + zz) + yy
 *   | _NORM_X_YZ xx + (yy + zz).
 * NRF_RAYGEN_CUDA_EAGER is the combination measured bit-identical - signed zeros included - to CUDA-eager PyTorch
 * 2.11 on B200 (scripts/raygen_probe.py over all 64 combinations; asserted in tests/test_gpu_bench_sizes.py). */
#define NRF_RAYGEN_DIR_SEPARATE 0
#define NRF_RAYGEN_DIR_FMA_ASC 1
#define NRF_RAYGEN_DIR_FMA_DESC 2
#define NRF_RAYGEN_DIR_SPLIT 3
#define NRF_RAYGEN_PIXEL_RECIP 4
#define NRF_RAYGEN_NORM_FMA 0
#define NRF_RAYGEN_NORM_XY_Z 8
#define NRF_RAYGEN_NORM_XZ_Y 16
#define NRF_RAYGEN_NORM_X_YZ 24
#define NRF_RAYGEN_CUDA_EAGER (NRF_RAYGEN_DIR_SPLIT | NRF_RAYGEN_PIXEL_RECIP | NRF_RAYGEN_NORM_XZ_Y)
int nrf_raygen_ex(const float* poses, int n_img, int W, int H, float fx, float fy, float cx, float cy,
                  float z_near, float z_far, float* rays_out, const float* intrinsics_dev, int flags, void* stream);

/* ---- stratified sampling (neural_rendering.py:159-176 sample_coarse) ------------------------
 * base (Kc) = linspace(0, 1-1/Kc, Kc); jitter (R,Kc) in [0,1) or NULL (perturb off). */
int nrf_sample_coarse(const float* rays, int R, int Kc, const float* base, const float* jitter,
                      int lindisp, float* z_out, void* stream);

/* ---- importance sampling (neural_rendering.py:179-207 sample_fine) ---------------------------
 * weights (R,Kc) coarse compositing weights, or cdf (R,Kc+1) when `cdf` is non-NULL (then weights
 * is ignored); u, jitter (R,Kf) (jitter NULL = 0).  Writes z_out[r*ldz + k], k<Kf, and, when
 * ind_out != NULL, the bin indices (R,Kf) as fp32 (the reference's `inds`). */
#define NRF_FINE_CUDA_EAGER 0x100 /* OR into `lindisp`: build the cdf from the weights in the association order of ATen's CUDA
                                    kernels (torch.sum: strided / float4 partials + shuffle-down tree; torch.cumsum: 32-element
                                    chunks, 16-thread Sklansky network) instead of the CPU back end's (double accumulation):
                                    bit-identical cdf, indices and depths to a GPU run of neural_rendering.py:189-207.
                                    Kc = 64 or 128 (measured: scripts/cdf_probe.py, tests/test_gpu_bench_sizes.py). */
int nrf_sample_fine(const float* rays, const float* weights, const float* cdf, int R, int Kc,
                    const float* u, const float* jitter, int Kf, int lindisp, float* z_out, int ldz,
                    float* ind_out, void* stream);

/* ---- per-ray ascending sort (neural_rendering.py:463 torch.sort) ----------------------------
 * z (R,K) sorted in place, K <= 1024; perm_out (R,K) int32 source positions or NULL. */
int nrf_sort_rows(float* z, int R, int K, int32_t* perm_out, void* stream);

/* ---- volume layout ---------------------------------------------------------------------------
 * (SB,C,V) <-> (SB,V,C), V = S0*S1*S2.  The reference reads the channel-first tensor through
 * F.grid_sample (models_embed.py:275); the gather kernel wants one contiguous C-vector per corner. */
int nrf_volume_to_channels_last(const float* src, float* dst, int SB, int C, int64_t V, void* stream);
int nrf_volume_to_channels_first(const float* src, float* dst, int SB, int C, int64_t V, void* stream);

/* ---- field input: points + canonicalise + trilinear gather + positional encoding -------------
 * Replaces neural_rendering.py:246-283 (points, viewdirs), models_embed.py:185-203
 * (world_to_canonical), :259-277 (grid_sample), utils.py:545-557 (PositionalEncoding) and the
 * concatenations at models_embed.py:366,405.  vol_cl is channels-last.  bounds = 6 floats (HOST).
 * out: (N, ld_out) fp32 (out_bf16 & 0xff == 0), bf16 (1) or fp16 (2); columns >= C+42 are zero-filled up to ld_out.
 * out_bf16 | 0x100: the eight corners are accumulated with FMAs, which is what ATen's CUDA grid_sampler_3d compiles to
 * (bit-identical to the reference run on a GPU); default: separately rounded multiply and add, ATen's CPU kernel
 * (bit-identical to the reference run on the CPU, which is what the golden fixtures were produced by).
 * out_bf16 | 0x200 (use_code_viewdirs, models_embed.py:370-372): the positional encoding runs over the six inputs
 * [canonical xyz | view direction] together, tail [in(6) | per frequency sin(in), cos(in)] = 6 + 12 num_freqs columns
 * (78), instead of [PE(xyz) (3 + 6 num_freqs) | view direction (3)] (:347-366, 42 columns).
 * rays_per_scene = R / SB. points_out (N,3) fp32 optional (debug / parity), may be NULL. */
int nrf_encode_points(const float* rays, const float* z, int R, int K, int rays_per_scene,
                      const float* vol_cl, int SB, int C, int S0, int S1, int S2,
                      const float* bounds_host, int num_freqs, float freq_factor,
                      void* out, int ld_out, int out_bf16, float* points_out, void* stream);
/* The same, and touch_flags[g] = 1 if any of the samples 32 g .. 32 g + 31 has a trilinear corner inside the grid, else 0
 * ((N + 31) / 32 bytes, every one written).  A sample without such a corner has an all-zero latent (zeros padding of
 * F.grid_sample, models_embed.py:275) and its dL/dlatent is never used: nrf_mlp_bwd skips whole sample tiles by these
 * flags (NrfMlpGrads.touch_flags).  In BASELINE config 2 three rays in four miss the 1 m box altogether. */
int nrf_encode_points_touch(const float* rays, const float* z, int R, int K, int rays_per_scene,
                            const float* vol_cl, int SB, int C, int S0, int S1, int S2,
                            const float* bounds_host, int num_freqs, float freq_factor,
                            void* out, int ld_out, int out_bf16, float* points_out, uint8_t* touch_flags,
                            void* stream);

/* ---- re-layout of the touched voxels only -----------------------------------------------------------------------
 * The gather reads the in-grid trilinear corners of its samples and nothing else.  When a step's rays cross a small part
 * of the volume (2048 rays of one 200^3 scene: 1.5 % of the voxels) the dense nrf_volume_to_channels_last pass is almost
 * all wasted traffic.  nrf_mark_voxels sets flags[scene * V + voxel] = 2 for every in-grid corner of the samples (R, K)
 * whose flag is still 0; nrf_volume_to_channels_last_marked moves every 32-voxel tile that holds a 2 and sets the
 * tile's flags to 1 (per_voxel == 0).  flags: SB * V bytes, zeroed by the caller once per volume; call the pair once per render pass.
 * Tiles never flagged stay unwritten in vol_cl. */
int nrf_mark_voxels(const float* rays, const float* z, int R, int K, int rays_per_scene, int SB, int S0, int S1,
                    int S2, const float* bounds_host, uint8_t* flags, void* stream);
int nrf_volume_to_channels_last_marked(const float* src, float* dst, int SB, int C, int64_t V, uint8_t* flags,
                                       int per_voxel, void* stream);
/* per_voxel != 0: only the flagged voxels themselves are moved (C strided 4 B reads each) and marked done; the rest of
 * their tiles stays unwritten.  The sparser choice: a ray crosses a 32-voxel run in one or two voxels. */

/* ---- volume gradient: transpose of the trilinear gather (autograd of models_embed.py:275) ----
 * dlatent (N, ld) fp32; grad_cl channels-last (SB,S0,S1,S2,C), accumulated into (caller zeroes). */
int nrf_scatter_volume_grad(const float* rays, const float* z, int R, int K, int rays_per_scene,
                            const float* dlatent, int ld, float* grad_cl, int SB, int C, int S0,
                            int S1, int S2, const float* bounds_host, void* stream);

/* Atomics-free, bit-reproducible variant: counting sort of the (sample, corner) entries by voxel, then one
 * warp per voxel sums its entries in a fixed order and writes the row once.  accumulate == 0: every voxel
 * row of grad_cl is written (zeros where untouched, no prior memset needed); accumulate != 0: touched rows
 * are read-modify-written.  workspace: nrf_scatter_sorted_workspace_bytes(N = R*K, SB, V = S0*S1*S2). */
int64_t nrf_scatter_sorted_workspace_bytes(int64_t N, int SB, int64_t V);
int nrf_scatter_volume_grad_sorted(const float* rays, const float* z, int R, int K, int rays_per_scene,
                                   const float* dlatent, int ld, float* grad_cl, int SB, int C, int S0,
                                   int S1, int S2, const float* bounds_host, int accumulate,
                                   void* workspace, void* stream);

/* Both render passes of a step in ONE counting sort, gradient written once and (channels_first != 0) directly
 * in the caller's (SB,C,S0,S1,S2) layout -- the autograd of F.grid_sample at models_embed.py:275 for the coarse
 * and the fine pass together (neural_rendering.py:446,466), without the reference's zero volume per 4096-point
 * chunk and without a re-layout pass.  Pass a: z_a (R,K_a), dlat_a (R*K_a, ld_a); pass b likewise or z_b = NULL
 * for a single pass.  A voxel's entries are added in a fixed order (pass a by sample, then pass b): bit-
 * reproducible, no float atomics; voxels without entries are written as zeros (grad needs no memset).
 * C must be 64 or 128.  workspace (16 B aligned): nrf_scatter_sorted_workspace_bytes(R*(K_a+K_b), SB, V). */
int nrf_scatter_volume_grad_merged(const float* rays, int R, int rays_per_scene, const float* z_a, int K_a,
                                   const float* dlat_a, int ld_a, const float* z_b, int K_b,
                                   const float* dlat_b, int ld_b, float* grad, int channels_first, int SB,
                                   int C, int S0, int S1, int S2, const float* bounds_host, void* workspace,
                                   void* stream);

/* ---- sparse exchange of a volume gradient between ranks (BASELINE config 5) --------------------
 * ONE scene whose rays are split over the GPUs: every rank's backward produces a full-size dL/dvoxel_feat of which
 * only the voxels its own rays crossed are non-zero.  The reference has no multi-GPU path (its ancestor shards rays
 * and replicates the network: featurenerf_robo/featurenerf/src/render/nerf_embed.py:412-429); these two calls are the
 * device side of summing such gradients by exchanging (voxel index, C-vector) rows instead of the whole volume.
 * grad: (SB,C,V) if channels_first else (SB,V,C), V = S0*S1*S2.  idx: n ascending, unique int64 `scene * V + voxel`.
 *   nrf_rows_gather: rows[i, 0:C] = grad[voxel idx[i]]                        rows (n, C) fp32, 16 B aligned
 *   nrf_rows_update: grad[voxel idx[i]] = rows[i] (add == 0), += rows[i] (add != 0), or = 0 (rows == NULL)
 * No atomics (unique indices): lists applied one call after the other in rank order give the same bits everywhere. */
int nrf_rows_gather(const float* grad, int channels_first, int C, int64_t V, const int64_t* idx, int64_t n,
                    float* rows, void* stream);
int nrf_rows_update(float* grad, int channels_first, int C, int64_t V, const int64_t* idx, int64_t n,
                    const float* rows, int add, void* stream);
/* All ranks' lists at once: rows (world, cap, C), idx (world, cap), counts_host[r] valid entries of rank r (HOST array).
 * Every voxel some rank lists is set to the sum of its rows in rank order (0.0 + r_0 + r_1 + ..: the same bits on every
 * rank); other voxels keep their value.  One pass, the volume is written once per touched 32-voxel tile, never read.
 * unlisted_are_zero != 0: the caller knows every unlisted voxel of `grad` holds 0 (a gradient fresh from the scatter, whose
 * untouched voxels are zero-filled): touched tiles are then written whole - full 32 B sectors instead of 4 B pieces. */
int nrf_rows_merge(float* grad, int channels_first, int C, int64_t V, int SB, const float* rows, const int64_t* idx,
                   int64_t cap, const int64_t* counts_host, int world, int unlisted_are_zero, void* stream);

/* ---- alpha compositing (neural_rendering.py:239-243,316-359) ---------------------------------
 * field_out (N, ldo) raw; heads sigmoid(rgb), relu(sigma) (models_embed.py:444-466) applied here.
 * weights (R,K), rgb (R,3), embed (R,D), depth (R).
 * sigma_noise (R,K) or NULL: the training-time density noise of neural_rendering.py:336-337, already scaled
 * by noise_std: alpha = 1 - exp(-delta * relu(relu(raw) + sigma_noise)). */
/* reuse (or NULL): the fine pass composites the ray's K = n_first + n_new sorted samples without re-evaluating the
 * first n_first of them (the coarse samples: same point, same view direction, same MLP when share_mlp,
 * models_embed.py:113-114 -> the same field output).  perm (R,K) int32 from nrf_sort_rows: sorted position k of ray r
 * is sample p = perm[r][k] of [coarse | new]; p < n_first -> row r*n_first + p of field_out (the coarse pass's
 * buffer), else row r*n_new + (p - n_first) of field_new.  Backward: gradient rows go to d_field (R*n_first rows) /
 * d_field_new by the same rule (every row written once); the coarse pass's own nrf_composite_bwd then runs with
 * accumulate = 1 on the same d_field, so a reused sample goes through the MLP backward once with the sum. */
typedef struct {
  const float* field_new;   /* (R*n_new, ldo) raw field outputs of the newly sampled points */
  const int32_t* perm;      /* (R, K) */
  int n_first;
  void* d_field_new;        /* backward only: (R*n_new, ldg), bf16 or fp32 like d_field */
} NrfCompositeReuse;

int nrf_composite_fwd(const float* field_out, int ldo, const float* z, const float* rays, int R, int K,
                      int D, int white_bkgd, float* weights, float* rgb, float* embed, float* depth,
                      const float* sigma_noise, const NrfCompositeReuse* reuse, void* stream);

/* Backward of the above (closed form).  d_weights and d_z may be NULL.
 * d_field (N, ldg): gradient w.r.t. the RAW MLP outputs, bf16 if out_bf16 else fp32; columns
 * [4+D, ldg) are zero-filled.  accumulate != 0: added to what d_field holds (rounded once more in bf16). */
int nrf_composite_bwd(const float* field_out, int ldo, const float* z, const float* rays, int R, int K,
                      int D, int white_bkgd, const float* d_rgb, const float* d_embed,
                      const float* d_depth, const float* d_weights, void* d_field, int ldg,
                      int out_bf16, float* d_z, const float* sigma_noise, const NrfCompositeReuse* reuse,
                      int accumulate, void* stream);

/* ---- rendering losses (neural_rendering.py:653-677 and their autograd) ---------------------------
 * terms[0..3] = mean((rgb_c - t)^2), mean((rgb_f - t)^2), mean((emb_c - e)^2), mean((emb_f - e)^2)   (F.mse_loss)
 * with t = gt_rgb[scene, idx[j]] (:676), e = gt_embed[scene, idx[j]] (:683) for ray r = scene*rays_per_scene + j;
 * idx == NULL: gt_rgb (R,3) / gt_embed (R,D) are already per ray.  rgb_* (R,3), emb_* (R,D), gt_rgb (SB,n_pix,3),
 * gt_embed (SB,n_pix,D), idx int64 (rays_per_scene).  d_* (same shapes as rgb_* / emb_*, any may be NULL) receive
 * d terms[i] / d input = 2 (x - t) / numel.  partial: (R,4) fp32 scratch, 16 B aligned.  Fixed summation order. */
int nrf_render_loss(const float* rgb_c, const float* rgb_f, const float* emb_c, const float* emb_f, int R, int D,
                    int rays_per_scene, const float* gt_rgb, const float* gt_embed, int64_t n_pix,
                    const int64_t* idx, float* partial, float* terms, float* d_rgb_c, float* d_rgb_f,
                    float* d_emb_c, float* d_emb_f, void* stream);

/* ---- voxelizer: the producer of the grid the volume is encoded from (voxel_grid_real.py:175-233) -----------
 * VoxelGrid.coords_to_bounding_voxel_grid: coords (B,N,3), feats (B,N,F) or NULL (F = 0), geom (B,6) device =
 * [bb_min - res | res + 1e-12] per scene (res = (bb_max - bb_min) / S, computed by the caller in fp32 as the
 * reference does, :176-186) -> out (B,S,S,S,3+F+3+1) = [mean xyz, mean feats, voxel index / S, occupancy].
 * Points of a voxel are added in ascending point index (the CPU reference's order); no float atomics.  F <= 13.
 * workspace: nrf_voxelize_workspace_bytes(B, N, S). */
int64_t nrf_voxelize_workspace_bytes(int B, int N, int S);
int nrf_voxelize(const float* coords, const float* feats, int B, int N, int F, const float* geom, int S,
                 float* out, void* workspace, void* stream);

/* ---- GEMM building block of the field MLP ----------------------------------------------------
 * v = resid + mask( [A0 | A1 | A2] . B^T + bias ),   B (N,K) row-major (nn.Linear layout)
 *   A[i] (M,K[i]) lda[i]: up to three operand matrices concatenated along K (K[i] = 0: unused)
 *   operand type: bf16 (NRF_PREC_BF16), fp16 (NRF_PREC_FP16) or fp32 (NRF_PREC_FP32); applies to A, B, mask_src,
 *   resid, out_act and out_act2
 *   bias (N) fp32 or NULL; mask_src (M,N) or NULL: v is zeroed where mask_src <= 0 (ReLU gate)
 *   resid (M,N) or NULL (may alias out_act: the update is element-wise in place)
 *   out_act / out_act2 (M,n_store) or NULL: v, ReLU'd when the matching relu flag is set
 *   out_f32 (M,n_store) fp32 or NULL: v.  bf16 mode: out_f32 excludes mask/resid/out_act*.
 * Only columns < n_store are written.  bf16 mode: every K[i] a multiple of 64, N a multiple of 128,
 * all pointers 16 B aligned with 16 B-multiple row pitches. */
typedef struct {
  const void* A[3]; int K[3]; int lda[3];
  const void* B;  int ldb;
  int M; int N; int n_store;
  const float* bias;
  const void* mask_src; int ldmask;
  const void* resid; int ldr;
  void* out_act; int ldact; int relu_act;
  void* out_act2; int ldact2; int relu_act2;
  float* out_f32; int ldo;
} NrfGemm;
int nrf_gemm(const NrfGemm* g, int precision, void* stream);

/* Weight gradient dW (N,K) += G^T . A with G (M,N) ldg, A (M,K) lda operand-typed; dW fp32 ldw.
 * workspace: fp32, >= nrf_wgrad_workspace_bytes(N,K) (bf16 mode; split-K partials), may be NULL
 * in fp32 mode.  Only rows < n_valid and columns < k_valid of dW are touched.
 * dbias (n_valid) += column sums of G when non-NULL.  NRF_PREC_FP16: G is bf16 (a gradient), A fp16 (an activation). */
int64_t nrf_wgrad_workspace_bytes(int N, int K);
int nrf_wgrad(const void* G, int ldg, const void* A, int lda, int M, int N, int K, int n_valid,
              int k_valid, float* dW, int ldw, float* dbias, void* workspace, int precision,
              void* stream);

/* ---- the ResnetFC field MLP (resnetfc.py:55-64,146-195) --------------------------------------
 * Parameters in the reference's own layout (fp32, nn.Linear (out,in) row-major).
 * lin_z has n_lin_z = min(combine_layer, n_blocks) entries. */
#define NRF_MAX_BLOCKS 8
typedef struct {
  int d_in;        /* 42 = PE(39) + viewdir(3)                     */
  int d_latent;    /* C                                             */
  int d_hidden;    /* 512                                           */
  int d_out;       /* 4 + D                                         */
  int n_blocks;    /* 5                                             */
  int n_lin_z;     /* 3                                             */
  const float* lin_in_w;  const float* lin_in_b;
  const float* lin_out_w; const float* lin_out_b;
  const float* fc0_w[NRF_MAX_BLOCKS]; const float* fc0_b[NRF_MAX_BLOCKS];
  const float* fc1_w[NRF_MAX_BLOCKS]; const float* fc1_b[NRF_MAX_BLOCKS];
  const float* lin_z_w[NRF_MAX_BLOCKS]; const float* lin_z_b[NRF_MAX_BLOCKS];
} NrfMlpParams;

/* Same shape as NrfMlpParams, pointing at fp32 gradient buffers that are ACCUMULATED into. */
typedef struct {
  float* lin_in_w;  float* lin_in_b;
  float* lin_out_w; float* lin_out_b;
  float* fc0_w[NRF_MAX_BLOCKS]; float* fc0_b[NRF_MAX_BLOCKS];
  float* fc1_w[NRF_MAX_BLOCKS]; float* fc1_b[NRF_MAX_BLOCKS];
  float* lin_z_w[NRF_MAX_BLOCKS]; float* lin_z_b[NRF_MAX_BLOCKS];
  int deterministic;   /* bf16 mode: reduce the sample splits of every weight gradient in a fixed order through the
                          scratch workspace (bit-reproducible, ~10 % slower wgrads) instead of fp32 atomics */
  const void* d_last;  /* NULL, or (N,d_hidden) gradient-typed: dL/dx_nb that reaches the last residual stream directly
                          (the MLP's second return value, resnetfc.py:192-195, composited instead of the embedding when
                          ret_last_feat: neural_rendering.py:332-334); nrf_mlp_bwd_layered only.  After
                          nrf_mlp_fwd_layered the last layer of `acts` holds x_nb itself, (N,d_hidden) operand-typed. */
  const void* touch_flags; /* NULL, or the (N + 31) / 32 bytes nrf_encode_points_touch wrote for these N samples: dlatent
                          is then only computed for 128-sample tiles with a flag set; the rows of the other tiles are
                          left unwritten (nothing reads them: the volume scatter visits in-grid corners only) */
  void* dlatent_ready_event; /* NULL, or a cudaEvent_t the call records on `stream` as soon as `dlatent` is complete: in
                          the default (fused, one-launch weight-gradient) path that is BEFORE the weight gradients are
                          enqueued - the dL/dz GEMM is then issued first - so that a caller's volume scatter on a second
                          stream runs under them; in every other path, when the whole backward is enqueued */
} NrfMlpGrads;

/* Sizes (bytes) of the caller-provided buffers for a given shape / precision. */
typedef struct {
  int kin_pad;            /* row length of the field-input matrix                 */
  int dout_pad;           /* row length of d_field (gradient of raw outputs)      */
  int64_t packed_bytes;   /* packed / transposed weight cache                     */
  int64_t fwd_bytes_per_sample;   /* activations kept by nrf_mlp_fwd              */
  int64_t bwd_bytes_per_sample;   /* scratch of nrf_mlp_bwd                       */
  int64_t bwd_fixed_bytes;        /* split-K partials etc.                        */
} NrfMlpSizes;
int nrf_mlp_sizes(const NrfMlpParams* p, int precision, NrfMlpSizes* out);

/* Packs the fp32 parameters into the operand layout the GEMMs read ([W_z0 | W_in] etc.). */
int nrf_mlp_pack(const NrfMlpParams* p, int precision, void* packed, void* stream);

/* Forward over N samples.  field_in (N,kin_pad) operand-typed; field_out (N, ldo) fp32 raw outputs with
 * ldo = d_out rounded up to a multiple of 4 (pad columns are written as zeros; ldo == d_out for the reference's 388 / 516).
 * acts: fwd_bytes_per_sample*N bytes, the operands kept for the backward: relu(x'_b) (b = 0..n_blocks), then
 * relu(net_b) (b < n_blocks), each (N,d_hidden) operand-typed, then one layer of scratch.
 * bf16 mode, when nrf_mlp_fused_supported(): ONE persistent tcgen05 kernel runs all layers per 256-sample tile
 * (csrc/mlp_fused.cu: activations never leave the SM; `acts` is written as a side effect and may be NULL for
 * inference, in which case nothing but field_out is written).  Otherwise a chain of fused-epilogue GEMMs, one
 * per layer (acts required).  nrf_mlp_fwd_layered always runs the chain (A/B timing, parity tests). */
int nrf_mlp_fused_supported(const NrfMlpParams* p, int precision);
int nrf_mlp_fwd(const NrfMlpParams* p, const void* packed, int precision, const void* field_in,
                int64_t N, void* acts, float* field_out, void* stream);
int nrf_mlp_fwd_layered(const NrfMlpParams* p, const void* packed, int precision, const void* field_in,
                        int64_t N, void* acts, float* field_out, void* stream);
/* nrf_mlp_fwd with the flags nrf_encode_points_touch wrote for these N samples ((N + 31) / 32 bytes, or NULL): in the
 * fused kernel a 256-sample tile without any flag set has an all-zero latent (zeros padding of F.grid_sample,
 * models_embed.py:275), so the k-panels that multiply the latent - two of the first layer's three, the lin_z tails of
 * fc_1 (resnetfc.py:181-182) - are neither loaded nor multiplied for it: the same accumulators bit for bit (they
 * would have added exact zeros), 6 % fewer MMAs on such a tile.  The layer-by-layer paths ignore the flags. */
int nrf_mlp_fwd_touch(const NrfMlpParams* p, const void* packed, int precision, const void* field_in,
                      int64_t N, void* acts, float* field_out, const uint8_t* touch_flags, void* stream);

/* Backward: d_field (N,dout_pad) operand-typed gradient of the raw outputs (from
 * nrf_composite_bwd); accumulates parameter gradients into `grads` and writes
 * dlatent (N,d_latent) fp32.  scratch: bwd_bytes_per_sample*N + bwd_fixed_bytes bytes. */
int nrf_mlp_bwd(const NrfMlpParams* p, const void* packed, int precision, const void* field_in,
                int64_t N, const void* acts, const void* d_field, const NrfMlpGrads* grads,
                float* dlatent, void* scratch, void* stream);
/* bf16 mode, when nrf_mlp_fused_supported(): nrf_mlp_bwd runs the whole data-gradient chain (lin_out^T, fc_1^T,
 * fc_0^T per block, ReLU gates from `acts`, residual gradient in registers) as ONE persistent tcgen05 kernel that
 * streams every dL/dx'_b and dL/dnet_b into `scratch`, followed by the weight-gradient GEMMs and one GEMM for
 * dL/dlatent.  It reads the ReLU gates bit-packed (64 B per sample and saved operand) from the last layer of
 * `acts`, where only the fused nrf_mlp_fwd writes them: pair nrf_mlp_fwd with nrf_mlp_bwd and nrf_mlp_fwd_layered
 * with nrf_mlp_bwd_layered (which always runs the per-layer chain). */
int nrf_mlp_bwd_layered(const NrfMlpParams* p, const void* packed, int precision, const void* field_in,
                        int64_t N, const void* acts, const void* d_field, const NrfMlpGrads* grads,
                        float* dlatent, void* scratch, void* stream);

/* ---- measurement hooks (bench.py) -------------------------------------------------------------
 * nrf_launch_count: number of kernels this library has launched in this process (monotonic).
 * nrf_timing_begin: from now on bracket every kernel launch with a CUDA event pair on its own stream.
 * nrf_timing_end:   synchronise, sum the durations per category into ms[cat] / launches[cat]
 *                   (arrays of NRF_TIMING_CATEGORIES), stop recording. */
#define NRF_CAT_GEMM 0        /* tcgen05 forward / dgrad GEMMs (gemm_tc_kernel)   */
#define NRF_CAT_WGRAD 1       /* tcgen05 weight-gradient GEMMs (wgrad_tc_kernel)  */
#define NRF_CAT_ENCODE 2      /* points + trilinear gather + positional encoding  */
#define NRF_CAT_COMPOSITE_FWD 3
#define NRF_CAT_COMPOSITE_BWD 4
#define NRF_CAT_SCATTER 5     /* volume-gradient scatter                          */
#define NRF_CAT_TRANSPOSE 6   /* volume re-layout                                 */
#define NRF_CAT_COLSUM 7      /* bias gradients                                   */
#define NRF_CAT_SAMPLING 8    /* raygen, coarse / fine sampling, sort             */
#define NRF_CAT_SIMT 9        /* fp32 parity-mode GEMMs                           */
#define NRF_CAT_MISC 10       /* weight packing, small vector kernels             */
#define NRF_CAT_FUSED_FWD 11  /* whole-MLP fused forward kernel (mlp_fused_fwd_kernel) */
#define NRF_CAT_FUSED_BWD 12  /* whole-MLP fused dgrad kernel                           */
#define NRF_TIMING_CATEGORIES 13
int64_t nrf_launch_count(void);
int nrf_timing_begin(void);
int nrf_timing_end(double* ms, int64_t* launches);

#ifdef __cplusplus
}
#endif
#endif /* NRF_B200_H */
