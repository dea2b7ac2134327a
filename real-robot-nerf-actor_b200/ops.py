"""Tensor-level wrappers over the C ABI (include/nrf_b200.h).  CUDA tensors in, CUDA tensors out.

Each function validates device / dtype / contiguity, allocates outputs through PyTorch's caching
allocator, and enqueues one C-ABI call on the current CUDA stream.  No CPU path exists: a CPU tensor
is an error.
"""
from __future__ import annotations

import ctypes as C
import functools

import torch

from . import _lib
from ._lib import NRF_PREC_BF16, NRF_PREC_BF16X3, NRF_PREC_FP16, NRF_PREC_FP32, check, ptr, stream_ptr

PRECISIONS = {"bf16": NRF_PREC_BF16, "fp32": NRF_PREC_FP32, "fp16": NRF_PREC_FP16, "bf16x3": NRF_PREC_BF16X3}


def _on_tensor_device(fn):
    """Runs `fn` with the CUDA device of its tensor arguments current (kernels are launched on the current device's
    current stream: without this a renderer on cuda:1 called while cuda:0 is current would launch on GPU 0 with GPU 1
    pointers).  Tensors on different devices raise.  One integer comparison per call when the device is already
    current (the one-process-per-GPU layout)."""
    @functools.wraps(fn)
    def wrapper(*args, **kw):
        dev = None
        for a in args:
            if isinstance(a, FieldMLP):
                a = a.params["lin_in.weight"]
            if isinstance(a, torch.Tensor) and a.is_cuda:
                if dev is None:
                    dev = a.device
                elif a.device != dev:
                    raise _lib.NrfError(f"{fn.__name__}: tensors on different devices ({dev} and {a.device})")
        if dev is not None and dev.index != torch.cuda.current_device():
            with torch.cuda.device(dev):
                return fn(*args, **kw)
        return fn(*args, **kw)
    return wrapper


def _f32(t: torch.Tensor, name: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise _lib.NrfError(f"{name}: expected a CUDA tensor (there is no CPU fallback)")
    if t.dtype != torch.float32:
        raise _lib.NrfError(f"{name}: expected float32, got {t.dtype}")
    return t.contiguous()


def act_dtype(precision: int):
    """dtype of the forward operands (field input, saved activations) of a precision mode."""
    return {NRF_PREC_BF16: torch.bfloat16, NRF_PREC_FP16: torch.float16}.get(precision, torch.float32)


def grad_dtype(precision: int):
    """dtype of the gradient operands (d_field ...): bf16 in both tensor-core modes (see NRF_PREC_FP16)."""
    return torch.bfloat16 if precision in (NRF_PREC_BF16, NRF_PREC_FP16) else torch.float32


_OUT_KIND = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}


# ------------------------------------------------------------------------------------- rays
_INTR_CACHE = {}


def _device_intrinsics(focal, c, width, height, device):
    """[fx, fy, cx, cy] as a 4-float device tensor, built with device ops only (no read-back).  The last result is kept
    per (focal, c) tensor object + version: a training loop passes the same tensors every step."""
    ver = lambda t: (id(t), t._version) if torch.is_tensor(t) else t
    key = (ver(focal), ver(c), width, height, str(device))
    hit = _INTR_CACHE.get("last")
    if hit is not None and hit[0] == key and hit[1] is focal and hit[2] is c:
        return hit[3]
    f = torch.as_tensor(focal, dtype=torch.float32, device=device).reshape(-1)
    intr = torch.empty(4, device=device, dtype=torch.float32)
    intr[0:2] = f[0:2] if f.numel() >= 2 else f[0]
    if c is None:
        intr[2].fill_(width * 0.5)            # fill kernels, not host -> device copies (CUDA-graph capturable)
        intr[3].fill_(height * 0.5)
    else:
        cc = torch.as_tensor(c, dtype=torch.float32, device=device).reshape(-1)
        intr[2:4] = cc[0:2] if cc.numel() >= 2 else cc[0]
    _INTR_CACHE["last"] = (key, focal, c, intr)          # holding focal / c keeps their ids from being reused
    return intr


# nrf_raygen_ex flags (include/nrf_b200.h)
NRF_RAYGEN_DIR_FMA_ASC, NRF_RAYGEN_DIR_FMA_DESC, NRF_RAYGEN_DIR_SPLIT, NRF_RAYGEN_PIXEL_RECIP = 1, 2, 3, 4
NRF_RAYGEN_NORM_XY_Z, NRF_RAYGEN_NORM_XZ_Y, NRF_RAYGEN_NORM_X_YZ = 8, 16, 24
NRF_RAYGEN_CUDA_EAGER = NRF_RAYGEN_DIR_SPLIT | NRF_RAYGEN_PIXEL_RECIP | NRF_RAYGEN_NORM_XZ_Y
RAYGEN_FLAGS = 0        # default rounding pattern of gen_rays: 0 = ATen's CPU kernels (what the golden fixtures hold);
                        # NRF_RAYGEN_CUDA_EAGER = bit-identical rays to the reference run on a GPU (like GATHER_FMA below)


@_on_tensor_device
def raygen(poses, width, height, focal, z_near, z_far, c=None, flags=None):
    """utils.py:477-506 gen_rays.  poses (B,4,4) -> rays (B,H,W,8).
    A `focal` / `c` that lives on the GPU stays there (4 floats handed to the kernel by pointer): reading it back would
    stall the host on everything queued before, once per step.
    flags: rounding pattern (NRF_RAYGEN_*); None = the module default RAYGEN_FLAGS."""
    flags = RAYGEN_FLAGS if flags is None else flags
    poses = _f32(poses, "poses")
    B = poses.shape[0]
    rays = torch.empty(B, height, width, 8, device=poses.device, dtype=torch.float32)
    on_dev = (torch.is_tensor(focal) and focal.is_cuda) or (torch.is_tensor(c) and c.is_cuda)
    if on_dev:
        intr = _device_intrinsics(focal, c, width, height, poses.device)
        check(_lib.load().nrf_raygen_ex(ptr(poses), B, width, height, 0.0, 0.0, 0.0, 0.0, float(z_near), float(z_far),
                                        ptr(rays), ptr(intr), int(flags), stream_ptr()), "nrf_raygen")
        return rays
    f = torch.as_tensor(focal, dtype=torch.float32).reshape(-1)
    fx, fy = (float(f[0]), float(f[0])) if f.numel() == 1 else (float(f[0]), float(f[1]))
    if c is None:
        cx, cy = width * 0.5, height * 0.5
    else:
        cc = torch.as_tensor(c, dtype=torch.float32).reshape(-1)
        cx, cy = (float(cc[0]), float(cc[0])) if cc.numel() == 1 else (float(cc[0]), float(cc[1]))
    check(_lib.load().nrf_raygen_ex(ptr(poses), B, width, height, fx, fy, cx, cy, float(z_near),
                                    float(z_far), ptr(rays), None, int(flags), stream_ptr()), "nrf_raygen")
    return rays


@_on_tensor_device
def sample_coarse(rays, n_coarse, jitter=None, lindisp=False):
    """neural_rendering.py:159-176.  rays (R,8) -> z (R,Kc)."""
    rays = _f32(rays, "rays")
    R = rays.shape[0]
    step = 1.0 / n_coarse
    base = torch.linspace(0, 1 - step, n_coarse, device=rays.device)     # the reference's own op
    if jitter is not None:
        jitter = _f32(jitter, "jitter")
        assert jitter.shape == (R, n_coarse)
    z = torch.empty(R, n_coarse, device=rays.device, dtype=torch.float32)
    check(_lib.load().nrf_sample_coarse(ptr(rays), R, n_coarse, ptr(base), ptr(jitter), int(lindisp),
                                        ptr(z), stream_ptr()), "nrf_sample_coarse")
    return z


FINE_CUDA_EAGER = False  # True: the in-kernel cdf in ATen's CUDA association order (Kc = 64 / 128), like RAYGEN_FLAGS / GATHER_FMA
NRF_FINE_CUDA_EAGER = 0x100


@_on_tensor_device
def sample_fine(rays, weights, n_coarse, u, jitter=None, lindisp=False, cdf=None, out=None,
                want_ind=False, cuda_eager=None):
    """neural_rendering.py:179-207.  Returns z (R,Kf) (written into out[:, :Kf] when given).
    cuda_eager: build the cdf the way a GPU run of the reference does (bit-identical indices and depths to CUDA-eager
    PyTorch for 64 / 128 coarse samples); default: the CPU back end's pattern (module default FINE_CUDA_EAGER)."""
    rays = _f32(rays, "rays")
    u = _f32(u, "u")
    R, Kf = u.shape
    weights = _f32(weights, "weights") if weights is not None else None
    cdf = _f32(cdf, "cdf") if cdf is not None else None
    jitter = _f32(jitter, "jitter") if jitter is not None else None
    if out is None:
        out = torch.empty(R, Kf, device=rays.device, dtype=torch.float32)
    assert out.is_contiguous() and out.shape[0] == R and out.shape[1] >= Kf
    ind = torch.empty(R, Kf, device=rays.device, dtype=torch.float32) if want_ind else None
    eager = FINE_CUDA_EAGER if cuda_eager is None else cuda_eager
    mode = int(bool(lindisp)) | (NRF_FINE_CUDA_EAGER if eager and cdf is None and n_coarse in (64, 128) else 0)
    check(_lib.load().nrf_sample_fine(ptr(rays), ptr(weights), ptr(cdf), R, n_coarse, ptr(u), ptr(jitter),
                                      Kf, mode, ptr(out), out.shape[1], ptr(ind), stream_ptr()),
          "nrf_sample_fine")
    return (out, ind) if want_ind else out


@_on_tensor_device
def sort_rows(z, want_perm=False):
    """neural_rendering.py:463.  In-place ascending sort of each row; optional int32 permutation."""
    assert z.is_cuda and z.dtype == torch.float32 and z.is_contiguous()
    R, K = z.shape
    perm = torch.empty(R, K, device=z.device, dtype=torch.int32) if want_perm else None
    check(_lib.load().nrf_sort_rows(ptr(z), R, K, ptr(perm), stream_ptr()), "nrf_sort_rows")
    return (z, perm) if want_perm else z


# ----------------------------------------------------------------------------------- volume
@_on_tensor_device
def volume_to_channels_last(vol):
    """(SB,C,S0,S1,S2) -> (SB,S0,S1,S2,C)."""
    vol = _f32(vol, "voxel_feat")
    SB, Cc, S0, S1, S2 = vol.shape
    out = torch.empty(SB, S0, S1, S2, Cc, device=vol.device, dtype=torch.float32)
    check(_lib.load().nrf_volume_to_channels_last(ptr(vol), ptr(out), SB, Cc, S0 * S1 * S2, stream_ptr()),
          "nrf_volume_to_channels_last")
    return out


class TouchedRelayout:
    """The channels-last copy of a volume, filled in only where the rays go (nrf_mark_voxels +
    nrf_volume_to_channels_last_marked): `add(rays, z, rays_per_scene)` before a pass's gather moves the 32-voxel tiles
    that pass touches and no earlier pass has moved.  `vol_cl` elsewhere is unwritten memory - which the gather never
    reads (it reads the in-grid corners of its own samples only)."""

    def __init__(self, vol, bounds, per_voxel=True):
        self.per_voxel = per_voxel          # move the flagged voxels themselves (default) or their whole 32-voxel tiles
        self.vol = _f32(vol, "voxel_feat")
        SB, Cc, S0, S1, S2 = self.vol.shape
        self.shape = (SB, S0, S1, S2, Cc)
        self.vol_cl = torch.empty(self.shape, device=vol.device, dtype=torch.float32)
        self.flags = torch.zeros(SB * S0 * S1 * S2, device=vol.device, dtype=torch.uint8)
        self.bh = _bounds_host(bounds)

    def add(self, rays, z, rays_per_scene):
        SB, S0, S1, S2, Cc = self.shape
        rays, z = _f32(rays, "rays"), _f32(z, "z")
        R, K = z.shape
        with torch.cuda.device(self.vol.device):
            lib = _lib.load()
            check(lib.nrf_mark_voxels(ptr(rays), ptr(z), R, K, rays_per_scene, SB, S0, S1, S2,
                                      C.cast(self.bh, C.c_void_p), ptr(self.flags), stream_ptr()), "nrf_mark_voxels")
            check(lib.nrf_volume_to_channels_last_marked(ptr(self.vol), ptr(self.vol_cl), SB, Cc, S0 * S1 * S2,
                                                         ptr(self.flags), int(self.per_voxel), stream_ptr()),
                  "nrf_volume_to_channels_last_marked")
        return self.vol_cl


@_on_tensor_device
def volume_to_channels_first(vol_cl):
    """(SB,S0,S1,S2,C) -> (SB,C,S0,S1,S2)."""
    vol_cl = _f32(vol_cl, "volume")
    SB, S0, S1, S2, Cc = vol_cl.shape
    out = torch.empty(SB, Cc, S0, S1, S2, device=vol_cl.device, dtype=torch.float32)
    check(_lib.load().nrf_volume_to_channels_first(ptr(vol_cl), ptr(out), SB, Cc, S0 * S1 * S2, stream_ptr()),
          "nrf_volume_to_channels_first")
    return out


_BOUNDS_CACHE = {}


def _bounds_host(bounds):
    """6 floats as a ctypes array; the last conversion is kept per tensor object (the renderer hands over the same
    cached host tensor several times a step)."""
    hit = _BOUNDS_CACHE.get("last")
    if hit is not None and hit[0] is bounds and (not torch.is_tensor(bounds) or hit[1] == bounds._version):
        return hit[2]
    b = torch.as_tensor(bounds, dtype=torch.float32).reshape(-1).cpu()
    assert b.numel() == 6
    arr = (C.c_float * 6)(*[float(v) for v in b])
    _BOUNDS_CACHE["last"] = (bounds, bounds._version if torch.is_tensor(bounds) else None, arr)
    return arr


GATHER_FMA = False      # True: accumulate the trilinear corners with FMAs (ATen's CUDA grid_sampler_3d rounding, i.e.
                        # bit-identical latents to the reference run on a GPU); default: ATen's CPU rounding


@_on_tensor_device
def encode_points(rays, z, rays_per_scene, vol_cl, bounds, num_freqs=6, freq_factor=1.5, ld_out=None,
                  precision=NRF_PREC_BF16, want_points=False, out=None, fma=None, want_touch=False,
                  code_viewdirs=False):
    """Field-input rows [latent | PE | viewdir | 0] for every sample (see nrf_encode_points).
    code_viewdirs: use_code_viewdirs (models_embed.py:370-372) - rows [latent | PE([xyz | viewdir]) | 0].
    want_touch: also return one uint8 per 32 consecutive samples, 1 if any of them has a corner inside the grid
    (nrf_encode_points_touch; what FieldMLP.backward(touch=...) skips whole dL/dlatent tiles by)."""
    rays = _f32(rays, "rays")
    z = _f32(z, "z")
    vol_cl = _f32(vol_cl, "volume")
    R, K = z.shape
    SB, S0, S1, S2, Cc = vol_cl.shape
    need = Cc + 6 + (12 if code_viewdirs else 6) * num_freqs
    if ld_out is None:
        ld_out = (need + 63) // 64 * 64
    if out is None:
        out = torch.empty(R * K, ld_out, device=rays.device, dtype=act_dtype(precision))
    pts = torch.empty(R * K, 3, device=rays.device, dtype=torch.float32) if want_points else None
    bh = _bounds_host(bounds)
    kind = _OUT_KIND[out.dtype] | (0x100 if (GATHER_FMA if fma is None else fma) else 0) | (0x200 if code_viewdirs else 0)
    if want_touch:
        touch = torch.empty((R * K + 31) // 32, device=rays.device, dtype=torch.uint8)
        check(_lib.load().nrf_encode_points_touch(ptr(rays), ptr(z), R, K, rays_per_scene, ptr(vol_cl), SB, Cc, S0, S1,
                                                  S2, C.cast(bh, C.c_void_p), num_freqs, float(freq_factor), ptr(out),
                                                  ld_out, kind, ptr(pts), ptr(touch), stream_ptr()),
              "nrf_encode_points_touch")
        return (out, pts, touch) if want_points else (out, touch)
    check(_lib.load().nrf_encode_points(ptr(rays), ptr(z), R, K, rays_per_scene, ptr(vol_cl), SB, Cc, S0, S1,
                                        S2, C.cast(bh, C.c_void_p), num_freqs, float(freq_factor), ptr(out),
                                        ld_out, kind, ptr(pts), stream_ptr()),
          "nrf_encode_points")
    return (out, pts) if want_points else out


@_on_tensor_device
def scatter_volume_grad(rays, z, rays_per_scene, dlatent, grad_cl, bounds):
    """grad_cl (SB,S0,S1,S2,C) += transpose-of-gather(dlatent (N,C))."""
    rays = _f32(rays, "rays")
    z = _f32(z, "z")
    dlatent = _f32(dlatent, "dlatent")
    assert grad_cl.is_cuda and grad_cl.dtype == torch.float32 and grad_cl.is_contiguous()
    R, K = z.shape
    SB, S0, S1, S2, Cc = grad_cl.shape
    bh = _bounds_host(bounds)
    check(_lib.load().nrf_scatter_volume_grad(ptr(rays), ptr(z), R, K, rays_per_scene, ptr(dlatent),
                                              dlatent.shape[1], ptr(grad_cl), SB, Cc, S0, S1, S2,
                                              C.cast(bh, C.c_void_p), stream_ptr()),
          "nrf_scatter_volume_grad")
    return grad_cl


@_on_tensor_device
def scatter_volume_grad_sorted(rays, z, rays_per_scene, dlatent, grad_cl, bounds, accumulate=False):
    """Atomics-free, bit-reproducible scatter (counting sort by voxel + one warp per voxel)."""
    rays = _f32(rays, "rays")
    z = _f32(z, "z")
    dlatent = _f32(dlatent, "dlatent")
    assert grad_cl.is_cuda and grad_cl.dtype == torch.float32 and grad_cl.is_contiguous()
    R, K = z.shape
    SB, S0, S1, S2, Cc = grad_cl.shape
    lib = _lib.load()
    ws = torch.empty(lib.nrf_scatter_sorted_workspace_bytes(R * K, SB, S0 * S1 * S2), device=z.device,
                     dtype=torch.uint8)
    bh = _bounds_host(bounds)
    check(lib.nrf_scatter_volume_grad_sorted(ptr(rays), ptr(z), R, K, rays_per_scene, ptr(dlatent),
                                             dlatent.shape[1], ptr(grad_cl), SB, Cc, S0, S1, S2,
                                             C.cast(bh, C.c_void_p), int(accumulate), ptr(ws), stream_ptr()),
          "nrf_scatter_volume_grad_sorted")
    return grad_cl


def _rows_args(grad, idx):
    assert grad.is_cuda and grad.dtype == torch.float32 and grad.dim() == 5
    assert idx.dtype == torch.int64 and idx.is_contiguous() and idx.device == grad.device
    SB, Cc = grad.shape[:2]
    V = grad[0, 0].numel()
    if grad.is_contiguous():
        cf = 1
    elif grad.is_contiguous(memory_format=torch.channels_last_3d):
        cf = 0
    else:
        raise ValueError("volume gradient must be contiguous or channels_last_3d")
    return cf, Cc, V


@_on_tensor_device
def rows_gather(grad, idx, out=None):
    """rows (n, C) fp32 = the voxel rows `idx` (ascending unique flat `scene * V + voxel`, int64) of the volume gradient
    `grad` (SB,C,S0,S1,S2), contiguous or channels_last_3d (nrf_rows_gather)."""
    cf, Cc, V = _rows_args(grad, idx)
    n = idx.numel()
    if out is None:
        out = torch.empty((n, Cc), device=grad.device, dtype=torch.float32)
    assert out.is_contiguous() and out.dtype == torch.float32 and out.shape[0] >= n and out.shape[1] == Cc
    check(_lib.load().nrf_rows_gather(ptr(grad), cf, Cc, V, ptr(idx), n, ptr(out), stream_ptr()), "nrf_rows_gather")
    return out


@_on_tensor_device
def rows_update(grad, idx, rows=None, add=True):
    """grad[voxel idx[i]] += rows[i] (add), = rows[i] (not add) or = 0 (rows None), in place, no atomics (nrf_rows_update)."""
    cf, Cc, V = _rows_args(grad, idx)
    n = idx.numel()
    if rows is not None:
        assert rows.is_contiguous() and rows.dtype == torch.float32 and rows.shape[0] >= n and rows.shape[1] == Cc
    check(_lib.load().nrf_rows_update(ptr(grad), cf, Cc, V, ptr(idx), n, ptr(rows) if rows is not None else None,
                                      int(bool(add)), stream_ptr()), "nrf_rows_update")
    return grad


@_on_tensor_device
def rows_merge(grad, all_rows, all_idx, counts, unlisted_are_zero=False):
    """grad[voxel] = sum in rank order of the ranks' rows for every listed voxel (nrf_rows_merge).  all_rows
    (world, cap, C) fp32, all_idx (world, cap) int64 ascending per rank, counts: list of world ints.
    unlisted_are_zero: every voxel of `grad` that no list names is known to hold 0 (whole-tile writes)."""
    assert all_idx.dtype == torch.int64 and all_idx.is_contiguous() and all_rows.is_contiguous()
    cf, Cc, V = _rows_args(grad, all_idx)
    world, cap = all_idx.shape
    assert all_rows.shape == (world, cap, Cc) and all_rows.dtype == torch.float32 and len(counts) == world
    cnt = (C.c_int64 * world)(*[int(c) for c in counts])
    check(_lib.load().nrf_rows_merge(ptr(grad), cf, Cc, V, grad.shape[0], ptr(all_rows), ptr(all_idx), cap,
                                     C.cast(cnt, C.c_void_p), world, int(bool(unlisted_are_zero)), stream_ptr()),
          "nrf_rows_merge")
    return grad


@_on_tensor_device
def scatter_volume_grad_merged(rays, rays_per_scene, passes, grad, channels_first, bounds, want_counts=False):
    """ONE atomics-free scatter for all render passes of a step (see nrf_scatter_volume_grad_merged).

    passes: [(z (R,K), dlatent (R*K, >=C) fp32), ...] (one or two); grad: (SB,C,S0,S1,S2) if channels_first
    else (SB,S0,S1,S2,C), fully overwritten (no memset needed).
    want_counts: also return the per-voxel entry counts (SB*V,) int32 the counting sort produced (the first array of the
    workspace): `counts > 0` marks exactly the voxels whose gradient row was computed rather than zero-filled - what
    parallel.sparse_allreduce_volume_grad needs, without a pass over the dense gradient."""
    rays = _f32(rays, "rays")
    assert 1 <= len(passes) <= 2
    assert grad.is_cuda and grad.dtype == torch.float32 and grad.is_contiguous()
    if channels_first:
        SB, Cc, S0, S1, S2 = grad.shape
    else:
        SB, S0, S1, S2, Cc = grad.shape
    R = rays.shape[0]
    za, da = _f32(passes[0][0], "z"), _f32(passes[0][1], "dlatent")
    zb, db = (_f32(passes[1][0], "z"), _f32(passes[1][1], "dlatent")) if len(passes) == 2 else (None, None)
    n_tot = R * (za.shape[1] + (zb.shape[1] if zb is not None else 0))
    lib = _lib.load()
    ws = torch.empty(lib.nrf_scatter_sorted_workspace_bytes(n_tot, SB, S0 * S1 * S2), device=rays.device,
                     dtype=torch.uint8)
    bh = _bounds_host(bounds)
    check(lib.nrf_scatter_volume_grad_merged(ptr(rays), R, rays_per_scene, ptr(za), za.shape[1], ptr(da),
                                             da.shape[1], ptr(zb), zb.shape[1] if zb is not None else 0, ptr(db),
                                             db.shape[1] if db is not None else 0, ptr(grad), int(channels_first),
                                             SB, Cc, S0, S1, S2, C.cast(bh, C.c_void_p), ptr(ws), stream_ptr()),
          "nrf_scatter_volume_grad_merged")
    if want_counts:
        return grad, ws[:4 * SB * S0 * S1 * S2].view(torch.int32).clone()
    return grad


# ------------------------------------------------------------------------------- compositing
def _reuse_struct(reuse, K, ld, d_field_new=None):
    """reuse = (field_new (R*n_new, ld) fp32, perm (R,K) int32, n_first) -> NrfCompositeReuse (kept alive by the caller)."""
    if reuse is None:
        return None
    field_new, perm, n_first = reuse
    field_new = _f32(field_new, "field_new")
    assert perm.is_cuda and perm.dtype == torch.int32 and perm.is_contiguous() and perm.shape[1] == K
    assert field_new.shape[1] == ld and 0 < n_first < K
    st = _lib.NrfCompositeReuse()
    st.field_new, st.perm, st.n_first = field_new.data_ptr(), perm.data_ptr(), int(n_first)
    st.d_field_new = d_field_new.data_ptr() if d_field_new is not None else None
    return st


@_on_tensor_device
def composite_fwd(field_out, z, rays, D, white_bkgd=False, sigma_noise=None, reuse=None):
    """neural_rendering.py:339-359 on RAW MLP outputs (N, 4+D).  -> weights, rgb, embed, depth.
    sigma_noise (R,K): training-time density noise, already scaled by noise_std (neural_rendering.py:336-337)."""
    if sigma_noise is not None:
        sigma_noise = _f32(sigma_noise, "sigma_noise")
        assert sigma_noise.shape == z.shape
    field_out = _f32(field_out, "field_out")
    z = _f32(z, "z")
    rays = _f32(rays, "rays")
    R, K = z.shape
    dev = z.device
    ru = _reuse_struct(reuse, K, field_out.shape[1])
    w = torch.empty(R, K, device=dev, dtype=torch.float32)
    rgb = torch.empty(R, 3, device=dev, dtype=torch.float32)
    emb = torch.empty(R, D, device=dev, dtype=torch.float32)
    dep = torch.empty(R, device=dev, dtype=torch.float32)
    check(_lib.load().nrf_composite_fwd(ptr(field_out), field_out.shape[1], ptr(z), ptr(rays), R, K, D,
                                        int(white_bkgd), ptr(w), ptr(rgb), ptr(emb), ptr(dep), ptr(sigma_noise),
                                        C.byref(ru) if ru is not None else None, stream_ptr()),
          "nrf_composite_fwd")
    return w, rgb, emb, dep


@_on_tensor_device
def composite_bwd(field_out, z, rays, D, d_rgb, d_embed, d_depth=None, d_weights=None, ldg=None,
                  precision=NRF_PREC_BF16, white_bkgd=False, want_dz=False, out=None, sigma_noise=None,
                  reuse=None, out_new=None, accumulate=False):
    """Closed-form backward; returns d_field (N, ldg) (operand-typed) and optionally d_z (R,K).
    reuse = (field_new, perm, n_first): the reused samples' gradient rows are written to `out` (R*n_first, ldg), the
    new samples' rows to `out_new` (R*(K-n_first), ldg); both returned.  accumulate: add to what `out` holds."""
    field_out = _f32(field_out, "field_out")
    z = _f32(z, "z")
    rays = _f32(rays, "rays")
    d_rgb = _f32(d_rgb, "d_rgb")
    d_embed = _f32(d_embed, "d_embed")
    d_depth = _f32(d_depth, "d_depth") if d_depth is not None else None
    d_weights = _f32(d_weights, "d_weights") if d_weights is not None else None
    sigma_noise = _f32(sigma_noise, "sigma_noise") if sigma_noise is not None else None
    R, K = z.shape
    if ldg is None:
        ldg = (4 + D + 63) // 64 * 64
    ru = None
    if reuse is not None:
        n_first = reuse[2]
        if out is None:
            out = torch.empty(R * n_first, ldg, device=z.device, dtype=grad_dtype(precision))
        assert out.shape == (R * n_first, ldg) and out.dtype == grad_dtype(precision) and not accumulate
        if out_new is None:
            out_new = torch.empty(R * (K - n_first), ldg, device=z.device, dtype=grad_dtype(precision))
        ru = _reuse_struct(reuse, K, field_out.shape[1], out_new)
    elif out is None:
        out = torch.empty(R * K, ldg, device=z.device, dtype=grad_dtype(precision))
    dz = torch.empty(R, K, device=z.device, dtype=torch.float32) if want_dz else None
    check(_lib.load().nrf_composite_bwd(ptr(field_out), field_out.shape[1], ptr(z), ptr(rays), R, K, D,
                                        int(white_bkgd), ptr(d_rgb), ptr(d_embed), ptr(d_depth),
                                        ptr(d_weights), ptr(out), ldg, int(out.dtype == torch.bfloat16),
                                        ptr(dz), ptr(sigma_noise), C.byref(ru) if ru is not None else None,
                                        int(accumulate), stream_ptr()), "nrf_composite_bwd")
    if reuse is not None:
        return (out, out_new, dz) if want_dz else (out, out_new)
    return (out, dz) if want_dz else out


@_on_tensor_device
def render_loss(rgb_c, rgb_f, emb_c, emb_f, rays_per_scene, gt_rgb, gt_embed, idx=None, want_grads=True):
    """The four F.mse_loss terms of neural_rendering.py:653-677 and their gradients in one pass (nrf_render_loss).

    rgb_* (R,3), emb_* (R,D); gt_rgb (SB,n_pix,3), gt_embed (SB,n_pix,D) with idx int64 (rays_per_scene) pixel per
    ray, or idx=None and per-ray targets (R,3) / (R,D).  -> terms (4,), (d_rgb_c, d_rgb_f, d_emb_c, d_emb_f) or None."""
    rgb_c, rgb_f = _f32(rgb_c, "rgb_c"), _f32(rgb_f, "rgb_f")
    emb_c, emb_f = _f32(emb_c, "emb_c"), _f32(emb_f, "emb_f")
    gt_rgb, gt_embed = _f32(gt_rgb, "gt_rgb"), _f32(gt_embed, "gt_embed")
    R, D = emb_c.shape
    assert rgb_c.shape == (R, 3) and rgb_f.shape == (R, 3) and emb_f.shape == (R, D) and gt_embed.shape[-1] == D
    if idx is not None:
        assert idx.is_cuda and idx.dtype == torch.int64 and idx.numel() == rays_per_scene
        idx = idx.contiguous()
        n_pix = gt_rgb.shape[-2]
        assert gt_rgb.numel() == (R // rays_per_scene) * n_pix * 3 and gt_embed.numel() == (R // rays_per_scene) * n_pix * D
    else:
        n_pix = 0
        assert gt_rgb.numel() == R * 3 and gt_embed.numel() == R * D
    dev = rgb_c.device
    partial = torch.empty(R, 4, device=dev, dtype=torch.float32)
    terms = torch.empty(4, device=dev, dtype=torch.float32)
    grads = None
    if want_grads:
        grads = (torch.empty_like(rgb_c), torch.empty_like(rgb_f), torch.empty_like(emb_c), torch.empty_like(emb_f))
    gp = [ptr(g) for g in grads] if grads is not None else [None] * 4
    check(_lib.load().nrf_render_loss(ptr(rgb_c), ptr(rgb_f), ptr(emb_c), ptr(emb_f), R, D, rays_per_scene,
                                      ptr(gt_rgb), ptr(gt_embed), n_pix, ptr(idx), ptr(partial), ptr(terms), *gp,
                                      stream_ptr()), "nrf_render_loss")
    return terms, grads


# ------------------------------------------------------------------------------------- GEMMs
@_on_tensor_device
def gemm(A1, B, *, A2=None, A3=None, bias=None, mask_src=None, resid=None, out_f32=None, out_act=None,
         relu_act=False, out_act2=None, relu_act2=False, n_store=None, precision=NRF_PREC_BF16):
    """v = resid + mask([A1|A2|A3].B^T + bias) -> out_act / out_act2 / out_f32; test hook over nrf_gemm."""
    g = _lib.NrfGemm()
    for i, A in enumerate((A1, A2, A3)):
        if A is not None:
            g.A[i], g.K[i], g.lda[i] = A.data_ptr(), A.shape[1], A.stride(0)
    g.B, g.ldb = ptr(B), B.stride(0)
    g.M, g.N = A1.shape[0], B.shape[0]
    g.n_store = n_store if n_store is not None else g.N
    g.bias = ptr(bias)
    if mask_src is not None:
        g.mask_src, g.ldmask = ptr(mask_src), mask_src.stride(0)
    if resid is not None:
        g.resid, g.ldr = ptr(resid), resid.stride(0)
    if out_f32 is not None:
        g.out_f32, g.ldo = ptr(out_f32), out_f32.stride(0)
    if out_act is not None:
        g.out_act, g.ldact, g.relu_act = ptr(out_act), out_act.stride(0), int(relu_act)
    if out_act2 is not None:
        g.out_act2, g.ldact2, g.relu_act2 = ptr(out_act2), out_act2.stride(0), int(relu_act2)
    check(_lib.load().nrf_gemm(C.byref(g), precision, stream_ptr()), "nrf_gemm")


@_on_tensor_device
def wgrad(G, A, dW, dbias=None, n_valid=None, k_valid=None, precision=NRF_PREC_BF16, deterministic=False):
    """dW (n_valid,k_valid) += G^T.A ; dbias += colsum(G).  deterministic: ordered reduction of the sample splits
    through a workspace instead of fp32 atomics."""
    M, N = G.shape
    K = A.shape[1]
    lib = _lib.load()
    ws = None
    if deterministic:
        ws = torch.empty(lib.nrf_wgrad_workspace_bytes(N, K), device=G.device, dtype=torch.uint8)
    check(lib.nrf_wgrad(ptr(G), G.stride(0), ptr(A), A.stride(0), M, N, K,
                        n_valid if n_valid is not None else N, k_valid if k_valid is not None else K,
                        ptr(dW), dW.stride(0), ptr(dbias), ptr(ws), precision, stream_ptr()), "nrf_wgrad")


# --------------------------------------------------------------------------------------- MLP
PARAM_ORDER_DOC = "lin_in.{weight,bias}, lin_out.*, blocks.b.fc_0.*, blocks.b.fc_1.*, lin_z.b.*"


class FieldMLP:
    """Host-side handle of the ResnetFC field MLP: parameter pointers, packed operand cache, sizes.

    `params` maps the reference's state_dict suffixes ('lin_in.weight', 'blocks.0.fc_1.bias',
    'lin_z.2.weight', ...) to fp32 CUDA tensors.  The packed (bf16 / transposed) copies are derived
    caches, refreshed by `pack()` whenever a parameter's version counter changes.
    """

    def __init__(self, params: dict, d_in: int, d_latent: int, d_hidden: int, d_out: int, n_blocks: int,
                 n_lin_z: int, precision: int = NRF_PREC_BF16):
        self.params = params
        self.dims = (d_in, d_latent, d_hidden, d_out, n_blocks, n_lin_z)
        self.precision = precision
        self._packed = None
        self._packed_key = None
        self.sizes = _lib.NrfMlpSizes()
        check(_lib.load().nrf_mlp_sizes(C.byref(self._cparams()), precision, C.byref(self.sizes)),
              "nrf_mlp_sizes")

    # names in a fixed order (used for flat gradient buffers too)
    def names(self):
        d_in, d_latent, d_hidden, d_out, nb, nz = self.dims
        out = ["lin_in.weight", "lin_in.bias", "lin_out.weight", "lin_out.bias"]
        for b in range(nb):
            out += [f"blocks.{b}.fc_0.weight", f"blocks.{b}.fc_0.bias",
                    f"blocks.{b}.fc_1.weight", f"blocks.{b}.fc_1.bias"]
        for b in range(nz):
            out += [f"lin_z.{b}.weight", f"lin_z.{b}.bias"]
        return out

    def _fill(self, st, get):
        d_in, d_latent, d_hidden, d_out, nb, nz = self.dims
        st.lin_in_w, st.lin_in_b = get("lin_in.weight"), get("lin_in.bias")
        st.lin_out_w, st.lin_out_b = get("lin_out.weight"), get("lin_out.bias")
        for b in range(nb):
            st.fc0_w[b], st.fc0_b[b] = get(f"blocks.{b}.fc_0.weight"), get(f"blocks.{b}.fc_0.bias")
            st.fc1_w[b], st.fc1_b[b] = get(f"blocks.{b}.fc_1.weight"), get(f"blocks.{b}.fc_1.bias")
        for b in range(nz):
            st.lin_z_w[b], st.lin_z_b[b] = get(f"lin_z.{b}.weight"), get(f"lin_z.{b}.bias")
        return st

    def _cparams(self):
        d_in, d_latent, d_hidden, d_out, nb, nz = self.dims
        st = _lib.NrfMlpParams()
        st.d_in, st.d_latent, st.d_hidden, st.d_out, st.n_blocks, st.n_lin_z = self.dims

        def get(name):
            t = self.params[name]
            if not t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous():
                raise _lib.NrfError(f"MLP parameter {name} must be a contiguous fp32 CUDA tensor")
            return t.data_ptr()
        return self._fill(st, get)

    @_on_tensor_device
    def pack(self, force=False):
        """The packed operand cache.  Refreshed when a parameter's (data_ptr, version) changes, and ALWAYS when
        `force` (every training forward passes it: in-place writes through `p.data` - apex / DeepSpeed fused optimizers,
        EMA swaps, manual clipping - do not bump the Parameter's version counter, and a stale cache would silently
        disagree with the biases read live from the parameters.  The pack is one 12 MB kernel launch)."""
        key = tuple((self.params[n].data_ptr(), self.params[n]._version) for n in self.names())
        if force or self._packed is None or key != self._packed_key:
            dev = self.params["lin_in.weight"].device
            if self._packed is None or self._packed.device != dev:
                self._packed = torch.empty(self.sizes.packed_bytes, device=dev, dtype=torch.uint8)
            check(_lib.load().nrf_mlp_pack(C.byref(self._cparams()), self.precision, ptr(self._packed),
                                           stream_ptr()), "nrf_mlp_pack")
            self._packed_key = key
        return self._packed

    @property
    def fused(self) -> bool:
        """True when nrf_mlp_fwd runs the whole MLP as one persistent tcgen05 kernel (csrc/mlp_fused.cu)."""
        return bool(_lib.load().nrf_mlp_fused_supported(C.byref(self._cparams()), self.precision))

    @_on_tensor_device
    def forward(self, field_in, acts=None, keep_acts=True, layered=False, repack=None, touch=None):
        """field_in (N,kin_pad) -> (field_out (N,d_out) fp32 raw, acts buffer).

        touch: the flags encode_points(want_touch=True) returned for these samples; the fused kernel then skips the
        latent k-panels of 256-sample tiles that lie outside the grid altogether (same results bit for bit).

        keep_acts=False (inference): the fused kernel keeps nothing (acts is None); the layer-by-layer chain
        still needs its buffer.  layered=True forces the chain (A/B timing, parity tests).
        repack: re-derive the packed weights first (default: whenever activations are kept, i.e. in training; the
        renderer packs once per forward_nerf and passes False for its passes)."""
        N = field_in.shape[0]
        dev = field_in.device
        packed = self.pack(force=keep_acts if repack is None else repack)
        lib = _lib.load()
        if acts is None and (keep_acts or layered or not self.fused):
            acts = torch.empty(self.sizes.fwd_bytes_per_sample * N, device=dev, dtype=torch.uint8)
        # rows are d_out rounded up to whole float4s; the pad columns come back as exact zeros
        out = torch.empty(N, (self.dims[3] + 3) // 4 * 4, device=dev, dtype=torch.float32)
        if touch is not None and not layered:
            assert touch.is_cuda and touch.dtype == torch.uint8 and touch.numel() == (N + 31) // 32
            check(lib.nrf_mlp_fwd_touch(C.byref(self._cparams()), ptr(packed), self.precision, ptr(field_in), N,
                                        ptr(acts), ptr(out), ptr(touch), stream_ptr()), "nrf_mlp_fwd_touch")
        else:
            fn = lib.nrf_mlp_fwd_layered if layered else lib.nrf_mlp_fwd
            check(fn(C.byref(self._cparams()), ptr(packed), self.precision, ptr(field_in), N,
                     ptr(acts), ptr(out), stream_ptr()), "nrf_mlp_fwd")
        if acts is not None:
            # the fused backward reads the bit-packed ReLU gates only the fused forward writes
            acts._nrf_layered = bool(layered or not self.fused)
        return out, acts

    @_on_tensor_device
    def last_feat(self, acts, N):
        """x_nb (N, d_hidden), the MLP's second return value (resnetfc.py:192-195), after a layered forward: a view of
        the last layer of `acts`."""
        H, nb = self.dims[2], self.dims[4]
        dt = act_dtype(self.precision)
        es = torch.empty(0, dtype=dt).element_size()
        off = (2 * nb + 1) * N * H * es
        return acts[off:off + N * H * es].view(dt).view(N, H)

    @_on_tensor_device
    def backward(self, field_in, acts, d_field, grads: dict, scratch=None, deterministic=False, layered=False,
                 d_last=None, touch=None, dlatent_event=None):
        """Accumulates parameter grads into `grads` (same keys as params); returns dlatent (N,C).
        deterministic: ordered (bit-reproducible) reduction of the weight-gradient sample splits.
        d_last (N, d_hidden): gradient w.r.t. the last residual stream (layered chain only).
        touch: the uint8 flags of encode_points(want_touch=True) for these samples: dlatent is then only computed for
        128-sample tiles with a sample inside the grid - rows of the other tiles are UNINITIALISED (the volume scatter
        never reads them).
        dlatent_event: a torch.cuda.Event (already recorded once, so that its handle exists) that the call records as
        soon as dlatent is complete - ahead of the weight gradients in the default path (NrfMlpGrads.dlatent_ready_event)."""
        N = field_in.shape[0]
        dev = field_in.device
        d_latent = self.dims[1]
        packed = self.pack()
        need = self.sizes.bwd_bytes_per_sample * N + self.sizes.bwd_fixed_bytes
        if scratch is None or scratch.numel() < need:
            scratch = torch.empty(need, device=dev, dtype=torch.uint8)
        dlatent = torch.empty(N, d_latent, device=dev, dtype=torch.float32)
        g = self._fill(_lib.NrfMlpGrads(), lambda n: grads[n].data_ptr() if grads.get(n) is not None else None)
        g.deterministic = int(bool(deterministic))
        if d_last is not None:
            assert d_last.shape == (N, self.dims[2]) and d_last.dtype == grad_dtype(self.precision) and d_last.is_contiguous()
            g.d_last = d_last.data_ptr()
        if touch is not None:
            assert touch.dtype == torch.uint8 and touch.is_contiguous() and touch.numel() == (N + 31) // 32
            g.touch_flags = touch.data_ptr()
        if dlatent_event is not None:
            handle = int(dlatent_event.cuda_event)
            if not handle:
                raise _lib.NrfError("dlatent_event has no CUDA handle yet: record it once before handing it over")
            g.dlatent_ready_event = handle
        lib = _lib.load()
        layered = layered or getattr(acts, "_nrf_layered", False)
        fn = lib.nrf_mlp_bwd_layered if layered else lib.nrf_mlp_bwd
        check(fn(C.byref(self._cparams()), ptr(packed), self.precision, ptr(field_in), N,
                 ptr(acts), ptr(d_field), C.byref(g), ptr(dlatent), ptr(scratch),
                 stream_ptr()), "nrf_mlp_bwd")
        return dlatent
