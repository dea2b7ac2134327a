"""Drop-in `NeuralRenderer` for the reference's neural_rendering.py:86-711, on the sm_100a C ABI.

Same constructor, methods, config keys, state_dict keys and loss dictionary as the reference class;
every stage of `forward_nerf` (sampling, points + trilinear gather + positional encoding, the ResnetFC
field MLP, alpha compositing, and the whole backward) runs in the hand-written CUDA kernels behind
include/nrf_b200.h.  There is no PyTorch/CPU fallback: CPU tensors raise.

Differences from the reference, all opt-in or invisible to its callers:
  * `precision` ("bf16" tensor-core mode, default; "fp32" parity mode), `perturb` (default True: the
    reference always draws sampling noise, neural_rendering.py:172,194,200,218), `scatter` ("atomic" |
    "sorted": atomics-free, bit-reproducible volume gradient) and `deterministic` (bit-reproducible MLP weight
    gradients too: ordered reduction instead of fp32 atomics) attributes;
  * `forward_nerf(rays, want_weights=False, noise=None)`: `noise` injects pre-drawn tensors
    (keys coarse / u / fine / depth), used by the parity tests;
  * the optional branches of nerfact.conf: coord / attention heads run on the default (fused) path; multi-scale voxels,
    ret_last_feat and use_code_viewdirs run composed.py (same kernels, three autograd nodes per pass, fp32 MLP);
    softplus activations (mlp.beta > 0) and SPADE modulation (mlp.use_spade) run there too, layer by layer on the fp32
    GEMM kernels (composed.mlp_general); normalize_z is the no-op it is in the reference; the two branches the reference
    itself cannot run (the depth-supervision volume, use_freenerf) raise NotImplementedError.
"""
from __future__ import annotations


import threading
import weakref

import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .utils import AttrDict, PositionalEncoding, gen_rays, pca_fit_transform

__all__ = ["NeuralRenderer", "PixelNeRFEmbedNet", "ResnetFC", "ResnetBlockFC", "PSNR_torch", "LossDict"]


class _trace:
    """The reference's profiler labels (torch.autograd.profiler.record_function: neural_rendering.py:236
    "renderer_composite", models_embed.py:306 "model_inference", resnetfc.py:153 "resnetfc_infer", resnetfc.py:56
    "resblock", utils.py:551 "positional_enc") around the stages that replace that code, as record_function ranges AND
    NVTX ranges (nsys / ncu --nvtx).  The five residual blocks run inside one fused kernel, so "resblock" brackets the
    same launch as "resnetfc_infer"; "positional_enc" brackets the fused gather + encoding kernel.
    `NeuralRenderer.trace_ranges = False` turns them off (each costs ~2 us of host time)."""
    enabled = True
    __slots__ = ("names", "rf")

    def __init__(self, *names):
        self.names = names

    def __enter__(self):
        if not _trace.enabled:
            return self
        self.rf = [torch.autograd.profiler.record_function(n) for n in self.names]
        for n, r in zip(self.names, self.rf):
            r.__enter__()
            torch.cuda.nvtx.range_push(n)
        return self

    def __exit__(self, *exc):
        if not _trace.enabled:
            return False
        for r in reversed(self.rf):
            torch.cuda.nvtx.range_pop()
            r.__exit__(*exc)
        return False


def _cfg_get(cfg, key, default=None):
    try:
        return cfg[key]
    except (KeyError, TypeError, AttributeError):
        return getattr(cfg, key, default)


def PSNR_torch(img1, img2, max_val=1):
    """neural_rendering.py:78-83."""
    mse = torch.mean((img1 - img2) ** 2)
    if mse == 0:
        return 100
    return 20 * torch.log10(max_val / torch.sqrt(mse))


class LossDict(dict):
    """What compute_rendering_loss returns (neural_rendering.py:697-707): 'loss' is the tensor to back-propagate, the
    other ten entries are Python floats.  The reference produces them with 13 `.item()` calls - host syncs between the
    forward and the backward of every step.  Here the seven scalars behind them are copied to pinned host memory
    asynchronously and become floats the first time any of them is read (or the dict is iterated / printed / copied /
    merged), so a loop that calls `out['loss'].backward()` before it logs never stalls the GPU.

    It IS a dict (isinstance, json.dumps, pickling work).  CPython's C fast paths for exact-dict layouts (`dict(d)`,
    `{**d}`, `other.update(d)`) are taken only while `__iter__` is dict's own, so overriding `__iter__` / `keys` routes
    them through `keys()` + `__getitem__`, which resolve first; `|`, `pop`, `setdefault`, ... are overridden too."""

    _KEYS = ("loss_rgb_coarse", "loss_rgb_fine", "loss_rgb", "loss_embed_coarse", "loss_embed_fine", "loss_embed",
             "loss_depth_coarse", "loss_depth_fine", "loss_depth", "psnr")
    _rings = {}                                # device index -> [pinned staging buffers, weak owners, next slot]
    _lock = threading.Lock()

    def __init__(self, loss, scalars):
        super().__init__(loss=loss)
        for k in self._KEYS:
            dict.__setitem__(self, k, None)
        cls = LossDict
        with cls._lock:
            ring = cls._rings.get(scalars.device.index)
            if ring is None:
                ring = cls._rings[scalars.device.index] = [
                    [torch.empty(7, dtype=torch.float32).pin_memory() for _ in range(8)], [None] * 8, 0]
            i = ring[2]
            ring[2] = (i + 1) % len(ring[0])
            prev = ring[1][i]() if ring[1][i] is not None else None
            ring[1][i] = weakref.ref(self)
        if prev is not None:
            prev._resolve()                    # its values leave the staging buffer before it is reused
        self._host = ring[0][i]
        self._host.copy_(scalars, non_blocking=True)
        self._event = torch.cuda.Event()
        self._event.record(torch.cuda.current_stream(scalars.device))

    def _resolve(self):
        if getattr(self, "_host", None) is None:
            return
        self._event.synchronize()
        v = self._host.tolist()
        self._host = None
        for k, x in zip(self._KEYS, (v[0], v[1], v[0] + v[1], v[2], v[3], v[2] + v[3], v[4], v[5], v[4] + v[5], v[6])):
            dict.__setitem__(self, k, x)

    def __getitem__(self, k):
        if k != "loss":
            self._resolve()
        return dict.__getitem__(self, k)

    def get(self, k, default=None):
        if k != "loss":
            self._resolve()
        return dict.get(self, k, default)

    def __iter__(self):
        return dict.__iter__(self)

    def keys(self):
        return dict.keys(self)

    def items(self):
        self._resolve()
        return dict.items(self)

    def values(self):
        self._resolve()
        return dict.values(self)

    def copy(self):
        self._resolve()
        return dict(self)

    def pop(self, k, *default):
        self._resolve()
        return dict.pop(self, k, *default)

    def popitem(self):
        self._resolve()
        return dict.popitem(self)

    def setdefault(self, k, default=None):
        self._resolve()
        return dict.setdefault(self, k, default)

    def __or__(self, other):
        return dict(self) | dict(other)

    def __ror__(self, other):
        return dict(other) | dict(self)

    def __repr__(self):
        self._resolve()
        return dict.__repr__(self)

    def __eq__(self, other):
        self._resolve()
        return dict.__eq__(self, other)

    __hash__ = None

    def __reduce__(self):
        self._resolve()
        return (dict, (dict(self),))


# ------------------------------------------------------------------ parameter containers
class ResnetBlockFC(nn.Module):
    """Parameters of resnetfc.py:12-64 (size_in == size_out == size_h; ReLU)."""

    def __init__(self, size):
        super().__init__()
        self.fc_0 = nn.Linear(size, size)
        self.fc_1 = nn.Linear(size, size)
        nn.init.constant_(self.fc_0.bias, 0.0)
        nn.init.kaiming_normal_(self.fc_0.weight, a=0, mode="fan_in")
        nn.init.constant_(self.fc_1.bias, 0.0)
        nn.init.zeros_(self.fc_1.weight)


class ResnetFC(nn.Module):
    """Parameters of resnetfc.py:67-209 with the reference's initialisation and state_dict keys.

    The arithmetic lives in the fused GEMM chain of csrc/mlp.cu; `forward(zx)` evaluates it
    (differentiable w.r.t. the latent part of zx and all parameters).
    """

    def __init__(self, d_in, d_out=4, n_blocks=5, d_latent=0, d_lang=0, d_hidden=128, beta=0.0,
                 combine_layer=1000, combine_type="average", use_spade=False, use_language=False):
        super().__init__()
        # resnetfc.py:43-46,138-141 softplus activations (beta > 0) and :130-136,184-186 SPADE modulation (x = sz * x + tz)
        # are off in nerfact.conf; they run layer by layer on the fp32 GEMM kernels (composed.mlp_general)
        self.beta = float(beta)
        self.general = self.beta > 0 or bool(use_spade)
        self.lin_in = nn.Linear(d_in, d_hidden)
        nn.init.constant_(self.lin_in.bias, 0.0)
        nn.init.kaiming_normal_(self.lin_in.weight, a=0, mode="fan_in")
        self.lin_out = nn.Linear(d_hidden, d_out)
        nn.init.constant_(self.lin_out.bias, 0.0)
        nn.init.kaiming_normal_(self.lin_out.weight, a=0, mode="fan_in")
        self.n_blocks, self.d_latent, self.d_lang = n_blocks, d_latent, d_lang
        self.d_in, self.d_out, self.d_hidden = d_in, d_out, d_hidden
        self.combine_layer, self.combine_type, self.use_spade = combine_layer, combine_type, use_spade
        self.use_language = False            # resnetfc.py:115 hard-wires this off
        self.blocks = nn.ModuleList([ResnetBlockFC(d_hidden) for _ in range(n_blocks)])
        if d_latent != 0:
            n_lin_z = min(combine_layer, n_blocks)
            self.lin_z = nn.ModuleList([nn.Linear(d_latent, d_hidden) for _ in range(n_lin_z)])
            for i in range(n_lin_z):
                nn.init.constant_(self.lin_z[i].bias, 0.0)
                nn.init.kaiming_normal_(self.lin_z[i].weight, a=0, mode="fan_in")
            if self.use_spade:                                        # resnetfc.py:130-136
                self.scale_z = nn.ModuleList([nn.Linear(d_latent, d_hidden) for _ in range(n_lin_z)])
                for i in range(n_lin_z):
                    nn.init.constant_(self.scale_z[i].bias, 0.0)
                    nn.init.kaiming_normal_(self.scale_z[i].weight, a=0, mode="fan_in")
        self._handles = {}

    @property
    def n_lin_z(self):
        return len(self.lin_z) if self.d_latent != 0 else 0

    def param_dict(self):
        """name -> Parameter, cached: walking the module tree costs ~0.1 ms and is needed several times per step
        (the step at the reference's own training shape is host-bound).  The cache remembers where every parameter
        hangs and is re-checked by identity (30 dictionary look-ups), so replacing a parameter object rebuilds it; it
        is also dropped whenever the module is converted (`_apply`: .to() / .cuda() / .float())."""
        c = self.__dict__.get("_param_dict_cache")
        if c is not None and all(m._parameters.get(n) is p for m, n, p in c[1]):
            return c[0]
        d, where = {}, []
        for mod_name, mod in self.named_modules():
            for n, p in mod._parameters.items():
                if p is not None:
                    d[(mod_name + "." if mod_name else "") + n] = p
                    where.append((mod, n, p))
        self.__dict__["_param_dict_cache"] = (d, where)
        return d

    def _apply(self, fn, *args, **kwargs):
        self.__dict__.pop("_param_dict_cache", None)
        return super()._apply(fn, *args, **kwargs)

    def handle(self, precision) -> ops.FieldMLP:
        """C-ABI handle for the given precision; rebuilt if parameters moved (e.g. after .to())."""
        params = self.param_dict()
        key = (precision, tuple(p.data_ptr() for p in params.values()))
        h = self._handles.get(precision)
        if h is None or h[0] != key:
            h = (key, ops.FieldMLP(params, self.d_in, self.d_latent, self.d_hidden, self.d_out,
                                   self.n_blocks, self.n_lin_z, precision))
            self._handles[precision] = h
        return h[1]

    def forward(self, zx, precision="bf16", **_ignored):
        """zx (..., d_latent + d_in) -> (out (..., d_out), None).  See `_MlpFn`."""
        lead = zx.shape[:-1]
        flat = zx.reshape(-1, zx.shape[-1])
        if self.general:
            from . import composed
            return composed.mlp_general(self, flat)[0].reshape(*lead, self.d_out), None
        prec = ops.PRECISIONS[precision] if isinstance(precision, str) else precision
        h = self.handle(prec)
        names = h.names()
        out = _MlpFn.apply(h, flat, *[self.param_dict()[n] for n in names])
        return out.reshape(*lead, self.d_out), None


class _MlpFn(torch.autograd.Function):
    """The field MLP alone: rows [latent | pe | dir] (fp32) -> raw outputs.  Test / API hook."""

    @staticmethod
    def forward(ctx, h: ops.FieldMLP, zx, *params):
        N, width = zx.shape
        kin = h.sizes.kin_pad
        fin = torch.zeros(N, kin, device=zx.device, dtype=ops.act_dtype(h.precision))
        fin[:, :width] = zx.detach().to(fin.dtype)
        out, acts = h.forward(fin)
        ctx.h, ctx.fin, ctx.acts, ctx.width = h, fin, acts, width
        return out

    @staticmethod
    def backward(ctx, d_out):
        h = ctx.h
        if ctx.acts is None:
            raise RuntimeError("ResnetFC.forward: backward a second time (retain_graph is not supported)")
        fin, acts = ctx.fin, ctx.acts
        ctx.acts = ctx.fin = None
        N = fin.shape[0]
        dpad = h.sizes.dout_pad
        dfield = torch.zeros(N, dpad, device=d_out.device, dtype=ops.grad_dtype(h.precision))
        dfield[:, :d_out.shape[1]] = d_out.to(dfield.dtype)
        names = h.names()
        grads = _zero_grads(h)
        dlat = h.backward(fin, acts, dfield, grads)
        dzx = torch.zeros(N, ctx.width, device=d_out.device, dtype=torch.float32)
        dzx[:, :dlat.shape[1]] = dlat
        return (None, dzx, *[grads[n] for n in names])


class PixelNeRFEmbedNet(nn.Module):
    """Parameter / config container mirroring models_embed.py:16-134 (default branch only)."""

    def __init__(self, conf, coordinate_bounds, stop_encoder_grad=False):
        super().__init__()
        self.conf = conf
        self.coordinate_bounds = coordinate_bounds
        g = lambda k, d=None: _cfg_get(conf, k, d)
        unsupported = dict(use_depth_supervision=g("use_depth_supervision", False),   # cannot run in the reference either
                                                                                      # (models_embed.py:151-154 never
                                                                                      # stores voxel_density, :289 samples None)
                           use_freenerf=g("use_freenerf", False))    # models_embed.py:350-352 is a debugger breakpoint
                                                                     # and an invalid statement: it cannot run there either
        for k, v in unsupported.items():
            if v:
                raise NotImplementedError(f"config option {k}=True is off in nerfact.conf and not built "
                                          "(SURVEY.md section 8f, 'next')")
        if not g("use_viewdirs", True):
            raise NotImplementedError("use_viewdirs=False (neural_rendering.py:294-295 raises too)")
        if not g("use_xyz", True) or not g("use_code", True):
            raise NotImplementedError("use_xyz / use_code must be True (nerfact.conf:69-71)")
        self._voxel_shape = g("voxel_shape")
        self.image_shape = (g("image_height"), g("image_width"))
        # models_embed.py:42 hard-codes canon_xyz = True, so xyz_rot IS xyz (:325-326) and normalize_z only picks
        # between two names of the same tensor (:337-340): the flag is accepted and changes nothing, as in the reference
        self.normalize_z = bool(g("normalize_z", False))
        self.canon_xyz = True
        self.stop_encoder_grad = stop_encoder_grad
        self.use_code, self.use_viewdirs, self.use_xyz = True, True, True
        # models_embed.py:86-95,:370-372: the view direction goes THROUGH the positional encoding (6 inputs, d_in = 78)
        # instead of behind it (d_in = 42); runs on the composed fp32 branch (composed.py)
        self.use_code_viewdirs = bool(g("use_code_viewdirs", False))
        self.use_freenerf = False
        self.regress_coord = bool(g("regress_coord", False))            # models_embed.py:63-68: 3 / 6 more outputs
        self.regress_attention = bool(g("regress_attention", False))
        self.use_depth_supervision = False
        self.use_multi_scale_voxel = bool(g("use_multi_scale_voxel", False))        # models_embed.py:69-72
        self.d_latent = d_latent = g("d_multi_scale_latent") if self.use_multi_scale_voxel else g("d_latent")
        self.d_lang = g("d_lang", 0)
        self.code = PositionalEncoding.from_conf(conf["code"], d_in=6 if self.use_code_viewdirs else 3)
        d_in = self.code.d_out + (0 if self.use_code_viewdirs else 3)
        d_out = 4 + conf["d_embed"] + (3 if self.regress_coord else 0) + (6 if self.regress_attention else 0)
        self.share_mlp = g("share_mlp", True)
        mlp = conf["mlp"] if not hasattr(conf, "mlp") else conf.mlp
        mk = lambda: ResnetFC(d_in=d_in, d_latent=d_latent, d_lang=self.d_lang, d_out=d_out,
                              d_hidden=_cfg_get(mlp, "d_hidden"), n_blocks=_cfg_get(mlp, "n_blocks"),
                              combine_layer=_cfg_get(mlp, "combine_layer"), beta=_cfg_get(mlp, "beta", 0.0),
                              use_spade=_cfg_get(mlp, "use_spade", False),
                              use_language=_cfg_get(mlp, "use_language", False))
        self.mlp_coarse = mk()
        self.mlp_fine = self.mlp_coarse if self.share_mlp else mk()
        self.register_buffer("poses", torch.empty(1, 3, 4), persistent=False)
        self.register_buffer("focal", torch.empty(1, 2), persistent=False)
        self.register_buffer("c", torch.empty(1, 2), persistent=False)
        self.d_in, self.d_out, self.d_embed = d_in, d_out, conf["d_embed"]
        self.num_objs, self.num_views_per_obj = 0, 1
        self.voxel_feat = None

    def encode(self, voxel_feat, lang, multi_scale_voxel_list, voxel_density, poses, focal, c=None):
        """models_embed.py:136-183: stores a reference to the volume; focal / c bookkeeping only.
        A bf16 / fp16 volume (a voxel encoder run under torch.autocast) is widened to fp32 here: what autocast itself
        does to the reference's F.grid_sample (models_embed.py:275; grid_sampler is on autocast's fp32 list) - the
        gather then interpolates in fp32 and autograd returns the volume gradient in the producer's dtype."""
        low = (torch.bfloat16, torch.float16)
        if torch.is_tensor(voxel_feat) and voxel_feat.dtype in low:
            voxel_feat = voxel_feat.float()       # keeps the memory format (channels_last_3d stays zero-copy below)
        self.voxel_feat = voxel_feat
        if self.use_multi_scale_voxel and multi_scale_voxel_list is not None:
            multi_scale_voxel_list = [v.float() if torch.is_tensor(v) and v.dtype in low else v
                                      for v in multi_scale_voxel_list]
        self.multi_scale_voxel_list = multi_scale_voxel_list if self.use_multi_scale_voxel else None   # :147-149
        self.voxel_density = None
        self.language = lang
        if focal is not None:
            focal = torch.as_tensor(focal)
            if focal.dim() == 0:
                focal = focal[None, None].repeat((1, 2))
            elif focal.dim() == 1:
                focal = focal.unsqueeze(-1).repeat((1, 2))
            else:
                focal = focal.clone()
            self.focal = focal.float()
            self.focal[..., 1] *= -1.0
        if c is not None:
            c = torch.as_tensor(c)
            if c.dim() == 0:
                c = c[None, None].repeat((1, 2))
            elif c.dim() == 1:
                c = c.unsqueeze(-1).repeat((1, 2))
        self.c = c

    def forward(self, xyz, coarse=True, viewdirs=None, far=False, ret_last_feat=False, precision="bf16"):
        """models_embed.py:295-471: the field at world points.  xyz, viewdirs (SB, B, 3) ->
        (output (SB, B, 4 + d_embed) = [sigmoid(rgb), relu(sigma), embed], point_density = None).  Call encode() first.
        Differentiable w.r.t. the encoded volume and the MLP parameters."""
        if viewdirs is None:
            raise NotImplementedError("use_viewdirs=False (neural_rendering.py:294-295 raises too)")
        if self.voxel_feat is None:
            raise RuntimeError("call encode() before forward()")
        SB, B, _ = xyz.shape
        if self.voxel_feat.shape[0] != SB:
            raise RuntimeError("grid_sampler(): expected grid and input to have same batch size, "
                               f"but got input with sizes {list(self.voxel_feat.shape)} and {SB} point batches")
        prec = ops.PRECISIONS[precision] if isinstance(precision, str) else precision
        mlp = self.mlp_coarse if coarse or self.mlp_fine is None else self.mlp_fine
        rays = torch.zeros(SB * B, 8, device=xyz.device, dtype=torch.float32)
        rays[:, 0:3] = xyz.reshape(-1, 3)
        rays[:, 3:6] = viewdirs.reshape(-1, 3)
        bounds = torch.as_tensor(self.coordinate_bounds, dtype=torch.float32).reshape(-1).cpu()
        last_feat = None
        if ret_last_feat or self.use_multi_scale_voxel or self.use_code_viewdirs or mlp.general:   # composed branch (composed.py): fp32 MLP, layer by layer
            from . import composed
            z0 = torch.zeros(SB * B, 1, device=xyz.device, dtype=torch.float32)
            vols = list(self.multi_scale_voxel_list or []) + [self.voxel_feat]
            shim = AttrDict(_bounds=bounds)
            raw, last_feat = composed.field_rows(shim, self, mlp, vols, rays, z0, B)
        else:
            h = mlp.handle(prec)
            ps = [mlp.param_dict()[n] for n in h.names()]
            keep = torch.is_grad_enabled() and (self.voxel_feat.requires_grad or any(p.requires_grad for p in ps))
            raw = _FieldFn.apply(self, h, bounds, self.voxel_feat, rays, SB, keep, *ps)
        parts = [torch.sigmoid(raw[:, :3]), torch.relu(raw[:, 3:4]), raw[:, 4:4 + self.d_embed]]   # :444-466
        o = 4 + self.d_embed
        if self.regress_coord:                                  # coord - canon_xyz (:449,:455); canon is @no_grad (:185)
            b6 = bounds.to(xyz.device)
            canon = (xyz.reshape(-1, 3).detach() - b6[:3]) / (b6[3:] - b6[:3])
            parts.append(raw[:, o:o + 3] - canon)
            o += 3
        if self.regress_attention:
            parts.append(raw[:, o:o + 6])
        out = torch.cat(parts, -1)
        if ret_last_feat:                                        # models_embed.py:468-471
            return out.reshape(SB, B, -1), last_feat.reshape(SB, B, -1), None
        return out.reshape(SB, B, -1), None


class _FieldFn(torch.autograd.Function):
    """The field at explicit points (PixelNeRFEmbedNet.forward, models_embed.py:295-471): points are handed to the
    encode kernel as zero-length rays (origin = point, direction = view direction, z = 0: o + 0 * d is exact), so
    gather, positional encoding, MLP and the volume-gradient scatter are the kernels of the render path.
    -> raw MLP outputs (SB*n, 4+D); no gradient w.r.t. the points (world_to_canonical is @no_grad, :185)."""

    @staticmethod
    def forward(ctx, model, h, bounds, voxel_feat, rays, sb, keep, *params):
        cl3d = _is_channels_last_3d(voxel_feat)
        vol_cl = voxel_feat.permute(0, 2, 3, 4, 1) if cl3d else ops.volume_to_channels_last(voxel_feat)
        n = rays.shape[0]
        z = torch.zeros(n, 1, device=rays.device, dtype=torch.float32)
        field_in = ops.encode_points(rays, z, n // sb, vol_cl, bounds, model.code.num_freqs,
                                     float(model.code.freq_factor), ld_out=h.sizes.kin_pad, precision=h.precision)
        out, acts = h.forward(field_in, keep_acts=keep)
        if keep:
            ctx.h, ctx.field_in, ctx.acts, ctx.rays, ctx.z = h, field_in, acts, rays, z
            ctx.vol_shape, ctx.cl3d, ctx.sb, ctx.bounds = tuple(vol_cl.shape), cl3d, sb, bounds
        return out[:, :model.d_out].contiguous() if out.shape[1] != model.d_out else out

    @staticmethod
    def backward(ctx, d_out):
        h = ctx.h
        if ctx.acts is None:
            raise RuntimeError("nerf_model.forward: backward a second time (retain_graph is not supported)")
        acts, field_in = ctx.acts, ctx.field_in
        ctx.acts = ctx.field_in = None
        n = ctx.rays.shape[0]
        d_field = torch.zeros(n, h.sizes.dout_pad, device=d_out.device, dtype=ops.grad_dtype(h.precision))
        d_field[:, :d_out.shape[1]] = d_out.to(d_field.dtype)
        names = h.names()
        grads = _zero_grads(h)
        dlat = h.backward(field_in, acts, d_field, grads)
        d_vol = None
        if ctx.needs_input_grad[3]:
            SB, S0, S1, S2, C = ctx.vol_shape
            if C in (64, 128):
                g = torch.empty(ctx.vol_shape if ctx.cl3d else (SB, C, S0, S1, S2), device=d_out.device,
                                dtype=torch.float32)
                ops.scatter_volume_grad_merged(ctx.rays, n // ctx.sb, [(ctx.z, dlat)], g, not ctx.cl3d, ctx.bounds)
                d_vol = g.permute(0, 4, 1, 2, 3) if ctx.cl3d else g
            else:
                g = torch.empty(ctx.vol_shape, device=d_out.device, dtype=torch.float32)
                ops.scatter_volume_grad_sorted(ctx.rays, ctx.z, n // ctx.sb, dlat, g, ctx.bounds)
                d_vol = g.permute(0, 4, 1, 2, 3) if ctx.cl3d else ops.volume_to_channels_first(g)
        return (None, None, None, d_vol, None, None, None, *[grads[k] for k in names])


# -------------------------------------------------------------------------- render passes
class _PassState:
    __slots__ = ("rays", "z", "field_in", "acts", "field_out", "rps", "mlp", "perm", "sig_noise", "z_sorted", "base",
                 "touch")


def _sigma_noise(ren, noise, key, R, K, device):
    """neural_rendering.py:336-337: sigmas + randn_like(sigmas) * noise_std, in training only.  `noise[key]` injects
    the standard-normal draw (parity tests); returns the scaled tensor the compositing kernels take, or None."""
    if not (ren.training and ren.noise_std and ren.noise_std > 0.0):
        return None
    nz = noise.get(key)
    if nz is None:
        nz = torch.randn(R, K, device=device)
    return (nz.to(torch.float32) * float(ren.noise_std)).contiguous()


def _pass_forward(ren, mlp: ops.FieldMLP, vol_cl, rays, z, rps, keep_acts=True, sig_noise=None, repack=None):
    """One composite pass (neural_rendering.py:224-395) over all samples of `z`.  keep_acts=False (no gradient
    will be asked for): the fused MLP kernel keeps nothing but the raw field outputs."""
    st = _PassState()
    st.rays, st.z, st.rps, st.mlp, st.perm, st.sig_noise = rays, z, rps, mlp, None, sig_noise
    st.z_sorted = st.base = None
    with _trace("renderer_composite"):
        with _trace("model_inference"):
            with _trace("positional_enc"):
                # training: one flag per 32 samples "some corner inside the grid" - the backward computes dL/dlatent only
                # for sample tiles that have one (three rays in four miss the box in the BASELINE camera set-up)
                st.field_in = ops.encode_points(rays, z, rps, vol_cl, ren._bounds, ren._num_freqs, ren._freq_factor,
                                                ld_out=mlp.sizes.kin_pad, precision=mlp.precision,
                                                want_touch=ren.skip_empty_latent_tiles)
                st.touch = None
                if ren.skip_empty_latent_tiles:
                    st.field_in, st.touch = st.field_in
            log = getattr(ren, "_touch_log", None)             # bench.py: which sample tiles were skipped (for its FLOP count)
            if log is not None and st.touch is not None:
                log.append(st.touch)
            with _trace("resnetfc_infer", "resblock"):
                # ... and the forward skips the latent k-panels of 256-sample tiles that lie outside the grid altogether
                st.field_out, st.acts = mlp.forward(st.field_in, keep_acts=keep_acts, repack=repack, touch=st.touch)
        outs = ops.composite_fwd(st.field_out, z, rays, ren._d_comp, ren.white_bkgd, sigma_noise=sig_noise)
    return st, outs


def _pass_forward_reuse(ren, mlp, vol_cl, rays, z_new, z_sorted, perm, base, rps, keep_acts=True, sig_noise=None):
    """The fine pass without re-evaluating the coarse samples (`reuse_coarse_evals`): the field is evaluated at the
    Kf NEW samples only; the compositing kernel reads the Kc coarse samples' outputs from the coarse pass's buffer
    through the sort permutation.  Same point, same view direction, same MLP (share_mlp) -> the rendered outputs are
    bit-identical to evaluating all Kc + Kf samples again (neural_rendering.py:463-468)."""
    st = _PassState()
    st.rays, st.z, st.rps, st.mlp, st.perm, st.sig_noise = rays, z_new, rps, mlp, perm, sig_noise
    st.z_sorted, st.base = z_sorted, base
    with _trace("renderer_composite"):
        with _trace("model_inference"):
            with _trace("positional_enc"):
                st.field_in = ops.encode_points(rays, z_new, rps, vol_cl, ren._bounds, ren._num_freqs, ren._freq_factor,
                                                ld_out=mlp.sizes.kin_pad, precision=mlp.precision,
                                                want_touch=ren.skip_empty_latent_tiles)
                st.touch = None
                if ren.skip_empty_latent_tiles:
                    st.field_in, st.touch = st.field_in
            with _trace("resnetfc_infer", "resblock"):
                st.field_out, st.acts = mlp.forward(st.field_in, keep_acts=keep_acts, repack=False, touch=st.touch)
        outs = ops.composite_fwd(base.field_out, z_sorted, rays, ren._d_comp, ren.white_bkgd, sigma_noise=sig_noise,
                                 reuse=(st.field_out, perm, base.z.shape[1]))
    return st, outs


def _backward_reuse(ren, st_c, st_f, gc, gf, grads, defer, depth_mask, grad_cl):
    """Backward of a coarse pass + a fine pass that reused its evaluations: the fine compositing backward writes the
    gradient rows of all Kc + Kf samples (coarse ones into the coarse buffer), the coarse compositing backward adds
    its own, and every sample goes through the MLP backward exactly once."""
    dev = st_c.rays.device
    R, D = st_c.rays.shape[0], ren._d_comp
    Kc, Kfd = st_c.z.shape[1], ren.n_fine_depth
    d_cw, d_crgb, d_cemb, d_cdep = gc
    d_fw, d_frgb, d_femb, d_fdep = gf
    mlp = st_c.mlp
    res = ops.composite_bwd(st_c.field_out, st_f.z_sorted, st_f.rays, D, _zeros_like_or(d_frgb, (R, 3), dev),
                            _zeros_like_or(d_femb, (R, D), dev), d_fdep, d_fw, ldg=mlp.sizes.dout_pad,
                            precision=mlp.precision, white_bkgd=ren.white_bkgd, want_dz=Kfd > 0,
                            sigma_noise=st_f.sig_noise, reuse=(st_f.field_out, st_f.perm, Kc))
    d_field_c, d_field_n = res[0], res[1]
    d_cdep = _zeros_like_or(d_cdep, (R,), dev)
    if Kfd > 0:
        K = st_f.z_sorted.shape[1]
        d_cat = torch.zeros(R, K, device=dev, dtype=torch.float32)
        d_cat.scatter_(1, st_f.perm.long(), res[2])
        d_cdep = d_cdep + (d_cat[:, K - Kfd:] * depth_mask).sum(-1)
    ops.composite_bwd(st_c.field_out, st_c.z, st_c.rays, D, _zeros_like_or(d_crgb, (R, 3), dev),
                      _zeros_like_or(d_cemb, (R, D), dev), d_cdep, d_cw, ldg=mlp.sizes.dout_pad,
                      precision=mlp.precision, white_bkgd=ren.white_bkgd, sigma_noise=st_c.sig_noise,
                      out=d_field_c, accumulate=True)
    for i, (st, d_field) in enumerate(((st_f, d_field_n), (st_c, d_field_c))):
        dlat = mlp.backward(st.field_in, st.acts, d_field, grads, deterministic=ren.deterministic, touch=st.touch)
        if defer is not None:
            defer.append((st.z, dlat))
        elif ren.scatter == "sorted":
            ops.scatter_volume_grad_sorted(st.rays, st.z, st.rps, dlat, grad_cl, ren._bounds, accumulate=i > 0)
        else:
            ops.scatter_volume_grad(st.rays, st.z, st.rps, dlat, grad_cl, ren._bounds)


def _coord_raw(ren, st):
    """regress_coord: the raw coordinate outputs of every sample (R, K, 3): they are averaged over the samples, not
    alpha-composited (neural_rendering.py:356-357), which forward_nerf does with plain torch ops."""
    R, K = st.z.shape
    o = 4 + ren._d_embed
    return st.field_out.view(R, K, -1)[:, :, o:o + 3].clone()


def _pass_backward(ren, st, d_rgb, d_embed, d_depth, d_weights, grads, grad_cl, want_dz=False, first=True,
                   defer=None, d_coord_raw=None, dlatent_event=None):
    res = ops.composite_bwd(st.field_out, st.z, st.rays, ren._d_comp, d_rgb, d_embed, d_depth, d_weights,
                            ldg=st.mlp.sizes.dout_pad, precision=st.mlp.precision,
                            white_bkgd=ren.white_bkgd, want_dz=want_dz, sigma_noise=st.sig_noise)
    d_field, d_z = res if want_dz else (res, None)
    if d_coord_raw is not None:           # gradient of the (uncomposited) coordinate outputs, straight into d_field
        R, K = st.z.shape
        o = 4 + ren._d_embed
        d_field.view(R, K, -1)[:, :, o:o + 3] += d_coord_raw.to(d_field.dtype)
    dlat = st.mlp.backward(st.field_in, st.acts, d_field, grads, deterministic=ren.deterministic, touch=st.touch,
                           dlatent_event=dlatent_event)
    if defer is not None:             # one merged scatter for all passes once the last one is through
        defer.append((st.z, dlat))
    elif ren.scatter == "sorted":     # atomics-free, bit-reproducible; the first pass writes every voxel row
        ops.scatter_volume_grad_sorted(st.rays, st.z, st.rps, dlat, grad_cl, ren._bounds, accumulate=not first)
    else:                             # fp32 vector reductions into a zeroed volume
        ops.scatter_volume_grad(st.rays, st.z, st.rps, dlat, grad_cl, ren._bounds)
    return d_z


_SCATTER_SIDE = {}      # device index -> (stream, event); per process, not per renderer: a backward records and waits on
                        # the event within one host call, and keeping CUDA handles out of the module keeps it picklable


def _scatter_side(ren, dev):
    """(stream, event) of `dev` for the overlapped volume scatter, made on first use; the event is recorded once so
    that its CUDA handle exists before the C ABI is given it."""
    hit = _SCATTER_SIDE.get(dev.index)
    if hit is None:
        with torch.cuda.device(dev):
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(dev))
            hit = _SCATTER_SIDE[dev.index] = (torch.cuda.Stream(device=dev), ev)
    return hit


def _merged_scatter_ok(ren, vol_shape_cl):
    """The merged scatter (all passes in one counting sort, gradient written once in the caller's layout) is
    built for 64 / 128 latent channels; other widths take the per-pass kernels + re-layout."""
    return ren.scatter == "sorted" and vol_shape_cl[-1] in (64, 128)


def _finish_volume_grad(ren, rays, rps, defer, vol_shape_cl, cl3d):
    """dL/dvoxel_feat in the caller's memory format from the deferred (z, dlatent) of every pass."""
    SB, S0, S1, S2, C = vol_shape_cl
    keep = bool(getattr(ren, "keep_voxel_counts", False))  # for parallel.sparse_allreduce_volume_grad (config 5)
    g = torch.empty(vol_shape_cl if cl3d else (SB, C, S0, S1, S2), device=rays.device, dtype=torch.float32)
    res = ops.scatter_volume_grad_merged(rays, rps, defer, g, not cl3d, ren._bounds, want_counts=keep)
    if keep:
        ren.last_voxel_counts = res[1]
    return g.permute(0, 4, 1, 2, 3) if cl3d else g


def _is_channels_last_3d(t):
    """(SB,C,S0,S1,S2) tensor whose memory is (SB,S0,S1,S2,C): what a conv3d stack run in torch.channels_last_3d
    hands over.  The kernels read exactly that layout, so no re-layout pass is needed."""
    return t.dim() == 5 and t.shape[1] > 1 and not t.is_contiguous() and \
        t.is_contiguous(memory_format=torch.channels_last_3d)


class _GradViews(dict):
    """name -> gradient view; `.flat` is the one buffer behind all of them (what the all-reduce takes)."""
    flat = None


def _zero_grads(mlp: ops.FieldMLP):
    """Zeroed gradient buffers for every MLP parameter: ONE flat allocation / fill, views per parameter."""
    names = mlp.names()
    sizes = [mlp.params[n].numel() for n in names]
    dev = mlp.params[names[0]].device
    # every view starts on a 16 B boundary (the weight-gradient kernels add float4s): sizes that are not multiples of 4
    # occur with the coord / attention heads (d_out = 397)
    padded = [(sz + 3) // 4 * 4 for sz in sizes]
    flat = torch.zeros(sum(padded), device=dev, dtype=torch.float32)
    out, off = _GradViews(), 0
    out.flat = flat
    for n, sz, psz in zip(names, sizes, padded):
        out[n] = flat[off:off + sz].view_as(mlp.params[n])
        off += psz
    return out


def _start_grad_allreduce(ren, grads_c, grads_f):
    """If the data-parallel wrapper installed a hook (parallel.overlap_mlp_grad_allreduce), hand it the flat MLP
    gradient buffer(s) of this step; it starts the (asynchronous) all-reduce and returns what to wait for."""
    hook = getattr(ren, "_grad_allreduce", None)
    if hook is None:
        return None
    flats = [grads_c.flat] + ([grads_f.flat] if grads_f is not grads_c else [])
    return hook(flats)


def _finish_grad_allreduce(pending):
    if pending is not None:
        pending()


def _zeros_like_or(t, ref_shape, device):
    return t.contiguous() if t is not None else torch.zeros(ref_shape, device=device, dtype=torch.float32)


class _ForwardNerfFn(torch.autograd.Function):
    """forward_nerf (neural_rendering.py:435-471) as one autograd node.

    inputs : ren, voxel_feat (SB,C,S,S,S), rays (R,8), sb, noise dict, n_param_coarse, keep (activations), *params
    outputs: z_coarse, cw, crgb, cemb, cdep[, z_fine, fw, frgb, femb, fdep]
    """

    @staticmethod
    def forward(ctx, ren, voxel_feat, rays, sb, noise, n_pc, keep, *params):
        R = rays.shape[0]
        rps = R // sb
        mlp_c = ren.nerf_model.mlp_coarse.handle(ren._prec)
        mlp_f = ren.nerf_model.mlp_fine.handle(ren._prec)
        mlp_c.pack(force=keep)                             # once per step (see FieldMLP.pack); the passes reuse it
        if mlp_f is not mlp_c:
            mlp_f.pack(force=keep)
        held = getattr(ren, "_vol_cl_held", None)          # rendering(): one re-layout for all ray chunks
        cl3d = _is_channels_last_3d(voxel_feat)
        Kc, Kf, Kfd = ren.n_coarse, ren.n_fine, ren.n_fine_depth
        touched = None                                     # re-layout of the voxels the rays touch, pass by pass
        if cl3d:                                           # producer ran in torch.channels_last_3d: zero-copy view
            vol_cl = voxel_feat.permute(0, 2, 3, 4, 1)
        elif held is not None and held[0] is voxel_feat:
            vol_cl = held[1]
        elif ren.sparse_relayout and voxel_feat.shape[1] in (64, 128) and \
                R * (Kc + (Kc + Kf if ren.using_fine else 0)) < ren.sparse_relayout_ratio * voxel_feat.shape[0] * voxel_feat[0, 0].numel():
            # far fewer samples than voxels (2048 rays of one scene on a 200^3 grid, BASELINE config 5 split over 8 GPUs:
            # 0.1 samples per voxel, 1.5 % of the voxels touched): a dense (C, V) -> (V, C) pass would move 8.2 GB of which
            # the gather reads 60 MB.  (At config 2, 0.4 samples per voxel, the rays already touch most 32-voxel tiles and
            # the dense pass is the faster one: measured, profiles/r02c_summary.md.)
            touched = ops.TouchedRelayout(voxel_feat, ren._bounds)
            vol_cl = touched.vol_cl
        else:
            vol_cl = ops.volume_to_channels_last(voxel_feat)
        ctx.cl3d = cl3d
        z_c = ops.sample_coarse(rays, Kc, noise.get("coarse"), ren.lindisp)
        if touched is not None:
            touched.add(rays, z_c, rps)
        st_c, (cw, crgb, cemb, cdep) = _pass_forward(ren, mlp_c, vol_cl, rays, z_c, rps, keep,
                                                     _sigma_noise(ren, noise, "sigma_c", R, Kc, rays.device),
                                                     repack=False)
        outs = [z_c, cw, crgb, cemb, cdep]
        st_f = None
        depth_mask = None
        if ren.using_fine:
            K = Kc + Kf
            z_all = torch.empty(R, K, device=rays.device, dtype=torch.float32)
            z_all[:, :Kc] = z_c
            kf = Kf - Kfd
            if kf > 0:
                zf = ops.sample_fine(rays, cw, Kc, noise["u"], noise.get("fine"), ren.lindisp)
                z_all[:, Kc:Kc + kf] = zf
            if Kfd > 0:
                # neural_rendering.py:210-221 (tiny; plain torch ops, incl. the clamp mask for backward)
                z0 = cdep.unsqueeze(1).repeat(1, Kfd)
                nz = noise.get("depth")
                if nz is not None:
                    z0 = z0 + nz * ren.depth_std
                near, far = rays[:, 6:7], rays[:, 7:8]
                depth_mask = ((z0 <= far) & (z0 >= near)).to(torch.float32)
                z_all[:, Kc + kf:] = torch.max(torch.min(z0, far), near)
            reuse = ren.reuse_coarse_evals and mlp_f is mlp_c and not ren.regress_coord
            z_new = z_all[:, Kc:].contiguous() if reuse else None
            if touched is not None:                        # what the fine samples touch beyond the coarse ones
                touched.add(rays, z_all[:, Kc:].contiguous(), rps)
            z_all, perm = ops.sort_rows(z_all, want_perm=True)
            sn_f = _sigma_noise(ren, noise, "sigma_f", R, K, rays.device)
            if reuse:
                st_f, (fw, frgb, femb, fdep) = _pass_forward_reuse(ren, mlp_f, vol_cl, rays, z_new, z_all, perm,
                                                                   st_c, rps, keep, sn_f)
            else:
                st_f, (fw, frgb, femb, fdep) = _pass_forward(ren, mlp_f, vol_cl, rays, z_all, rps, keep, sn_f,
                                                             repack=False)
                st_f.perm = perm
            outs += [z_all, fw, frgb, femb, fdep]
        if ren.regress_coord:                              # appended: the raw coordinate outputs of each pass
            outs.append(_coord_raw(ren, st_c))
            if st_f is not None:
                outs.append(_coord_raw(ren, st_f))
        ctx.ren, ctx.st_c, ctx.st_f, ctx.sb, ctx.n_pc = ren, st_c, st_f, sb, n_pc
        ctx.vol_shape = tuple(vol_cl.shape)
        ctx.depth_mask = depth_mask
        ctx.n_params = len(params)
        # one call only: a second mark_non_differentiable() would replace the first, give z_c a grad_fn
        # pointing at this node while st_c.z holds it, and leak every activation through that cycle
        ctx.mark_non_differentiable(*([z_c] + ([outs[5]] if st_f is not None else [])))
        if not keep:                           # nothing to back-propagate: drop what a backward would have used
            st_c.acts = st_c.field_in = st_c.field_out = None
            if st_f is not None:
                st_f.acts = st_f.field_in = st_f.field_out = None
        return tuple(outs)

    @staticmethod
    def backward(ctx, *g):
        ren, st_c, st_f = ctx.ren, ctx.st_c, ctx.st_f
        if st_c is None:
            raise RuntimeError("forward_nerf: backward a second time (its buffers are released by the first backward; "
                               "retain_graph is not supported)")
        # this node lives as long as any output tensor does (a training loop keeps the loss dict for logging): drop the
        # activations now rather than when `out['loss']` goes away - the next step's forward would otherwise coexist
        # with them (2 x 70 GB at config 4)
        ctx.st_c = ctx.st_f = None
        dev = st_c.rays.device
        R = st_c.rays.shape[0]
        D = ren._d_comp
        n_base = 5 if st_f is None else 10
        d_ccoord = g[n_base] if ren.regress_coord else None
        d_fcoord = g[n_base + 1] if ren.regress_coord and st_f is not None else None
        shared = ren.nerf_model.mlp_fine is ren.nerf_model.mlp_coarse
        names_c = st_c.mlp.names()
        grads_c = _zero_grads(st_c.mlp)
        grads_f = grads_c if shared or st_f is None else _zero_grads(st_f.mlp)
        want_vol = ctx.needs_input_grad[1]
        merged = _merged_scatter_ok(ren, ctx.vol_shape)
        defer = [] if merged else None
        grad_cl = None
        if not merged:
            alloc = torch.empty if ren.scatter == "sorted" else torch.zeros
            grad_cl = alloc(ctx.vol_shape, device=dev, dtype=torch.float32)
        _, d_cw, d_crgb, d_cemb, d_cdep = g[:5]
        if st_f is not None and st_f.base is not None:        # the fine pass reused the coarse evaluations
            _backward_reuse(ren, st_c, st_f, (d_cw, d_crgb, d_cemb, d_cdep), g[6:10], grads_c, defer, ctx.depth_mask,
                            grad_cl)
            pending = _start_grad_allreduce(ren, grads_c, grads_c)
            d_vol = None
            if want_vol and merged:
                d_vol = _finish_volume_grad(ren, st_c.rays, st_c.rps, defer, ctx.vol_shape, ctx.cl3d)
            elif want_vol:
                d_vol = grad_cl.permute(0, 4, 1, 2, 3) if ctx.cl3d else ops.volume_to_channels_first(grad_cl)
            _finish_grad_allreduce(pending)
            return (None, d_vol, None, None, None, None, None, *[grads_c[n] for n in names_c])
        d_cdep = _zeros_like_or(d_cdep, (R,), dev)
        if st_f is not None:
            _, d_fw, d_frgb, d_femb, d_fdep = g[5:10]
            Kfd = ren.n_fine_depth
            d_z = _pass_backward(ren, st_f, _zeros_like_or(d_frgb, (R, 3), dev),
                                 _zeros_like_or(d_femb, (R, D), dev), d_fdep, d_fw, grads_f, grad_cl,
                                 want_dz=Kfd > 0, defer=defer, d_coord_raw=d_fcoord)
            if Kfd > 0:
                # route dL/dz of the depth-guided samples back through sort and clamp to coarse depth
                K = st_f.z.shape[1]
                d_cat = torch.zeros(R, K, device=dev, dtype=torch.float32)
                d_cat.scatter_(1, st_f.perm.long(), d_z)
                d_cdep = d_cdep + (d_cat[:, K - Kfd:] * ctx.depth_mask).sum(-1)
        # overlap_scatter: the last pass's backward records an event as soon as its dL/dlatent is complete - ahead of its
        # weight gradients - and the merged volume scatter (HBM-bound, one 18 KB CTA fits beside a weight-gradient CTA)
        # runs on a stream of its own under them
        overlap = merged and want_vol and ren.overlap_scatter and not torch.cuda.is_current_stream_capturing()
        ev = _scatter_side(ren, dev)[1] if overlap else None
        _pass_backward(ren, st_c, _zeros_like_or(d_crgb, (R, 3), dev), _zeros_like_or(d_cemb, (R, D), dev),
                       d_cdep, d_cw, grads_c, grad_cl, first=st_f is None, defer=defer, d_coord_raw=d_ccoord,
                       dlatent_event=ev)
        # every MLP gradient of this step is final: start their all-reduce now, it overlaps the volume scatter
        pending = _start_grad_allreduce(ren, grads_c, grads_f)
        d_vol = None
        if want_vol:
            if overlap:
                main = torch.cuda.current_stream(dev)
                side = _scatter_side(ren, dev)[0]
                side.wait_event(ev)               # both passes' dlatent (and everything enqueued before) are complete
                with torch.cuda.stream(side):
                    d_vol = _finish_volume_grad(ren, st_c.rays, st_c.rps, defer, ctx.vol_shape, ctx.cl3d)
                main.wait_stream(side)
                d_vol.record_stream(main)         # allocated on the side stream, consumed (and freed) on this one
                if ren.last_voxel_counts is not None and getattr(ren, "keep_voxel_counts", False):
                    ren.last_voxel_counts.record_stream(main)
            elif merged:
                d_vol = _finish_volume_grad(ren, st_c.rays, st_c.rps, defer, ctx.vol_shape, ctx.cl3d)
            else:
                d_vol = grad_cl.permute(0, 4, 1, 2, 3) if ctx.cl3d else ops.volume_to_channels_first(grad_cl)
        _finish_grad_allreduce(pending)
        pg = [grads_c[n] for n in names_c]
        if not shared and st_f is not None:
            pg += [grads_f[n] for n in st_f.mlp.names()]
        return (None, d_vol, None, None, None, None, None, *pg)


class _RenderLossFn(torch.autograd.Function):
    """The rgb / embed mean-squared errors of both passes (neural_rendering.py:664-685) as one autograd node:
    one kernel produces the four terms and, in the same read, their gradients; backward only scales them."""

    @staticmethod
    def forward(ctx, rgb_c, rgb_f, emb_c, emb_f, gt_rgb, gt_embed, idx, rps):
        need = any(ctx.needs_input_grad[:4])
        terms, grads = ops.render_loss(rgb_c, rgb_f, emb_c, emb_f, rps, gt_rgb, gt_embed, idx, want_grads=need)
        ctx.grads = grads
        return terms

    @staticmethod
    def backward(ctx, g):
        d = ctx.grads
        out = [None] * 8
        if d is not None:
            for i in range(4):
                if ctx.needs_input_grad[i]:
                    out[i] = d[i] * g[i]                       # terms: rgb_c, rgb_f, emb_c, emb_f
        return tuple(out)


# ------------------------------------------------------------------------------ the renderer
class NeuralRenderer(nn.Module):
    """take a voxel, camera pose, and camera intrinsics as input, and output a rendered image
    (neural_rendering.py:86-711)."""

    def __init__(self, cfg, coordinate_bounds, precision="bf16", feature_extractor=None):
        super().__init__()
        self.cfg = cfg
        self.coordinate_bounds = coordinate_bounds
        g = lambda k, d=None: _cfg_get(cfg, k, d)
        self.W, self.H = g("image_width"), g("image_height")
        self.z_near, self.z_far = g("z_near"), g("z_far")
        self.regress_coord, self.regress_attention = bool(g("regress_coord", False)), bool(g("regress_attention", False))
        self.n_coarse, self.n_fine, self.n_fine_depth = g("n_coarse"), g("n_fine"), g("n_fine_depth", 0)
        self.lindisp = g("lindisp", False)
        self.using_fine = self.n_fine > 0
        self.eval_batch_size = g("eval_batch_size", 4096)
        self.ret_last_feat = g("ret_last_feat", False)
        self.noise_std, self.white_bkgd, self.depth_std = g("noise_std", 0.0), g("white_bkgd", False), g("depth_std", 0.001)
        self.nerf_model = PixelNeRFEmbedNet(cfg, coordinate_bounds)
        # multi-scale voxels / ret_last_feat change the MLP's input / what is composited: composed.py
        self._composed = bool(self.ret_last_feat) or self.nerf_model.use_multi_scale_voxel or \
            self.nerf_model.use_code_viewdirs or self.nerf_model.mlp_coarse.general
        self.model_name = g("foundation_model_name", None)
        if self.model_name not in ("odise", "diffusion", "dinov2", "deepfloyd", None):
            raise NotImplementedError(f"foundation model {self.model_name} is not implemented")
        # target-feature extractors are outside the render path: pass `gt_embed`,
        # or supply `feature_extractor(gt_rgb, lang_goal) -> (B,D,H,W)`.
        self.feature_extractor = feature_extractor
        self.lambda_embed = g("lambda_embed", 0.01)
        self.lambda_depth = g("lambda_depth", 0.0)
        self.threshold_depth_supervision = g("threshold_depth_supervision", 0.8)
        self.precision = precision
        self.scatter = "sorted"                # volume-gradient scatter: "sorted" (atomics-free, default) | "atomic"
        self.deterministic = False             # True: ordered split reduction of the MLP weight gradients as well
        self.target_ready_event = None         # optional torch.cuda.Event: gt_rgb / gt_embed were copied on a side
                                               # stream; awaited right before the losses read them
        self.perturb = True
        self.reuse_coarse_evals = False        # True: the fine pass evaluates only its Kf new samples and composites
                                               # the Kc coarse ones from the coarse pass's outputs (bit-identical
                                               # rendering, 1/3 fewer MLP evaluations at Kc = Kf; needs share_mlp)
        self._grad_allreduce = None            # set by parallel.overlap_mlp_grad_allreduce(): the backward all-reduces
                                               # the flat MLP gradient buffer itself, under the volume scatter
        self.fused_loss = True                 # rgb / embed losses + their gradients in one kernel (nrf_render_loss)
        self.render_chunk_rays = 4096          # neural_rendering.py:482
        self.trace_ranges = True               # the reference's five profiler labels as record_function + NVTX ranges
        # sample tiles without a sample inside the grid have an all-zero latent: the forward skips their latent k-panels,
        # the backward their dL/dlatent and lin_z weight-gradient blocks (NRF_SKIP_EMPTY_TILES=0: every tile in full)
        self.skip_empty_latent_tiles = os.environ.get("NRF_SKIP_EMPTY_TILES", "1") != "0"
        # a training step with fewer samples than voxels re-lays out only the 32-voxel tiles its rays touch
        # (ops.TouchedRelayout; NRF_SPARSE_RELAYOUT=0: always the dense pass)
        self.sparse_relayout = os.environ.get("NRF_SPARSE_RELAYOUT", "1") != "0"
        self.sparse_relayout_ratio = float(os.environ.get("NRF_SPARSE_RELAYOUT_RATIO", "0.25"))   # samples per voxel below which it pays
        # the volume scatter of a training step on a side stream under the last pass's weight gradients (see
        # _ForwardNerfFn.backward); NRF_SCATTER_OVERLAP=1 switches it on
        self.overlap_scatter = os.environ.get("NRF_SCATTER_OVERLAP", "0") == "1"
        self.keep_voxel_counts = False         # True: the backward leaves the per-voxel entry counts of its scatter in
        self.last_voxel_counts = None          # `last_voxel_counts` (what the sparse volume-gradient exchange sends)
        self._num_freqs = self.nerf_model.code.num_freqs
        self._freq_factor = float(self.nerf_model.code.freq_factor)
        self._d_embed = self.nerf_model.d_embed
        # channels the compositing kernels weight-sum after [rgb, sigma]: embed (+ coord, which is then ignored) + attention,
        # rounded up to whole float4s (the field-output rows are padded with exact zeros)
        n_heads = (3 if self.regress_coord else 0) + (6 if self.regress_attention else 0)
        self._d_comp = (self._d_embed + n_heads + 3) // 4 * 4
        if not self.nerf_model.code.include_input:
            raise NotImplementedError("code.include_input=False is not built")

    # ---- internals
    @property
    def _prec(self):
        return ops.PRECISIONS[self.precision]

    @property
    def _bounds(self):
        """Host copy of the 6 bounds, read back once per `coordinate_bounds` object (callers pass a CUDA tensor: a
        `.cpu()` per kernel call would stall the host on the stream several times a step)."""
        cached = getattr(self, "_bounds_cache", None)
        if cached is None or cached[0] is not self.coordinate_bounds:
            cached = (self.coordinate_bounds,
                      torch.as_tensor(self.coordinate_bounds, dtype=torch.float32).reshape(-1).cpu())
            self._bounds_cache = cached
        return cached[1]

    def _draw_noise(self, R, device):
        """Noise in the reference's draw order (SURVEY 8b 'RNG'); zeros / a fixed grid if not perturb."""
        Kc, kf, Kfd = self.n_coarse, self.n_fine - self.n_fine_depth, self.n_fine_depth
        n = {}
        sig = self.training and self.noise_std and self.noise_std > 0.0
        if self.perturb:
            n["coarse"] = torch.rand(R, Kc, device=device)
            if sig:                                        # drawn inside the coarse composite (:336-337)
                n["sigma_c"] = torch.randn(R, Kc, device=device)
            if self.using_fine and kf > 0:
                n["u"] = torch.rand(R, kf, dtype=torch.float32, device=device)
                n["fine"] = torch.rand(R, kf, device=device)
            if self.using_fine and Kfd > 0:
                n["depth"] = torch.randn(R, Kfd, device=device)
            if sig and self.using_fine:
                n["sigma_f"] = torch.randn(R, Kc + self.n_fine, device=device)
        elif self.using_fine and kf > 0:
            n["u"] = ((torch.arange(kf, device=device, dtype=torch.float32) + 0.5) / kf).repeat(R, 1)
        return n

    def _params_flat(self):
        m = self.nerf_model
        pc = m.mlp_coarse.param_dict()
        names = m.mlp_coarse.handle(self._prec).names()
        ps = [pc[n] for n in names]
        n_pc = len(ps)
        if m.mlp_fine is not m.mlp_coarse:
            pf = m.mlp_fine.param_dict()
            ps += [pf[n] for n in m.mlp_fine.handle(self._prec).names()]
        return n_pc, ps

    # ---- reference API
    def sample_coarse(self, rays, jitter="draw"):
        """neural_rendering.py:159-176."""
        if isinstance(jitter, str):
            jitter = torch.rand(rays.shape[0], self.n_coarse, device=rays.device) if self.perturb else None
        return ops.sample_coarse(rays, self.n_coarse, jitter, self.lindisp)

    def sample_fine(self, rays, weights, u=None, jitter="draw"):
        """neural_rendering.py:179-207."""
        R, kf = rays.shape[0], self.n_fine - self.n_fine_depth
        if u is None:
            u = torch.rand(R, kf, dtype=torch.float32, device=rays.device)
        if isinstance(jitter, str):
            jitter = torch.rand(R, kf, device=rays.device) if self.perturb else None
        return ops.sample_fine(rays, weights.detach(), self.n_coarse, u, jitter, self.lindisp)

    def sample_fine_depth(self, rays, depth, noise="draw"):
        """neural_rendering.py:210-221 (plain torch: R*Kfd elements)."""
        z = depth.unsqueeze(1).repeat((1, self.n_fine_depth))
        if isinstance(noise, str):
            noise = torch.randn_like(z) if self.perturb else None
        if noise is not None:
            z = z + noise * self.depth_std
        return torch.max(torch.min(z, rays[:, -1:]), rays[:, -2:-1])

    def encode(self, multi_scale_voxel_list, voxel_density, lang, voxel_feat, poses, focal, c=None):
        """neural_rendering.py:428-432."""
        self.nerf_model.encode(multi_scale_voxel_list=multi_scale_voxel_list, voxel_density=voxel_density,
                               lang=lang, voxel_feat=voxel_feat, poses=poses, focal=focal, c=c)

    def composite(self, model, rays, z_samp, coarse=True, sb=0):
        """neural_rendering.py:224-395 for externally supplied samples -> weights, rgb, embed, depth."""
        model = model if model is not None else self.nerf_model
        sb = max(int(sb), 1)
        if self._composed:
            from . import composed
            return composed.composite_pass(self, model, rays.contiguous(), z_samp.contiguous(), coarse, sb,
                                           _sigma_noise(self, {}, None, z_samp.shape[0], z_samp.shape[1], z_samp.device))
        mlp = (model.mlp_coarse if coarse or model.mlp_fine is None else model.mlp_fine)
        h = mlp.handle(self._prec)
        names = h.names()
        ps = [mlp.param_dict()[n] for n in names]
        rays, z_samp = rays.contiguous(), z_samp.contiguous()
        res = _CompositeFn.apply(self, h, model.voxel_feat, rays, z_samp, sb, *ps)
        return self._split_heads(*res[:4], res[4] if self.regress_coord else None, rays, z_samp)

    def _format_outputs(self, rendered_outputs, superbatch_size, want_weights=False):
        """neural_rendering.py:398-426: (weights, rgb, embed, [coord], [attention], depth) -> AttrDict."""
        rendered_outputs = list(rendered_outputs)
        weights, rgb, embed = rendered_outputs[:3]
        depth = rendered_outputs[-1]
        rest = rendered_outputs[3:-1]
        coord = rest.pop(0) if self.regress_coord else None
        attention = rest.pop(0) if self.regress_attention else None
        if superbatch_size > 0:
            rgb = rgb.reshape(superbatch_size, -1, 3)
            embed = embed.reshape(superbatch_size, -1, embed.shape[-1])
            depth = depth.reshape(superbatch_size, -1)
            weights = weights.reshape(superbatch_size, -1, weights.shape[-1])
            if coord is not None:
                coord = coord.reshape(superbatch_size, -1, 3)
            if attention is not None:
                attention = attention.reshape(superbatch_size, -1, attention.shape[-1])
        ret = AttrDict(rgb=rgb, embed=embed, depth=depth)
        if want_weights:
            ret.weights = weights
        if coord is not None:
            ret.coord = coord
        if attention is not None:
            ret.attention = attention
        return ret

    def _split_heads(self, w, rgb, emb_all, dep, coord_raw, rays, z, d_embed=None):
        """(weights, rgb, embed', depth) of a compositing pass + the raw coordinate outputs -> the reference's tuple
        (weights, rgb, embed, [coord], [attention], depth) (neural_rendering.py:388-395)."""
        D = self._d_embed if d_embed is None else d_embed
        out = [w, rgb, emb_all[:, :D] if emb_all.shape[1] != D else emb_all]
        o = D
        if self.regress_coord:
            # coord_final = mean over the samples of (coord - canon_xyz) (models_embed.py:449, neural_rendering.py:356-357)
            b6 = torch.as_tensor(self._bounds, dtype=torch.float32).to(rays.device, non_blocking=True)
            pts = rays[:, None, :3] + z.unsqueeze(2) * rays[:, None, 3:6]
            canon = (pts - b6[:3]) / (b6[3:] - b6[:3])
            out.append(torch.mean(coord_raw - canon, -2))
            o += 3
        if self.regress_attention:
            out.append(emb_all[:, o:o + 6])
        out.append(dep)
        return tuple(out)

    def extract_radience(self, model, rays, z_samp, coarse=True, sb=0, ret_last_feat=False):
        """The ancestor renderer's per-sample field values (nerf_embed.py:432-516): points, rgbs, sigmas, embeds."""
        from . import extract
        return extract.extract_radience(self, model, rays, z_samp, coarse, sb, ret_last_feat)

    def forward_nerf(self, rays, want_weights=False, noise=None, extract_radience=False, ret_last_feat=False):
        """neural_rendering.py:435-471.  rays (SB,B,8) -> AttrDict(coarse=..., fine=...).
        extract_radience (the ancestor's switch, nerf_embed.py:338-342): run the coarse pass, draw the fine samples and
        return the field values (points, rgbs, sigmas, embeds) at the sorted sample set instead of compositing them."""
        if extract_radience:
            with torch.no_grad():
                out = self.forward_nerf(rays, want_weights=True, noise=noise)
                z = out.fine.z if self.using_fine else out.coarse.z
                return self.extract_radience(self.nerf_model, rays.reshape(-1, 8), z, coarse=not self.using_fine,
                                             sb=rays.shape[0], ret_last_feat=ret_last_feat)
        assert len(rays.shape) == 3
        sb = rays.shape[0]
        vol = self.nerf_model.voxel_feat
        if vol is None:
            raise RuntimeError("call encode() before forward_nerf()")
        if vol.shape[0] != sb:
            raise RuntimeError("grid_sampler(): expected grid and input to have same batch size, "
                               f"but got input with sizes {list(vol.shape)} and {sb} ray batches")
        flat = rays.reshape(-1, 8).contiguous()
        _trace.enabled = bool(self.trace_ranges)
        if noise is None:
            noise = self._draw_noise(flat.shape[0], flat.device)
        if self._composed:
            from . import composed
            return composed.forward_nerf(self, flat, sb, noise, want_weights)
        n_pc, ps = self._params_flat()
        # decided here: inside Function.forward grad mode is always off and needs_input_grad ignores no_grad()
        keep = torch.is_grad_enabled() and (vol.requires_grad or any(p.requires_grad for p in ps))
        outs = _ForwardNerfFn.apply(self, vol, flat, sb, noise, n_pc, keep, *ps)
        n_base = 10 if self.using_fine else 5
        craw = outs[n_base] if self.regress_coord else None
        outputs = AttrDict(coarse=self._format_outputs(self._split_heads(*outs[1:5], craw, flat, outs[0]), sb, want_weights))
        outputs.coarse.z = outs[0]
        if self.using_fine:
            fraw = outs[n_base + 1] if self.regress_coord else None
            outputs.fine = self._format_outputs(self._split_heads(*outs[6:10], fraw, flat, outs[5]), sb, want_weights)
            outputs.fine.z = outs[5]
        return outputs

    @torch.no_grad()
    def rendering(self, voxel_feat, language, multi_scale_voxel_list, voxel_density, voxel_pose, focal,
                  tgt_pose, c=None):
        """neural_rendering.py:474-502: full images, fine outputs, SB forced to 1."""
        rays = gen_rays(tgt_pose, self.W, self.H, focal, self.z_near, self.z_far, c=c)
        self.encode(multi_scale_voxel_list=multi_scale_voxel_list, voxel_density=voxel_density, lang=language,
                    voxel_feat=voxel_feat, poses=voxel_pose, focal=focal, c=c)
        B, H, W, _ = rays.shape
        rays = rays.reshape(B * H * W, 8)
        rgbs, embeds, depths = [], [], []
        # the volume is re-laid out channels-last once for all ray chunks (the tensor is held, so identity is safe)
        voxel_feat = self.nerf_model.voxel_feat            # as encode() keeps it (fp32 also for a bf16 / fp16 hand-over)
        self._vol_cl_held = None if _is_channels_last_3d(voxel_feat) else \
            (voxel_feat, ops.volume_to_channels_last(voxel_feat))
        try:
            for i in range(0, rays.shape[0], self.render_chunk_rays):
                out = self.forward_nerf(rays[i:i + self.render_chunk_rays].unsqueeze(0))
                fine = out.fine
                rgbs.append(fine.rgb.squeeze(0))
                embeds.append(fine.embed.squeeze(0))
                depths.append(fine.depth.squeeze(0))
        finally:
            self._vol_cl_held = None
        rgbs = torch.cat(rgbs, dim=0).reshape(B, H, W, 3)
        embeds = torch.cat(embeds, dim=0).reshape(B, H, W, -1)
        depths = torch.cat(depths, dim=0).reshape(B, H, W)
        return rgbs, embeds, depths

    def extract_foundation_model_feature(self, gt_rgb, lang_goal):
        """neural_rendering.py:505-592 is target preprocessing, outside the render path."""
        if self.feature_extractor is None:
            raise NotImplementedError(
                "target-feature extraction (odise / diffusion / dinov2 / deepfloyd) is outside the render "
                "hot path: pass gt_embed=..., or construct NeuralRenderer(..., feature_extractor=fn)")
        feat = self.feature_extractor(gt_rgb, lang_goal)
        return F.interpolate(feat, size=(self.H, self.W), mode="bilinear", align_corners=False)

    def compute_rendering_loss(self, multi_scale_voxel_list, voxel_density, language, voxel_feat, voxel_poses,
                               focal, gt_rgb, gt_depth, gt_pose, c=None, lang_goal=None, gt_embed=None):
        """neural_rendering.py:595-707."""
        loss, scalars = self._loss_tensors(multi_scale_voxel_list, voxel_density, language, voxel_feat, voxel_poses,
                                           focal, gt_rgb, gt_depth, gt_pose, c, lang_goal, gt_embed)
        return LossDict(loss, scalars)

    def _loss_tensors(self, multi_scale_voxel_list, voxel_density, language, voxel_feat, voxel_poses,
                      focal, gt_rgb, gt_depth, gt_pose, c=None, lang_goal=None, gt_embed=None):
        """The body of compute_rendering_loss on device tensors only: -> (loss, the 7 scalars behind the float entries
        of the loss dict).  No host interaction at all, so the whole step can be captured into a CUDA graph (graphed.py)."""
        rays = gen_rays(gt_pose, self.W, self.H, focal, self.z_near, self.z_far, c=c)
        self.encode(multi_scale_voxel_list=multi_scale_voxel_list, voxel_density=voxel_density, lang=language,
                    voxel_feat=voxel_feat, poses=voxel_poses, focal=focal, c=c)
        B, H, W, DimRay = rays.shape
        rays = rays.reshape(B, H * W, DimRay)
        chunk_size = _cfg_get(self.cfg, "ray_chunk_size")
        idx = torch.randint(H * W, (chunk_size,), device=rays.device)      # shared across scenes (:608)
        sampled_rays = rays[:, idx, :]
        outputs = self.forward_nerf(sampled_rays)
        if self.target_ready_event is not None:            # targets prefetched on a side stream (see bench.py)
            torch.cuda.current_stream(rays.device).wait_event(self.target_ready_event)
            self.target_ready_event = None
        if gt_embed is None:
            with torch.no_grad():
                gt_embed = self.extract_foundation_model_feature(gt_rgb, lang_goal)
                embed_dim = _cfg_get(self.cfg, "d_embed")
                if embed_dim < 512 and gt_embed.shape[1] != embed_dim:
                    # neural_rendering.py:640-646 (sklearn PCA on the CPU) on the device: no host round trip / stall
                    Bq, Dq, Hq, Wq = gt_embed.shape
                    flat = gt_embed.permute(0, 2, 3, 1).reshape(Bq * Hq * Wq, Dq)
                    flat = pca_fit_transform(flat.float(), embed_dim)
                    gt_embed = flat.reshape(Bq, Hq, Wq, embed_dim).permute(0, 3, 1, 2)
                gt_embed = gt_embed.permute(0, 2, 3, 1)
        D = gt_embed.shape[-1]
        if self.fused_loss and gt_rgb.is_cuda and gt_rgb.dtype == torch.float32 and gt_embed.dtype == torch.float32:
            # one kernel: target gather [:, idx], the four mse terms and their gradients (nrf_render_loss)
            t = _RenderLossFn.apply(outputs.coarse.rgb.reshape(-1, 3), outputs.fine.rgb.reshape(-1, 3),
                                    outputs.coarse.embed.reshape(-1, D), outputs.fine.embed.reshape(-1, D),
                                    gt_rgb.reshape(B, H * W, 3), gt_embed.reshape(B, H * W, D), idx, chunk_size)
            loss_rgb_coarse, loss_rgb_fine = t[0], t[1]
            loss_embed_coarse, loss_embed_fine = self.lambda_embed * t[2], self.lambda_embed * t[3]
            mse_fine = t[1].detach()
        else:
            gt_rgb = gt_rgb.reshape(B, H * W, 3)[:, idx, :]
            loss_rgb_coarse = F.mse_loss(outputs.coarse.rgb, gt_rgb)
            loss_rgb_fine = F.mse_loss(outputs.fine.rgb, gt_rgb)
            mse_fine = torch.mean((outputs.fine.rgb.detach() - gt_rgb) ** 2)
            gt_embed = gt_embed.reshape(B, H * W, -1)[:, idx, :]
            loss_embed_coarse = self.lambda_embed * F.mse_loss(outputs.coarse.embed, gt_embed)
            loss_embed_fine = self.lambda_embed * F.mse_loss(outputs.fine.embed, gt_embed)
        loss = loss_rgb_coarse + loss_rgb_fine
        psnr = torch.where(mse_fine == 0, torch.full_like(mse_fine, 100.0),          # PSNR_torch without its host-side branch
                           20 * torch.log10(1.0 / torch.sqrt(mse_fine.clamp_min(1e-45))))
        loss = loss + loss_embed_coarse + loss_embed_fine
        if gt_depth is not None:
            gt_depth = gt_depth.reshape(B, H * W)[:, idx]
            far_mask = gt_depth < self.z_far
            loss_depth_coarse = self.lambda_depth * F.mse_loss(gt_depth[far_mask], outputs.coarse.depth[far_mask])
            loss_depth_fine = self.lambda_depth * F.mse_loss(gt_depth[far_mask], outputs.fine.depth[far_mask])
            loss = loss + loss_depth_coarse + loss_depth_fine
        else:
            loss_depth_coarse = torch.zeros((), device=loss.device)
            loss_depth_fine = torch.zeros((), device=loss.device)
        # no host sync here (the reference has 13 .item() calls): see LossDict
        psnr_t = psnr if torch.is_tensor(psnr) else torch.tensor(float(psnr), device=loss.device)
        scalars = torch.stack([v.detach().float().reshape(()) for v in
                               (loss_rgb_coarse, loss_rgb_fine, loss_embed_coarse, loss_embed_fine,
                                loss_depth_coarse, loss_depth_fine, psnr_t)])
        return loss, scalars

    def forward(self, multi_scale_voxel_list, voxel_density, language, voxel_feat, voxel_poses, focal, gt_rgb,
                gt_depth, gt_pose, c=None, lang_goal=None, gt_embed=None):
        """neural_rendering.py:710-711."""
        return self.compute_rendering_loss(multi_scale_voxel_list, voxel_density, language, voxel_feat,
                                           voxel_poses, focal, gt_rgb, gt_depth, gt_pose, c, lang_goal, gt_embed)


class _CompositeFn(torch.autograd.Function):
    """A single composite pass with caller-supplied samples (public `NeuralRenderer.composite`)."""

    @staticmethod
    def forward(ctx, ren, h, voxel_feat, rays, z, sb, *params):
        cl3d = _is_channels_last_3d(voxel_feat)
        vol_cl = voxel_feat.permute(0, 2, 3, 4, 1) if cl3d else ops.volume_to_channels_last(voxel_feat)
        st, outs = _pass_forward(ren, h, vol_cl, rays, z, rays.shape[0] // sb,
                                 sig_noise=_sigma_noise(ren, {}, None, z.shape[0], z.shape[1], z.device))
        ctx.ren, ctx.st, ctx.vol_shape, ctx.cl3d = ren, st, tuple(vol_cl.shape), cl3d
        return (*outs, _coord_raw(ren, st)) if ren.regress_coord else outs

    @staticmethod
    def backward(ctx, d_w, d_rgb, d_emb, d_dep, d_coord_raw=None):
        ren, st = ctx.ren, ctx.st
        if st is None:
            raise RuntimeError("composite: backward a second time (retain_graph is not supported)")
        ctx.st = None
        dev = st.rays.device
        R, D = st.rays.shape[0], ren._d_comp
        names = st.mlp.names()
        grads = _zero_grads(st.mlp)
        merged = _merged_scatter_ok(ren, ctx.vol_shape)
        defer = [] if merged else None
        grad_cl = None
        if not merged:
            alloc = torch.empty if ren.scatter == "sorted" else torch.zeros
            grad_cl = alloc(ctx.vol_shape, device=dev, dtype=torch.float32)
        d_z = _pass_backward(ren, st, _zeros_like_or(d_rgb, (R, 3), dev), _zeros_like_or(d_emb, (R, D), dev),
                             d_dep, d_w, grads, grad_cl, want_dz=ctx.needs_input_grad[4], defer=defer,
                             d_coord_raw=d_coord_raw)
        d_vol = None
        if ctx.needs_input_grad[2]:
            if merged:
                d_vol = _finish_volume_grad(ren, st.rays, st.rps, defer, ctx.vol_shape, ctx.cl3d)
            else:
                d_vol = grad_cl.permute(0, 4, 1, 2, 3) if ctx.cl3d else ops.volume_to_channels_first(grad_cl)
        return (None, None, d_vol, None, d_z, None, *[grads[n] for n in names])
