"""The optional model branches that change the field MLP's INPUT or what is composited:

  * `use_multi_scale_voxel` (models_embed.py:279-286; nerfact.conf:30-32, d_multi_scale_latent = 266): the latent is the
    concatenation of trilinear gathers from several volumes of different resolution and channel count
    ([*multi_scale_voxel_list, voxel_feat]);
  * `ret_last_feat` (neural_rendering.py:285-293,332-334; resnetfc.py:192-195): the MLP's last residual stream x_nb
    (d_hidden channels) is alpha-composited in place of the embedding head;
  * `use_code_viewdirs` (models_embed.py:86-95,:355-372): the view direction goes through the positional encoding
    together with the point (nrf_encode_points with out_bf16 | 0x200: tail of 78 columns, MLP d_in = 78).

  * softplus activations (`mlp.beta > 0`) and SPADE modulation (`mlp.use_spade`), resnetfc.py:43-46,130-141,184-186:
    the MLP layer by layer (`mlp_general`: one LinearFn node per layer on nrf_gemm / nrf_wgrad, fp32).

All are off in nerfact.conf.  They run on the same kernels as the default path (nrf_encode_points per volume, the
field MLP, nrf_composite_*, nrf_scatter_volume_grad_sorted per volume), composed at the torch level as three autograd
nodes per pass instead of one per forward_nerf - so the general shapes (266 latent channels, volumes of 10 channels,
an extra 512-wide output) need no kernel of their own.  The MLP runs in the fp32 parity mode here (d_latent = 266 is
not a whole number of tensor-core k-blocks, and x_nb is only kept by the layer-by-layer chain).
"""
from __future__ import annotations

import torch

from . import ops
from ._lib import NRF_PREC_FP32


class GatherFn(torch.autograd.Function):
    """One volume's contribution to the field input at the samples of `z`: rows [latent (C) | PE (39) | viewdir (3)],
    fp32 (models_embed.py:259-277 grid_sample + :185-203 world_to_canonical + utils.py:545-557).  Backward: the
    atomics-free sorted scatter into that volume; no gradient reaches rays / z (world_to_canonical is @no_grad)."""

    @staticmethod
    def forward(ctx, vol, rays, z, rps, bounds, num_freqs, freq_factor, code_viewdirs=False):
        SB, C = vol.shape[:2]
        Cp = (C + 3) // 4 * 4                      # the kernels move whole float4s: pad the channels with zeros
        v = vol.detach()
        if Cp != C:
            v = torch.cat([v, v.new_zeros(SB, Cp - C, *v.shape[2:])], 1)
        vol_cl = ops.volume_to_channels_last(v.contiguous())
        tail = 6 + (12 if code_viewdirs else 6) * num_freqs
        rows = ops.encode_points(rays, z, rps, vol_cl, bounds, num_freqs, freq_factor, precision=NRF_PREC_FP32,
                                 code_viewdirs=code_viewdirs)
        ctx.save_for_backward(rays, z)
        ctx.meta = (rps, bounds, tuple(vol_cl.shape), C, Cp)
        out = torch.cat([rows[:, :C], rows[:, Cp:Cp + tail]], 1)
        return out

    @staticmethod
    def backward(ctx, d_rows):
        rays, z = ctx.saved_tensors
        rps, bounds, shape_cl, C, Cp = ctx.meta
        if not ctx.needs_input_grad[0]:
            return (None,) * 8
        d_lat = d_rows.new_zeros(d_rows.shape[0], Cp)
        d_lat[:, :C] = d_rows[:, :C]
        g = torch.empty(shape_cl, device=d_rows.device, dtype=torch.float32)
        ops.scatter_volume_grad_sorted(rays, z, rps, d_lat, g, bounds)
        return (ops.volume_to_channels_first(g)[:, :C], None, None, None, None, None, None, None)


class MlpLastFn(torch.autograd.Function):
    """The field MLP through the layer-by-layer chain, returning (raw outputs (N, d_out), x_nb (N, d_hidden)) - the two
    return values of ResnetFC.forward (resnetfc.py:192-195).  zx (N, d_latent + d_in) fp32."""

    @staticmethod
    def forward(ctx, h, zx, *params):
        N, width = zx.shape
        fin = torch.zeros(N, h.sizes.kin_pad, device=zx.device, dtype=ops.act_dtype(h.precision))
        fin[:, :width] = zx.detach().to(fin.dtype)
        out, acts = h.forward(fin, layered=True)
        last = h.last_feat(acts, N)
        ctx.h, ctx.fin, ctx.acts, ctx.width = h, fin, acts, width
        return out[:, :h.dims[3]].contiguous(), last.to(torch.float32).clone()

    @staticmethod
    def backward(ctx, d_out, d_last):
        from .neural_rendering import _zero_grads
        h = ctx.h
        if ctx.acts is None:
            raise RuntimeError("field MLP: backward a second time (retain_graph is not supported)")
        fin, acts = ctx.fin, ctx.acts
        ctx.fin = ctx.acts = None
        N = fin.shape[0]
        gd = ops.grad_dtype(h.precision)
        d_field = torch.zeros(N, h.sizes.dout_pad, device=fin.device, dtype=gd)
        if d_out is not None:
            d_field[:, :d_out.shape[1]] = d_out.to(gd)
        dl = d_last.to(gd).contiguous() if d_last is not None else None
        grads = _zero_grads(h)
        dlat = h.backward(fin, acts, d_field, grads, layered=True, d_last=dl)
        dzx = torch.zeros(N, ctx.width, device=fin.device, dtype=torch.float32)
        dzx[:, :dlat.shape[1]] = dlat
        return (None, dzx, *[grads[n] for n in h.names()])


class LinearFn(torch.autograd.Function):
    """y = x . W^T + b on the fp32 GEMM kernels: nrf_gemm forward and data gradient, nrf_wgrad for dW / db - one layer of
    the field MLP as an autograd node of its own, for the MLP variants whose glue between the layers is not the fused
    chain's (mlp_general)."""

    @staticmethod
    def forward(ctx, x, w, b):
        x, w = x.detach().contiguous(), w.detach().contiguous()
        out = torch.empty(x.shape[0], w.shape[0], device=x.device, dtype=torch.float32)
        ops.gemm(x, w, bias=b.detach().contiguous(), out_f32=out, precision=NRF_PREC_FP32)
        ctx.save_for_backward(x, w)
        return out

    @staticmethod
    def backward(ctx, g):
        x, w = ctx.saved_tensors
        g = g.contiguous()
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            ops.gemm(g, w.t().contiguous(), out_f32=dx, precision=NRF_PREC_FP32)
        if ctx.needs_input_grad[1] or ctx.needs_input_grad[2]:
            dw = torch.zeros_like(w)
            db = torch.zeros(w.shape[0], device=w.device, dtype=torch.float32)
            ops.wgrad(g, x, dw, db, precision=NRF_PREC_FP32)
        return dx, dw, db


def mlp_general(mlp, zx):
    """ResnetFC.forward (resnetfc.py:146-195) with the two options the fused chain is not built for: softplus
    activations (beta > 0, :43-46,:138-141) and SPADE modulation of the residual stream by the latent (use_spade,
    :130-136,:184-186: x = scale_z(z) * x + lin_z(z)).  zx (N, d_latent + d_in) fp32 -> (raw outputs (N, d_out), x_nb (N,
    d_hidden)).  Every linear layer is a LinearFn node; the activations and the modulation are elementwise torch ops."""
    if zx.shape[1] != mlp.d_latent + mlp.d_in:
        raise RuntimeError(f"field MLP input has {zx.shape[1]} columns, expected d_latent + d_in = "
                           f"{mlp.d_latent} + {mlp.d_in}")
    act = (lambda t: torch.nn.functional.softplus(t, beta=mlp.beta)) if mlp.beta > 0 else torch.relu
    lin = lambda layer, t: LinearFn.apply(t, layer.weight, layer.bias)
    zx = zx.to(torch.float32)
    z, x = zx[:, :mlp.d_latent], zx[:, mlp.d_latent:]
    x = lin(mlp.lin_in, x)
    for b in range(mlp.n_blocks):
        if b < mlp.n_lin_z:                                   # resnetfc.py:181-188 (combine_interleaved is the identity
            tz = lin(mlp.lin_z[b], z)                         # for one view, :176-179)
            x = lin(mlp.scale_z[b], z) * x + tz if mlp.use_spade else x + tz
        blk = mlp.blocks[b]                                   # resnetfc.py:55-64, identity shortcut
        net = lin(blk.fc_0, act(x))
        x = x + lin(blk.fc_1, act(net))
    return lin(mlp.lin_out, act(x)), x


class CompositeFieldFn(torch.autograd.Function):
    """Alpha compositing (neural_rendering.py:339-359) of given RAW field rows (N, ld) [rgb | sigma | D channels]:
    -> weights (R,K), rgb (R,3), composited channels (R,D), depth (R).  Differentiable w.r.t. the rows and z."""

    @staticmethod
    def forward(ctx, field, z, rays, D, white_bkgd, sigma_noise):
        field = field.detach().contiguous()
        outs = ops.composite_fwd(field, z.detach(), rays, D, white_bkgd, sigma_noise=sigma_noise)
        ctx.save_for_backward(field, z.detach(), rays)
        ctx.meta = (D, white_bkgd, sigma_noise)
        return outs

    @staticmethod
    def backward(ctx, d_w, d_rgb, d_emb, d_dep):
        field, z, rays = ctx.saved_tensors
        D, white_bkgd, sigma_noise = ctx.meta
        R = z.shape[0]
        zeros = lambda t, shape: t.contiguous() if t is not None else torch.zeros(shape, device=z.device)
        want_dz = ctx.needs_input_grad[1]
        res = ops.composite_bwd(field, z, rays, D, zeros(d_rgb, (R, 3)), zeros(d_emb, (R, D)), d_dep, d_w,
                                ldg=field.shape[1], precision=NRF_PREC_FP32, white_bkgd=white_bkgd, want_dz=want_dz,
                                sigma_noise=sigma_noise)
        d_field, d_z = res if want_dz else (res, None)
        return d_field, d_z, None, None, None, None


def field_rows(ren, model, mlp, vols, rays, z, rps):
    """[*multi-scale latents | main latent | PE | viewdir] -> (raw MLP outputs (N, d_out), x_nb (N, d_hidden))."""
    nf, ff = model.code.num_freqs, float(model.code.freq_factor)
    cv = bool(getattr(model, "use_code_viewdirs", False))
    tail = 6 + (12 if cv else 6) * nf
    rows = [GatherFn.apply(v, rays, z, rps, ren._bounds, nf, ff, cv) for v in vols]
    zx = torch.cat([r[:, :r.shape[1] - tail] for r in rows] + [rows[-1][:, -tail:]], 1)
    if zx.shape[1] != model.d_latent + model.d_in:
        raise RuntimeError(f"multi-scale latent has {zx.shape[1] - model.d_in} channels, the model expects d_latent = "
                           f"{model.d_latent} (d_multi_scale_latent)")
    if mlp.general:                                           # softplus / SPADE: layer by layer at the torch level
        return mlp_general(mlp, zx)
    h = mlp.handle(NRF_PREC_FP32)
    ps = [mlp.param_dict()[n] for n in h.names()]
    return MlpLastFn.apply(h, zx, *ps)


def composite_pass(ren, model, rays, z, coarse, sb, sigma_noise=None):
    """One composite pass (neural_rendering.py:224-395) -> the reference's tuple (weights, rgb, embed, [coord],
    [attention], depth), every stage an autograd node of its own."""
    rps = rays.shape[0] // max(int(sb), 1)
    vols = list(model.multi_scale_voxel_list or []) + [model.voxel_feat]
    mlp = model.mlp_coarse if coarse or model.mlp_fine is None else model.mlp_fine
    out, last = field_rows(ren, model, mlp, vols, rays, z, rps)
    D = ren._d_embed
    extras = out[:, 4 + D:]                                           # coord (3) / attention (6) columns, if any
    emb = last if ren.ret_last_feat else out[:, 4:4 + D]               # neural_rendering.py:332-334
    d_comp = (emb.shape[1] + extras.shape[1] + 3) // 4 * 4
    pad = out.new_zeros(out.shape[0], d_comp - emb.shape[1] - extras.shape[1])
    field = torch.cat([out[:, :4], emb, extras, pad], 1)
    w, rgb, emb_all, dep = CompositeFieldFn.apply(field, z, rays, d_comp, ren.white_bkgd, sigma_noise)
    R, K = z.shape
    o = 4 + emb.shape[1]
    coord_raw = field.view(R, K, -1)[:, :, o:o + 3] if ren.regress_coord else None
    return ren._split_heads(w, rgb, emb_all, dep, coord_raw, rays, z, d_embed=emb.shape[1])


def forward_nerf(ren, rays_flat, sb, noise, want_weights):
    """neural_rendering.py:435-471 over the composed passes (sampling kernels as in the default path; the depth-guided
    samples and the sort stay torch ops so that autograd carries the fine -> coarse gradient path)."""
    from .neural_rendering import AttrDict, _sigma_noise
    model = ren.nerf_model
    R = rays_flat.shape[0]
    Kc, Kf, Kfd = ren.n_coarse, ren.n_fine, ren.n_fine_depth
    z_c = ops.sample_coarse(rays_flat, Kc, noise.get("coarse"), ren.lindisp)
    coarse = composite_pass(ren, model, rays_flat, z_c, True, sb, _sigma_noise(ren, noise, "sigma_c", R, Kc, rays_flat.device))
    outputs = AttrDict(coarse=ren._format_outputs(coarse, sb, want_weights))
    outputs.coarse.z = z_c
    if ren.using_fine:
        samps = [z_c]
        if Kf - Kfd > 0:
            samps.append(ops.sample_fine(rays_flat, coarse[0].detach(), Kc, noise["u"], noise.get("fine"), ren.lindisp))
        if Kfd > 0:
            samps.append(ren.sample_fine_depth(rays_flat, coarse[-1], noise.get("depth")))
        z_all, _ = torch.sort(torch.cat(samps, -1), -1)
        fine = composite_pass(ren, model, rays_flat, z_all, False, sb,
                              _sigma_noise(ren, noise, "sigma_f", R, Kc + Kf, rays_flat.device))
        outputs.fine = ren._format_outputs(fine, sb, want_weights)
        outputs.fine.z = z_all.detach()
    return outputs
