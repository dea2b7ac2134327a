"""Radiance / point-cloud extraction on top of the field kernels.

The ancestor renderer can return the per-sample field values instead of compositing them
(featurenerf_robo/featurenerf/src/render/nerf_embed.py:432-516 `extract_radience`, switched on by
`forward(..., extract_radience=True)` at :338-342), and the training script turns them into a coloured, feature-carrying
point cloud by density / brightness masks (train_nerfact_multi_kitchen.py:849-1077 `extract_nerf_feat`, :985-1040).
Both are restated here over this repo's kernels: the field is evaluated by nrf_encode_points + the fused MLP (inference
variant: nothing saved), the masks are device-side torch ops.
"""
from __future__ import annotations

import torch

from . import ops


@torch.no_grad()
def extract_radience(ren, model, rays, z_samp, coarse=True, sb=0, ret_last_feat=False):
    """nerf_embed.py:432-516.  rays (B,8), z_samp (B,K) -> points (SB, B'K, 3) (or (BK, 3) when sb == 0),
    rgbs (..., 3) after the sigmoid, sigmas (...) after the ReLU, embeds (..., D) - or, with ret_last_feat, the MLP's
    last residual stream reshaped (B, K, d_hidden) as the ancestor does (:513-514)."""
    from .neural_rendering import _is_channels_last_3d
    model = model if model is not None else ren.nerf_model
    rays, z_samp = rays.contiguous(), z_samp.contiguous()
    B, K = z_samp.shape
    rps = B // max(int(sb), 1)
    mlp = model.mlp_coarse if coarse or model.mlp_fine is None else model.mlp_fine
    vol = model.voxel_feat
    last = None
    if ret_last_feat or ren._composed:
        from . import composed
        vols = list(model.multi_scale_voxel_list or []) + [vol]
        raw, last = composed.field_rows(ren, model, mlp, vols, rays, z_samp, rps)
        pts = rays[:, None, :3] + z_samp.unsqueeze(2) * rays[:, None, 3:6]
        pts = pts.reshape(-1, 3)
    else:
        h = mlp.handle(ren._prec)
        vol_cl = vol.permute(0, 2, 3, 4, 1) if _is_channels_last_3d(vol) else ops.volume_to_channels_last(vol)
        field_in, pts = ops.encode_points(rays, z_samp, rps, vol_cl, ren._bounds, ren._num_freqs, ren._freq_factor,
                                          ld_out=h.sizes.kin_pad, precision=h.precision, want_points=True)
        raw, _ = h.forward(field_in, keep_acts=False)
    D = ren._d_embed
    rgbs, sigmas = torch.sigmoid(raw[:, :3]), torch.relu(raw[:, 3])
    embeds = raw[:, 4:4 + D]                         # the ancestor keeps [4:-3] with regress_coord, i.e. the embedding
    if sb > 0:
        pts, rgbs, sigmas, embeds = (pts.reshape(sb, -1, 3), rgbs.reshape(sb, -1, 3), sigmas.reshape(sb, -1),
                                     embeds.reshape(sb, -1, D))
    if ret_last_feat:
        embeds = last.reshape(B, K, -1)
    return pts, rgbs, sigmas, embeds


@torch.no_grad()
def extract_point_cloud(pnts, rgbs, sigmas, embeds, lower_bound=50000, upper_bound=70000, world_to_base=None,
                        max_iters=1000):
    """train_nerfact_multi_kitchen.py:985-1040: keep the samples that are dense (sigma above a fraction `step` of the
    maximum, `step` walked by -0.01 / +0.02 from 0.1 until between `lower_bound` and `upper_bound` samples survive) and
    brighter than average (rgb sum above its mean); optionally map the points into the robot base frame.
    -> (points (n,3), rgbs (n,3), embeds (n,D), step).  `max_iters` bounds the reference's unbounded search loop."""
    mask2 = rgbs.sum(-1) > rgbs.sum(-1).mean()
    step, num_current, it = 0.1, 0, 0
    smax = sigmas.max()
    mask1 = sigmas > smax * step
    while num_current < lower_bound or num_current > upper_bound:
        mask1 = sigmas > (smax * step)
        num_current = int((mask1 & mask2).sum())
        if num_current < lower_bound:
            step -= 0.01
        elif num_current > upper_bound:
            step += 0.02
        else:
            break
        it += 1
        if it >= max_iters:
            break
    mask = mask1 & mask2
    p, c, e = pnts[mask], rgbs[mask], embeds.reshape(*mask.shape, -1)[mask]
    if world_to_base is not None:
        w2b = torch.as_tensor(world_to_base, dtype=p.dtype, device=p.device)
        p = p @ w2b[:3, :3].T + w2b[:3, 3]
    return p, c, e, step
