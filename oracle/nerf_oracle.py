"""CPU/torch-eager restatement of the reference feature-NeRF render path.

TEST INFRASTRUCTURE ONLY -- never imported by the product package
(`real-robot-nerf-actor_b200/`).  Allowed users: tests/, tests/golden/make_golden.py,
__graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` leg.

Parity status: PINNED.  The reference ships no golden vectors or tests
(SURVEY.md section 4), so this restatement is pinned by outputs of the reference
itself: tests/golden/make_golden.py imports the unmodified reference through
oracle/ref_loader.py in the build container and commits its outputs under
tests/golden/; tests/test_oracle_golden.py checks every function here against them.

All arithmetic is PyTorch ATen fp32 (the reference's only arithmetic dependency;
installed torch 2.11.0, the reference pins no version).  Each function cites the
reference file:line it restates.  Functions are pure: parameters come in as a flat
dict with the reference's state_dict key suffixes (`lin_in.weight`, `blocks.0.fc_0.bias`,
`lin_z.1.weight`, ...), noise comes in as explicit tensors (the reference always draws
it, neural_rendering.py:172,194,200,218; "perturb off" = zeros).
"""
from __future__ import annotations

import math
from typing import Dict

import torch
import torch.nn.functional as F

Params = Dict[str, torch.Tensor]


# --------------------------------------------------------------------------- rays
def unproj_map(width, height, f, c=None, device="cpu"):
    """utils.py:444-474. Unit ray direction per pixel in the OpenGL camera frame, (H,W,3)."""
    if c is None:
        cx, cy = width * 0.5, height * 0.5
    else:
        c = c.squeeze()
        cx, cy = float(c[0]), float(c[1])
    f = torch.as_tensor(f, dtype=torch.float32).reshape(-1)
    fx, fy = (float(f[0]), float(f[0])) if f.numel() == 1 else (float(f[0]), float(f[1]))
    ys = torch.arange(height, dtype=torch.float32) - cy
    xs = torch.arange(width, dtype=torch.float32) - cx
    Y, X = torch.meshgrid(ys, xs, indexing="ij")
    X = X.to(device) / fx
    Y = Y.to(device) / fy
    d = torch.stack((X, -Y, -torch.ones_like(X)), dim=-1)
    return d / torch.norm(d, dim=-1).unsqueeze(-1)


def gen_rays(poses, width, height, focal, z_near, z_far, c=None):
    """utils.py:477-506. (B,4,4) cam->world poses -> (B,H,W,8) = [origin, dir, near, far]."""
    B = poses.shape[0]
    dev = poses.device
    cam = unproj_map(width, height, focal.squeeze() if torch.is_tensor(focal) else focal,
                     c=c, device=dev)
    cam = cam.unsqueeze(0).repeat(B, 1, 1, 1)
    origin = poses[:, None, None, :3, 3].expand(-1, height, width, -1)
    dirs = torch.matmul(poses[:, None, None, :3, :3], cam.unsqueeze(-1))[..., 0]
    near = torch.full((B, height, width, 1), float(z_near), dtype=torch.float32, device=dev)
    far = torch.full((B, height, width, 1), float(z_far), dtype=torch.float32, device=dev)
    return torch.cat((origin, dirs, near, far), dim=-1)


# ----------------------------------------------------------------------- sampling
def sample_coarse(rays, n_coarse, jitter=None, lindisp=False):
    """neural_rendering.py:159-176. rays (R,8) -> z (R,Kc). jitter (R,Kc) in [0,1) or None=0."""
    near, far = rays[:, -2:-1], rays[:, -1:]
    step = 1.0 / n_coarse
    R = rays.shape[0]
    z_steps = torch.linspace(0, 1 - step, n_coarse, device=rays.device)
    z_steps = z_steps.unsqueeze(0).repeat(R, 1)
    if jitter is not None:
        z_steps = z_steps + jitter * step
    if not lindisp:
        return near * (1 - z_steps) + far * z_steps
    return 1 / (1 / near * (1 - z_steps) + 1 / far * z_steps)


def fine_cdf(weights):
    """neural_rendering.py:189-192. weights (R,Kc) -> cdf (R,Kc+1) with leading 0."""
    w = weights.detach() + 1e-5
    pdf = w / torch.sum(w, -1, keepdim=True)
    cdf = torch.cumsum(pdf, -1)
    return torch.cat([torch.zeros_like(cdf[:, :1]), cdf], -1)


def sample_fine_from_cdf(rays, cdf, n_coarse, u, jitter, lindisp=False):
    """neural_rendering.py:194-207. cdf (R,Kc+1), u (R,Kf), jitter (R,Kf) -> (ind, z)."""
    inds = torch.searchsorted(cdf, u.contiguous(), right=True).float() - 1.0
    inds = torch.clamp_min(inds, 0.0)          # no upper clamp -- replicated
    z_steps = (inds + jitter) / n_coarse
    near, far = rays[:, -2:-1], rays[:, -1:]
    if not lindisp:
        z = near * (1 - z_steps) + far * z_steps
    else:
        z = 1 / (1 / near * (1 - z_steps) + 1 / far * z_steps)
    return inds, z


def sample_fine(rays, weights, n_coarse, u, jitter, lindisp=False):
    """neural_rendering.py:179-207."""
    return sample_fine_from_cdf(rays, fine_cdf(weights), n_coarse, u, jitter, lindisp)[1]


def sample_fine_depth(rays, depth, n_fine_depth, noise, depth_std):
    """neural_rendering.py:210-221. Samples around `depth`; NOT detached."""
    z = depth.unsqueeze(1).repeat((1, n_fine_depth))
    z = z + noise * depth_std
    return torch.max(torch.min(z, rays[:, -1:]), rays[:, -2:-1])


# -------------------------------------------------------------------- field model
def positional_encoding(x, num_freqs=6, freq_factor=1.5, include_input=True):
    """utils.py:521-557. (n,d) -> (n, d + 2*num_freqs*d); per frequency: sin(xyz) then cos(xyz)."""
    freqs = freq_factor * 2.0 ** torch.arange(0, num_freqs)
    _freqs = torch.repeat_interleave(freqs, 2).view(1, -1, 1).to(x.device)
    _phases = torch.zeros(2 * num_freqs)
    _phases[1::2] = math.pi * 0.5
    _phases = _phases.view(1, -1, 1).to(x.device)
    e = x.unsqueeze(1).repeat(1, num_freqs * 2, 1)
    e = torch.sin(torch.addcmul(_phases, e, _freqs)).view(x.shape[0], -1)
    return torch.cat((x, e), dim=-1) if include_input else e


def world_to_canonical(xyz, bounds):
    """models_embed.py:185-203 (@no_grad). (xyz - bb_min) / (bb_max - bb_min)."""
    with torch.no_grad():
        b = torch.as_tensor(bounds, dtype=torch.float32)
        bb_min = b[:3].reshape(1, 1, 3).to(xyz.device)
        bb_max = b[3:].reshape(1, 1, 3).to(xyz.device)
        out = xyz.clone()
        out -= bb_min
        out /= (bb_max - bb_min)
        return out


def trilinear_gather(voxel_feat, canon):
    """models_embed.py:259-277. voxel_feat (SB,C,D0,D1,D2), canon (SB,n,3) in [0,1] -> (SB,n,C).

    Axis quirk kept: canonical x indexes the LAST spatial dim, z the first.
    """
    g = canon.clone() * 2 - 1.0
    g = g.unsqueeze(1).unsqueeze(1)
    out = F.grid_sample(voxel_feat, g, align_corners=True, mode="bilinear")
    return out.squeeze(2).squeeze(2).permute(0, 2, 1)


def trilinear_gather_explicit(voxel_feat, canon):
    """Explicit-arithmetic restatement of ATen's CPU grid_sampler_3d, no autograd.

    Used to pin the CUDA gather bit-for-bit: per axis i = ((g+1)/2)*(S-1) with g = c*2-1,
    corner weights as three-factor products (x)(y)(z), accumulation in ATen's corner order with
    separately rounded multiply and add, corners outside [0,S-1] skipped.
    """
    SB, C, D0, D1, D2 = voxel_feat.shape
    g = canon * 2 - 1.0
    ix = ((g[..., 0] + 1) / 2) * (D2 - 1)
    iy = ((g[..., 1] + 1) / 2) * (D1 - 1)
    iz = ((g[..., 2] + 1) / 2) * (D0 - 1)
    x0, y0, z0 = torch.floor(ix), torch.floor(iy), torch.floor(iz)
    x1, y1, z1 = x0 + 1, y0 + 1, z0 + 1
    wx0, wx1 = x1 - ix, ix - x0
    wy0, wy1 = y1 - iy, iy - y0
    wz0, wz1 = z1 - iz, iz - z0
    vol = voxel_feat.permute(0, 2, 3, 4, 1)  # (SB,D0,D1,D2,C)
    out = torch.zeros(SB, canon.shape[1], C, dtype=voxel_feat.dtype, device=voxel_feat.device)
    bidx = torch.arange(SB, device=canon.device).view(SB, 1).expand(SB, canon.shape[1])
    corners = [  # ATen order: tnw, tne, tsw, tse, bnw, bne, bsw, bse
        (x0, y0, z0, wx0, wy0, wz0), (x1, y0, z0, wx1, wy0, wz0),
        (x0, y1, z0, wx0, wy1, wz0), (x1, y1, z0, wx1, wy1, wz0),
        (x0, y0, z1, wx0, wy0, wz1), (x1, y0, z1, wx1, wy0, wz1),
        (x0, y1, z1, wx0, wy1, wz1), (x1, y1, z1, wx1, wy1, wz1)]
    for cx, cy, cz, wx, wy, wz in corners:
        w = (wx * wy) * wz
        ok = (cx >= 0) & (cx <= D2 - 1) & (cy >= 0) & (cy <= D1 - 1) & (cz >= 0) & (cz <= D0 - 1)
        xi = cx.clamp(0, D2 - 1).long()
        yi = cy.clamp(0, D1 - 1).long()
        zi = cz.clamp(0, D0 - 1).long()
        v = vol[bidx, zi, yi, xi]                     # (SB,n,C)
        contrib = v * w.unsqueeze(-1)
        out = torch.where(ok.unsqueeze(-1), out + contrib, out)
    return out


def _linear(x, w, b, operand_dtype=None):
    """nn.Linear; `operand_dtype` rounds both GEMM operands (fp32 accumulate), SURVEY section 10."""
    if operand_dtype is not None:
        x = x.to(operand_dtype).to(torch.float32)
        w = w.to(operand_dtype).to(torch.float32)
    return F.linear(x, w, b)


def resnetfc(p: Params, zx, d_latent, n_blocks=5, combine_layer=3, operand_dtype=None, ret_last=False, beta=0.0,
             use_spade=False):
    """resnetfc.py:146-195 with ResnetBlockFC.forward :55-64 (no shortcut: size_in == size_out).

    zx (n, d_latent + d_in) -> (n, d_out).  combine_interleaved (utils.py:509-519) at
    blkid == combine_layer averages over a size-1 dim (num_views_per_obj = 1): identity.
    beta > 0: softplus activations (:43-46,:138-141); use_spade: x = scale_z(z) * x + lin_z(z) (:130-136,:184-186).
    """
    z, x = zx[..., :d_latent], zx[..., d_latent:]
    lin = lambda name, t: _linear(t, p[name + ".weight"], p[name + ".bias"], operand_dtype)
    act = (lambda t: F.softplus(t, beta=beta)) if beta > 0 else torch.relu
    x = lin("lin_in", x)
    for b in range(n_blocks):
        if d_latent > 0 and b < combine_layer:
            tz = lin(f"lin_z.{b}", z)
            x = lin(f"scale_z.{b}", z) * x + tz if use_spade else x + tz
        net = lin(f"blocks.{b}.fc_0", act(x))
        dx = lin(f"blocks.{b}.fc_1", act(net))
        x = x + dx
    out = lin("lin_out", act(x))
    return (out, x) if ret_last else out            # resnetfc.py:192-195: (out, x); ret_last_feat concatenates them


def field(p: Params, voxel_feat, xyz, viewdirs, bounds, code=(6, 1.5, True),
          n_blocks=5, combine_layer=3, operand_dtype=None, return_mlp_input=False, regress_coord=False,
          regress_attention=False, multi_scale_voxel_list=None, ret_last_feat=False, code_viewdirs=False, beta=0.0,
          use_spade=False):
    """models_embed.py:295-471 default branch. xyz, viewdirs (SB,n,3) -> (SB,n,4+D).

    mlp_input = [latent(C) | PE(xyz)(39) | viewdir(3)] (:366,:405); heads sigmoid(rgb),
    relu(sigma), raw embed (:444-466).  No gradient reaches xyz (world_to_canonical is no_grad).
    """
    SB, n, _ = xyz.shape
    canon = world_to_canonical(xyz, bounds)
    if code_viewdirs:                                # models_embed.py:355-372: [xyz | viewdir] through the encoding together
        zf = positional_encoding(torch.cat((canon.reshape(-1, 3), viewdirs.reshape(-1, 3)), dim=1), *code)
    else:                                            # :347-366: PE(xyz), then the raw view direction
        zf = positional_encoding(canon.reshape(-1, 3), *code)
        zf = torch.cat((zf, viewdirs.reshape(-1, 3)), dim=1)
    latent = trilinear_gather(voxel_feat, canon)
    if multi_scale_voxel_list:                       # models_embed.py:279-286: [*multi-scale, main]
        latent = torch.cat([trilinear_gather(v, canon) for v in multi_scale_voxel_list] + [latent], dim=-1)
    C = latent.shape[-1]
    mlp_input = torch.cat((latent.reshape(-1, C), zf), dim=-1)
    if return_mlp_input:
        return mlp_input
    out = resnetfc(p, mlp_input, C, n_blocks, combine_layer, operand_dtype, ret_last=ret_last_feat, beta=beta,
                   use_spade=use_spade)
    last = None
    if ret_last_feat:
        out, last = out
        last = last.reshape(SB, n, -1)
    out = out.reshape(-1, n, out.shape[-1])
    # models_embed.py:444-466: optional coordinate head (last 3, or [-9:-6] with attention: the RESIDUAL to the
    # canonical point is returned) and attention head (last 6)
    n_att = 6 if regress_attention else 0
    n_tail = n_att + (3 if regress_coord else 0)
    parts = [torch.sigmoid(out[..., :3]), torch.relu(out[..., 3:4]), out[..., 4:out.shape[-1] - n_tail]]
    if regress_coord:
        parts.append(out[..., out.shape[-1] - n_tail:out.shape[-1] - n_att] - canon)
    if regress_attention:
        parts.append(out[..., out.shape[-1] - 6:])
    out = torch.cat(parts, -1)
    return (out.reshape(SB, n, -1), last) if ret_last_feat else out.reshape(SB, n, -1)


# ------------------------------------------------------------------- compositing
def composite_weights(sigmas, z_samp, far, sigma_noise=None):
    """neural_rendering.py:239-243,336-346. sigmas, z (R,K); far (R,1) -> weights (R,K).
    sigma_noise: the training-time `randn_like(sigmas) * noise_std` of :336-337 (already scaled), or None."""
    deltas = torch.cat([z_samp[:, 1:] - z_samp[:, :-1], far - z_samp[:, -1:]], -1)
    if sigma_noise is not None:
        sigmas = sigmas + sigma_noise
    alphas = 1 - torch.exp(-deltas * torch.relu(sigmas))
    shifted = torch.cat([torch.ones_like(alphas[:, :1]), 1 - alphas + 1e-10], -1)
    T = torch.cumprod(shifted, -1)
    return alphas * T[:, :-1]


def composite_from_field(out, z_samp, far, white_bkgd=False, sigma_noise=None, regress_coord=False,
                         regress_attention=False, last_feat=None):
    """neural_rendering.py:316-359,383-395. out (R,K,4+D[+3][+6]) -> weights, rgb, embed, [coord], [attention], depth.
    The attention head is alpha-composited like the embedding (:353-354), the coordinate head is the plain MEAN over
    the samples (:356-357)."""
    n_att = 6 if regress_attention else 0
    n_tail = n_att + (3 if regress_coord else 0)
    last = out.shape[-1]
    w = composite_weights(out[..., 3], z_samp, far, sigma_noise)
    rgb = torch.sum(w.unsqueeze(-1) * out[..., :3], -2)
    embeds = out[..., 4:last - n_tail] if last_feat is None else last_feat      # neural_rendering.py:332-334
    embed = torch.sum(w.unsqueeze(-1) * embeds, -2)
    depth = torch.sum(w * z_samp, -1)
    if white_bkgd:
        rgb = rgb + 1 - w.sum(dim=1).unsqueeze(-1)
    res = [w, rgb, embed]
    if regress_coord:
        res.append(torch.mean(out[..., last - n_tail:last - n_att], -2))
    if regress_attention:
        res.append(torch.sum(w.unsqueeze(-1) * out[..., last - 6:], -2))
    return (*res, depth)


def composite(p, voxel_feat, rays, z_samp, sb, bounds, eval_batch_size=4096, white_bkgd=False, sigma_noise=None,
              **fkw):
    """neural_rendering.py:224-395. rays (R,8), z (R,K), sb scenes -> weights, rgb, embed, depth.

    Points o + z*d (:246), viewdirs = d broadcast (:276); the field is evaluated in chunks of
    (eval_batch_size-1)//sb+1 points per scene (:267,:273) -- chunking does not change results.
    """
    R, K = z_samp.shape
    pts = rays[:, None, :3] + z_samp.unsqueeze(2) * rays[:, None, 3:6]
    pts = pts.reshape(sb, -1, 3)
    dirs = rays[:, None, 3:6].expand(-1, K, -1).reshape(sb, -1, 3)
    chunk = (eval_batch_size - 1) // sb + 1
    vals = [field(p, voxel_feat, a, b, bounds, **fkw)
            for a, b in zip(torch.split(pts, chunk, dim=1), torch.split(dirs, chunk, dim=1))]
    last = None
    if fkw.get("ret_last_feat", False):
        last = torch.cat([v[1] for v in vals], dim=1)
        last = last.reshape(R, K, -1)
        vals = [v[0] for v in vals]
    out = torch.cat(vals, dim=1).reshape(R, K, -1)
    return composite_from_field(out, z_samp, rays[:, -1:], white_bkgd, sigma_noise, fkw.get("regress_coord", False),
                                fkw.get("regress_attention", False), last)


def forward_nerf(p, voxel_feat, rays, bounds, n_coarse, n_fine, n_fine_depth=0, noise=None,
                 lindisp=False, depth_std=0.001, eval_batch_size=4096, white_bkgd=False, noise_std=0.0, **fkw):
    """neural_rendering.py:435-471. rays (SB,B,8) -> dict(coarse=..., fine=..., z_coarse, z_fine).

    noise: dict with optional 'coarse' (R,Kc), 'u' (R,Kf-Kfd), 'fine' (R,Kf-Kfd), 'depth' (R,Kfd);
    missing entries mean zeros (perturb off).  noise_std > 0 (training mode of :336-337): 'sigma_c' (R,Kc) and
    'sigma_f' (R,Kc+Kf) are the standard-normal draws added, times noise_std, to the densities of each pass.
    """
    noise = noise or {}
    SB = rays.shape[0]
    r = rays.reshape(-1, 8)
    R = r.shape[0]
    z_c = sample_coarse(r, n_coarse, noise.get("coarse"), lindisp)
    sn = lambda k: noise[k] * noise_std if noise_std > 0 and k in noise else None
    has_c, has_a = fkw.get("regress_coord", False), fkw.get("regress_attention", False)

    def fmt(t):                                   # neural_rendering.py:398-426
        t = list(t)
        d = dict(rgb=t[1].reshape(SB, -1, 3), embed=t[2].reshape(SB, -1, t[2].shape[-1]),
                 depth=t[-1].reshape(SB, -1), weights=t[0].reshape(SB, -1, t[0].shape[-1]))
        rest = t[3:-1]
        if has_c:
            d["coord"] = rest.pop(0).reshape(SB, -1, 3)
        if has_a:
            d["attention"] = rest.pop(0).reshape(SB, -1, 6)
        return d
    comp_c = composite(p, voxel_feat, r, z_c, SB, bounds, eval_batch_size, white_bkgd, sn("sigma_c"), **fkw)
    wc, dep_c = comp_c[0], comp_c[-1]
    res = dict(coarse=fmt(comp_c), z_coarse=z_c)
    if n_fine > 0:
        samps = [z_c]
        kf = n_fine - n_fine_depth
        if kf > 0:
            zeros = torch.zeros(R, kf, device=r.device)
            samps.append(sample_fine(r, wc.detach(), n_coarse, noise.get("u", zeros),
                                     noise.get("fine", zeros), lindisp))
        if n_fine_depth > 0:
            nz = noise.get("depth", torch.zeros(R, n_fine_depth, device=r.device))
            samps.append(sample_fine_depth(r, dep_c, n_fine_depth, nz, depth_std))
        z_all, _ = torch.sort(torch.cat(samps, dim=-1), dim=-1)
        res["fine"] = fmt(composite(p, voxel_feat, r, z_all, SB, bounds, eval_batch_size, white_bkgd, sn("sigma_f"),
                                    **fkw))
        res["z_fine"] = z_all
    return res


def rendering_loss(outputs, gt_rgb, gt_embed, gt_depth=None, lambda_embed=0.01, lambda_depth=0.0,
                   z_far=4.0):
    """neural_rendering.py:653-707 on already-subsampled targets (SB,chunk,.)."""
    c, f = outputs["coarse"], outputs["fine"]
    l_rgb_c, l_rgb_f = F.mse_loss(c["rgb"], gt_rgb), F.mse_loss(f["rgb"], gt_rgb)
    l_e_c = lambda_embed * F.mse_loss(c["embed"], gt_embed)
    l_e_f = lambda_embed * F.mse_loss(f["embed"], gt_embed)
    loss = l_rgb_c + l_rgb_f + l_e_c + l_e_f
    l_d_c = l_d_f = torch.tensor(0.)
    if gt_depth is not None:
        m = gt_depth < z_far
        l_d_c = lambda_depth * F.mse_loss(gt_depth[m], c["depth"][m])
        l_d_f = lambda_depth * F.mse_loss(gt_depth[m], f["depth"][m])
        loss = loss + l_d_c + l_d_f
    mse = torch.mean((f["rgb"] - gt_rgb) ** 2)
    psnr = 20 * torch.log10(1.0 / torch.sqrt(mse))
    return dict(loss=loss, loss_rgb_coarse=l_rgb_c, loss_rgb_fine=l_rgb_f, loss_embed_coarse=l_e_c,
                loss_embed_fine=l_e_f, loss_depth_coarse=l_d_c, loss_depth_fine=l_d_f, psnr=psnr)


# -------------------------------------------------------------------- parameters
def init_params(d_in=42, d_latent=128, d_hidden=512, d_out=388, n_blocks=5, combine_layer=3,
                seed=0, randomize_fc1=True, device="cpu", use_spade=False) -> Params:
    """Reference init (resnetfc.py:38-41,92-98,126-128): kaiming-normal fan-in, zero biases,
    fc_1.weight = 0.  `randomize_fc1` overwrites fc_1.weight ~ N(0, 2/d_hidden) so the blocks
    are not identities.  Drawn from a CPU generator: identical on every machine.
    """
    g = torch.Generator().manual_seed(seed)
    kaiming = lambda o, i: torch.randn(o, i, generator=g) * math.sqrt(2.0 / i)
    p = {"lin_in.weight": kaiming(d_hidden, d_in), "lin_in.bias": torch.zeros(d_hidden),
         "lin_out.weight": kaiming(d_out, d_hidden), "lin_out.bias": torch.zeros(d_out)}
    for b in range(n_blocks):
        p[f"blocks.{b}.fc_0.weight"] = kaiming(d_hidden, d_hidden)
        p[f"blocks.{b}.fc_0.bias"] = torch.zeros(d_hidden)
        p[f"blocks.{b}.fc_1.weight"] = (kaiming(d_hidden, d_hidden) if randomize_fc1
                                        else torch.zeros(d_hidden, d_hidden))
        p[f"blocks.{b}.fc_1.bias"] = torch.zeros(d_hidden)
    for b in range(min(combine_layer, n_blocks)):
        p[f"lin_z.{b}.weight"] = kaiming(d_hidden, d_latent)
        p[f"lin_z.{b}.bias"] = torch.zeros(d_hidden)
    if use_spade:                                    # resnetfc.py:130-136 (drawn last: the other tensors keep their values)
        for b in range(min(combine_layer, n_blocks)):
            p[f"scale_z.{b}.weight"] = kaiming(d_hidden, d_latent)
            p[f"scale_z.{b}.bias"] = torch.zeros(d_hidden)
    return {k: v.to(device) for k, v in p.items()}


def params_from_state_dict(sd, prefix="nerf_model.mlp_coarse.") -> Params:
    return {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}


# ------------------------------------------------------------------- extraction
def extract_radience(p, voxel_feat, rays, z_samp, sb, bounds, **fkw):
    """featurenerf_robo/featurenerf/src/render/nerf_embed.py:432-516: the field at every sample, not composited.
    rays (R,8), z (R,K) -> points (sb, R'K, 3), rgbs (sb, R'K, 3), sigmas (sb, R'K), embeds (sb, R'K, D)."""
    R, K = z_samp.shape
    pts = (rays[:, None, :3] + z_samp.unsqueeze(2) * rays[:, None, 3:6]).reshape(sb, -1, 3)
    dirs = rays[:, None, 3:6].expand(-1, K, -1).reshape(sb, -1, 3)
    out = field(p, voxel_feat, pts, dirs, bounds, **fkw)
    return pts, out[..., :3], out[..., 3], out[..., 4:]


def point_cloud_masks(rgbs, sigmas, lower_bound=50000, upper_bound=70000, max_iters=1000):
    """train_nerfact_multi_kitchen.py:985-1013: brighter-than-average AND sigma > step * max, `step` walked from 0.1 by
    -0.01 / +0.02 until the survivor count is inside [lower_bound, upper_bound]."""
    mask2 = rgbs.sum(-1) > rgbs.sum(-1).mean()
    step, num, it = 0.1, 0, 0
    mask1 = sigmas > sigmas.max() * step
    while num < lower_bound or num > upper_bound:
        mask1 = sigmas > (sigmas.max() * step)
        num = int((mask1 & mask2).sum())
        if num < lower_bound:
            step -= 0.01
        elif num > upper_bound:
            step += 0.02
        else:
            break
        it += 1
        if it >= max_iters:
            break
    return mask1 & mask2, step
