"""Generates tests/golden/*.npz by running the UNMODIFIED reference (build container only).

    python tests/golden/make_golden.py

The reference (/root/reference) ships no tests or golden vectors, so these fixtures -- outputs of
the reference's own NeuralRenderer on seeded synthetic inputs -- are what pins parity:
tests/test_oracle_golden.py checks oracle/nerf_oracle.py against them on CPU, and the `-m gpu`
tests check the CUDA path against the same files on the B200 box (where /root/reference is absent).

Cases
  small_kfd0 / small_kfd4 : tiny dims (S=12, C=16, D=24, hidden 64); every input is stored.
  small_noise_wb          : same dims with the optional branches noise_std=0.5 (training-time density noise,
                            neural_rendering.py:336-337), white_bkgd (:383-386) and lindisp (:175-176,205-206).
  full_s32                : BASELINE dims (C=128, D=384, hidden 512, Kc=Kf=64), 32^3 volume;
                            inputs are regenerated from seeds (synthetic.py / init_params), only
                            outputs and gradient projections are stored.
  small_heads             : regress_coord + regress_attention (3 + 6 extra outputs, d_out = 37): all outputs of
                            forward_nerf incl. coord / attention and the gradients of a probe loss over all of them.
  small_multiscale        : use_multi_scale_voxel (three volumes of 10 / 8 / 16 channels at 12^3 / 6^3 / 12^3) with
                            ret_last_feat (the MLP's last residual stream composited) and depth-guided samples.
  small_codeviewdirs      : use_code_viewdirs (view direction through the positional encoding, d_in = 78) with
                            normalize_z = True (a no-op under the hard-coded canon_xyz) and depth-guided samples.
  small_softplus_spade    : mlp.beta = 10 (softplus activations) with mlp.use_spade (latent-modulated residual stream).
  extract_small           : the ancestor renderer's extract_radience (nerf_embed.py:432-516, imported unmodified) over the
                            reference's field model: points / rgbs / sigmas / embeds at given sorted sample depths.
  raygen                  : gen_rays for 60x80 (focal 76.18187) and rows of 128x128 (focal 153).
  voxelize_small          : the reference's VoxelGrid.coords_to_bounding_voxel_grid (voxel_grid_real.py) on a seeded
                            clustered point cloud (inputs regenerated from the seed by synthetic.voxelizer_points).
"""
import importlib
import os
import sys
from unittest import mock

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_loader as L          # noqa: E402
from oracle import nerf_oracle as O         # noqa: E402

syn = importlib.import_module("real-robot-nerf-actor_b200.synthetic")


def load_params_into(renderer, params):
    sd = renderer.state_dict()
    for k, v in params.items():
        sd["nerf_model.mlp_coarse." + k].copy_(v)


def run_case(name, S, C, D, hidden, SB, n_rays, Kc, Kf, Kfd, H, W, focal, store_inputs, seed=0,
             perturb=True, noise_std=0.0, white_bkgd=False, lindisp=False):
    torch.manual_seed(seed)
    cfg = L.default_cfg(d_embed=D, d_latent=C, voxel_shape=S, image_width=W, image_height=H,
                        n_coarse=Kc, n_fine=Kf, n_fine_depth=Kfd, ray_chunk_size=n_rays,
                        mlp=dict(d_hidden=hidden), eval_batch_size=1024, noise_std=noise_std,
                        white_bkgd=white_bkgd, lindisp=lindisp)
    bounds = torch.tensor(syn.BOUNDS)
    ren = L.build_reference_renderer(cfg, bounds)
    params = O.init_params(d_in=42, d_latent=C, d_hidden=hidden, d_out=4 + D, seed=seed)
    # non-zero biases so that the bias path is exercised
    g = torch.Generator().manual_seed(77 + seed)
    for k in params:
        if k.endswith(".bias"):
            params[k] = 0.05 * torch.randn(params[k].shape, generator=g)
    load_params_into(ren, params)
    vol = syn.make_volume(SB, C, S, seed=seed).requires_grad_(True)
    poses = syn.arc_poses(SB)
    focal_t = torch.tensor(focal, dtype=torch.float32)
    idx = syn.pick_ray_indices(H * W, n_rays, seed=seed)
    R = SB * n_rays
    noise = syn.make_noise(R, Kc, Kf - Kfd, seed=seed, perturb=perturb)
    gd = torch.Generator().manual_seed(5000 + seed)
    noise_depth = torch.randn(R, Kfd, generator=gd) if Kfd > 0 else None
    gt_rgb_img = torch.rand(SB, H, W, 3, generator=gd)
    gt_embed_img = torch.randn(SB, H, W, D, generator=gd)

    if noise_std > 0:          # the reference draws randn_like(sigmas) inside each composite (training mode)
        noise["sigma_c"] = torch.randn(R, Kc, generator=gd)
        noise["sigma_f"] = torch.randn(R, Kc + Kf, generator=gd)
    draws = [noise.get("coarse")]
    if noise_std > 0:
        draws += [noise["sigma_c"]]
    if Kf - Kfd > 0:
        draws += [noise["u"], noise.get("fine")]
    if Kfd > 0:
        draws += [noise_depth]
    if noise_std > 0:
        draws += [noise["sigma_f"]]
    ren.train()
    with L.inject_noise(draws), \
            mock.patch.object(torch, "randint", lambda *a, **k: idx.clone()):
        out = ren(multi_scale_voxel_list=None, voxel_density=None, language=None, voxel_feat=vol,
                  voxel_poses=poses, focal=focal_t, gt_rgb=gt_rgb_img, gt_depth=None,
                  gt_pose=poses, c=None, lang_goal=None, gt_embed=gt_embed_img)
    out["loss"].backward()
    grads = {k: p.grad.clone() for k, p in ren.nerf_model.mlp_coarse.named_parameters()}
    vgrad = vol.grad.clone()

    # second, no-grad pass for the intermediate tensors (same noise; still in training mode: noise_std applies)
    rays_full = ren_rays = None
    with torch.no_grad():
        U = sys.modules["_nrf_reference_utils"]
        rays_full = U.gen_rays(poses, W, H, focal_t, cfg.z_near, cfg.z_far, c=None)
        rays = rays_full.reshape(SB, H * W, 8)[:, idx]
        ren.encode(None, None, None, vol.detach(), poses, focal_t, None)
        with L.inject_noise(list(draws)):
            o = ren.forward_nerf(rays, want_weights=True)
        # z samples, by replaying the samplers with the same noise
        r = rays.reshape(-1, 8)
        with L.inject_noise([noise.get("coarse")]):
            z_c = ren.sample_coarse(r)

    fx = {"meta": np.array([S, C, D, hidden, SB, n_rays, Kc, Kf, Kfd, H, W, seed]),
          "opts": np.array([noise_std, float(white_bkgd), float(lindisp)], dtype=np.float32),
          "focal": np.float32(focal), "idx": idx.numpy(), "rays": rays.numpy(), "z_coarse": z_c.numpy(),
          "loss": np.float32(out["loss"].item()),
          "loss_items": np.array([out[k] for k in ("loss_rgb_coarse", "loss_rgb_fine", "loss_embed_coarse",
                                                   "loss_embed_fine", "psnr")], dtype=np.float32)}
    for lvl in ("coarse", "fine"):
        for k in ("rgb", "embed", "depth", "weights"):
            fx[f"{lvl}_{k}"] = o[lvl][k].numpy()
    if noise_depth is not None:
        fx["noise_depth"] = noise_depth.numpy()
    gsign = torch.Generator().manual_seed(99)
    sign = (torch.randint(0, 2, (C,), generator=gsign) * 2 - 1).float()
    if store_inputs:
        fx["vol"] = vol.detach().numpy()
        fx["vgrad"] = vgrad.numpy()
        for k, v in params.items():
            fx["param." + k] = v.numpy()
        for k, v in grads.items():
            fx["grad." + k] = v.numpy()
        for k, v in noise.items():
            fx["noise_" + k] = v.numpy()
        fx["gt_rgb_img"] = gt_rgb_img.numpy()
        fx["gt_embed_img"] = gt_embed_img.numpy()
        fx["poses"] = poses.numpy()
    else:
        fx["vgrad_sum"] = vgrad.sum(1).numpy()
        fx["vgrad_sign"] = (vgrad * sign.view(1, C, 1, 1, 1)).sum(1).numpy()
        fx["vgrad_norm"] = np.float32(vgrad.norm().item())
        for k, v in grads.items():
            fx["gradnorm." + k] = np.float32(v.norm().item())
            if v.numel() <= 512 * 42 or k.endswith(".bias"):
                fx["grad." + k] = v.numpy()
            else:                      # 64 seeded rows of each big matrix
                rows = torch.randperm(v.shape[0], generator=torch.Generator().manual_seed(5))[:64]
                fx["gradrows." + k] = v[rows].numpy()
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **fx)
    print(name, "->", path, os.path.getsize(path) // 1024, "KiB", "loss", float(out["loss"]))


def run_heads_case(name="small_heads", S=12, C=16, D=24, hidden=64, SB=2, n_rays=40, Kc=16, Kf=16, H=16, W=16,
                   focal=19.0, seed=6):
    """regress_coord + regress_attention (models_embed.py:63-68,447-461, neural_rendering.py:318-329,353-357,388-395):
    d_out = 4 + D + 3 + 6 = 37 outputs; forward_nerf of the reference + a loss over ALL of its outputs (the reference's
    own loss ignores coord / attention), gradients into the volume and the MLP."""
    torch.manual_seed(seed)
    cfg = L.default_cfg(d_embed=D, d_latent=C, voxel_shape=S, image_width=W, image_height=H, n_coarse=Kc, n_fine=Kf,
                        n_fine_depth=0, ray_chunk_size=n_rays, mlp=dict(d_hidden=hidden), eval_batch_size=1024,
                        regress_coord=True, regress_attention=True)
    ren = L.build_reference_renderer(cfg, torch.tensor(syn.BOUNDS))
    d_out = 4 + D + 9
    params = O.init_params(d_in=42, d_latent=C, d_hidden=hidden, d_out=d_out, seed=seed)
    g = torch.Generator().manual_seed(77 + seed)
    for k in params:
        if k.endswith(".bias"):
            params[k] = 0.05 * torch.randn(params[k].shape, generator=g)
    load_params_into(ren, params)
    vol = syn.make_volume(SB, C, S, seed=seed).requires_grad_(True)
    poses = syn.arc_poses(SB)
    focal_t = torch.tensor(focal, dtype=torch.float32)
    idx = syn.pick_ray_indices(H * W, n_rays, seed=seed)
    R = SB * n_rays
    noise = syn.make_noise(R, Kc, Kf, seed=seed)
    U = sys.modules["_nrf_reference_utils"]
    rays = U.gen_rays(poses, W, H, focal_t, cfg.z_near, cfg.z_far, c=None).reshape(SB, H * W, 8)[:, idx]
    ren.train()
    ren.encode(None, None, None, vol, poses, focal_t, None)
    with L.inject_noise([noise["coarse"], noise["u"], noise["fine"]]):
        o = ren.forward_nerf(rays, want_weights=True)
    gw = torch.Generator().manual_seed(900 + seed)
    loss = 0.0
    probes = {}
    for lvl in ("coarse", "fine"):
        for k in ("rgb", "embed", "depth", "coord", "attention"):
            t = o[lvl][k]
            probes[f"{lvl}_{k}"] = torch.randn(t.shape, generator=gw)
            loss = loss + (t * probes[f"{lvl}_{k}"]).sum()
    loss.backward()
    fx = {"meta": np.array([S, C, D, hidden, SB, n_rays, Kc, Kf, 0, H, W, seed]), "focal": np.float32(focal),
          "idx": idx.numpy(), "rays": rays.numpy(), "loss": np.float32(loss.item()), "vol": vol.detach().numpy(),
          "vgrad": vol.grad.numpy(), "poses": poses.numpy()}
    for lvl in ("coarse", "fine"):
        for k in ("rgb", "embed", "depth", "coord", "attention", "weights"):
            fx[f"{lvl}_{k}"] = o[lvl][k].detach().numpy()
    for k, v in probes.items():
        fx["probe_" + k] = v.numpy()
    for k, v in params.items():
        fx["param." + k] = v.numpy()
    for k, p in ren.nerf_model.mlp_coarse.named_parameters():
        fx["grad." + k] = p.grad.numpy()
    for k, v in noise.items():
        fx["noise_" + k] = v.numpy()
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **fx)
    print(name, "->", path, os.path.getsize(path) // 1024, "KiB", "loss", float(loss))


def run_multiscale_case(name="small_multiscale", S=12, C=16, D=24, hidden=64, SB=2, n_rays=40, Kc=16, Kf=16, Kfd=2,
                        H=16, W=16, focal=19.0, seed=8):
    """use_multi_scale_voxel (models_embed.py:279-286: latent = [10-ch 12^3 | 8-ch 6^3 | main 16-ch 12^3] = 34 channels)
    together with ret_last_feat (neural_rendering.py:285-293,332-334: the 64-d last residual stream is composited in
    place of the embedding head) and 2 depth-guided samples; forward_nerf of the reference + a probe loss over its
    outputs, gradients into all three volumes and the MLP."""
    torch.manual_seed(seed)
    ms_shapes = [(10, S), (8, S // 2)]
    cfg = L.default_cfg(d_embed=D, d_latent=C, voxel_shape=S, image_width=W, image_height=H, n_coarse=Kc, n_fine=Kf,
                        n_fine_depth=Kfd, ray_chunk_size=n_rays, mlp=dict(d_hidden=hidden), eval_batch_size=1024,
                        use_multi_scale_voxel=True, d_multi_scale_latent=C + sum(c for c, _ in ms_shapes),
                        ret_last_feat=True)
    ren = L.build_reference_renderer(cfg, torch.tensor(syn.BOUNDS))
    params = O.init_params(d_in=42, d_latent=cfg.d_multi_scale_latent, d_hidden=hidden, d_out=4 + D, seed=seed)
    g = torch.Generator().manual_seed(77 + seed)
    for k in params:
        if k.endswith(".bias"):
            params[k] = 0.05 * torch.randn(params[k].shape, generator=g)
    load_params_into(ren, params)
    vol = syn.make_volume(SB, C, S, seed=seed).requires_grad_(True)
    ms = [(torch.randn(SB, c, s_, s_, s_, generator=g) * 0.1).requires_grad_(True) for c, s_ in ms_shapes]
    poses = syn.arc_poses(SB)
    focal_t = torch.tensor(focal, dtype=torch.float32)
    idx = syn.pick_ray_indices(H * W, n_rays, seed=seed)
    R = SB * n_rays
    noise = syn.make_noise(R, Kc, Kf - Kfd, seed=seed)
    noise["depth"] = torch.randn(R, Kfd, generator=g)
    U = sys.modules["_nrf_reference_utils"]
    rays = U.gen_rays(poses, W, H, focal_t, cfg.z_near, cfg.z_far, c=None).reshape(SB, H * W, 8)[:, idx]
    ren.train()
    ren.encode(ms, None, None, vol, poses, focal_t, None)
    with L.inject_noise([noise["coarse"], noise["u"], noise["fine"], noise["depth"]]):
        o = ren.forward_nerf(rays, want_weights=True)
    gw = torch.Generator().manual_seed(900 + seed)
    loss = 0.0
    probes = {}
    for lvl in ("coarse", "fine"):
        for k in ("rgb", "embed", "depth"):
            t = o[lvl][k]
            probes[f"{lvl}_{k}"] = torch.randn(t.shape, generator=gw)
            loss = loss + (t * probes[f"{lvl}_{k}"]).sum()
    loss.backward()
    assert o["coarse"]["embed"].shape[-1] == hidden
    fx = {"meta": np.array([S, C, D, hidden, SB, n_rays, Kc, Kf, Kfd, H, W, seed]), "focal": np.float32(focal),
          "idx": idx.numpy(), "rays": rays.numpy(), "loss": np.float32(loss.item()), "vol": vol.detach().numpy(),
          "vgrad": vol.grad.numpy(), "poses": poses.numpy(), "n_ms": np.int64(len(ms))}
    for i, v in enumerate(ms):
        fx[f"ms{i}"] = v.detach().numpy()
        fx[f"ms{i}_grad"] = v.grad.numpy()
    for lvl in ("coarse", "fine"):
        for k in ("rgb", "embed", "depth", "weights"):
            fx[f"{lvl}_{k}"] = o[lvl][k].detach().numpy()
    for k, v in probes.items():
        fx["probe_" + k] = v.numpy()
    for k, v in params.items():
        fx["param." + k] = v.numpy()
    for k, p in ren.nerf_model.mlp_coarse.named_parameters():
        fx["grad." + k] = p.grad.numpy()
    for k, v in noise.items():
        fx["noise_" + k] = v.numpy()
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **fx)
    print(name, "->", path, os.path.getsize(path) // 1024, "KiB", "loss", float(loss))


def run_code_viewdirs_case(name="small_codeviewdirs", S=12, C=16, D=24, hidden=64, SB=2, n_rays=40, Kc=16, Kf=16, Kfd=2,
                           H=16, W=16, focal=19.0, seed=9):
    """use_code_viewdirs (models_embed.py:86-95,:355-372: the view direction goes through the positional encoding with
    the point, d_in = 6 + 12 * 6 = 78) together with normalize_z = True (:337-340: a no-op under the hard-coded
    canon_xyz, :42 - this case is what pins that claim); forward_nerf of the reference + a probe loss over its
    outputs, gradients into the volume and the MLP.  Also the reference's PositionalEncoding on 6-d inputs."""
    torch.manual_seed(seed)
    cfg = L.default_cfg(d_embed=D, d_latent=C, voxel_shape=S, image_width=W, image_height=H, n_coarse=Kc, n_fine=Kf,
                        n_fine_depth=Kfd, ray_chunk_size=n_rays, mlp=dict(d_hidden=hidden), eval_batch_size=1024,
                        use_code_viewdirs=True, normalize_z=True)
    ren = L.build_reference_renderer(cfg, torch.tensor(syn.BOUNDS))
    assert ren.nerf_model.d_in == 78 and ren.nerf_model.mlp_coarse.lin_in.weight.shape == (hidden, 78)
    params = O.init_params(d_in=78, d_latent=C, d_hidden=hidden, d_out=4 + D, seed=seed)
    g = torch.Generator().manual_seed(77 + seed)
    for k in params:
        if k.endswith(".bias"):
            params[k] = 0.05 * torch.randn(params[k].shape, generator=g)
    load_params_into(ren, params)
    vol = syn.make_volume(SB, C, S, seed=seed).requires_grad_(True)
    poses = syn.arc_poses(SB)
    focal_t = torch.tensor(focal, dtype=torch.float32)
    idx = syn.pick_ray_indices(H * W, n_rays, seed=seed)
    R = SB * n_rays
    noise = syn.make_noise(R, Kc, Kf - Kfd, seed=seed)
    noise["depth"] = torch.randn(R, Kfd, generator=g)
    U = sys.modules["_nrf_reference_utils"]
    rays = U.gen_rays(poses, W, H, focal_t, cfg.z_near, cfg.z_far, c=None).reshape(SB, H * W, 8)[:, idx]
    ren.train()
    ren.encode(None, None, None, vol, poses, focal_t, None)
    with L.inject_noise([noise["coarse"], noise["u"], noise["fine"], noise["depth"]]):
        o = ren.forward_nerf(rays, want_weights=True)
    gw = torch.Generator().manual_seed(900 + seed)
    loss = 0.0
    probes = {}
    for lvl in ("coarse", "fine"):
        for k in ("rgb", "embed", "depth"):
            t = o[lvl][k]
            probes[f"{lvl}_{k}"] = torch.randn(t.shape, generator=gw)
            loss = loss + (t * probes[f"{lvl}_{k}"]).sum()
    loss.backward()
    fx = {"meta": np.array([S, C, D, hidden, SB, n_rays, Kc, Kf, Kfd, H, W, seed]), "focal": np.float32(focal),
          "idx": idx.numpy(), "rays": rays.numpy(), "loss": np.float32(loss.item()), "vol": vol.detach().numpy(),
          "vgrad": vol.grad.numpy(), "poses": poses.numpy()}
    for lvl in ("coarse", "fine"):
        for k in ("rgb", "embed", "depth", "weights"):
            fx[f"{lvl}_{k}"] = o[lvl][k].detach().numpy()
    for k, v in probes.items():
        fx["probe_" + k] = v.numpy()
    for k, v in params.items():
        fx["param." + k] = v.numpy()
    for k, p in ren.nerf_model.mlp_coarse.named_parameters():
        fx["grad." + k] = p.grad.numpy()
    for k, v in noise.items():
        fx["noise_" + k] = v.numpy()
    x6 = torch.rand(129, 6, generator=torch.Generator().manual_seed(5)) * 2.0 - 1.0
    fx["pe6_x"] = x6.numpy()
    fx["pe6_out"] = ren.nerf_model.code(x6).detach().numpy()
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **fx)
    print(name, "->", path, os.path.getsize(path) // 1024, "KiB", "loss", float(loss))


def run_softplus_spade_case(name="small_softplus_spade", S=12, C=16, D=24, hidden=64, SB=2, n_rays=40, Kc=16, Kf=16,
                            H=16, W=16, focal=19.0, seed=10, beta=10.0):
    """mlp.beta > 0 (softplus activations, resnetfc.py:43-46,:138-141) together with mlp.use_spade (the latent modulates
    the residual stream: x = scale_z(z) * x + lin_z(z), :130-136,:184-186; six more parameters); forward_nerf of the
    reference + a probe loss over its outputs, gradients into the volume and every MLP parameter."""
    torch.manual_seed(seed)
    cfg = L.default_cfg(d_embed=D, d_latent=C, voxel_shape=S, image_width=W, image_height=H, n_coarse=Kc, n_fine=Kf,
                        n_fine_depth=0, ray_chunk_size=n_rays, mlp=dict(d_hidden=hidden, beta=beta, use_spade=True),
                        eval_batch_size=1024)
    ren = L.build_reference_renderer(cfg, torch.tensor(syn.BOUNDS))
    assert len(ren.nerf_model.mlp_coarse.scale_z) == 3 and len(ren.state_dict()) == 74
    params = O.init_params(d_in=42, d_latent=C, d_hidden=hidden, d_out=4 + D, seed=seed, use_spade=True)
    g = torch.Generator().manual_seed(77 + seed)
    for k in params:
        if k.endswith(".bias"):
            params[k] = 0.05 * torch.randn(params[k].shape, generator=g)
        if k.startswith("scale_z") and k.endswith(".bias"):
            params[k] = params[k] + 1.0                  # a scale around one keeps the stream's magnitude
    load_params_into(ren, params)
    vol = syn.make_volume(SB, C, S, seed=seed).requires_grad_(True)
    poses = syn.arc_poses(SB)
    focal_t = torch.tensor(focal, dtype=torch.float32)
    idx = syn.pick_ray_indices(H * W, n_rays, seed=seed)
    R = SB * n_rays
    noise = syn.make_noise(R, Kc, Kf, seed=seed)
    U = sys.modules["_nrf_reference_utils"]
    rays = U.gen_rays(poses, W, H, focal_t, cfg.z_near, cfg.z_far, c=None).reshape(SB, H * W, 8)[:, idx]
    ren.train()
    ren.encode(None, None, None, vol, poses, focal_t, None)
    with L.inject_noise([noise["coarse"], noise["u"], noise["fine"]]):
        o = ren.forward_nerf(rays, want_weights=True)
    gw = torch.Generator().manual_seed(900 + seed)
    loss = 0.0
    probes = {}
    for lvl in ("coarse", "fine"):
        for k in ("rgb", "embed", "depth"):
            t = o[lvl][k]
            probes[f"{lvl}_{k}"] = torch.randn(t.shape, generator=gw)
            loss = loss + (t * probes[f"{lvl}_{k}"]).sum()
    loss.backward()
    fx = {"meta": np.array([S, C, D, hidden, SB, n_rays, Kc, Kf, 0, H, W, seed]), "focal": np.float32(focal),
          "beta": np.float32(beta), "idx": idx.numpy(), "rays": rays.numpy(), "loss": np.float32(loss.item()),
          "vol": vol.detach().numpy(), "vgrad": vol.grad.numpy(), "poses": poses.numpy()}
    for lvl in ("coarse", "fine"):
        for k in ("rgb", "embed", "depth", "weights"):
            fx[f"{lvl}_{k}"] = o[lvl][k].detach().numpy()
    for k, v in probes.items():
        fx["probe_" + k] = v.numpy()
    for k, v in params.items():
        fx["param." + k] = v.numpy()
    for k, p in ren.nerf_model.mlp_coarse.named_parameters():
        fx["grad." + k] = p.grad.numpy()
    for k, v in noise.items():
        fx["noise_" + k] = v.numpy()
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **fx)
    print(name, "->", path, os.path.getsize(path) // 1024, "KiB", "loss", float(loss))


def run_extract_case(name="extract_small", S=12, C=16, D=24, hidden=64, SB=2, n_rays=24, K=20, H=16, W=16, focal=19.0,
                     seed=12):
    """The ancestor renderer's `extract_radience` (featurenerf_robo/featurenerf/src/render/nerf_embed.py:432-516), imported
    unmodified (its `util` import stubbed: the method does not use it), driven with the reference's own field model
    (PixelNeRFEmbedNet after encode(); the ancestor expects a model that returns the value tensor, this repo's returns
    (value, point_density): the wrapper takes element 0) on sorted sample depths: points, rgbs, sigmas, embeds."""
    import importlib.util
    torch.manual_seed(seed)
    cfg = L.default_cfg(d_embed=D, d_latent=C, voxel_shape=S, image_width=W, image_height=H, n_coarse=K, n_fine=0,
                        n_fine_depth=0, ray_chunk_size=n_rays, mlp=dict(d_hidden=hidden), eval_batch_size=256)
    ren = L.build_reference_renderer(cfg, torch.tensor(syn.BOUNDS))
    params = O.init_params(d_in=42, d_latent=C, d_hidden=hidden, d_out=4 + D, seed=seed)
    g = torch.Generator().manual_seed(77 + seed)
    for k in params:
        if k.endswith(".bias"):
            params[k] = 0.05 * torch.randn(params[k].shape, generator=g)
    load_params_into(ren, params)
    sys.modules.setdefault("util", mock.MagicMock())
    spec = importlib.util.spec_from_file_location(
        "_nrf_ancestor_nerf_embed", "/root/reference/featurenerf_robo/featurenerf/src/render/nerf_embed.py")
    anc_mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(anc_mod)
    anc = anc_mod.NeRFEmbedRenderer(n_coarse=K, n_fine=0, eval_batch_size=256)

    class Model:
        use_viewdirs = True

        def __call__(self, pnts, coarse=True, viewdirs=None):
            return ren.nerf_model(pnts, coarse=coarse, viewdirs=viewdirs)[0]

    vol = syn.make_volume(SB, C, S, seed=seed)
    poses = syn.arc_poses(SB)
    focal_t = torch.tensor(focal, dtype=torch.float32)
    idx = syn.pick_ray_indices(H * W, n_rays, seed=seed)
    U = sys.modules["_nrf_reference_utils"]
    rays = U.gen_rays(poses, W, H, focal_t, cfg.z_near, cfg.z_far, c=None).reshape(SB, H * W, 8)[:, idx].reshape(-1, 8)
    R = rays.shape[0]
    z = rays[:, 6:7] + (rays[:, 7:8] - rays[:, 6:7]) * torch.rand(R, K, generator=g)
    z, _ = torch.sort(z, dim=-1)
    ren.eval()
    ren.encode(None, None, None, vol, poses, focal_t, None)
    with torch.no_grad():
        pts, rgbs, sigmas, embeds = anc.extract_radience(Model(), rays, z, coarse=True, sb=SB)
    assert pts.shape == (SB, n_rays * K, 3) and embeds.shape == (SB, n_rays * K, D)   # (sb = 0 hands the model 2-D points,
                                                                                     # which models_embed.py:307 rejects)
    fx = {"meta": np.array([S, C, D, hidden, SB, n_rays, K, 0, 0, H, W, seed]), "rays": rays.numpy(), "z": z.numpy(),
          "vol": vol.numpy(), "points": pts.numpy(), "rgbs": rgbs.numpy(), "sigmas": sigmas.numpy(),
          "embeds": embeds.numpy()}
    for k, v in params.items():
        fx["param." + k] = v.numpy()
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **fx)
    print(name, "->", path, os.path.getsize(path) // 1024, "KiB", "sigma max", float(sigmas.max()))


def run_raygen():
    L.load_reference()
    U = sys.modules["_nrf_reference_utils"]
    poses = syn.arc_poses(3)
    fx = {"poses": poses.numpy()}
    r = U.gen_rays(poses, 80, 60, torch.tensor(76.18187, dtype=torch.float32), 1.2, 4.0)
    fx["rays_60x80"] = r.numpy()
    r = U.gen_rays(poses[:2], 128, 128, torch.tensor(153.0, dtype=torch.float32), 1.2, 4.0)
    fx["rays_128_rows"] = r[:, ::16].numpy()
    c = torch.tensor([70.5, 61.25])
    r = U.gen_rays(poses[:1], 128, 128, torch.tensor([150.0, 151.0]), 0.5, 3.0, c=c)
    fx["rays_128_c_rows"] = r[:, ::32].numpy()
    pe = U.PositionalEncoding(6, 3, 1.5, True)
    x = torch.rand(257, 3, generator=torch.Generator().manual_seed(3)) * 1.4 - 0.2
    fx["pe_x"] = x.numpy()
    fx["pe_out"] = pe(x).numpy()
    np.savez_compressed(os.path.join(HERE, "raygen_pe.npz"), **fx)
    print("raygen_pe done")


def run_voxelizer():
    """voxelize_small.npz: the reference's VoxelGrid (voxel_grid_real.py, imports torch + numpy only) on CPU."""
    sys.path.insert(0, "/root/reference")
    import voxel_grid_real as V
    B, N, F, S = 2, 6000, 3, 10
    coords, feats = syn.voxelizer_points(B, N, F, seed=11)
    vg = V.VoxelGrid(coord_bounds=syn.BOUNDS, voxel_size=S, device="cpu", batch_size=B, feature_size=F,
                     max_num_coords=N)
    out = vg.coords_to_bounding_voxel_grid(coords, coord_features=feats, coord_bounds=torch.tensor(syn.BOUNDS)[None])
    out_default_bounds = vg.coords_to_bounding_voxel_grid(coords, coord_features=feats)
    assert torch.equal(out, out_default_bounds)
    np.savez_compressed(os.path.join(HERE, "voxelize_small.npz"), meta=np.array([B, N, F, S, 11]),
                        out=out.numpy())
    print("voxelize_small done", tuple(out.shape), "occupied", int(out[..., -1].sum()))


if __name__ == "__main__":
    if "--heads-only" in sys.argv:
        run_heads_case()
        sys.exit(0)
    if "--multiscale-only" in sys.argv:
        run_multiscale_case()
        sys.exit(0)
    if "--codeviewdirs-only" in sys.argv:
        run_code_viewdirs_case()
        sys.exit(0)
    if "--extract-only" in sys.argv:
        run_extract_case()
        sys.exit(0)
    if "--softplus-spade-only" in sys.argv:
        run_softplus_spade_case()
        sys.exit(0)
    run_raygen()
    run_heads_case()
    run_multiscale_case()
    run_code_viewdirs_case()
    run_softplus_spade_case()
    run_extract_case()
    run_voxelizer()
    run_case("small_kfd0", S=12, C=16, D=24, hidden=64, SB=2, n_rays=48, Kc=16, Kf=16, Kfd=0,
             H=16, W=16, focal=19.0, store_inputs=True)
    run_case("small_kfd4", S=12, C=16, D=24, hidden=64, SB=2, n_rays=48, Kc=16, Kf=16, Kfd=4,
             H=16, W=16, focal=19.0, store_inputs=True, seed=1)
    run_case("small_noperturb", S=12, C=16, D=24, hidden=64, SB=1, n_rays=40, Kc=16, Kf=8, Kfd=0,
             H=16, W=16, focal=19.0, store_inputs=True, seed=2, perturb=False)
    run_case("small_noise_wb", S=12, C=16, D=24, hidden=64, SB=2, n_rays=48, Kc=16, Kf=16, Kfd=0,
             H=16, W=16, focal=19.0, store_inputs=True, seed=4, noise_std=0.5, white_bkgd=True, lindisp=True)
    run_case("full_s32", S=32, C=128, D=384, hidden=512, SB=2, n_rays=64, Kc=64, Kf=64, Kfd=0,
             H=128, W=128, focal=153.0, store_inputs=False, seed=3)
