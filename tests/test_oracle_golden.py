"""Pins oracle/nerf_oracle.py against outputs of the unmodified reference (tests/golden/*.npz,
produced by tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import nerf_oracle as O
from tests.conftest import golden, load_pkg

syn = load_pkg("synthetic")
T = torch.from_numpy


def _rel(a, b):
    return float((a - b).norm() / (b.norm() + 1e-30))


def test_raygen_bitexact_cpu():
    fx = golden("raygen_pe")
    poses = T(fx["poses"])
    r = O.gen_rays(poses, 80, 60, torch.tensor(76.18187), 1.2, 4.0)
    assert torch.equal(r, T(fx["rays_60x80"]))
    r = O.gen_rays(poses[:2], 128, 128, torch.tensor(153.0), 1.2, 4.0)
    assert torch.equal(r[:, ::16], T(fx["rays_128_rows"]))
    r = O.gen_rays(poses[:1], 128, 128, torch.tensor([150.0, 151.0]), 0.5, 3.0,
                   c=torch.tensor([70.5, 61.25]))
    assert torch.equal(r[:, ::32], T(fx["rays_128_c_rows"]))


def test_positional_encoding_bitexact_cpu():
    fx = golden("raygen_pe")
    assert torch.equal(O.positional_encoding(T(fx["pe_x"])), T(fx["pe_out"]))
    # known-answer from SURVEY.md 8(a7)
    pe = O.positional_encoding(torch.tensor([[.1, .2, .3]]))[0, :9]
    assert torch.allclose(pe, torch.tensor([.1, .2, .3, .149438, .295520, .434966,
                                            .988771, .955337, .900447]), atol=1e-6)


def test_explicit_trilinear_matches_grid_sample_bitwise():
    g = torch.Generator().manual_seed(0)
    vol = torch.randn(2, 6, 9, 10, 11, generator=g)
    canon = torch.rand(2, 5000, 3, generator=g) * 1.3 - 0.15     # includes out-of-box points
    a = O.trilinear_gather(vol, canon)
    b = O.trilinear_gather_explicit(vol, canon)
    assert torch.equal(a, b)


def _case_inputs(fx):
    S, C, D, hidden, SB, n_rays, Kc, Kf, Kfd, H, W, seed = [int(v) for v in fx["meta"]]
    params = {k[6:]: T(fx[k]) for k in fx.files if k.startswith("param.")}
    noise = {k[6:]: T(fx[k]) for k in fx.files if k.startswith("noise_") and k != "noise_depth"}
    if "noise_depth" in fx.files:
        noise["depth"] = T(fx["noise_depth"])
    return dict(S=S, C=C, D=D, hidden=hidden, SB=SB, n_rays=n_rays, Kc=Kc, Kf=Kf, Kfd=Kfd, H=H, W=W,
                seed=seed, params=params, noise=noise)


def _opts(fx):
    """noise_std, white_bkgd, lindisp of a case (absent in the first fixtures: all off)."""
    if "opts" not in fx.files:
        return dict(noise_std=0.0, white_bkgd=False, lindisp=False)
    o = fx["opts"]
    return dict(noise_std=float(o[0]), white_bkgd=bool(o[1]), lindisp=bool(o[2]))


@pytest.mark.parametrize("name", ["small_kfd0", "small_kfd4", "small_noperturb", "small_noise_wb"])
def test_small_cases_forward_backward(name):
    fx = golden(name)
    ci = _case_inputs(fx)
    params = {k: v.clone().requires_grad_(True) for k, v in ci["params"].items()}
    vol = T(fx["vol"]).clone().requires_grad_(True)
    poses = T(fx["poses"])
    rays_full = O.gen_rays(poses, ci["W"], ci["H"], torch.tensor(float(fx["focal"])), 1.2, 4.0)
    idx = T(fx["idx"])
    rays = rays_full.reshape(ci["SB"], -1, 8)[:, idx]
    assert torch.equal(rays, T(fx["rays"]))
    out = O.forward_nerf(params, vol, rays, syn.BOUNDS, ci["Kc"], ci["Kf"], ci["Kfd"],
                         noise=ci["noise"], eval_batch_size=1024, **_opts(fx))
    assert torch.equal(out["z_coarse"], T(fx["z_coarse"]))
    for lvl in ("coarse", "fine"):
        for k in ("rgb", "embed", "depth", "weights"):
            ref = T(fx[f"{lvl}_{k}"])
            assert _rel(out[lvl][k].detach(), ref) < 2e-6, (lvl, k)
    gt_rgb = T(fx["gt_rgb_img"]).reshape(ci["SB"], -1, 3)[:, idx]
    gt_emb = T(fx["gt_embed_img"]).reshape(ci["SB"], -1, ci["D"])[:, idx]
    L = O.rendering_loss(out, gt_rgb, gt_emb)
    assert abs(float(L["loss"]) - float(fx["loss"])) < 1e-6 * max(1.0, abs(float(fx["loss"])))
    items = [float(L[k]) for k in ("loss_rgb_coarse", "loss_rgb_fine", "loss_embed_coarse",
                                   "loss_embed_fine", "psnr")]
    assert np.allclose(items, fx["loss_items"], rtol=1e-5)
    L["loss"].backward()
    assert _rel(vol.grad, T(fx["vgrad"])) < 1e-5
    for k, p in params.items():
        assert _rel(p.grad, T(fx["grad." + k])) < 1e-5, k


def test_full_dims_forward_backward():
    """BASELINE dims (C=128, D=384, hidden 512, Kc=Kf=64) on a 32^3 volume; inputs from seeds."""
    fx = golden("full_s32")
    S, C, D, hidden, SB, n_rays, Kc, Kf, Kfd, H, W, seed = [int(v) for v in fx["meta"]]
    inp = syn_case_inputs(fx)
    params = {k: v.clone().requires_grad_(True) for k, v in inp["params"].items()}
    vol = inp["vol"].requires_grad_(True)
    assert torch.equal(inp["rays"], T(fx["rays"]))
    out = O.forward_nerf(params, vol, inp["rays"], syn.BOUNDS, Kc, Kf, Kfd, noise=inp["noise"])
    assert torch.equal(out["z_coarse"], T(fx["z_coarse"]))
    for lvl in ("coarse", "fine"):
        for k in ("rgb", "embed", "depth", "weights"):
            assert _rel(out[lvl][k].detach(), T(fx[f"{lvl}_{k}"])) < 5e-6, (lvl, k)
    L = O.rendering_loss(out, inp["gt_rgb"], inp["gt_embed"])
    assert abs(float(L["loss"]) - float(fx["loss"])) < 2e-6
    L["loss"].backward()
    sign = inp["sign"]
    assert _rel(vol.grad.sum(1), T(fx["vgrad_sum"])) < 2e-5
    assert _rel((vol.grad * sign.view(1, C, 1, 1, 1)).sum(1), T(fx["vgrad_sign"])) < 2e-5
    for k, p in params.items():
        assert abs(float(p.grad.norm()) - float(fx["gradnorm." + k])) < 2e-5 * float(fx["gradnorm." + k]) + 1e-9, k
        if "grad." + k in fx.files:
            assert _rel(p.grad, T(fx["grad." + k])) < 2e-5, k
        else:
            rows = torch.randperm(p.shape[0], generator=torch.Generator().manual_seed(5))[:64]
            assert _rel(p.grad[rows], T(fx["gradrows." + k])) < 2e-5, k


def syn_case_inputs(fx):
    """Rebuilds the seeded inputs of a `store_inputs=False` golden case (see make_golden.run_case)."""
    S, C, D, hidden, SB, n_rays, Kc, Kf, Kfd, H, W, seed = [int(v) for v in fx["meta"]]
    params = O.init_params(d_in=42, d_latent=C, d_hidden=hidden, d_out=4 + D, seed=seed)
    g = torch.Generator().manual_seed(77 + seed)
    for k in params:
        if k.endswith(".bias"):
            params[k] = 0.05 * torch.randn(params[k].shape, generator=g)
    vol = syn.make_volume(SB, C, S, seed=seed)
    poses = syn.arc_poses(SB)
    idx = syn.pick_ray_indices(H * W, n_rays, seed=seed)
    noise = syn.make_noise(SB * n_rays, Kc, Kf - Kfd, seed=seed, perturb=True)
    gd = torch.Generator().manual_seed(5000 + seed)
    if Kfd > 0:
        noise["depth"] = torch.randn(SB * n_rays, Kfd, generator=gd)
    gt_rgb_img = torch.rand(SB, H, W, 3, generator=gd)
    gt_embed_img = torch.randn(SB, H, W, D, generator=gd)
    rays_full = O.gen_rays(poses, W, H, torch.tensor(float(fx["focal"])), 1.2, 4.0)
    rays = rays_full.reshape(SB, -1, 8)[:, idx]
    sign = (torch.randint(0, 2, (C,), generator=torch.Generator().manual_seed(99)) * 2 - 1).float()
    return dict(params=params, vol=vol, poses=poses, idx=idx, noise=noise, rays=rays,
                gt_rgb=gt_rgb_img.reshape(SB, -1, 3)[:, idx],
                gt_embed=gt_embed_img.reshape(SB, -1, D)[:, idx], sign=sign,
                gt_rgb_img=gt_rgb_img, gt_embed_img=gt_embed_img)


def test_voxelizer_oracle_matches_the_reference_bitwise():
    """oracle/voxel_oracle.py against the output of the reference's own VoxelGrid (tests/golden/voxelize_small.npz)."""
    from oracle import voxel_oracle as VO
    fx = golden("voxelize_small")
    B, N, F, S, seed = [int(v) for v in fx["meta"]]
    coords, feats = syn.voxelizer_points(B, N, F, seed)
    out = VO.voxelize(coords, feats, syn.BOUNDS, S)
    ref = T(fx["out"])
    assert out.shape == ref.shape == (B, S, S, S, 3 + F + 4)
    assert torch.equal(out, ref)
    assert int(ref[..., -1].sum()) > 100                       # the case is not degenerate


def test_oracle_heads_match_reference():
    """regress_coord + regress_attention (models_embed.py:447-461, neural_rendering.py:318-329,353-357) against the
    reference's own outputs and the gradients of the probe loss over all of them (small_heads.npz)."""
    fx = golden("small_heads")
    ci = _case_inputs(fx)
    pr = {k: v.clone().requires_grad_(True) for k, v in ci["params"].items()}
    vol = T(fx["vol"]).clone().requires_grad_(True)
    out = O.forward_nerf(pr, vol, T(fx["rays"]), syn.BOUNDS, ci["Kc"], ci["Kf"], noise=ci["noise"],
                         eval_batch_size=1024, regress_coord=True, regress_attention=True)
    loss = 0.0
    for lvl in ("coarse", "fine"):
        for k in ("rgb", "embed", "depth", "coord", "attention"):
            assert _rel(out[lvl][k], T(fx[f"{lvl}_{k}"])) < 2e-6, (lvl, k)
            loss = loss + (out[lvl][k] * T(fx[f"probe_{lvl}_{k}"])).sum()
    assert abs(float(loss) - float(fx["loss"])) < 1e-4 * max(1.0, abs(float(fx["loss"])))
    loss.backward()
    assert _rel(vol.grad, T(fx["vgrad"])) < 1e-5
    for k, v in pr.items():
        assert _rel(v.grad, T(fx["grad." + k])) < 1e-4, k


def test_oracle_multiscale_and_last_feat_match_reference():
    """use_multi_scale_voxel + ret_last_feat + depth-guided samples against the reference (small_multiscale.npz)."""
    fx = golden("small_multiscale")
    ci = _case_inputs(fx)
    pr = {k: v.clone().requires_grad_(True) for k, v in ci["params"].items()}
    vol = T(fx["vol"]).clone().requires_grad_(True)
    ms = [T(fx[f"ms{i}"]).clone().requires_grad_(True) for i in range(int(fx["n_ms"]))]
    out = O.forward_nerf(pr, vol, T(fx["rays"]), syn.BOUNDS, ci["Kc"], ci["Kf"], ci["Kfd"], noise=ci["noise"],
                         eval_batch_size=1024, multi_scale_voxel_list=ms, ret_last_feat=True)
    loss = 0.0
    for lvl in ("coarse", "fine"):
        assert out[lvl]["embed"].shape[-1] == ci["hidden"]
        for k in ("rgb", "embed", "depth"):
            assert _rel(out[lvl][k], T(fx[f"{lvl}_{k}"])) < 2e-6, (lvl, k)
            loss = loss + (out[lvl][k] * T(fx[f"probe_{lvl}_{k}"])).sum()
    loss.backward()
    assert _rel(vol.grad, T(fx["vgrad"])) < 1e-5
    for i, v in enumerate(ms):
        assert _rel(v.grad, T(fx[f"ms{i}_grad"])) < 1e-5, i
    for k, v in pr.items():
        assert _rel(v.grad, T(fx["grad." + k])) < 1e-4, k


def test_oracle_code_viewdirs_matches_reference():
    """use_code_viewdirs (models_embed.py:86-95,:355-372: [xyz | viewdir] through the positional encoding together, d_in =
    78) with normalize_z = True - which the fixture shows to be the no-op models_embed.py:42,:337-340 make it: the oracle
    has no such switch and still reproduces the reference (small_codeviewdirs.npz)."""
    fx = golden("small_codeviewdirs")
    ci = _case_inputs(fx)
    assert ci["params"]["lin_in.weight"].shape[1] == 78
    assert _rel(O.positional_encoding(T(fx["pe6_x"]), 6, 1.5, True), T(fx["pe6_out"])) < 1e-7
    pr = {k: v.clone().requires_grad_(True) for k, v in ci["params"].items()}
    vol = T(fx["vol"]).clone().requires_grad_(True)
    out = O.forward_nerf(pr, vol, T(fx["rays"]), syn.BOUNDS, ci["Kc"], ci["Kf"], ci["Kfd"], noise=ci["noise"],
                         eval_batch_size=1024, code_viewdirs=True)
    loss = 0.0
    for lvl in ("coarse", "fine"):
        for k in ("rgb", "embed", "depth"):
            assert _rel(out[lvl][k], T(fx[f"{lvl}_{k}"])) < 2e-6, (lvl, k)
            loss = loss + (out[lvl][k] * T(fx[f"probe_{lvl}_{k}"])).sum()
    assert abs(float(loss) - float(fx["loss"])) < 1e-4 * max(1.0, abs(float(fx["loss"])))
    loss.backward()
    assert _rel(vol.grad, T(fx["vgrad"])) < 1e-5
    for k, v in pr.items():
        assert _rel(v.grad, T(fx["grad." + k])) < 1e-4, k


def test_oracle_softplus_and_spade_match_reference():
    """mlp.beta = 10 (softplus, resnetfc.py:43-46,:138-141) with mlp.use_spade (x = scale_z(z) * x + lin_z(z),
    :130-136,:184-186) against the reference's outputs and gradients (small_softplus_spade.npz)."""
    fx = golden("small_softplus_spade")
    ci = _case_inputs(fx)
    beta = float(fx["beta"])
    assert "scale_z.2.weight" in ci["params"] and beta == 10.0
    pr = {k: v.clone().requires_grad_(True) for k, v in ci["params"].items()}
    vol = T(fx["vol"]).clone().requires_grad_(True)
    out = O.forward_nerf(pr, vol, T(fx["rays"]), syn.BOUNDS, ci["Kc"], ci["Kf"], noise=ci["noise"],
                         eval_batch_size=1024, beta=beta, use_spade=True)
    loss = 0.0
    for lvl in ("coarse", "fine"):
        for k in ("rgb", "embed", "depth"):
            assert _rel(out[lvl][k], T(fx[f"{lvl}_{k}"])) < 2e-6, (lvl, k)
            loss = loss + (out[lvl][k] * T(fx[f"probe_{lvl}_{k}"])).sum()
    assert abs(float(loss) - float(fx["loss"])) < 1e-4 * max(1.0, abs(float(fx["loss"])))
    loss.backward()
    assert _rel(vol.grad, T(fx["vgrad"])) < 1e-5
    for k, v in pr.items():
        assert _rel(v.grad, T(fx["grad." + k])) < 1e-4, k


def test_oracle_extract_radience_matches_the_ancestor_renderer():
    """oracle.extract_radience against the ancestor's own method (featurenerf_robo/featurenerf/src/render/nerf_embed.py:
    432-516, run unmodified over the reference's field model by make_golden.py: extract_small.npz)."""
    fx = golden("extract_small")
    ci = _case_inputs(fx)
    with torch.no_grad():
        pts, rgbs, sigmas, embeds = O.extract_radience(ci["params"], T(fx["vol"]), T(fx["rays"]), T(fx["z"]), ci["SB"],
                                                       syn.BOUNDS)
    assert torch.equal(pts, T(fx["points"]))
    assert rgbs.shape == T(fx["rgbs"]).shape and sigmas.shape == T(fx["sigmas"]).shape
    assert _rel(rgbs, T(fx["rgbs"])) < 2e-6 and _rel(sigmas, T(fx["sigmas"])) < 2e-6
    assert _rel(embeds, T(fx["embeds"])) < 2e-6 and float(T(fx["sigmas"]).max()) > 0
