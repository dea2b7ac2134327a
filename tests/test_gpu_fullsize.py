"""Parity at BASELINE.json's full config-2 size (2 scenes x 2048 rays, 64 + 64 samples, 100^3 x 128-channel volume,
786 432 field evaluations), where the CPU oracle cannot follow: size-independent properties of the path.

  * coarse sample depths bit-exact against the oracle formula (cheap on the CPU even at this size);
  * ray independence: a ray's outputs do not depend on which other rays are in the batch (bit-identical);
  * sampling / compositing invariants (sorted depths, weights in [0,1], sum of weights <= 1, rgb in [0,1]);
  * the volume-gradient scatter is the exact adjoint of the trilinear gather (<gather(V), Y> == <V, scatter(Y)>);
  * the backward is linear in the upstream gradient: scaling the loss by 2 doubles every gradient EXACTLY in the
    reproducible mode (powers of two commute with every rounding on the path);
  * the fused MLP kernels equal the layer-by-layer chain bit for bit at the fine pass's size (524 288 samples).
"""
import pytest
import torch

from oracle import nerf_oracle as O
from tests.conftest import load_pkg

pytestmark = pytest.mark.gpu
syn = load_pkg("synthetic")


@pytest.fixture(scope="module")
def env():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    ops, NR, U = load_pkg("ops"), load_pkg("neural_rendering"), load_pkg("utils")
    wl = syn.CONFIGS["config2"]
    cfg = U.default_config(voxel_shape=wl.S, d_latent=wl.C, d_embed=wl.D, n_coarse=wl.n_coarse, n_fine=wl.n_fine,
                           ray_chunk_size=wl.rays_per_scene)
    ren = NR.NeuralRenderer(cfg, torch.tensor(syn.BOUNDS), precision="bf16")
    syn.init_mlp_(ren.nerf_model.mlp_coarse, seed=0)
    ren = ren.cuda()
    ren.deterministic = True
    g = torch.Generator(device="cuda").manual_seed(11)
    vol = torch.randn(wl.SB, wl.C, wl.S, wl.S, wl.S, device="cuda", generator=g) * 0.1
    poses = syn.arc_poses(wl.SB)
    rays_all = O.gen_rays(poses, wl.W, wl.H, torch.tensor(wl.focal), 1.2, 4.0).reshape(wl.SB, -1, 8)
    idx = syn.pick_ray_indices(wl.W * wl.H, wl.rays_per_scene, seed=3)
    rays = rays_all[:, idx].contiguous()                                   # (SB, 2048, 8) CPU
    noise = syn.make_noise(wl.SB * wl.rays_per_scene, wl.n_coarse, wl.n_fine, seed=5)
    return dict(ops=ops, NR=NR, ren=ren, wl=wl, vol=vol, rays=rays, noise=noise)


def _render(env, rays, noise, vol=None, want_grad=False, scale=1.0):
    ren = env["ren"]
    vol = (env["vol"] if vol is None else vol).detach().clone().requires_grad_(want_grad)
    for p in ren.parameters():
        p.grad = None
    ren.encode(None, None, None, vol, None, None, None)
    ctx = torch.enable_grad() if want_grad else torch.no_grad()
    with ctx:
        out = ren.forward_nerf(rays.cuda(), want_weights=True, noise={k: v.cuda() for k, v in noise.items()})
        if want_grad:
            g = torch.Generator(device="cuda").manual_seed(99)
            loss = 0.0
            for lvl in ("coarse", "fine"):
                for k in ("rgb", "embed", "depth"):
                    t = out[lvl][k]
                    loss = loss + (t * torch.randn(t.shape, device="cuda", generator=g)).sum()
            (loss * scale).backward()
    grads = {k: p.grad.clone() for k, p in ren.named_parameters() if p.grad is not None} if want_grad else None
    return out, (vol.grad if want_grad else None), grads


def test_coarse_depths_bit_exact_and_invariants_at_full_size(env):
    wl = env["wl"]
    out, _, _ = _render(env, env["rays"], env["noise"])
    R = wl.SB * wl.rays_per_scene
    flat = env["rays"].reshape(R, 8)
    z_ref = O.sample_coarse(flat, wl.n_coarse, jitter=env["noise"]["coarse"])
    assert torch.equal(out.coarse.z.cpu(), z_ref), "coarse sample depths must be bit-exact"
    zf = out.fine.z
    assert zf.shape == (R, wl.n_coarse + wl.n_fine)
    assert bool((zf[:, 1:] >= zf[:, :-1]).all()), "fine-pass depths are sorted"
    assert float(zf.min()) >= 1.2 and float(zf.max()) <= 4.0 + 1e-3
    for lvl in ("coarse", "fine"):
        w = out[lvl].weights.reshape(R, -1)
        assert float(w.min()) >= 0.0 and float(w.sum(-1).max()) <= 1.0 + 1e-4
        rgb = out[lvl].rgb
        assert float(rgb.min()) >= 0.0 and float(rgb.max()) <= 1.0 + 1e-5
        assert bool(torch.isfinite(out[lvl].embed).all()) and bool(torch.isfinite(out[lvl].depth).all())
        acc = w.sum(-1)
        d = out[lvl].depth.reshape(R)
        assert bool((d <= acc * 4.0 + 1e-3).all()) and bool((d >= acc * 1.2 - 1e-3).all())


def test_a_rays_outputs_do_not_depend_on_the_batch(env):
    """Bit-identical outputs for a 2 x 200-ray subset rendered alone (other tiles, other CTAs, other chunk order)."""
    wl = env["wl"]
    full, _, _ = _render(env, env["rays"], env["noise"])
    sel = torch.arange(37, 37 + 200)
    rows = torch.cat([sel, wl.rays_per_scene + sel])
    sub_noise = {k: v[rows].contiguous() for k, v in env["noise"].items()}
    sub, _, _ = _render(env, env["rays"][:, sel].contiguous(), sub_noise)
    for lvl in ("coarse", "fine"):
        for k in ("rgb", "embed", "depth", "weights"):
            a = full[lvl][k][:, sel]
            assert torch.equal(a, sub[lvl][k]), (lvl, k)
    assert torch.equal(full.fine.z[rows.cuda()], sub.fine.z)


def test_scatter_is_the_adjoint_of_the_gather_at_full_size(env):
    ops, wl = env["ops"], env["wl"]
    R = wl.SB * wl.rays_per_scene
    rays = env["rays"].reshape(R, 8).cuda()
    z = ops.sample_coarse(rays, wl.n_coarse, env["noise"]["coarse"].cuda())
    vol_cl = ops.volume_to_channels_last(env["vol"])
    fin = ops.encode_points(rays, z, wl.rays_per_scene, vol_cl, torch.tensor(syn.BOUNDS), precision=ops.NRF_PREC_FP32)
    lat = fin[:, :wl.C].double()
    g = torch.Generator(device="cuda").manual_seed(4)
    Y = torch.randn(R * wl.n_coarse, wl.C, device="cuda", generator=g)
    lhs = float((lat * Y.double()).sum())
    for variant in ("sorted", "atomic"):
        if variant == "sorted":
            grad = torch.empty_like(vol_cl)
            ops.scatter_volume_grad_sorted(rays, z, wl.rays_per_scene, Y, grad, torch.tensor(syn.BOUNDS))
        else:
            grad = torch.zeros_like(vol_cl)
            ops.scatter_volume_grad(rays, z, wl.rays_per_scene, Y, grad, torch.tensor(syn.BOUNDS))
        rhs = float((grad.double() * vol_cl.double()).sum())
        scale = float((lat.abs() * Y.double().abs()).sum())
        assert abs(lhs - rhs) <= 2e-6 * scale, (variant, lhs, rhs)
    # a sample inside the box really gathers something: the test is not vacuous
    assert float(lat.abs().sum()) > 0
    # the merged scatter (the default of the training step): coarse + fine-sized pass in ONE sort, channel-first out
    u = torch.rand(R, wl.n_fine, device="cuda", generator=g)
    z2 = ops.sample_fine(rays, torch.rand(R, wl.n_coarse, device="cuda", generator=g), wl.n_coarse, u)
    fin2 = ops.encode_points(rays, z2, wl.rays_per_scene, vol_cl, torch.tensor(syn.BOUNDS), precision=ops.NRF_PREC_FP32)
    Y2 = torch.randn(R * wl.n_fine, wl.C, device="cuda", generator=g)
    lhs2 = lhs + float((fin2[:, :wl.C].double() * Y2.double()).sum())
    grad_cf = torch.full_like(env["vol"], 7.0)                               # garbage: every element must be rewritten
    ops.scatter_volume_grad_merged(rays, wl.rays_per_scene, [(z, Y), (z2, Y2)], grad_cf, True, torch.tensor(syn.BOUNDS))
    rhs2 = float((grad_cf.double() * env["vol"].double()).sum())
    scale2 = scale + float((fin2[:, :wl.C].double().abs() * Y2.double().abs()).sum())
    assert abs(lhs2 - rhs2) <= 2e-6 * scale2, (lhs2, rhs2)
    again = torch.empty_like(grad_cf)
    ops.scatter_volume_grad_merged(rays, wl.rays_per_scene, [(z, Y), (z2, Y2)], again, True, torch.tensor(syn.BOUNDS))
    assert torch.equal(again, grad_cf), "bit-reproducible at full size"


def test_backward_is_exactly_linear_in_the_upstream_gradient(env):
    """2 x loss -> exactly 2 x every gradient, bit for bit, in the reproducible mode: every operation of the backward
    is linear in the upstream gradient and scaling by a power of two commutes with fp32 and bf16 rounding."""
    _, vg1, pg1 = _render(env, env["rays"], env["noise"], want_grad=True, scale=1.0)
    _, vg2, pg2 = _render(env, env["rays"], env["noise"], want_grad=True, scale=2.0)
    assert float(vg1.abs().sum()) > 0
    assert torch.equal(vg1 * 2.0, vg2)
    for k in pg1:
        assert torch.equal(pg1[k] * 2.0, pg2[k]), k


def test_fused_mlp_equals_layered_chain_at_fine_pass_size(env):
    ops, NR, ren = env["ops"], env["NR"], env["ren"]
    h = ren.nerf_model.mlp_coarse.handle(ops.NRF_PREC_BF16)
    assert h.fused
    N = 524288
    g = torch.Generator(device="cuda").manual_seed(8)
    fin = torch.zeros(N, h.sizes.kin_pad, device="cuda", dtype=torch.bfloat16)
    fin[:, :170] = (torch.randn(N, 170, device="cuda", generator=g) * 0.5).to(torch.bfloat16)
    dfield = torch.zeros(N, h.sizes.dout_pad, device="cuda", dtype=torch.bfloat16)
    dfield[:, :388] = (torch.randn(N, 388, device="cuda", generator=g) * 0.1).to(torch.bfloat16)
    out_l, acts_l = h.forward(fin, layered=True)
    out_f, acts_f = h.forward(fin)
    assert torch.equal(out_l, out_f)
    n = 11 * N * 512
    assert torch.equal(acts_l.view(torch.bfloat16)[:n], acts_f.view(torch.bfloat16)[:n])
    del acts_l
    ga, gb = NR._zero_grads(h), NR._zero_grads(h)
    da = h.backward(fin, acts_f, dfield, ga, deterministic=True, layered=True)
    db = h.backward(fin, acts_f, dfield, gb, deterministic=True)
    assert torch.equal(da, db)
    assert all(torch.equal(ga[k], gb[k]) for k in h.names())
