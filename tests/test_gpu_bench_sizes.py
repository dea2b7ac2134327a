"""Oracle-VALUE parity at the sizes bench.py runs (VERDICT r1 "what's weak" 2 and 3).

Rays are independent (tests/test_gpu_fullsize.py proves a ray's outputs do not depend on the batch), so a small ray
subset of each BASELINE config is enough for the CPU oracle to follow at the REAL volume size - which is what pins the
index arithmetic `((g + 1) / 2) * (S - 1)` at S = 100 / 200 (18 % mismatch risk if simplified):

  * config 2: 2 scenes x 64 of the step's rays, 64 + 64 samples, 100^3 x 128-ch volume, forward + backward, every
    precision mode against the fp32 oracle (fp32: 1e-4 / 3e-4; bf16x3: 1e-3; fp16 / bf16: their measured levels);
  * config 3: one 128-pixel image row of each of the 5 cameras out of a full `rendering()` call (81 920 rays);
  * config 5: 32 rays x (128 + 256) samples on the 200^3 x 128-ch volume (4.1 GB), forward + backward.

And the SURVEY 7.1b probes: the oracle's own ATen ops run ON THE DEVICE (CUDA eager, what the reference executes on
a GPU) against the kernels, stage by stage, recording which stages are bit-identical to CUDA-eager.
"""
import pytest
import torch

from oracle import nerf_oracle as O
from tests.conftest import load_pkg

pytestmark = pytest.mark.gpu
syn = load_pkg("synthetic")


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def bits_equal_frac(a, b):
    a, b = a.detach().cpu().contiguous(), b.detach().cpu().contiguous()
    return float((a.view(torch.int32) == b.view(torch.int32)).float().mean())


@pytest.fixture(scope="module")
def mods():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return load_pkg("ops"), load_pkg("neural_rendering"), load_pkg("utils")


def _renderer(mods, wl, precision, n_rays, **opts):
    ops, NR, U = mods
    cfg = U.default_config(voxel_shape=wl.S, d_latent=wl.C, d_embed=wl.D, n_coarse=wl.n_coarse, n_fine=wl.n_fine,
                           ray_chunk_size=n_rays, image_width=wl.W, image_height=wl.H, **opts)
    ren = NR.NeuralRenderer(cfg, torch.tensor(syn.BOUNDS), precision=precision)
    params = O.init_params(d_in=42, d_latent=wl.C, d_hidden=512, d_out=4 + wl.D, seed=0)
    sd = ren.state_dict()
    for k, v in params.items():
        sd["nerf_model.mlp_coarse." + k].copy_(v)
    return ren.cuda(), params


def _loss(out, gt_rgb, gt_emb, levels=("coarse", "fine")):
    return sum(((out[l]["rgb"] - gt_rgb) ** 2).mean() + 0.01 * ((out[l]["embed"] - gt_emb) ** 2).mean() +
               0.05 * out[l]["depth"].mean() for l in levels)


def _oracle_step(params, vol, rays, wl, noise, gt_rgb, gt_emb, eval_batch_size=4096):
    pr = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    vr = vol.requires_grad_(True)
    ref = O.forward_nerf(pr, vr, rays, syn.BOUNDS, wl.n_coarse, wl.n_fine, noise=noise, eval_batch_size=eval_batch_size)
    _loss(ref, gt_rgb, gt_emb).backward()
    return ref, vr.grad, {k: v.grad for k, v in pr.items()}


def _cuda_step(ren, vol_cuda, rays, noise, gt_rgb, gt_emb):
    for p in ren.parameters():
        p.grad = None
    volc = vol_cuda.detach().requires_grad_(True)
    ren.encode(None, None, None, volc, None, None, None)
    out = ren.forward_nerf(rays.cuda(), want_weights=True, noise={k: v.cuda() for k, v in noise.items()})
    _loss(out, gt_rgb.cuda(), gt_emb.cuda()).backward()
    grads = {k[len("nerf_model.mlp_coarse."):]: v.grad for k, v in ren.named_parameters()
             if k.startswith("nerf_model.mlp_coarse.")}
    return out, volc.grad, grads


# (outputs, depth, volume gradient, parameter gradients) relative-L2 bounds per precision mode against the fp32 oracle
# Measured (config-2 subset): fp32 2-4e-7 / dvoxel 2e-6; bf16x3 coarse <= 1.8e-5, fine 3-5e-4 and gradients 1e-2 (an
# importance-sampling bin that flips moves one fine sample: a discontinuity of the reference algorithm itself; the
# config-5 subset, without a flip, shows 5e-6 / 1e-4); fp16 4-15e-4 / 5e-2; bf16 3-9e-3 / 1.4e-1.
# The oracle runs LIVE on the GPU box's host CPU here (its result depends on the host: MKL partitions the GEMMs by core
# count, 16 on some boxes of the pool and 32 on others), and the reference's gradient is discontinuous in its own
# rounding: one ReLU gate whose pre-activation sits within fp32 noise of zero moved dvoxel / dparam from 2e-6 / 1e-4 to
# 1.5e-3 / 1.6e-3 on one box in ten runs (outputs 6e-6), the same floor DESIGN.md section 2 records for the oracle
# against an fp64 evaluation of itself (3.7e-3).  The fp32 gradient bounds below are therefore 5e-3; the bounds against
# the COMMITTED fixtures (tests/test_gpu_render.py, produced once by the reference) stay at 3e-4.
_BOUNDS = {"fp32": (1e-4, 1e-4, 5e-3, 5e-3), "bf16x3": (1e-3, 1e-3, 2e-2, 2e-2), "fp16": (3e-3, 1e-3, 1.2e-1, 8e-2),
           "bf16": (1.5e-2, 5e-3, 2.5e-1, 2e-1)}


@pytest.fixture(scope="module")
def config2_case(mods):
    """Inputs + the oracle's result for the config-2 subset (computed once, shared by the four precision modes)."""
    wl = syn.CONFIGS["config2"]
    n = 64
    vol = syn.make_volume(wl.SB, wl.C, wl.S, seed=2)                        # (2,128,100,100,100): the real size
    poses = syn.arc_poses(wl.SB)
    rays_all = O.gen_rays(poses, wl.W, wl.H, torch.tensor(wl.focal), 1.2, 4.0).reshape(wl.SB, -1, 8)
    rays = rays_all[:, syn.pick_ray_indices(wl.W * wl.H, wl.rays_per_scene, seed=3)[:n]].contiguous()
    noise = syn.make_noise(wl.SB * n, wl.n_coarse, wl.n_fine, seed=5)
    gt_rgb, gt_emb = syn.make_targets(wl.SB, n, wl.D)
    params = O.init_params(d_in=42, d_latent=wl.C, d_hidden=512, d_out=4 + wl.D, seed=0)
    ref, vg_r, pg_r = _oracle_step(params, vol.clone(), rays, wl, noise, gt_rgb, gt_emb)
    return dict(wl=wl, n=n, vol=vol, rays=rays, noise=noise, gt=(gt_rgb, gt_emb), ref=ref, vg=vg_r, pg=pg_r)


@pytest.mark.parametrize("precision", ["fp32", "bf16x3", "fp16", "bf16"])
def test_config2_ray_subset_matches_the_oracle_at_s100(mods, config2_case, precision):
    c = config2_case
    wl = c["wl"]
    ren, _ = _renderer(mods, wl, precision, c["n"])
    out, vg, pg = _cuda_step(ren, c["vol"].cuda(), c["rays"], c["noise"], *c["gt"])
    ref = c["ref"]
    assert torch.equal(out.coarse.z.cpu(), ref["z_coarse"]), "coarse sample depths must be bit-exact"
    b_out, b_dep, b_vol, b_par = _BOUNDS[precision]
    errs = {(l, k): rel(out[l][k], ref[l][k]) for l in ("coarse", "fine") for k in ("rgb", "embed", "depth")}
    e_vol = rel(vg, c["vg"])
    touched = c["vg"].abs().sum(1) > 0                                      # (SB,S,S,S) voxels the oracle wrote
    e_touched = rel(vg.cpu().permute(0, 2, 3, 4, 1)[touched], c["vg"].permute(0, 2, 3, 4, 1)[touched])
    stray = float(vg.cpu().permute(0, 2, 3, 4, 1)[~touched].abs().max())
    e_par = {k: rel(pg[k], c["pg"][k]) for k in c["pg"]}
    print(f"[config2 S=100 subset, {precision}] " + ", ".join(f"{a}.{b}={v:.1e}" for (a, b), v in errs.items()) +
          f"; dvoxel {e_vol:.1e} (touched voxels {e_touched:.1e}, {int(touched.sum())} of them, stray max {stray:.1e});"
          f" worst dparam {max(e_par.values()):.1e}")
    for (l, k), v in errs.items():
        assert v < (b_dep if k == "depth" else b_out), (l, k, v)
    assert e_vol < b_vol and e_touched < b_vol
    assert stray == 0.0 if precision in ("fp32", "bf16x3") else stray < 1e-3   # same voxels touched (flips aside)
    assert max(e_par.values()) < b_par, e_par


def test_config3_image_rows_match_the_oracle(mods):
    """rendering() over the full config-3 workload (5 cameras x 128 x 128 px, SB forced to 1, 4096-ray chunks); one
    image row per camera is checked against the oracle (perturb off: zero jitter, the fixed u grid)."""
    wl = syn.CONFIGS["config3"]
    vol = syn.make_volume(1, wl.C, wl.S, seed=4)
    poses = syn.arc_poses(wl.n_cams)
    focal = torch.tensor(wl.focal)
    kf = wl.n_fine
    for precision, tol in (("fp32", 1e-4), ("bf16x3", 1e-3), ("fp16", 3e-3)):
        ren, params = _renderer(mods, wl, precision, 4096)
        ren.perturb = False
        ren.eval()
        rgb, emb, dep = ren.rendering(vol.cuda(), None, None, None, None, focal.cuda(), poses.cuda())
        assert rgb.shape == (wl.n_cams, wl.H, wl.W, 3) and emb.shape == (wl.n_cams, wl.H, wl.W, wl.D)
        if precision == "fp32":
            rays_all = O.gen_rays(poses, wl.W, wl.H, focal, 1.2, 4.0)               # (5,H,W,8)
            rows = [17 + 23 * i for i in range(wl.n_cams)]
            rays = torch.stack([rays_all[i, r] for i, r in enumerate(rows)]).reshape(1, -1, 8)   # (1, 5*128, 8)
            u = ((torch.arange(kf, dtype=torch.float32) + 0.5) / kf).repeat(rays.shape[1], 1)
            with torch.no_grad():
                ref = O.forward_nerf(params, vol, rays, syn.BOUNDS, wl.n_coarse, wl.n_fine, noise={"u": u})
        pick = lambda img: torch.stack([img[i, r] for i, r in enumerate(rows)]).reshape(1, -1, *img.shape[3:])
        e = {k: rel(pick(v), ref["fine"][k]) for k, v in (("rgb", rgb), ("embed", emb), ("depth", dep))}
        print(f"[config3 image rows, {precision}] " + ", ".join(f"{k}={v:.1e}" for k, v in e.items()))
        assert max(e.values()) < tol, (precision, e)


def test_config5_ray_subset_matches_the_oracle_at_s200(mods):
    """The stress shape's volume: 200^3 x 128 channels (4.1 GB fp32), 128 + 128 samples; 32 rays forward + backward."""
    wl = syn.CONFIGS["config5"]
    n = 32
    g = torch.Generator(device="cuda").manual_seed(21)
    vol_c = torch.randn(1, wl.C, wl.S, wl.S, wl.S, device="cuda", generator=g) * 0.1
    vol = vol_c.cpu()
    poses = syn.arc_poses(1)
    rays_all = O.gen_rays(poses, wl.W, wl.H, torch.tensor(wl.focal), 1.2, 4.0).reshape(1, -1, 8)
    rays = rays_all[:, syn.pick_ray_indices(wl.W * wl.H, n, seed=8)].contiguous()
    noise = syn.make_noise(n, wl.n_coarse, wl.n_fine, seed=9)
    gt_rgb, gt_emb = syn.make_targets(1, n, wl.D)
    params = O.init_params(d_in=42, d_latent=wl.C, d_hidden=512, d_out=4 + wl.D, seed=0)
    # one chunk: the reference's 4096-point chunks would zero-fill the 4.1 GB gradient volume three times over
    ref, vg_r, pg_r = _oracle_step(params, vol, rays, wl, noise, gt_rgb, gt_emb, eval_batch_size=1 << 20)
    touched = vg_r.abs().sum(1) > 0
    vg_r_t = vg_r.permute(0, 2, 3, 4, 1)[touched].clone()
    del vg_r, vol
    for precision in ("fp32", "bf16x3"):
        ren, _ = _renderer(mods, wl, precision, n)
        out, vg, pg = _cuda_step(ren, vol_c, rays, noise, gt_rgb, gt_emb)
        assert torch.equal(out.coarse.z.cpu(), ref["z_coarse"])
        b_out, b_dep, b_vol, b_par = _BOUNDS[precision]
        errs = {(l, k): rel(out[l][k], ref[l][k]) for l in ("coarse", "fine") for k in ("rgb", "embed", "depth")}
        vg_t = vg.permute(0, 2, 3, 4, 1)[touched.cuda()]
        e_touched = rel(vg_t, vg_r_t)
        total = float(vg.double().norm())
        outside = (total ** 2 - float(vg_t.double().norm()) ** 2)             # gradient energy on untouched voxels
        e_par = max(rel(pg[k], pg_r[k]) for k in pg_r)
        print(f"[config5 S=200 subset, {precision}] " + ", ".join(f"{a}.{b}={v:.1e}" for (a, b), v in errs.items()) +
              f"; dvoxel on {int(touched.sum())} touched voxels {e_touched:.1e}; worst dparam {e_par:.1e}")
        for (l, k), v in errs.items():
            assert v < (b_dep if k == "depth" else b_out), (l, k, v)
        assert e_touched < b_vol and abs(outside) <= 1e-12 * max(total ** 2, 1e-30) + 1e-20
        assert e_par < b_par
        del vg, out, pg, ren
        torch.cuda.empty_cache()


# ------------------------------------------------------------------ against CUDA-eager ATen, on the device
def test_stages_against_the_oracle_run_on_the_device(mods):
    """The reference on a GPU executes ATen's CUDA kernels; the golden fixtures were produced by its CPU kernels.  Here
    the oracle's own ops run on the device and every stage is compared with the kernels: which stages are bit-identical
    to CUDA-eager is printed (and asserted where it must hold)."""
    ops, NR, U = mods
    dev = torch.device("cuda")
    wl = syn.CONFIGS["config2"]
    poses = syn.arc_poses(2).to(dev)
    focal = torch.tensor(wl.focal, device=dev)
    report = {}
    # gen_rays (utils.py:444-506): norm + a K=3 matmul
    r_ref = O.gen_rays(poses, wl.W, wl.H, focal, 1.2, 4.0)
    r = ops.raygen(poses, wl.W, wl.H, focal, 1.2, 4.0)
    report["gen_rays"] = bits_equal_frac(r, r_ref)
    assert torch.equal(r[..., :3], r_ref[..., :3]) and torch.equal(r[..., 6:], r_ref[..., 6:])
    assert float((r - r_ref).abs().max()) <= 2.5e-7
    # ... and in the rounding pattern CUDA-eager executes (pixel * (1 / f), norm as (xx + zz) + yy, cuBLAS' two-accumulator
    # K = 3 product: scripts/raygen_probe.py tried all 64 combinations) every ray is bit-identical, signed zeros included
    same_bits = lambda a, b: torch.equal(a.contiguous().view(torch.int32), b.contiguous().view(torch.int32))
    for P, W_, H_, f_, c_ in ((poses, wl.W, wl.H, focal, None),
                              (torch.eye(4, device=dev)[None].repeat(2, 1, 1), wl.W, wl.H, focal, None),
                              (syn.arc_poses(5).to(dev), 160, 120, torch.tensor([201.3, 199.1], device=dev),
                               torch.tensor([77.2, 61.9], device=dev)),
                              (syn.arc_poses(3).to(dev), 80, 60, torch.tensor(76.18187, device=dev), None)):
        ref_ = O.gen_rays(P, W_, H_, f_, 1.2, 4.0, c=c_)
        got_ = ops.raygen(P, W_, H_, f_, 1.2, 4.0, c=c_, flags=ops.NRF_RAYGEN_CUDA_EAGER)
        assert same_bits(got_, ref_), (W_, H_)
    report["gen_rays_cuda_eager_mode"] = 1.0
    rays = r_ref.reshape(2, -1, 8)[:, syn.pick_ray_indices(wl.W * wl.H, 512, seed=1).to(dev)].reshape(-1, 8).contiguous()
    R = rays.shape[0]
    noise = {k: v.to(dev) for k, v in syn.make_noise(R, 64, 64, seed=2).items()}
    # sample_coarse (neural_rendering.py:159-176): linspace, lerp
    for Kc in (64, 128):
        jit = torch.rand(R, Kc, device=dev)
        report[f"sample_coarse_{Kc}"] = bits_equal_frac(ops.sample_coarse(rays, Kc, jit), O.sample_coarse(rays, Kc, jit))
        assert report[f"sample_coarse_{Kc}"] == 1.0
    z = O.sample_coarse(rays, 64, noise["coarse"])
    # points o + z d, canonical coordinates, grid_sample, positional encoding (models_embed.py:185-277, utils.py:545-557)
    S, C = 100, 128
    g = torch.Generator(device=dev).manual_seed(3)
    vol = torch.randn(2, C, S, S, S, device=dev, generator=g) * 0.1
    pts = (rays[:, None, :3] + z.unsqueeze(2) * rays[:, None, 3:6]).reshape(2, -1, 3)
    canon = O.world_to_canonical(pts, syn.BOUNDS)
    lat_ref = O.trilinear_gather(vol, canon).reshape(-1, C)
    pe_ref = O.positional_encoding(canon.reshape(-1, 3))
    field_in, pts_k = ops.encode_points(rays, z, R // 2, ops.volume_to_channels_last(vol), torch.tensor(syn.BOUNDS),
                                        precision=ops.NRF_PREC_FP32, want_points=True)
    report["points"] = bits_equal_frac(pts_k, pts.reshape(-1, 3))
    assert report["points"] == 1.0
    inside = lat_ref.abs().sum(1) > 0
    report["grid_sample_latent"] = bits_equal_frac(field_in[:, :C][inside], lat_ref[inside])
    report["grid_sample_latent_max_abs"] = float((field_in[:, :C] - lat_ref).abs().max())
    report["samples_inside_box"] = float(inside.float().mean())
    assert report["grid_sample_latent_max_abs"] <= 2e-7 and torch.equal(field_in[:, :C][~inside], lat_ref[~inside])
    # the FMA form of the corner accumulation is what nvcc makes of ATen's CUDA kernel: bit-identical to CUDA-eager
    field_fma = ops.encode_points(rays, z, R // 2, ops.volume_to_channels_last(vol), torch.tensor(syn.BOUNDS),
                                  precision=ops.NRF_PREC_FP32, fma=True)
    report["grid_sample_latent_fma_mode"] = bits_equal_frac(field_fma[:, :C][inside], lat_ref[inside])
    assert report["grid_sample_latent_fma_mode"] == 1.0 and torch.equal(field_fma[:, :C], lat_ref)
    report["positional_encoding_xyz"] = bits_equal_frac(field_in[:, C:C + 3], pe_ref[:, :3])
    assert report["positional_encoding_xyz"] == 1.0                             # canonical coordinates: IEEE sub / div
    report["positional_encoding_sin"] = bits_equal_frac(field_in[:, C + 3:C + 39], pe_ref[:, 3:])
    report["positional_encoding_sin_max_abs"] = float((field_in[:, C + 3:C + 39] - pe_ref[:, 3:]).abs().max())
    assert report["positional_encoding_sin_max_abs"] <= 5e-7
    assert torch.equal(field_in[:, C + 39:C + 42], rays[:, None, 3:6].expand(-1, 64, -1).reshape(-1, 3))
    # compositing (neural_rendering.py:339-359): exp, cumprod, sums
    D = 16
    out = torch.randn(R, 64, 4 + D, device=dev, generator=g)
    out[..., 3] = out[..., 3] * 3
    head = torch.cat([torch.sigmoid(out[..., :3]), torch.relu(out[..., 3:4]), out[..., 4:]], -1)
    w_ref, rgb_ref, emb_ref, dep_ref = O.composite_from_field(head, z, rays[:, -1:])
    w, rgb, emb, dep = ops.composite_fwd(out.reshape(-1, 4 + D), z, rays, D)
    report["composite_weights"] = bits_equal_frac(w, w_ref)
    for name, a, b in (("weights", w, w_ref), ("rgb", rgb, rgb_ref), ("embed", emb, emb_ref), ("depth", dep, dep_ref)):
        assert rel(a, b) < 2e-6, name
    # importance sampling (neural_rendering.py:179-207): sum, cumsum, searchsorted
    cdf = O.fine_cdf(w_ref)
    ind_ref, zf_ref = O.sample_fine_from_cdf(rays, cdf, 64, noise["u"], noise["fine"])
    zf, ind = ops.sample_fine(rays, None, 64, noise["u"], noise["fine"], cdf=cdf, want_ind=True)
    report["fine_ind_given_cdf"] = float((ind == ind_ref).float().mean())
    report["fine_z_given_cdf"] = bits_equal_frac(zf, zf_ref)
    assert report["fine_ind_given_cdf"] == 1.0 and report["fine_z_given_cdf"] == 1.0
    zf2, ind2 = ops.sample_fine(rays, w_ref, 64, noise["u"], noise["fine"], want_ind=True)
    report["fine_ind_in_kernel_cdf"] = float((ind2 == ind_ref).float().mean())
    assert report["fine_ind_in_kernel_cdf"] > 0.997
    # the in-kernel cdf in the association order of ATen's CUDA sum / cumsum (scripts/cdf_probe.py): every index and
    # depth bit-identical to CUDA-eager, for both coarse sample counts of the BASELINE configs
    for Kc_ in (64, 128):
        w_ = w_ref if Kc_ == 64 else (torch.rand(R, Kc_, device=dev, generator=g) ** 4 * 0.2).contiguous()
        u_ = torch.rand(R, 96, device=dev, generator=g)
        j_ = torch.rand(R, 96, device=dev, generator=g)
        ind_e, z_e = O.sample_fine_from_cdf(rays, O.fine_cdf(w_), Kc_, u_, j_)
        z_k, ind_k = ops.sample_fine(rays, w_, Kc_, u_, j_, want_ind=True, cuda_eager=True)
        assert torch.equal(ind_k, ind_e) and same_bits(z_k, z_e), Kc_
    report["fine_in_kernel_cdf_cuda_eager_mode"] = 1.0
    # sort (neural_rendering.py:463)
    z_all = torch.cat([z, zf_ref], -1)
    report["sort"] = bits_equal_frac(ops.sort_rows(z_all.clone()), torch.sort(z_all, dim=-1)[0])
    assert report["sort"] == 1.0
    print("bit-identical fraction vs CUDA-eager ATen on this device: " +
          ", ".join(f"{k}={v:.6g}" for k, v in report.items()))
