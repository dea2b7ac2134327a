"""Per-kernel parity tests: every C-ABI entry point against the oracle (oracle/nerf_oracle.py, torch
CPU fp32) on the same seeded inputs.  Run on the B200 box: `pytest -m gpu`."""
import math

import pytest
import torch

from oracle import nerf_oracle as O
from tests.conftest import golden, load_pkg

pytestmark = pytest.mark.gpu

syn = load_pkg("synthetic")
T = torch.from_numpy


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return load_pkg("ops")


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def bits_equal_frac(a, b):
    a, b = a.cpu().contiguous(), b.cpu().contiguous()
    return float((a.view(torch.int32) == b.view(torch.int32)).float().mean())


# ------------------------------------------------------------------------------------ rays
def test_raygen_matches_reference(ops):
    fx = golden("raygen_pe")
    poses = T(fx["poses"]).cuda()
    r = ops.raygen(poses, 80, 60, 76.18187, 1.2, 4.0).cpu()
    ref = T(fx["rays_60x80"])
    assert torch.equal(r[..., :3], ref[..., :3]) and torch.equal(r[..., 6:], ref[..., 6:])
    frac = bits_equal_frac(r, ref)
    print(f"raygen 60x80: bit-identical fraction vs CPU reference = {frac:.6f}")
    assert (r - ref).abs().max() <= 2.5e-7
    r = ops.raygen(poses[:2], 128, 128, torch.tensor(153.0), 1.2, 4.0).cpu()
    assert (r[:, ::16] - T(fx["rays_128_rows"])).abs().max() <= 2.5e-7
    r = ops.raygen(poses[:1], 128, 128, torch.tensor([150.0, 151.0]), 0.5, 3.0,
                   c=torch.tensor([70.5, 61.25])).cpu()
    assert (r[:, ::32] - T(fx["rays_128_c_rows"])).abs().max() <= 2.5e-7
    assert frac > 0.99
    # intrinsics that live on the GPU are handed over by pointer (no host read-back): same bits as by value
    for f_, c_ in ((torch.tensor(153.0), None), (torch.tensor([150.0, 151.0]), torch.tensor([70.5, 61.25])),
                   (torch.tensor([76.18187]), None)):
        host = ops.raygen(poses[:2], 128, 96, f_, 1.2, 4.0, c=c_)
        dev = ops.raygen(poses[:2], 128, 96, f_.cuda(), 1.2, 4.0, c=None if c_ is None else c_.cuda())
        assert torch.equal(host, dev)


@pytest.mark.parametrize("Kc,lindisp,jit", [(64, False, False), (64, False, True), (128, False, True),
                                            (96, False, True), (64, True, True), (17, False, True)])
def test_sample_coarse_bit_exact(ops, Kc, lindisp, jit):
    g = torch.Generator().manual_seed(Kc)
    R = 333
    rays = torch.rand(R, 8, generator=g)
    rays[:, 6] = 0.5 + rays[:, 6]
    rays[:, 7] = 3.0 + rays[:, 7]
    jitter = torch.rand(R, Kc, generator=g) if jit else None
    ref = O.sample_coarse(rays, Kc, jitter, lindisp)
    out = ops.sample_coarse(rays.cuda(), Kc, jitter.cuda() if jit else None, lindisp).cpu()
    if Kc in (64, 128):
        assert torch.equal(out, ref)          # linspace == i/Kc exactly
    else:                                     # linspace bits may differ between CPU and CUDA ATen
        assert (out - ref).abs().max() <= 5e-7


@pytest.mark.parametrize("Kc,Kf,lindisp", [(64, 64, False), (64, 32, False), (128, 128, False), (16, 12, True)])
def test_sample_fine_bit_exact_given_cdf(ops, Kc, Kf, lindisp):
    g = torch.Generator().manual_seed(Kc + Kf)
    R = 257
    rays = torch.rand(R, 8, generator=g)
    rays[:, 6], rays[:, 7] = 1.2, 4.0
    w = torch.rand(R, Kc, generator=g) ** 4
    w[::7] = 0.0                                               # degenerate rays: uniform pdf
    u = torch.rand(R, Kf, generator=g)
    u[0, 0], u[1, 0] = 0.0, 0.99999994                          # edge draws (no upper clamp)
    jitter = torch.rand(R, Kf, generator=g)
    cdf = O.fine_cdf(w)
    ind_ref, z_ref = O.sample_fine_from_cdf(rays, cdf, Kc, u, jitter, lindisp)
    z, ind = ops.sample_fine(rays.cuda(), None, Kc, u.cuda(), jitter.cuda(), lindisp, cdf=cdf.cuda(),
                             want_ind=True)
    assert torch.equal(ind.cpu(), ind_ref)
    assert torch.equal(z.cpu(), z_ref)
    # cdf built in-kernel from the weights: identical up to rare ulp flips at bin edges
    z2, ind2 = ops.sample_fine(rays.cuda(), w.cuda(), Kc, u.cuda(), jitter.cuda(), lindisp, want_ind=True)
    mism = float((ind2.cpu() != ind_ref).float().mean())
    print(f"sample_fine Kc={Kc}: bin mismatches with in-kernel cdf = {mism:.2e}")
    assert mism < 2e-3
    same = ind2.cpu() == ind_ref
    assert torch.equal(z2.cpu()[same], z_ref[same])


@pytest.mark.parametrize("K", [128, 96, 112, 256, 33])
def test_sort_rows(ops, K):
    g = torch.Generator().manual_seed(K)
    z = torch.rand(301, K, generator=g)
    z[:, 5] = z[:, 3]                                          # ties
    ref, _ = torch.sort(z, dim=-1)
    out, perm = ops.sort_rows(z.cuda().clone(), want_perm=True)
    assert torch.equal(out.cpu(), ref)
    assert torch.equal(torch.gather(z, 1, perm.cpu().long()), ref)
    assert torch.equal(torch.sort(perm.cpu().long(), dim=1)[0], torch.arange(K).expand(301, K))


# ---------------------------------------------------------------------------------- volume
@pytest.mark.parametrize("C,dims", [(20, (7, 9, 11)), (128, (7, 9, 11)), (256, (8, 8, 8)), (128, (5, 3, 2))])
def test_volume_transpose_roundtrip(ops, C, dims):
    """C % 128 == 0 takes the 128-channel x 32-voxel kernel (ragged voxel tails included), other widths the 32x32 one."""
    g = torch.Generator().manual_seed(0)
    v = torch.randn(2, C, *dims, generator=g).cuda()
    cl = ops.volume_to_channels_last(v)
    assert torch.equal(cl, v.permute(0, 2, 3, 4, 1).contiguous())
    assert torch.equal(ops.volume_to_channels_first(cl), v)


def _scene_inputs(SB, C, S, R_per, K, seed=0):
    g = torch.Generator().manual_seed(seed)
    vol = syn.make_volume(SB, C, S, seed=seed)
    poses = syn.arc_poses(SB)
    rays = O.gen_rays(poses, 32, 32, torch.tensor(38.0), 1.2, 4.0).reshape(SB, -1, 8)
    idx = torch.randint(32 * 32, (R_per,), generator=g)
    rays = rays[:, idx].reshape(-1, 8).contiguous()
    z = O.sample_coarse(rays, K, torch.rand(SB * R_per, K, generator=g))
    return vol, rays, z


@pytest.mark.parametrize("S,fma,C", [(24, False, 128), (100, False, 128), (7, True, 128), (17, False, 64), (12, True, 64)])
def test_encode_points_tma_box_gather_is_bit_identical(ops, monkeypatch, S, fma, C):
    """encode_points_tma_kernel (NRF_ENCODE_TMA=1: a sample's eight corners as one 5-D TMA box, out-of-grid corners
    zero-filled by the TMA unit) against encode_points_w32_kernel, which computes corner offsets and skips corners
    outside the grid: every latent, tail value and touch flag bit for bit - rays that cross the faces of the box, end
    inside it, or miss it; all three output types."""
    SB, R_per, K = 2, 70, 48
    vol, rays, z = _scene_inputs(SB, C, S, R_per, K, seed=S)
    g = torch.Generator().manual_seed(S)
    b = torch.tensor(syn.BOUNDS)
    n_in = 40                                   # ... plus rays that start inside the box and run through its faces
    o = b[:3] + (b[3:] - b[:3]) * torch.rand(SB * n_in, 3, generator=g)
    d = torch.nn.functional.normalize(torch.randn(SB * n_in, 3, generator=g), dim=-1) * 0.9
    inside = torch.cat([o, d, torch.zeros(SB * n_in, 1), torch.ones(SB * n_in, 1)], 1)
    rays = torch.cat([rays.view(SB, R_per, 8), inside.view(SB, n_in, 8)], 1).reshape(-1, 8).contiguous()
    z = torch.cat([z.view(SB, R_per, K), torch.rand(SB, n_in, K, generator=g).sort(-1)[0]], 1).reshape(-1, K).contiguous()
    R_per += n_in
    vol_cl = ops.volume_to_channels_last(vol.cuda())
    for prec in (ops.NRF_PREC_FP32, ops.NRF_PREC_BF16, ops.NRF_PREC_FP16):
        res = {}
        for mode in ("0", "1"):
            monkeypatch.setenv("NRF_ENCODE_TMA", mode)
            res[mode] = ops.encode_points(rays.cuda(), z.cuda(), R_per, vol_cl, syn.BOUNDS, precision=prec, fma=fma,
                                          want_points=True, want_touch=True)
        for x, y in zip(res["0"], res["1"]):
            assert torch.equal(x, y)
        lat = res["0"][0][:, :C].float()
        assert 0.05 < float((lat.abs().sum(1) > 0).float().mean()) < 0.95      # samples inside AND outside


@pytest.mark.parametrize("C,S", [(128, 24), (16, 12), (64, 17)])
def test_encode_points_fp32_matches_oracle(ops, C, S):
    SB, R_per, K = 2, 50, 24
    vol, rays, z = _scene_inputs(SB, C, S, R_per, K, seed=C)
    pts = rays[:, None, :3] + z.unsqueeze(2) * rays[:, None, 3:6]
    dirs = rays[:, None, 3:6].expand(-1, K, -1)
    ref = O.field(None, vol, pts.reshape(SB, -1, 3), dirs.reshape(SB, -1, 3), syn.BOUNDS,
                  return_mlp_input=True)
    vol_cl = ops.volume_to_channels_last(vol.cuda())
    out, p = ops.encode_points(rays.cuda(), z.cuda(), R_per, vol_cl, syn.BOUNDS,
                               precision=ops.NRF_PREC_FP32, want_points=True)
    out, p = out.cpu(), p.cpu()
    assert torch.equal(p, pts.reshape(-1, 3)), "sample positions must be bit-exact"
    lat_frac = bits_equal_frac(out[:, :C], ref[:, :C])
    print(f"encode C={C}: latent bit-identical fraction = {lat_frac:.6f}")
    assert (out[:, :C] - ref[:, :C]).abs().max() <= 1e-6
    assert lat_frac > 0.999
    assert torch.equal(out[:, C:C + 3], ref[:, C:C + 3])                 # canonical xyz: bit-exact
    assert (out[:, C + 3:C + 39] - ref[:, C + 3:C + 39]).abs().max() <= 5e-7   # sin(): ulp-level
    assert torch.equal(out[:, C + 39:C + 42], ref[:, C + 39:C + 42])     # view direction
    assert (out[:, C + 42:] == 0).all()
    # bf16 operand mode is the fp32 row rounded to nearest-even
    out16 = ops.encode_points(rays.cuda(), z.cuda(), R_per, vol_cl, syn.BOUNDS, precision=ops.NRF_PREC_BF16)
    assert torch.equal(out16.cpu(), out.to(torch.bfloat16))


@pytest.mark.parametrize("C", [8, 64, 128])       # 8: one warp per sample; 64 / 128: 32 samples per warp iteration
def test_encode_points_outside_box_and_edges(ops, C):
    """Zero padding outside the grid, exact hits on the faces, far-away points."""
    S = 5
    g = torch.Generator().manual_seed(1)
    vol = torch.randn(1, C, S, S, S, generator=g)
    b = torch.tensor(syn.BOUNDS)
    canon = torch.tensor([[0., 0., 0.], [1., 1., 1.], [0.5, 0.5, 0.5], [-0.01, 0.5, 0.5], [1.3, 0.3, 0.1],
                          [0.25, 0.5, 0.75], [50., -50., 3.], [1.0, 0.0, 0.5]])
    world = canon * (b[3:] - b[:3]) + b[:3]
    n = world.shape[0]
    rays = torch.zeros(n, 8)
    rays[:, :3] = world
    rays[:, 3:6] = torch.tensor([0., 0., 1.])
    rays[:, 6], rays[:, 7] = 0.0, 1.0
    z = torch.zeros(n, 1)
    ref = O.field(None, vol, world.reshape(1, n, 3), rays[:, 3:6].reshape(1, n, 3), syn.BOUNDS,
                  return_mlp_input=True)
    out = ops.encode_points(rays.cuda(), z.cuda(), n, ops.volume_to_channels_last(vol.cuda()), syn.BOUNDS,
                            precision=ops.NRF_PREC_FP32).cpu()
    assert (out[:, :C] - ref[:, :C]).abs().max() <= 1e-6
    assert (out[6, :C] == 0).all() and (out[4, :C] == 0).all()


def test_scatter_volume_grad_matches_autograd(ops):
    SB, C, S, R_per, K = 2, 32, 10, 40, 16
    vol, rays, z = _scene_inputs(SB, C, S, R_per, K, seed=5)
    vol.requires_grad_(True)
    pts = (rays[:, None, :3] + z.unsqueeze(2) * rays[:, None, 3:6]).reshape(SB, -1, 3)
    lat = O.trilinear_gather(vol, O.world_to_canonical(pts, syn.BOUNDS))
    g = torch.Generator().manual_seed(9)
    dl = torch.randn(lat.shape, generator=g)
    lat.backward(dl)
    grad_cl = torch.zeros(SB, S, S, S, C, device="cuda")
    ops.scatter_volume_grad(rays.cuda(), z.cuda(), R_per, dl.reshape(-1, C).cuda(), grad_cl, syn.BOUNDS)
    got = ops.volume_to_channels_first(grad_cl).cpu()
    assert rel(got, vol.grad) < 1e-6


# ------------------------------------------------------------------------------ compositing
@pytest.mark.parametrize("K,D,white,noise_std", [(64, 384, False, 0.0), (128, 384, False, 0.0), (96, 24, True, 0.0),
                                                 (40, 8, False, 0.0), (128, 384, True, 0.0), (80, 512, False, 0.0),
                                                 (256, 384, False, 0.0), (128, 384, True, 1.5), (48, 24, False, 1.5)])
def test_composite_fwd_bwd_matches_oracle(ops, K, D, white, noise_std):
    """noise_std > 0: the training-time density noise of neural_rendering.py:336-337 (both ReLUs gate the gradient)."""
    g = torch.Generator().manual_seed(K + D)
    R = 37
    raw = torch.randn(R, K, 4 + D, generator=g)
    raw[..., 3] = raw[..., 3] * 3.0                      # sigma: mix of <0 (gated) and large values
    raw.requires_grad_(True)
    rays = torch.rand(R, 8, generator=g)
    rays[:, 6], rays[:, 7] = 1.2, 4.0
    z = O.sample_coarse(rays, K, torch.rand(R, K, generator=g)).requires_grad_(True)
    act = torch.cat([torch.sigmoid(raw[..., :3]), torch.relu(raw[..., 3:4]), raw[..., 4:]], -1)
    sn = torch.randn(R, K, generator=g) * noise_std if noise_std > 0 else None
    snc = sn.cuda() if sn is not None else None
    w, rgb, emb, dep = O.composite_from_field(act, z, rays[:, -1:], white_bkgd=white, sigma_noise=sn)
    d_rgb, d_emb = torch.randn(R, 3, generator=g), torch.randn(R, D, generator=g)
    d_dep, d_w = torch.randn(R, generator=g), torch.randn(R, K, generator=g) * 0.1
    (rgb * d_rgb).sum().add((emb * d_emb).sum()).add((dep * d_dep).sum()).add((w * d_w).sum()).backward()

    raw_c = raw.detach().reshape(R * K, 4 + D).cuda()
    w2, rgb2, emb2, dep2 = ops.composite_fwd(raw_c, z.detach().cuda(), rays.cuda(), D, white, sigma_noise=snc)
    assert rel(w2, w.detach()) < 2e-6 and rel(rgb2, rgb.detach()) < 2e-6
    assert rel(emb2, emb.detach()) < 2e-6 and rel(dep2, dep.detach()) < 2e-6
    dfield, dz = ops.composite_bwd(raw_c, z.detach().cuda(), rays.cuda(), D, d_rgb.cuda(), d_emb.cuda(),
                                   d_dep.cuda(), d_w.cuda(), precision=ops.NRF_PREC_FP32, white_bkgd=white,
                                   want_dz=True, sigma_noise=snc)
    ldg = dfield.shape[1]
    assert ldg % 64 == 0 and (dfield[:, 4 + D:] == 0).all()
    assert rel(dfield[:, :4 + D].reshape(R, K, -1), raw.grad) < 1e-5
    assert rel(dz, z.grad) < 1e-5
    df16 = ops.composite_bwd(raw_c, z.detach().cuda(), rays.cuda(), D, d_rgb.cuda(), d_emb.cuda(),
                             d_dep.cuda(), d_w.cuda(), precision=ops.NRF_PREC_BF16, white_bkgd=white,
                             sigma_noise=snc)
    assert torch.equal(df16.cpu(), dfield.cpu().to(torch.bfloat16))


# ------------------------------------------------------------------------------------ GEMMs
def _bf16r(t):
    return t.to(torch.bfloat16).to(torch.float32)


@pytest.mark.parametrize("M,N,K1,K2,K3", [(128, 256, 64, 0, 0), (300, 512, 512, 0, 0), (1000, 512, 512, 128, 0),
                                          (257, 128, 512, 512, 512), (4096 + 77, 512, 192, 0, 0),
                                          (130, 512, 448, 0, 0), (148 * 128 * 3 + 5, 512, 512, 0, 0)])
def test_gemm_tc_against_fp32_matmul_of_bf16_operands(ops, M, N, K1, K2, K3):
    g = torch.Generator().manual_seed(M + N + K1 + K2)
    K = K1 + K2 + K3
    As = [torch.randn(M, k, generator=g).to(torch.bfloat16) for k in (K1, K2, K3) if k]
    B = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(torch.bfloat16)
    bias = torch.randn(N, generator=g)
    resid = torch.randn(M, N, generator=g).to(torch.bfloat16)
    mask = torch.randn(M, N, generator=g).relu().to(torch.bfloat16)
    acc = torch.cat([a.float() for a in As], 1) @ B.float().t()
    Ac = [a.cuda() for a in As] + [None, None]
    kw = dict(A2=Ac[1], A3=Ac[2])

    # plain fp32 output (the lin_out / dL/dz form)
    out = torch.empty(M, N, device="cuda")
    ops.gemm(Ac[0], B.cuda(), out_f32=out, **kw)
    assert rel(out, acc) < 2e-6, "accumulation must be fp32-grade"
    # forward-style epilogue: bias + residual updated in place + relu'd second copy
    xres = resid.clone().cuda()
    act = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    ops.gemm(Ac[0], B.cuda(), bias=bias.cuda(), resid=xres, out_act=xres, out_act2=act, relu_act2=True, **kw)
    want = acc + bias + resid.float()
    assert rel(xres.float(), want) < 3e-3                      # one bf16 rounding of the result
    assert torch.equal(act.cpu(), xres.cpu().float().relu().to(torch.bfloat16)) or \
        rel(act.float(), want.relu()) < 3e-3
    assert (xres.cpu().float() - want).abs().max() <= want.abs().max() * 2 ** -8
    # dgrad-style epilogue: ReLU gate from a saved activation + residual
    act2 = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    ops.gemm(Ac[0], B.cuda(), mask_src=mask.cuda(), resid=resid.cuda(), out_act=act2, **kw)
    want2 = torch.where(mask.float() > 0, acc, torch.zeros_like(acc)) + resid.float()
    assert rel(act2.float(), want2) < 3e-3
    gated = (mask.float() <= 0)
    assert torch.equal(act2.cpu()[gated], resid[gated])        # gated entries pass the residual through exactly
    # relu'd single output (fc_0 form)
    act3 = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    ops.gemm(Ac[0], B.cuda(), bias=bias.cuda(), out_act=act3, relu_act=True, **kw)
    assert rel(act3.float(), (acc + bias).relu()) < 3e-3


def test_gemm_tc_partial_store(ops):
    """lin_out: N padded to 512, only 388 columns stored into a 388-wide fp32 matrix."""
    g = torch.Generator().manual_seed(3)
    M, K, N, ns = 333, 512, 512, 388
    A = torch.randn(M, K, generator=g).to(torch.bfloat16)
    B = torch.zeros(N, K)
    B[:ns] = torch.randn(ns, K, generator=g) / math.sqrt(K)
    B = B.to(torch.bfloat16)
    bias = torch.zeros(N)
    bias[:ns] = torch.randn(ns, generator=g)
    out = torch.full((M, ns), 7.0, device="cuda")
    ops.gemm(A.cuda(), B.cuda(), bias=bias.cuda(), out_f32=out, n_store=ns)
    assert rel(out, A.float() @ B.float().t()[:, :ns] + bias[:ns]) < 2e-6


@pytest.mark.parametrize("M,N,K,nv,kv", [(64, 128, 64, 128, 64), (1000, 512, 512, 512, 512),
                                         (5000, 448, 512, 388, 512), (777, 512, 128, 512, 128),
                                         (20000, 512, 64, 512, 42), (4099, 512, 256, 512, 256)])
def test_wgrad_tc(ops, M, N, K, nv, kv):
    g = torch.Generator().manual_seed(M + N + K)
    G = torch.randn(M, N, generator=g).to(torch.bfloat16)
    A = torch.randn(M, K, generator=g).to(torch.bfloat16)
    dW = torch.ones(nv, kv, device="cuda")
    db = torch.ones(nv, device="cuda")
    ops.wgrad(G.cuda(), A.cuda(), dW, db, n_valid=nv, k_valid=kv)
    want = (G.float().t() @ A.float())[:nv, :kv] + 1.0
    assert rel(dW, want) < 5e-6
    assert rel(db, G.float().sum(0)[:nv] + 1.0) < 5e-6
    # deterministic mode: ordered reduction of the sample splits, bit-identical run to run
    outs = []
    for _ in range(2):
        dW2 = torch.ones(nv, kv, device="cuda")
        db2 = torch.ones(nv, device="cuda")
        ops.wgrad(G.cuda(), A.cuda(), dW2, db2, n_valid=nv, k_valid=kv, deterministic=True)
        outs.append((dW2, db2))
    assert rel(outs[0][0], want) < 5e-6 and rel(outs[0][1], G.float().sum(0)[:nv] + 1.0) < 5e-6
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])


@pytest.mark.parametrize("M,N,K1,K2", [(100, 64, 58, 0), (513, 388, 512, 0), (300, 512, 512, 128), (65, 70, 33, 16)])
def test_gemm_simt_fp32(ops, M, N, K1, K2):
    g = torch.Generator().manual_seed(M + N)
    A1 = torch.randn(M, K1, generator=g)
    A2 = torch.randn(M, K2, generator=g) if K2 else None
    B = torch.randn(N, K1 + K2, generator=g) / math.sqrt(K1 + K2)
    bias, resid = torch.randn(N, generator=g), torch.randn(M, N, generator=g)
    mask = torch.randn(M, N, generator=g).relu()
    A = A1 if A2 is None else torch.cat([A1, A2], 1)
    acc = (A.double() @ B.double().t()).float()
    out = torch.empty(M, N, device="cuda")
    act = torch.empty(M, N, device="cuda")
    act2 = torch.empty(M, N, device="cuda")
    ops.gemm(A1.cuda(), B.cuda(), A2=None if A2 is None else A2.cuda(), bias=bias.cuda(), mask_src=mask.cuda(),
             resid=resid.cuda(), out_f32=out, out_act=act, out_act2=act2, relu_act2=True,
             precision=ops.NRF_PREC_FP32)
    want = torch.where(mask > 0, acc + bias, torch.zeros_like(acc)) + resid
    assert rel(out, want) < 1e-6 and rel(act, want) < 1e-6
    assert rel(act2, want.relu()) < 1e-6
    dW = torch.zeros(N, K1, device="cuda")
    db = torch.zeros(N, device="cuda")
    G = torch.randn(M, N, generator=g)
    ops.wgrad(G.cuda(), A1.cuda(), dW, db, precision=ops.NRF_PREC_FP32)
    assert rel(dW, (G.double().t() @ A1.double()).float()) < 1e-6
    assert rel(db, G.sum(0)) < 1e-6


def test_scatter_sorted_is_atomics_free_and_reproducible(ops):
    """Counting-sort scatter: equals the autograd of the gather, bit-identical run to run, accumulate mode."""
    SB, C, S, R_per, K = 2, 128, 12, 300, 48          # many samples per voxel: long lists, ties, both scenes
    vol, rays, z = _scene_inputs(SB, C, S, R_per, K, seed=11)
    vol.requires_grad_(True)
    pts = (rays[:, None, :3] + z.unsqueeze(2) * rays[:, None, 3:6]).reshape(SB, -1, 3)
    lat = O.trilinear_gather(vol, O.world_to_canonical(pts, syn.BOUNDS))
    dl = torch.randn(lat.shape, generator=torch.Generator().manual_seed(9))
    lat.backward(dl)
    dlc = dl.reshape(-1, C).cuda()
    outs = []
    for _ in range(3):
        g = torch.full((SB, S, S, S, C), 7.0, device="cuda")         # garbage: every row must be rewritten
        ops.scatter_volume_grad_sorted(rays.cuda(), z.cuda(), R_per, dlc, g, syn.BOUNDS)
        outs.append(g)
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2]), "must be bit-reproducible"
    got = ops.volume_to_channels_first(outs[0]).cpu()
    assert rel(got, vol.grad) < 1e-6
    g2 = outs[0].clone()
    ops.scatter_volume_grad_sorted(rays.cuda(), z.cuda(), R_per, dlc, g2, syn.BOUNDS, accumulate=True)
    assert rel(g2, 2 * outs[0]) < 1e-6
    ga = torch.zeros(SB, S, S, S, C, device="cuda")
    ops.scatter_volume_grad(rays.cuda(), z.cuda(), R_per, dlc, ga, syn.BOUNDS)
    assert rel(ga, outs[0]) < 1e-6


@pytest.mark.parametrize("C,channels_first,S", [(128, True, 11), (128, True, 12), (64, True, 12), (64, True, 9),
                                                 (128, False, 11)])      # S = 12: V % 32 == 0, vector write-out
def test_scatter_merged_two_passes_one_sort(ops, C, channels_first, S):
    """nrf_scatter_volume_grad_merged: coarse + fine pass in one counting sort, gradient written once in the caller's
    layout; equals the autograd of both gathers, every element rewritten, bit-identical run to run."""
    SB, R_per, Ka, Kb = 2, 150, 16, 40
    vol, rays, za = _scene_inputs(SB, C, S, R_per, Ka, seed=21)
    zb = O.sample_coarse(rays, Kb, torch.rand(SB * R_per, Kb, generator=torch.Generator().manual_seed(3)))
    vol.requires_grad_(True)
    dls = []
    for i, z in enumerate((za, zb)):
        pts = (rays[:, None, :3] + z.unsqueeze(2) * rays[:, None, 3:6]).reshape(SB, -1, 3)
        lat = O.trilinear_gather(vol, O.world_to_canonical(pts, syn.BOUNDS))
        dl = torch.randn(lat.shape, generator=torch.Generator().manual_seed(30 + i))
        lat.backward(dl)
        dls.append(dl.reshape(-1, C).cuda())
    shape = (SB, C, S, S, S) if channels_first else (SB, S, S, S, C)
    outs = []
    for _ in range(2):
        g = torch.full(shape, 7.0, device="cuda")
        ops.scatter_volume_grad_merged(rays.cuda(), R_per, [(za.cuda(), dls[0]), (zb.cuda(), dls[1])], g,
                                       channels_first, syn.BOUNDS)
        outs.append(g)
    assert torch.equal(outs[0], outs[1]), "must be bit-reproducible"
    got = outs[0] if channels_first else outs[0].permute(0, 4, 1, 2, 3)
    assert rel(got.cpu(), vol.grad) < 1e-6
    # single pass = the per-pass sorted scatter, bit for bit in channels-last (same entry order)
    g1 = torch.full(shape, 7.0, device="cuda")
    ops.scatter_volume_grad_merged(rays.cuda(), R_per, [(za.cuda(), dls[0])], g1, channels_first, syn.BOUNDS)
    gs = torch.empty(SB, S, S, S, C, device="cuda")
    ops.scatter_volume_grad_sorted(rays.cuda(), za.cuda(), R_per, dls[0], gs, syn.BOUNDS)
    g1_cl = g1.permute(0, 2, 3, 4, 1) if channels_first else g1
    assert torch.equal(g1_cl, gs)


@pytest.mark.parametrize("C,S,R,K", [(128, 16, 96, 48), (64, 16, 300, 40), (128, 8, 200, 64), (128, 32, 64, 96)])
def test_scatter_batched_reduce_equals_the_one_voxel_per_warp_kernel(ops, monkeypatch, C, S, R, K):
    """scatter_reduce_cf2_kernel (a warp sums 8 consecutive voxels from one batched read of their entries, rows as one
    vector per lane) against scatter_reduce_cf_kernel (NRF_SCATTER_RCF=1), rays INSIDE the box: a few to a few hundred
    entries per voxel, so batches of whole voxels, voxels that split a batch, empty voxels between them and the
    more-than-a-warp fallback all occur.  Same ascending-entry fmaf chain per channel: bit for bit."""
    g = torch.Generator().manual_seed(S * 1000 + C)
    b = torch.tensor(syn.BOUNDS)
    o = (b[:3] + (b[3:] - b[:3]) * (0.15 + 0.7 * torch.rand(2 * R, 3, generator=g)))
    d = torch.nn.functional.normalize(torch.randn(2 * R, 3, generator=g), dim=-1) * 0.25
    rays = torch.cat([o, d, torch.zeros(2 * R, 1), torch.ones(2 * R, 1)], 1).cuda()
    za = torch.rand(2 * R, K, generator=g).sort(-1)[0].cuda()
    zb = torch.rand(2 * R, K + 8, generator=g).sort(-1)[0].cuda()
    dla = torch.randn(2 * R * K, C, generator=g).cuda()
    dlb = torch.randn(2 * R * (K + 8), C, generator=g).cuda()
    outs = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("NRF_SCATTER_RCF", mode)
        grad = torch.full((2, C, S, S, S), 7.0, device="cuda")
        _, counts = ops.scatter_volume_grad_merged(rays, R, [(za, dla), (zb, dlb)], grad, True, b, want_counts=True)
        outs[mode] = grad
    assert torch.equal(outs["0"], outs["1"])
    assert float(outs["0"].abs().sum()) > 0
    c = counts.cpu()
    print(f"S={S} C={C}: entries per touched voxel mean {float(c[c > 0].float().mean()):.1f}, max {int(c.max())}, "
          f"touched {float((c > 0).float().mean()):.2f} of the voxels")


# ------------------------------------------------------------------ whole-MLP fused forward kernel
def _bf16_mlp(C=128, D=384, seed=0):
    NR = load_pkg("neural_rendering")
    mlp = NR.ResnetFC(d_in=42, d_out=4 + D, n_blocks=5, d_latent=C, d_hidden=512, combine_layer=3)
    syn.init_mlp_(mlp, seed=seed)
    g = torch.Generator().manual_seed(5 + seed)
    with torch.no_grad():
        for n, p in mlp.named_parameters():          # non-zero biases: the bias path must be exercised
            if n.endswith(".bias"):
                p.copy_(torch.randn(p.shape, generator=g) * 0.1)
    return mlp.cuda()


@pytest.mark.parametrize("C,N,D", [(128, 256, 384), (128, 1000, 384), (128, 40000, 384), (64, 777, 384), (128, 1, 384),
                                   (64, 1000, 512), (128, 40000, 512), (64, 3, 512)])
def test_fused_mlp_forward_is_bit_identical_to_the_layered_chain(ops, C, N, D):
    """csrc/mlp_fused.cu (one persistent tcgen05 kernel, activations on chip) against the layer-by-layer GEMM chain
    on the same packed weights: raw outputs and every operand saved for the backward, ragged N included; the
    inference variant (nothing saved) gives the same outputs.  D = 512 (nerfact.conf:22): lin_out has 516 outputs =
    five 128-wide chunks instead of four."""
    mlp = _bf16_mlp(C=C, D=D)
    h = mlp.handle(ops.NRF_PREC_BF16)
    assert h.fused, "the BASELINE shape must take the fused kernel"
    g = torch.Generator().manual_seed(N)
    fin = torch.zeros(N, h.sizes.kin_pad, dtype=torch.bfloat16)
    fin[:, :C + 42] = (torch.randn(N, C + 42, generator=g) * 0.5).to(torch.bfloat16)
    fin = fin.cuda()
    out_l, acts_l = h.forward(fin, layered=True)
    out_f, acts_f = h.forward(fin)
    out_i, acts_i = h.forward(fin, keep_acts=False)
    assert acts_i is None
    assert torch.equal(out_f, out_l) and torch.equal(out_i, out_l)
    n = 11 * N * 512
    assert torch.equal(acts_f.view(torch.bfloat16)[:n], acts_l.view(torch.bfloat16)[:n])
    # and against fp32 math on the bf16-rounded operands (independent of both CUDA paths)
    p = {k: v.detach().float().cpu() for k, v in mlp.named_parameters()}
    ref = O.resnetfc(p, fin[:, :C + 42].float().cpu(), d_latent=C, operand_dtype=torch.bfloat16)
    ref = ref[0] if isinstance(ref, tuple) else ref
    assert rel(out_f, ref) < 3e-2


@pytest.mark.parametrize("C,N,D,prec", [(128, 100000, 384, "bf16"), (128, 77777, 384, "fp16"), (64, 50000, 384, "bf16"),
                                        (128, 40001, 512, "bf16"), (64, 3000, 512, "fp16"), (128, 200, 384, "bf16")])
def test_fused_forward_skips_latent_panels_of_tiles_outside_the_grid(ops, C, N, D, prec):
    """nrf_mlp_fwd_touch: a 256-sample tile whose flags say "no sample has a corner inside the grid" has an all-zero
    latent; the fused forward neither loads nor multiplies the k-panels that meet it (two of the first layer's three,
    the lin_z tails of fc_1; with a 64-channel latent one panel of a mixed unit).  Raw outputs, every saved operand and
    the gate bits equal the unflagged run bit for bit - tiles that are dead, live, and dead only in part (those must
    not be skipped), training and inference variants."""
    mlp = _bf16_mlp(C=C, D=D, seed=3)
    precision = ops.NRF_PREC_BF16 if prec == "bf16" else ops.NRF_PREC_FP16
    h = mlp.handle(precision)
    assert h.fused
    dt = torch.bfloat16 if prec == "bf16" else torch.float16
    g = torch.Generator().manual_seed(N)
    fin = torch.zeros(N, h.sizes.kin_pad, dtype=dt)
    fin[:, :C + 42] = (torch.randn(N, C + 42, generator=g) * 0.5).to(dt)
    # runs of samples without a latent, the way rays that miss the box produce them: whole 32-sample groups
    groups = (N + 31) // 32
    live = torch.rand(groups, generator=g) < 0.5
    run = torch.rand(groups // 24 + 1, generator=g) < 0.6                  # long dead stretches (whole tiles) ...
    live &= ~run.repeat_interleave(24)[:groups]
    live[-1] = N % 256 == 200                                              # ... and a ragged last tile either way
    rows_live = live.repeat_interleave(32)[:N]
    fin[~rows_live, :C] = 0
    fin = fin.cuda()
    touch = live.to(torch.uint8).cuda()
    tiles = (N + 255) // 256
    pad = torch.zeros(tiles * 8, dtype=torch.bool)
    pad[:groups] = live
    dead_tiles = int((~pad.view(tiles, 8).any(1)).sum())
    assert dead_tiles > 0 or N < 1000
    out0, acts0 = h.forward(fin)
    out1, acts1 = h.forward(fin, touch=touch)
    outi, _ = h.forward(fin, keep_acts=False, touch=touch)
    print(f"C={C} N={N}: {dead_tiles} of {tiles} tiles dead")
    assert torch.equal(out0, out1) and torch.equal(out0, outi)
    n_ops = 11 * N * 512 * 2                                               # bytes: 11 saved operands, then the gate bits
    n_gate = 11 * N * 64
    assert torch.equal(acts0[:n_ops + n_gate], acts1[:n_ops + n_gate])


@pytest.mark.parametrize("C,N,D", [(128, 256, 384), (128, 1000, 384), (128, 33000, 384), (64, 777, 384), (128, 3, 384),
                                   (64, 1000, 512), (128, 33000, 512), (64, 5, 512)])
def test_fused_mlp_backward_is_bit_identical_to_the_layered_chain(ops, C, N, D):
    """The fused data-gradient kernel (residual gradient in registers, bit-packed ReLU gates written by the fused
    forward) + weight-gradient GEMMs against the per-layer backward on the same saved operands: dL/dlatent and, with
    the ordered split reduction, every parameter gradient.  D = 512: d_field has 576 columns = 9 k-panels, six of them
    in the shared-memory operand panel and three through the weight ring."""
    NR = load_pkg("neural_rendering")
    mlp = _bf16_mlp(C=C, D=D, seed=1)
    h = mlp.handle(ops.NRF_PREC_BF16)
    g = torch.Generator().manual_seed(N + 1)
    fin = torch.zeros(N, h.sizes.kin_pad, dtype=torch.bfloat16)
    fin[:, :C + 42] = (torch.randn(N, C + 42, generator=g) * 0.5).to(torch.bfloat16)
    dfield = torch.zeros(N, h.sizes.dout_pad, dtype=torch.bfloat16)
    dfield[:, :4 + D] = (torch.randn(N, 4 + D, generator=g) * 0.1).to(torch.bfloat16)
    fin, dfield = fin.cuda(), dfield.cuda()
    out, acts = h.forward(fin)
    res = {}
    for layered in (True, False):
        grads = NR._zero_grads(h)
        dlat = h.backward(fin, acts, dfield, grads, deterministic=True, layered=layered)
        res[layered] = (dlat, grads)
    assert torch.equal(res[True][0], res[False][0])
    for n in h.names():
        assert torch.equal(res[True][1][n], res[False][1][n]), n
    # and the gate bits really are what the backward thinks they are: fp32 autograd on the bf16-rounded weights
    p = {k: v.detach().float().cpu().to(torch.bfloat16).float().requires_grad_(True) for k, v in mlp.named_parameters()}
    x = fin[:, :C + 42].float().cpu().requires_grad_(True)
    ref = O.resnetfc(p, x, d_latent=C, operand_dtype=torch.bfloat16)
    ref = ref[0] if isinstance(ref, tuple) else ref
    ref.backward(dfield[:, :4 + D].float().cpu())
    # bf16 gradient operands against fp32 autograd: SURVEY section 10 measured ~1e-1 on dL/dlatent (ReLU-gate flips)
    cos = torch.nn.functional.cosine_similarity(res[False][0].cpu().flatten(), x.grad[:, :C].flatten(), dim=0)
    assert rel(res[False][0], x.grad[:, :C]) < 2e-1 and cos > 0.98
    assert rel(res[False][1]["lin_out.weight"], p["lin_out.weight"].grad) < 3e-2


@pytest.mark.parametrize("C,N,D", [(128, 256, 384), (128, 1000, 384), (128, 33000, 384), (64, 777, 384), (128, 3, 384),
                                   (64, 1000, 512), (128, 70001, 512), (64, 5, 512), (128, 300000, 384)])
def test_one_launch_weight_gradients_match_the_per_gemm_kernel(ops, C, N, D):
    """wgrad_multi_kernel (every weight / bias gradient of the pass in one persistent launch; bias sums on CUDA cores;
    lin_in + lin_z[0] in one tile) against the per-GEMM kernel with its ordered split reduction on the same operands:
    same bf16 products, fp32 sums in a different order.  dL/dlatent does not depend on the choice: bit-identical."""
    NR = load_pkg("neural_rendering")
    mlp = _bf16_mlp(C=C, D=D, seed=2)
    h = mlp.handle(ops.NRF_PREC_BF16)
    g = torch.Generator().manual_seed(N + 7)
    fin = torch.zeros(N, h.sizes.kin_pad, dtype=torch.bfloat16)
    fin[:, :C + 42] = (torch.randn(N, C + 42, generator=g) * 0.5).to(torch.bfloat16)
    dfield = torch.zeros(N, h.sizes.dout_pad, dtype=torch.bfloat16)
    dfield[:, :4 + D] = (torch.randn(N, 4 + D, generator=g) * 0.1).to(torch.bfloat16)
    fin, dfield = fin.cuda(), dfield.cuda()
    out, acts = h.forward(fin)
    res = {}
    for det in (True, False):
        grads = NR._zero_grads(h)
        n0 = load_pkg("_lib").launch_count()
        dlat = h.backward(fin, acts, dfield, grads, deterministic=det)
        res[det] = (dlat, grads, load_pkg("_lib").launch_count() - n0)
    assert res[False][2] <= 4 < res[True][2], f"fused backward + ONE weight-gradient launch + dL/dz GEMM, got {res[False][2]}"
    assert torch.equal(res[True][0], res[False][0])
    for n in h.names():
        a, b = res[False][1][n], res[True][1][n]
        assert a.shape == b.shape
        err = float((a.double() - b.double()).abs().max() / (b.double().abs().max() + 1e-30))
        assert err < 2e-5, (n, err)
    # accumulation: a second pass adds on top (the gradient buffers are not overwritten)
    grads = res[False][1]
    before = {n: grads[n].clone() for n in h.names()}
    h.backward(fin, acts, dfield, grads, deterministic=False)
    for n in h.names():
        err = float((grads[n].double() - 2 * before[n].double()).abs().max() / (before[n].double().abs().max() + 1e-30))
        assert err < 2e-5, (n, err)


def test_acts_of_the_layered_forward_take_the_layered_backward(ops):
    """Only the fused forward writes the gate bits; FieldMLP routes a backward over layered activations to the chain."""
    NR = load_pkg("neural_rendering")
    mlp = _bf16_mlp()
    h = mlp.handle(ops.NRF_PREC_BF16)
    N = 500
    fin = torch.zeros(N, h.sizes.kin_pad, dtype=torch.bfloat16, device="cuda")
    fin[:, :170] = (torch.randn(N, 170, device="cuda") * 0.5).to(torch.bfloat16)
    dfield = torch.zeros(N, h.sizes.dout_pad, dtype=torch.bfloat16, device="cuda")
    dfield[:, :388] = (torch.randn(N, 388, device="cuda") * 0.1).to(torch.bfloat16)
    _, acts_l = h.forward(fin, layered=True)
    _, acts_f = h.forward(fin)
    ga, gb = NR._zero_grads(h), NR._zero_grads(h)
    da = h.backward(fin, acts_l, dfield, ga, deterministic=True)      # silently layered
    db = h.backward(fin, acts_f, dfield, gb, deterministic=True)      # fused
    assert torch.equal(da, db)
    assert all(torch.equal(ga[n], gb[n]) for n in h.names())


# ------------------------------------------------------------------------------ fused rendering losses
@pytest.mark.parametrize("D,with_idx", [(384, True), (24, True), (512, False)])
def test_render_loss_matches_mse_loss_and_its_autograd(ops, D, with_idx):
    """nrf_render_loss = the four F.mse_loss terms of neural_rendering.py:664-685 with the [:, idx] target gather,
    plus their gradients, bit-reproducible."""
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(D)
    SB, rps, n_pix = 2, 150, 400
    R = SB * rps
    xs = [torch.randn(R, 3, generator=g), torch.randn(R, 3, generator=g), torch.randn(R, D, generator=g),
          torch.randn(R, D, generator=g)]
    xs = [x.requires_grad_(True) for x in xs]
    gt_rgb, gt_emb = torch.rand(SB, n_pix, 3, generator=g), torch.randn(SB, n_pix, D, generator=g)
    idx = torch.randint(n_pix, (rps,), generator=g)
    t_rgb, t_emb = gt_rgb[:, idx].reshape(R, 3), gt_emb[:, idx].reshape(R, D)
    ref = torch.stack([F.mse_loss(xs[0], t_rgb), F.mse_loss(xs[1], t_rgb), F.mse_loss(xs[2], t_emb),
                       F.mse_loss(xs[3], t_emb)])
    coef = torch.tensor([1.0, 0.5, 0.01, 0.02])
    (ref * coef).sum().backward()
    xc = [x.detach().cuda() for x in xs]
    if with_idx:
        terms, grads = ops.render_loss(*xc, rps, gt_rgb.cuda(), gt_emb.cuda(), idx.cuda())
    else:
        terms, grads = ops.render_loss(*xc, rps, t_rgb.cuda(), t_emb.cuda(), None)
    assert rel(terms, ref.detach()) < 1e-6
    for i in range(4):
        assert rel(grads[i] * float(coef[i]), xs[i].grad) < 1e-6
    terms2, grads2 = ops.render_loss(*xc, rps, gt_rgb.cuda() if with_idx else t_rgb.cuda(),
                                     gt_emb.cuda() if with_idx else t_emb.cuda(), idx.cuda() if with_idx else None)
    assert torch.equal(terms, terms2) and all(torch.equal(a, b) for a, b in zip(grads, grads2))
    terms3, none = ops.render_loss(*xc, rps, gt_rgb.cuda() if with_idx else t_rgb.cuda(),
                                   gt_emb.cuda() if with_idx else t_emb.cuda(), idx.cuda() if with_idx else None,
                                   want_grads=False)
    assert none is None and torch.equal(terms, terms3)


# ------------------------------------------------------------------------------ voxelizer
def test_voxelizer_matches_the_reference_fixture_bitwise(ops):
    """VoxelGrid.coords_to_bounding_voxel_grid on the GPU == the reference's own output (CPU), bit for bit."""
    VG = load_pkg("voxel_grid")
    fx = golden("voxelize_small")
    B, N, F, S, seed = [int(v) for v in fx["meta"]]
    coords, feats = syn.voxelizer_points(B, N, F, seed)
    vg = VG.VoxelGrid(coord_bounds=syn.BOUNDS, voxel_size=S, device="cuda", batch_size=B, feature_size=F,
                      max_num_coords=N).cuda()
    out = vg.coords_to_bounding_voxel_grid(coords.cuda(), coord_features=feats.cuda())
    assert torch.equal(out.cpu(), torch.from_numpy(fx["out"]))
    only = vg.coords_to_bounding_voxel_grid(coords.cuda(), coord_features=feats.cuda(), only_features=True)
    assert only.shape[-1] == 3 + F - 3 and torch.equal(only, out[..., :-7])
    with pytest.raises(Exception):
        vg.coords_to_bounding_voxel_grid(coords, coord_features=feats)          # CPU tensors: no fallback


@pytest.mark.parametrize("B,N,F,S", [(2, 220000, 3, 100), (1, 5000, 0, 16), (3, 40000, 7, 33)])
def test_voxelizer_matches_oracle_at_real_sizes(ops, B, N, F, S):
    """PerAct sizes (<= 220 k points into 100^3, voxel_grid_real.py / train_nerfact_multi_kitchen.py:1131), per-scene
    bounds, no features; bit-identical to the CPU oracle and run to run."""
    from oracle import voxel_oracle as VO
    VG = load_pkg("voxel_grid")
    coords, feats = syn.voxelizer_points(B, N, max(F, 1), seed=N % 97)
    feats = feats[..., :F] if F > 0 else None
    bounds = torch.tensor(syn.BOUNDS).repeat(B, 1)
    bounds[:, 3:] += 0.05 * torch.arange(B).view(B, 1)                           # different box per scene
    ref = VO.voxelize(coords, feats, bounds, S)
    vg = VG.VoxelGrid(coord_bounds=syn.BOUNDS, voxel_size=S, device="cuda", batch_size=B, feature_size=F,
                      max_num_coords=N).cuda()
    args = (coords.cuda(), feats.cuda() if feats is not None else None, bounds.cuda())
    out = vg.coords_to_bounding_voxel_grid(*args)
    assert torch.equal(out.cpu(), ref)
    assert torch.equal(out, vg.coords_to_bounding_voxel_grid(*args))


def test_sort_rows_with_nan_keeps_a_valid_permutation(ops):
    """A NaN depth sorts last (torch.sort's convention) and never pulls a padding slot into the row (K = 96 pads to 128)."""
    g = torch.Generator().manual_seed(0)
    z = torch.rand(50, 96, generator=g)
    z[3, 10] = float("nan")
    z[7, 0] = float("nan")
    z[7, 95] = float("nan")
    out, perm = ops.sort_rows(z.cuda().clone(), want_perm=True)
    ref, _ = torch.sort(z, dim=-1)
    assert torch.equal(torch.nan_to_num(out.cpu(), nan=9.0), torch.nan_to_num(ref, nan=9.0))
    assert int(perm.max()) < 96 and int(perm.min()) >= 0
    assert torch.equal(torch.sort(perm.cpu().long(), dim=1)[0], torch.arange(96).expand(50, 96))


def test_scatter_and_voxelizer_with_long_per_voxel_lists(ops):
    """Hundreds / thousands of entries in ONE voxel (coarse grid under many samples; a point cloud collapsed onto a few
    voxels): the per-voxel ordering is O(n log n) now and still the fixed ascending order - bit-identical run to run and
    equal to the oracle's sums."""
    # scatter: 4^3 grid, 64 rays x 64 samples inside the box -> ~500 entries per voxel
    S, C, R, K = 4, 64, 64, 64
    g = torch.Generator().manual_seed(1)
    b = torch.tensor(syn.BOUNDS)
    o = (b[:3] + (b[3:] - b[:3]) * (0.2 + 0.6 * torch.rand(R, 3, generator=g)))
    d = torch.nn.functional.normalize(torch.randn(R, 3, generator=g), dim=-1) * 0.05
    rays = torch.cat([o, d, torch.zeros(R, 1), torch.ones(R, 1)], 1).cuda()
    z = torch.rand(R, K, generator=g).sort(-1)[0].cuda()
    dlat = torch.randn(R * K, C, generator=g).cuda()
    outs = []
    for _ in range(2):
        grad = torch.empty(1, C, S, S, S, device="cuda")
        ops.scatter_volume_grad_merged(rays, R, [(z, dlat)], grad, True, b)
        outs.append(grad)
    assert torch.equal(outs[0], outs[1])
    vol = torch.zeros(1, C, S, S, S, requires_grad=True)
    pts = (rays[:, None, :3] + z.unsqueeze(2) * rays[:, None, 3:6]).reshape(1, -1, 3).cpu()
    lat = O.trilinear_gather(vol, O.world_to_canonical(pts, syn.BOUNDS))
    (lat.reshape(-1, C) * dlat.cpu()).sum().backward()
    assert rel(outs[0], vol.grad) < 2e-6
    # voxelizer: 3000 points in a handful of voxels, against the oracle's ascending-index sums (bit for bit)
    VG = load_pkg("voxel_grid")
    from oracle import voxel_oracle as VO
    coords = b[:3] + (b[3:] - b[:3]) * (0.5 + 0.02 * torch.rand(1, 3000, 3, generator=g))
    feats = torch.rand(1, 3000, 3, generator=g)
    vg = VG.VoxelGrid(coord_bounds=syn.BOUNDS, voxel_size=10, device="cuda", batch_size=1, feature_size=3, max_num_coords=3000)
    got = vg.coords_to_bounding_voxel_grid(coords.cuda(), coord_features=feats.cuda())
    ref = VO.voxelize(coords, feats, syn.BOUNDS, 10)
    assert int((got[..., -1] > 0).sum()) <= 8 and torch.equal(got.cpu(), ref)


# ------------------------------------------------------------------- voxel rows <-> compact list (sparse exchange)
@pytest.mark.parametrize("C,dims,SB,cl3d", [(128, (9, 7, 11), 2, False), (128, (9, 7, 11), 2, True), (64, (5, 6, 7), 1, False),
                                            (12, (4, 4, 5), 3, False), (12, (4, 4, 5), 3, True)])
def test_voxel_rows_gather_and_update(ops, C, dims, SB, cl3d):
    """nrf_rows_gather / nrf_rows_update against torch indexing, both memory formats: gather, clear, overwrite, add -
    exact (no arithmetic but one add), untouched voxels untouched."""
    g = torch.Generator().manual_seed(C + SB)
    grad = torch.randn(SB, C, *dims, generator=g).cuda()
    if cl3d:
        grad = grad.contiguous(memory_format=torch.channels_last_3d)
    V = dims[0] * dims[1] * dims[2]
    flat = lambda t: t.permute(0, 2, 3, 4, 1).reshape(SB * V, C)        # (scene * V + voxel, C) copy
    for n in (0, 1, 31, 32, 33, SB * V // 3, SB * V):
        idx = torch.randperm(SB * V, generator=g)[:n].sort()[0].cuda()
        rows = ops.rows_gather(grad, idx)
        assert torch.equal(rows, flat(grad)[idx])
        upd = torch.randn(n, C, generator=g).cuda()
        ref = flat(grad).clone()
        work = grad.clone(memory_format=torch.preserve_format)
        ops.rows_update(work, idx, upd, add=True)
        ref[idx] += upd
        assert torch.equal(flat(work), ref)
        ops.rows_update(work, idx, upd, add=False)
        ref[idx] = upd
        assert torch.equal(flat(work), ref)
        ops.rows_update(work, idx, None)
        ref[idx] = 0
        assert torch.equal(flat(work), ref)
        assert work.stride() == grad.stride()


@pytest.mark.parametrize("C,dims,SB,cl3d,world", [(128, (9, 7, 11), 2, False, 3), (128, (9, 7, 11), 1, True, 8),
                                                  (64, (5, 6, 7), 1, False, 2), (12, (4, 4, 5), 3, False, 5),
                                                  (128, (40, 40, 40), 1, False, 8)])
def test_voxel_rows_merge_sums_the_lists_in_rank_order(ops, C, dims, SB, cl3d, world):
    """nrf_rows_merge: every voxel some rank lists = ((0 + r_0) + r_1) + .. in rank order (bit-exact against the same
    order in torch), everything else untouched; ragged lists, empty ranks, voxel counts that are not multiples of 32."""
    g = torch.Generator().manual_seed(C + world)
    V = dims[0] * dims[1] * dims[2]
    total = SB * V
    grad = torch.randn(SB, C, *dims, generator=g).cuda()
    if cl3d:
        grad = grad.contiguous(memory_format=torch.channels_last_3d)
    flat = lambda t: t.permute(0, 2, 3, 4, 1).reshape(total, C)
    counts = [int(torch.randint(0, max(2, total // 4), (1,), generator=g)) for _ in range(world)]
    counts[world // 2] = 0
    counts[0] = min(total, max(counts[0], 40))
    cap = max(counts) + 3
    all_idx = torch.zeros(world, cap, dtype=torch.int64)
    all_rows = torch.randn(world, cap, C, generator=g)
    for r in range(world):
        all_idx[r, :counts[r]] = torch.randperm(total, generator=g)[:counts[r]].sort()[0]
    ref = flat(grad).cpu().clone()
    listed = torch.zeros(total, dtype=torch.bool)
    for r in range(world):
        listed[all_idx[r, :counts[r]]] = True
    ref[listed] = 0.0
    for r in range(world):
        ref[all_idx[r, :counts[r]]] += all_rows[r, :counts[r]]
    work = grad.clone(memory_format=torch.preserve_format)
    ops.rows_merge(work, all_rows.cuda(), all_idx.cuda(), counts)
    assert torch.equal(flat(work).cpu(), ref)
    # whole-tile mode: the caller vouches that every unlisted voxel holds 0 (a gradient fresh from the scatter)
    sparse = grad.clone(memory_format=torch.preserve_format)
    fl = flat(sparse)
    fl[~listed.cuda()] = 0.0
    sparse = fl.reshape(SB, *dims, C).permute(0, 4, 1, 2, 3)
    sparse = sparse.contiguous(memory_format=torch.channels_last_3d) if cl3d else sparse.contiguous()
    ref0 = ref.clone()
    ref0[~listed] = 0.0
    ops.rows_merge(sparse, all_rows.cuda(), all_idx.cuda(), counts, unlisted_are_zero=True)
    assert torch.equal(flat(sparse).cpu(), ref0)
