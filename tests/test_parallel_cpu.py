"""Host-side logic of the multi-GPU path on CPU: ray sharding and the gradient all-reduce with
world_size 2 over gloo.  Also checks that the C-ABI library loads and exports every declared symbol."""
import os
import re
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.conftest import ROOT, load_pkg

par = load_pkg("parallel")


def test_shard_bounds_cover_and_balance():
    for n in (0, 1, 7, 4096, 81920, 81921):
        for w in (1, 2, 3, 4, 8):
            spans = [par.shard_bounds(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    rays = torch.arange(10 * 8, dtype=torch.float32).reshape(10, 8)
    assert torch.equal(torch.cat([par.shard_rays(rays, 3, r) for r in range(3)]), rays)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        lin = torch.nn.Linear(5, 3)
        shared = torch.nn.Linear(3, 3)
        model = torch.nn.ModuleDict({"a": lin, "coarse": shared, "fine": shared, "nograd": torch.nn.Linear(2, 2)})
        x = torch.full((4, 5), float(rank + 1))
        (shared(lin(x)).sum() * (rank + 1)).backward()          # 'nograd' never receives a gradient
        local = [p.grad.clone() for p in (lin.weight, lin.bias, shared.weight, shared.bias)]
        nbytes = par.allreduce_mlp_grads(model)
        n_unique = sum(p.numel() for p in model.parameters())
        assert nbytes == 4 * n_unique                            # aliased module counted once
        gathered = [[torch.zeros_like(g) for _ in range(world)] for g in local]
        for g, dst in zip(local, gathered):
            dist.all_gather(dst, g)
        for p, parts in zip((lin.weight, lin.bias, shared.weight, shared.bias), gathered):
            assert torch.allclose(p.grad, sum(parts))
        assert torch.equal(model["nograd"].weight.grad, torch.zeros(2, 2))
        # the hook the renderer's backward calls with its flat gradient buffer(s)
        class _R:
            pass
        r = _R()
        par.overlap_mlp_grad_allreduce(r, average=True)
        flats = [torch.full((5,), float(rank + 1)), torch.arange(3, dtype=torch.float32) * (rank + 1)]
        finish = r._grad_allreduce(flats)
        finish()
        tot = float(sum(range(1, world + 1)))
        assert torch.allclose(flats[0], torch.full((5,), tot / world))
        assert torch.allclose(flats[1], torch.arange(3, dtype=torch.float32) * tot / world)
        par.overlap_mlp_grad_allreduce(r, enabled=False)
        assert r._grad_allreduce is None
        vol = torch.full((2, 3), float(rank + 1))
        par.allreduce_volume_grad(vol)
        assert torch.equal(vol, torch.full((2, 3), float(sum(range(1, world + 1)))))
        # one scene's rays split over the ranks: the sparse exchange of touched voxel rows equals the dense sum, for a
        # contiguous and a channels_last_3d gradient, with overlapping and disjoint footprints, and two scenes
        for SB, fmt in ((1, torch.contiguous_format), (2, torch.channels_last_3d)):
            g = torch.Generator().manual_seed(100 + rank)
            C_, S_ = 8, 6
            local = torch.zeros(SB, C_, S_, S_, S_)
            hit = torch.randint(0, S_ ** 3, (40,), generator=g)
            hit[:5] = torch.arange(5)                            # five voxels every rank touches
            for b in range(SB):
                local[b].reshape(C_, -1)[:, hit] = torch.randn(C_, 40, generator=g)
            local = local.contiguous(memory_format=fmt)
            dense = local.clone()
            dist.all_reduce(dense)
            sparse = local.clone()
            stats = par.sparse_allreduce_volume_grad(sparse)
            assert sparse.is_contiguous(memory_format=fmt)
            assert torch.allclose(sparse, dense, atol=1e-6) and len(stats["rows"]) == world
            assert stats["bytes"] < dense.numel() * 4
            both = [torch.empty_like(sparse) for _ in range(world)]
            dist.all_gather(both, sparse.contiguous())
            assert torch.equal(both[0], both[1])                 # bit-identical replicas
        out.put((rank, "ok"))
    except Exception as e:                                        # pragma: no cover
        import traceback
        out.put((rank, traceback.format_exc()[-800:]))
    finally:
        dist.destroy_process_group()


def test_grad_allreduce_world2_gloo():
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    res = dict(out.get() for _ in range(2))
    assert res == {0: "ok", 1: "ok"}, res


def test_c_abi_library_exports_every_declared_symbol():
    """include/nrf_b200.h <-> libnrf_b200.so <-> the ctypes binding agree (no compute calls: no GPU here)."""
    lib_mod = load_pkg("_lib")
    if not os.path.exists(lib_mod.LIB_PATH):
        load_pkg("_build").build()
    lib = lib_mod.load()
    header = open(os.path.join(ROOT, "include", "nrf_b200.h")).read()
    declared = set(re.findall(r"\b(nrf_[a-z0-9_]+)\s*\(", header))
    assert declared == set(lib_mod.EXPORTS), declared ^ set(lib_mod.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    assert b"sm_100a" in lib.nrf_version()


def test_missing_library_fails_loudly(monkeypatch):
    lib_mod = load_pkg("_lib")
    monkeypatch.setattr(lib_mod, "_lib", None)
    monkeypatch.setattr(lib_mod, "LIB_PATH", "/nonexistent/libnrf_b200.so")
    with pytest.raises(lib_mod.NrfError):
        lib_mod.load()


def test_renderer_keeps_reference_surface():
    """Constructor / config / state_dict surface of the drop-in (no GPU needed)."""
    NR = load_pkg("neural_rendering")
    U = load_pkg("utils")
    ren = NR.NeuralRenderer(U.default_config(), torch.tensor([-0.1, -0.3, -0.2, 0.8, 0.7, 0.7]))
    sd = ren.state_dict()
    assert len(sd) == 62 and sum(p.numel() for p in ren.parameters()) == 3_045_764
    assert ren.nerf_model.mlp_fine is ren.nerf_model.mlp_coarse
    assert "nerf_model.code._freqs" in sd and "nerf_model.mlp_fine.lin_z.2.bias" in sd
    assert float(ren.nerf_model.mlp_coarse.blocks[0].fc_1.weight.abs().max()) == 0.0     # resnetfc.py:41
    for name in ("forward", "rendering", "encode", "forward_nerf", "sample_coarse", "sample_fine",
                 "sample_fine_depth", "composite", "compute_rendering_loss"):
        assert callable(getattr(ren, name))
    with pytest.raises(NotImplementedError):      # the reference cannot run it either (models_embed.py:151-154,:289)
        NR.NeuralRenderer(U.default_config(use_depth_supervision=True), torch.zeros(6))
    ms = NR.NeuralRenderer(U.default_config(use_multi_scale_voxel=True, ret_last_feat=True), torch.zeros(6))
    assert ms._composed and ms.nerf_model.d_latent == 266 and ms.nerf_model.mlp_coarse.lin_z[0].weight.shape == (512, 266)
    heads = NR.NeuralRenderer(U.default_config(regress_coord=True, regress_attention=True), torch.zeros(6))
    assert heads.nerf_model.d_out == 4 + 384 + 3 + 6 and heads._d_comp == 396          # models_embed.py:96-103
    assert heads.nerf_model.mlp_coarse.lin_out.weight.shape == (397, 512)
    cv = NR.NeuralRenderer(U.default_config(use_code_viewdirs=True, normalize_z=True), torch.zeros(6))
    assert cv._composed and cv.nerf_model.d_in == 78 and cv.nerf_model.code.d_out == 78      # models_embed.py:86-95
    assert cv.nerf_model.mlp_coarse.lin_in.weight.shape == (512, 78) and len(cv.state_dict()) == 62
    sp = NR.NeuralRenderer(U.default_config(mlp=dict(beta=10.0, use_spade=True)), torch.zeros(6))   # resnetfc.py:130-141
    assert sp._composed and sp.nerf_model.mlp_coarse.general and len(sp.state_dict()) == 62 + 2 * 6
    assert sp.nerf_model.mlp_coarse.scale_z[2].weight.shape == (512, 128)
    with pytest.raises(NotImplementedError):      # a debugger breakpoint in the reference (models_embed.py:350-352)
        NR.NeuralRenderer(U.default_config(use_freenerf=True), torch.zeros(6))
    with pytest.raises(NotImplementedError):
        NR.NeuralRenderer(U.default_config(foundation_model_name="nope"), torch.zeros(6))
    with pytest.raises(Exception):                # CPU tensors are rejected: there is no CPU fallback
        ren.encode(None, None, None, torch.zeros(1, 128, 4, 4, 4), None, None)
        ren.forward_nerf(torch.zeros(1, 4, 8))


def test_dropin_modules_resolve_under_the_reference_names():
    """With dropin/ first on sys.path the reference scripts' own import lines (`from neural_rendering import
    NeuralRenderer`, `from voxel_grid_real import VoxelGrid`) get the B200 classes; CPU tensors are refused."""
    import importlib
    import sys
    d = os.path.join(ROOT, "real-robot-nerf-actor_b200", "dropin")
    saved = {k: sys.modules.pop(k, None) for k in ("neural_rendering", "voxel_grid_real")}
    sys.path.insert(0, d)
    try:
        nr = importlib.import_module("neural_rendering")
        vg = importlib.import_module("voxel_grid_real")
        assert nr.NeuralRenderer is load_pkg("neural_rendering").NeuralRenderer
        assert vg.VoxelGrid is load_pkg("voxel_grid").VoxelGrid
        grid = vg.VoxelGrid(coord_bounds=[-0.1, -0.3, -0.2, 0.8, 0.7, 0.7], voxel_size=10, device="cpu", batch_size=1,
                            feature_size=3, max_num_coords=100)
        with pytest.raises(Exception, match="CUDA"):
            grid.coords_to_bounding_voxel_grid(torch.zeros(1, 100, 3), coord_features=torch.zeros(1, 100, 3))
    finally:
        sys.path.remove(d)
        for k, v in saved.items():
            sys.modules.pop(k, None)
            if v is not None:
                sys.modules[k] = v


def test_param_dict_cache_follows_the_module():
    """ResnetFC.param_dict(): same names / order / objects as named_parameters(), cached, rebuilt when a parameter
    object is replaced or the module is converted."""
    NR = load_pkg("neural_rendering")
    m = NR.ResnetFC(d_in=42, d_out=388, n_blocks=5, d_latent=128, d_hidden=512, combine_layer=3)
    d, ref = m.param_dict(), dict(m.named_parameters())
    assert list(d) == list(ref) and all(d[k] is ref[k] for k in ref) and len(d) == 30
    assert m.param_dict() is d
    m.lin_in.weight = torch.nn.Parameter(torch.zeros_like(m.lin_in.weight))
    d2 = m.param_dict()
    assert d2 is not d and d2["lin_in.weight"] is m.lin_in.weight
    m.double()
    assert m.param_dict() is not d2 and m.param_dict()["lin_out.bias"].dtype == torch.float64
