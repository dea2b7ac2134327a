"""Two-GPU checks of the data-parallel path over NCCL (skipped on a single-GPU box): the MLP-gradient all-reduce
started inside the backward gives exactly the gradients of the separate all-reduce after it."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.conftest import load_pkg

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        NR, U, syn, par = (load_pkg(m) for m in ("neural_rendering", "utils", "synthetic", "parallel"))
        S, SB, n_rays = 16, 2, 64
        cfg = U.default_config(voxel_shape=S, ray_chunk_size=n_rays, image_width=32, image_height=32)
        res = {}
        for mode in ("post", "overlap"):
            ren = NR.NeuralRenderer(cfg, torch.tensor(syn.BOUNDS), precision="bf16")
            syn.init_mlp_(ren.nerf_model.mlp_coarse, seed=0)             # replicated MLP
            ren = ren.cuda().train()
            ren.deterministic = True
            if mode == "overlap":
                par.overlap_mlp_grad_allreduce(ren)
            g = torch.Generator(device="cuda").manual_seed(100 + rank)  # every rank has its own scenes
            vol = (torch.randn(SB, 128, S, S, S, device="cuda", generator=g) * 0.1).requires_grad_(True)
            poses = syn.arc_poses(SB).cuda()
            gt_rgb = torch.rand(SB, 32, 32, 3, device="cuda", generator=g)
            gt_emb = torch.randn(SB, 32, 32, 384, device="cuda", generator=g)
            torch.manual_seed(7 + rank)
            out_d = ren(multi_scale_voxel_list=None, voxel_density=None, language=None, voxel_feat=vol,
                        voxel_poses=poses, focal=torch.tensor(38.0, device="cuda"), gt_rgb=gt_rgb, gt_depth=None,
                        gt_pose=poses, c=None, lang_goal=None, gt_embed=gt_emb)
            out_d["loss"].backward()
            if mode == "post":
                par.allreduce_mlp_grads(ren)
            torch.cuda.synchronize()
            res[mode] = ({k: p.grad.clone() for k, p in ren.named_parameters()}, vol.grad.clone())
        for k in res["post"][0]:
            assert torch.equal(res["post"][0][k], res["overlap"][0][k]), k
        assert torch.equal(res["post"][1], res["overlap"][1])          # the volume gradient stays local
        # and the reduced gradient really is the sum over ranks: it is identical on both ranks
        w = res["overlap"][0]["nerf_model.mlp_coarse.lin_out.weight"]
        both = [torch.empty_like(w) for _ in range(world)]
        dist.all_gather(both, w)
        assert torch.equal(both[0], both[1]) and float(w.abs().sum()) > 0
        out.put((rank, "ok"))
    except Exception as e:                                              # pragma: no cover
        import traceback
        out.put((rank, traceback.format_exc()[-1500:]))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_overlapped_grad_allreduce_equals_the_separate_one():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    results = [out.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(msg == "ok" for _, msg in results), results
