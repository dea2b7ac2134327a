"""Phase timings of the sparse volume-gradient exchange (config 5: one scene's rays split over the ranks).
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/exchange_probe.py
Per rank: one config-5 training step (16384 / N rays), then parallel.sparse_allreduce_volume_grad and the dense
all-reduce on that step's gradient, each timed as a whole and, for the sparse form, phase by phase (CUDA events)."""
import importlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
PKG = "real-robot-nerf-actor_b200"
NR = importlib.import_module(PKG + ".neural_rendering"); U = importlib.import_module(PKG + ".utils")
syn = importlib.import_module(PKG + ".synthetic"); par = importlib.import_module(PKG + ".parallel")

world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
wl = syn.CONFIGS["config5"]
n_rays = wl.rays_per_scene // world
cfg = U.default_config(voxel_shape=wl.S, d_latent=wl.C, d_embed=wl.D, n_coarse=wl.n_coarse, n_fine=wl.n_fine,
                       ray_chunk_size=n_rays, image_width=wl.W, image_height=wl.H)
ren = NR.NeuralRenderer(cfg, torch.tensor(syn.BOUNDS), precision="bf16")
syn.init_mlp_(ren.nerf_model.mlp_coarse, seed=0)
ren = ren.to(dev).train()
ren.keep_voxel_counts = True
g = torch.Generator(device=dev).manual_seed(1234)
vol = (torch.randn(1, wl.C, wl.S, wl.S, wl.S, device=dev, generator=g) * 0.1).requires_grad_(True)
poses = syn.arc_poses(1).to(dev); focal = torch.tensor(wl.focal, device=dev)
torch.manual_seed(100 + rank)                                     # every rank draws its own rays
gt_rgb = torch.rand(1, wl.H, wl.W, 3, device=dev); gt_emb = torch.randn(1, wl.H, wl.W, wl.D, device=dev)
for _ in range(2):
    vol.grad = None
    out = ren(multi_scale_voxel_list=None, voxel_density=None, language=None, voxel_feat=vol, voxel_poses=poses,
              focal=focal, gt_rgb=gt_rgb, gt_depth=None, gt_pose=poses, c=None, lang_goal=None, gt_embed=gt_emb)
    out["loss"].backward()
torch.cuda.synchronize()
grad0, counts = vol.grad.detach().clone(), ren.last_voxel_counts


def timed(fn, n=3):
    dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n


res = {"world": world, "rows_this_rank": int((counts > 0).sum())}
gbuf = grad0.clone()
par.sparse_allreduce_volume_grad(gbuf.copy_(grad0), counts=counts)      # warm-up: allocator blocks, NCCL channels
res["sparse_ms"] = timed(lambda: par.sparse_allreduce_volume_grad(gbuf.copy_(grad0), counts=counts))
res["copy_ms"] = timed(lambda: gbuf.copy_(grad0))
res["dense_ms"] = timed(lambda: par.allreduce_volume_grad(gbuf))
if hasattr(par, "sparse_allreduce_phases"):
    res["phases"] = par.sparse_allreduce_phases(gbuf.copy_(grad0), counts)
a = grad0.clone(); b_ = grad0.clone()
par.sparse_allreduce_volume_grad(a, counts=counts)
par.allreduce_volume_grad(b_)
res["sparse_vs_dense_rel"] = float((a.double() - b_.double()).norm() / b_.double().norm())
allr = [None] * world
dist.all_gather_object(allr, res)
if rank == 0:
    print(json.dumps(allr[0]))
    print(json.dumps([r["sparse_ms"] for r in allr]))
dist.destroy_process_group()
