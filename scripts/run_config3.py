"""BASELINE config 3: novel-view full-image render, 5 cameras x 128x128 px, 64 + 64 samples, inference only,
rays sharded over the ranks.  python scripts/run_config3.py   (or under torchrun for N > 1)"""
import importlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
PKG = "real-robot-nerf-actor_b200"
NR = importlib.import_module(PKG + ".neural_rendering"); U = importlib.import_module(PKG + ".utils")
syn = importlib.import_module(PKG + ".synthetic"); par = importlib.import_module(PKG + ".parallel")
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
wl = syn.CONFIGS["config3"]
ren = NR.NeuralRenderer(U.default_config(voxel_shape=wl.S), torch.tensor(syn.BOUNDS))
syn.init_mlp_(ren.nerf_model.mlp_coarse); ren = ren.to(dev).eval()
ren.render_chunk_rays = int(os.environ.get("NRF_RENDER_CHUNK", "4096"))
vol = torch.randn(1, wl.C, wl.S, wl.S, wl.S, device=dev, generator=torch.Generator(device=dev).manual_seed(0)) * 0.1
poses = syn.arc_poses(5).to(dev); focal = torch.tensor(wl.focal, device=dev)
def run():
    return par.render_sharded(ren, vol, focal, poses, gather=True)
for _ in range(2): out = run()
torch.cuda.synchronize()
if world > 1: dist.barrier()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); n = 3
for _ in range(n): rgb, emb, dep = run()
b.record(); torch.cuda.synchronize()
ms = torch.tensor([a.elapsed_time(b) / n], device=dev)
if world > 1: dist.all_reduce(ms, op=dist.ReduceOp.MAX)
# opt-in: the fine pass reuses the coarse pass's field evaluations (bit-identical images)
ren.perturb = False
ref_imgs = run()
ren.reuse_coarse_evals = True
same = all(torch.equal(x, y) for x, y in zip(ref_imgs, run()))
ren.perturb = True
torch.cuda.synchronize()
if world > 1: dist.barrier()
a.record()
for _ in range(n): run()
b.record(); torch.cuda.synchronize()
ms_reuse = torch.tensor([a.elapsed_time(b) / n], device=dev)
if world > 1: dist.all_reduce(ms_reuse, op=dist.ReduceOp.MAX)
if rank == 0:
    evals = wl.evals
    print(json.dumps({"workload": "config3: 5 cams x 128x128 px, 64+64 samples, inference, rays sharded", "n_gpus": world,
                      "ms_per_render": round(float(ms), 2), "ray_samples_per_s": round(evals / (float(ms) * 1e-3), 1),
                      "reuse_coarse_evals": {"ms_per_render": round(float(ms_reuse), 2), "images_bit_identical": same},
                      "shapes": [list(rgb.shape), list(emb.shape), list(dep.shape)],
                      "finite": bool(torch.isfinite(rgb).all() and torch.isfinite(emb).all()),
                      "peak_mem_gib": round(torch.cuda.max_memory_allocated() / 2**30, 1)}))
if world > 1: dist.destroy_process_group()
