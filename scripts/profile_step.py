"""Host-side profile of one bench step (where does the CPU time go?).  python scripts/profile_step.py"""
import cProfile
import importlib
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

PKG = "real-robot-nerf-actor_b200"
NR = importlib.import_module(PKG + ".neural_rendering")
U = importlib.import_module(PKG + ".utils")
syn = importlib.import_module(PKG + ".synthetic")

wl = syn.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "config2"]
dev = torch.device("cuda", 0)
cfg = U.default_config(voxel_shape=wl.S, d_latent=wl.C, d_embed=wl.D, n_coarse=wl.n_coarse, n_fine=wl.n_fine,
                       ray_chunk_size=wl.rays_per_scene)
ren = NR.NeuralRenderer(cfg, torch.tensor(syn.BOUNDS))
syn.init_mlp_(ren.nerf_model.mlp_coarse)
ren = ren.to(dev).train()
SB = wl.SB
vol = (torch.randn(SB, wl.C, wl.S, wl.S, wl.S, device=dev) * 0.1).requires_grad_(True)
poses = syn.arc_poses(SB).to(dev)
focal = torch.tensor(wl.focal, device=dev)
gt_rgb = torch.rand(SB, wl.H, wl.W, 3, device=dev)
gt_emb = torch.randn(SB, wl.H, wl.W, wl.D, device=dev)


def step(sync_phases=False):
    vol.grad = None
    for p in ren.parameters():
        p.grad = None
    t0 = time.perf_counter()
    out = ren(multi_scale_voxel_list=None, voxel_density=None, language=None, voxel_feat=vol, voxel_poses=poses,
              focal=focal, gt_rgb=gt_rgb, gt_depth=None, gt_pose=poses, c=None, lang_goal=None, gt_embed=gt_emb)
    if sync_phases:
        torch.cuda.synchronize()
    t1 = time.perf_counter()
    out["loss"].backward()
    t2 = time.perf_counter()
    if sync_phases:
        torch.cuda.synchronize()
    t3 = time.perf_counter()
    return t1 - t0, t2 - t1, t3 - t2


for _ in range(3):
    step()
torch.cuda.synchronize()
for _ in range(3):
    f, b, s = step(True)
    print(f"forward {f*1e3:.1f} ms (synced)  backward launch {b*1e3:.1f} ms  backward drain {s*1e3:.1f} ms")
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    step()
torch.cuda.synchronize()
print(f"5 steps unsynced: {(time.perf_counter()-t0)/5*1e3:.1f} ms/step; peak mem {torch.cuda.max_memory_allocated()/2**30:.1f} GiB")
pr = cProfile.Profile()
pr.enable()
for _ in range(3):
    step()
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(28)
