"""Times the voxelizer (nrf_voxelize) at PerAct sizes on cuda:0.  python scripts/voxelize_bench.py"""
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

PKG = "real-robot-nerf-actor_b200"
VG = importlib.import_module(PKG + ".voxel_grid")
syn = importlib.import_module(PKG + ".synthetic")

B, N, F, S = 2, 220000, 3, 100
coords, feats = syn.voxelizer_points(B, N, F, seed=1)
coords, feats = coords.cuda(), feats.cuda()
vg = VG.VoxelGrid(syn.BOUNDS, S, "cuda", B, F, N).cuda()
for _ in range(3):
    out = vg.coords_to_bounding_voxel_grid(coords, coord_features=feats)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20):
    out = vg.coords_to_bounding_voxel_grid(coords, coord_features=feats)
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / 20
T = B * S ** 3
bytes_alg = out.numel() * 4 + B * N * (3 + F) * 4            # grid written once + points read once
print(json.dumps({"voxelize": {"B": B, "N": N, "F": F, "S": S, "ms": round(ms, 4),
                               "algorithmic_GB": round(bytes_alg / 1e9, 4),
                               "achieved_GBps": round(bytes_alg / ms / 1e6, 1),
                               "occupied_voxels": int(out[..., -1].sum())}}))
