"""A few config-2 training steps for ncu (kernel filters pick the launches).  python scripts/one_step.py [precision] [inbox] [steps]"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
PKG = "real-robot-nerf-actor_b200"
NR = importlib.import_module(PKG + ".neural_rendering"); U = importlib.import_module(PKG + ".utils")
syn = importlib.import_module(PKG + ".synthetic")
precision = sys.argv[1] if len(sys.argv) > 1 else "bf16"
inbox = len(sys.argv) > 2 and sys.argv[2] == "inbox"
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
wl = syn.CONFIGS["config2"]
dev = torch.device("cuda", 0)
near, far, focal = (2.4, 3.2, 500.0) if inbox else (1.2, 4.0, wl.focal)
cfg = U.default_config(voxel_shape=wl.S, d_latent=wl.C, d_embed=wl.D, n_coarse=wl.n_coarse, n_fine=wl.n_fine,
                       ray_chunk_size=wl.rays_per_scene, z_near=near, z_far=far)
ren = NR.NeuralRenderer(cfg, torch.tensor(syn.BOUNDS), precision=precision)
syn.init_mlp_(ren.nerf_model.mlp_coarse, seed=0)
ren = ren.to(dev).train()
g = torch.Generator(device=dev).manual_seed(1234)
vol = (torch.randn(wl.SB, wl.C, wl.S, wl.S, wl.S, device=dev, generator=g) * 0.1).requires_grad_(True)
poses = syn.arc_poses(wl.SB).to(dev); focal = torch.tensor(focal, device=dev)
gt_rgb = torch.rand(wl.SB, wl.H, wl.W, 3, device=dev); gt_emb = torch.randn(wl.SB, wl.H, wl.W, wl.D, device=dev)
for _ in range(steps):
    vol.grad = None
    for p in ren.parameters(): p.grad = None
    out = ren(multi_scale_voxel_list=None, voxel_density=None, language=None, voxel_feat=vol, voxel_poses=poses,
              focal=focal, gt_rgb=gt_rgb, gt_depth=None, gt_pose=poses, c=None, lang_goal=None, gt_embed=gt_emb)
    out["loss"].backward()
torch.cuda.synchronize()
print("ran", steps, "steps", precision, "inbox" if inbox else "", float(out["loss"]))
