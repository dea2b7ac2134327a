"""Eager vs CUDA-graph step at the reference's own training shape (nerfact.conf: 1 scene x 512 rays, 64 + 64 samples of
which 16 depth-guided, 64 latent channels, 512-d features).  python scripts/graphed_step_probe.py"""
import importlib, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
PKG = "real-robot-nerf-actor_b200"
NR = importlib.import_module(PKG + ".neural_rendering"); U = importlib.import_module(PKG + ".utils")
syn = importlib.import_module(PKG + ".synthetic"); G = importlib.import_module(PKG + ".graphed")
dev = torch.device("cuda", 0)
cfg = U.default_config(voxel_shape=100, d_latent=64, d_embed=512, n_coarse=64, n_fine=64, n_fine_depth=16,
                       ray_chunk_size=512, image_width=128, image_height=128)
ren = NR.NeuralRenderer(cfg, torch.tensor(syn.BOUNDS))
syn.init_mlp_(ren.nerf_model.mlp_coarse, seed=0)
ren = ren.to(dev).train()
vol = (torch.randn(1, 64, 100, 100, 100, device=dev) * 0.1).requires_grad_(True)
poses = syn.arc_poses(1).to(dev); focal = torch.tensor(153.0, device=dev)
gt_rgb = torch.rand(1, 128, 128, 3, device=dev); gt_emb = torch.randn(1, 128, 128, 512, device=dev)
params = list(ren.parameters())
def eager():
    return ren(multi_scale_voxel_list=None, voxel_density=None, language=None, voxel_feat=vol, voxel_poses=poses,
               focal=focal, gt_rgb=gt_rgb, gt_depth=None, gt_pose=poses, c=None, lang_goal=None, gt_embed=gt_emb)
graphed = G.GraphedRenderLoss(ren, vol, poses, focal, gt_rgb, gt_emb)
def run(fn, n=50):
    def step():
        vol.grad = None
        for p in params: p.grad = None
        out = fn()
        out["loss"].backward()
        return out
    for _ in range(5): step()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); a.record()
    for _ in range(n): out = step()
    b.record(); torch.cuda.synchronize()
    return {"ms_per_step": round(a.elapsed_time(b) / n, 3), "wall_ms_per_step": round((time.perf_counter() - t0) / n * 1e3, 3),
            "loss": float(out["loss"]), "psnr": out["psnr"], "vol_grad_norm": float(vol.grad.norm()),
            "lin_out_grad_norm": float(ren.nerf_model.mlp_coarse.lin_out.weight.grad.norm())}
res = {"eager": run(eager), "cuda_graph": run(lambda: graphed(vol, poses, focal, gt_rgb, gt_emb)), "eager_again": run(eager)}
print(json.dumps(res))
