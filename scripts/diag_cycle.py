import gc, importlib, os, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
PKG = "real-robot-nerf-actor_b200"
NR = importlib.import_module(PKG + ".neural_rendering"); U = importlib.import_module(PKG + ".utils")
syn = importlib.import_module(PKG + ".synthetic")
dev = torch.device("cuda", 0)
cfg = U.default_config(voxel_shape=16, n_coarse=64, n_fine=64, ray_chunk_size=64)
ren = NR.NeuralRenderer(cfg, torch.tensor(syn.BOUNDS)); ren = ren.to(dev).train()
SB = 2
vol = (torch.randn(SB, 128, 16, 16, 16, device=dev) * 0.1).requires_grad_(True)
poses = syn.arc_poses(SB).to(dev); focal = torch.tensor(153.0, device=dev)
gt_rgb = torch.rand(SB, 128, 128, 3, device=dev); gt_emb = torch.randn(SB, 128, 128, 384, device=dev)
def step():
    out = ren(multi_scale_voxel_list=None, voxel_density=None, language=None, voxel_feat=vol, voxel_poses=poses,
              focal=focal, gt_rgb=gt_rgb, gt_depth=None, gt_pose=poses, c=None, lang_goal=None, gt_embed=gt_emb)
    out["loss"].backward()
step(); gc.collect()
gc.set_debug(gc.DEBUG_SAVEALL)
step()
n = gc.collect()
print("collected", n, "garbage", len(gc.garbage))
cnt = collections.Counter(type(o).__name__ for o in gc.garbage)
print(cnt.most_common(20))
for o in gc.garbage:
    tn = type(o).__name__
    if "Backward" in tn or tn in ("_PassState",):
        print("==", tn)
        for r in gc.get_referrers(o):
            if r is gc.garbage: continue
            print("   referrer:", type(r).__name__, (list(r.keys())[:8] if isinstance(r, dict) else ""))
