import gc, importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
PKG = "real-robot-nerf-actor_b200"
NR = importlib.import_module(PKG + ".neural_rendering"); U = importlib.import_module(PKG + ".utils")
syn = importlib.import_module(PKG + ".synthetic"); lib = importlib.import_module(PKG + "._lib")
wl = syn.CONFIGS["config2"]; dev = torch.device("cuda", 0)
cfg = U.default_config(voxel_shape=wl.S, n_coarse=wl.n_coarse, n_fine=wl.n_fine, ray_chunk_size=wl.rays_per_scene)
ren = NR.NeuralRenderer(cfg, torch.tensor(syn.BOUNDS)); syn.init_mlp_(ren.nerf_model.mlp_coarse); ren = ren.to(dev).train()
SB = wl.SB
vol = (torch.randn(SB, wl.C, wl.S, wl.S, wl.S, device=dev) * 0.1).requires_grad_(True)
poses = syn.arc_poses(SB).to(dev); focal = torch.tensor(wl.focal, device=dev)
gt_rgb = torch.rand(SB, wl.H, wl.W, 3, device=dev); gt_emb = torch.randn(SB, wl.H, wl.W, wl.D, device=dev)
GiB = 2.0**30
def mem(tag):
    print(f"{tag}: allocated {torch.cuda.memory_allocated()/GiB:.2f} GiB reserved {torch.cuda.memory_reserved()/GiB:.2f} peak {torch.cuda.max_memory_allocated()/GiB:.2f}", flush=True)
def step():
    vol.grad = None
    for p in ren.parameters(): p.grad = None
    out = ren(multi_scale_voxel_list=None, voxel_density=None, language=None, voxel_feat=vol, voxel_poses=poses,
              focal=focal, gt_rgb=gt_rgb, gt_depth=None, gt_pose=poses, c=None, lang_goal=None, gt_embed=gt_emb)
    mem("  after fwd")
    out["loss"].backward()
    mem("  after bwd")
mem("start")
for i in range(4):
    step(); mem(f"step {i}")
gc.collect(); mem("after gc")
s = lib.load(); 
sz = ren.nerf_model.mlp_coarse.handle(0).sizes
print("sizes", sz.kin_pad, sz.dout_pad, sz.packed_bytes, sz.fwd_bytes_per_sample, sz.bwd_bytes_per_sample, sz.bwd_fixed_bytes)
def timeit(n, timing):
    torch.cuda.synchronize(); 
    if timing: lib.timing_begin()
    t0 = time.perf_counter()
    for _ in range(n):
        vol.grad = None
        for p in ren.parameters(): p.grad = None
        out = ren(multi_scale_voxel_list=None, voxel_density=None, language=None, voxel_feat=vol, voxel_poses=poses,
              focal=focal, gt_rgb=gt_rgb, gt_depth=None, gt_pose=poses, c=None, lang_goal=None, gt_embed=gt_emb)
        out["loss"].backward()
    torch.cuda.synchronize(); dt = (time.perf_counter()-t0)/n*1e3
    if timing: lib.timing_end()
    return dt
print("no timing:", timeit(5, False), "ms/step")
print("with per-kernel events:", timeit(5, True), "ms/step")
print("no timing again:", timeit(5, False), "ms/step")
