"""Step time at the reference's own training shape (nerfact.conf: 512-ray chunks, 64 + 64 samples of which 16
depth-guided, 64 latent channels, 512-d features): a launch-bound step.  python scripts/small_step_probe.py [SB]"""
import importlib, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
PKG = "real-robot-nerf-actor_b200"
NR = importlib.import_module(PKG + ".neural_rendering"); U = importlib.import_module(PKG + ".utils")
syn = importlib.import_module(PKG + ".synthetic"); lib = importlib.import_module(PKG + "._lib")
SB = int(sys.argv[1]) if len(sys.argv) > 1 else 1
dev = torch.device("cuda", 0)
cfg = U.default_config(voxel_shape=100, d_latent=64, d_embed=512, n_coarse=64, n_fine=64, n_fine_depth=16,
                       ray_chunk_size=512, image_width=128, image_height=128)
res = {}
for reuse in (False, True):
    ren = NR.NeuralRenderer(cfg, torch.tensor(syn.BOUNDS))
    syn.init_mlp_(ren.nerf_model.mlp_coarse, seed=0)
    ren = ren.to(dev).train()
    ren.reuse_coarse_evals = reuse
    vol = (torch.randn(SB, 64, 100, 100, 100, device=dev) * 0.1).requires_grad_(True)
    poses = syn.arc_poses(SB).to(dev); focal = torch.tensor(153.0, device=dev)
    gt_rgb = torch.rand(SB, 128, 128, 3, device=dev); gt_emb = torch.randn(SB, 128, 128, 512, device=dev)
    def step():
        vol.grad = None
        for p in ren.parameters(): p.grad = None
        out = ren(multi_scale_voxel_list=None, voxel_density=None, language=None, voxel_feat=vol, voxel_poses=poses,
                  focal=focal, gt_rgb=gt_rgb, gt_depth=None, gt_pose=poses, c=None, lang_goal=None, gt_embed=gt_emb)
        out["loss"].backward()
    for _ in range(5): step()
    torch.cuda.synchronize()
    n0 = lib.launch_count(); t0 = time.perf_counter()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20): step()
    b.record(); torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / 20 * 1e3
    lib.timing_begin(); step(); torch.cuda.synchronize(); kern = lib.timing_end()
    res["reuse" if reuse else "reference schedule"] = {
        "ms_per_step": round(a.elapsed_time(b) / 20, 3), "wall_ms_per_step": round(wall, 3),
        "nrf_launches_per_step": (lib.launch_count() - n0) // 21,
        "kernel_ms_sum": round(sum(v[0] for v in kern.values()), 3),
        "kernel_ms": {k: [round(v[0], 3), v[1]] for k, v in kern.items() if v[1] > 0},
        "evals_per_step": SB * 512 * 192}
# the same step captured into CUDA graphs (graphed.py): the host enqueues 2 graph launches instead of ~56 kernels + ~40 torch ops
G = importlib.import_module(PKG + ".graphed")
ren = NR.NeuralRenderer(cfg, torch.tensor(syn.BOUNDS))
syn.init_mlp_(ren.nerf_model.mlp_coarse, seed=0)
ren = ren.to(dev).train()
vol = (torch.randn(SB, 64, 100, 100, 100, device=dev) * 0.1).requires_grad_(True)
graphed = G.GraphedRenderLoss(ren, vol, poses, focal, gt_rgb, gt_emb)
def gstep():
    vol.grad = None
    for p in ren.parameters(): p.grad = None
    graphed(vol, poses, focal, gt_rgb, gt_emb)["loss"].backward()
for _ in range(5): gstep()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); a.record()
for _ in range(50): gstep()
b.record(); torch.cuda.synchronize()
res["cuda graph (reference schedule)"] = {"ms_per_step": round(a.elapsed_time(b) / 50, 3),
                                          "wall_ms_per_step": round((time.perf_counter() - t0) / 50 * 1e3, 3),
                                          "evals_per_step": SB * 512 * 192}
print(json.dumps({"nerfact.conf shape, SB=%d" % SB: res}))
