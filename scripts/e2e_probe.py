"""Where does the gap between bench.py's `value` (inputs resident in HBM) and `e2e` (inputs from pinned host memory,
losses read every step) come from?  Same config-2 step, variants of the host side.  python scripts/e2e_probe.py"""
import argparse, importlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench

PKG = bench.PKG
mods = {"NR": importlib.import_module(PKG + ".neural_rendering"), "U": importlib.import_module(PKG + ".utils"),
        "syn": importlib.import_module(PKG + ".synthetic"), "par": importlib.import_module(PKG + ".parallel")}
args = argparse.Namespace(scatter="sorted", volume_layout="contiguous")
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
wl = mods["syn"].CONFIGS["config2"]
c = bench.Case(mods, wl, "bf16", dev, 0, 1, args)
K = 20

# preallocated, double-buffered device staging for the step's inputs
stage = [[torch.empty_like(t, device=dev) for t in (c.poses_h, c.focal_h, c.gt_rgb_h, c.gt_emb_h)] for _ in range(2)]
done = [None, None]
count = [0]


def step_prealloc():
    ren = c.ren
    c.vol.grad = None
    for p in c.params:
        p.grad = None
    i = count[0] % 2
    count[0] += 1
    main = torch.cuda.current_stream(dev)
    if done[i] is not None:
        c.copy_stream.wait_event(done[i])               # the step that last read these buffers is through
    with torch.cuda.stream(c.copy_stream):
        stage[i][0].copy_(c.poses_h, non_blocking=True)
        stage[i][1].copy_(c.focal_h, non_blocking=True)
        ev_small = torch.cuda.Event(); ev_small.record(c.copy_stream)
        stage[i][2].copy_(c.gt_rgb_h, non_blocking=True)
        stage[i][3].copy_(c.gt_emb_h, non_blocking=True)
        ev_big = torch.cuda.Event(); ev_big.record(c.copy_stream)
    main.wait_event(ev_small)
    ren.target_ready_event = ev_big
    poses, focal, gt_rgb, gt_emb = stage[i]
    out = ren(multi_scale_voxel_list=None, voxel_density=None, language=None, voxel_feat=c.vol, voxel_poses=poses,
              focal=focal, gt_rgb=gt_rgb, gt_depth=None, gt_pose=poses, c=None, lang_goal=None, gt_embed=gt_emb)
    out["loss"].backward()
    done[i] = torch.cuda.Event(); done[i].record(main)
    return out


variants = {
    "device inputs, no read (= value)": (lambda: c.step(), False),
    "device inputs, losses read": (lambda: c.step(), True),
    "host inputs (.to per step), no read": (lambda: c.step(host_inputs=True), False),
    "host inputs (.to per step), losses read (= e2e)": (lambda: c.step(host_inputs=True), True),
    "host inputs (preallocated double buffer), losses read": (step_prealloc, True),
}
res = {}
for rep in range(2):
    for name, (fn, read) in variants.items():
        for _ in range(5):
            fn()
        ms, _ = bench.timed_region(fn, K, dev, 1, read_losses=read)
        res.setdefault(name, []).append(round(ms / K, 3))
for k, v in res.items():
    print(f"{k:60s} {v}")
if len(sys.argv) > 1:
    json.dump(res, open(sys.argv[1], "w"), indent=1)
