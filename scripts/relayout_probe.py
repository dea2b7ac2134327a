"""Config 5 as ONE rank of N sees it (16384 / N rays of the 200^3 x 128 scene): the step with the dense re-layout of the
input volume against the re-layout of the touched voxels only.  python scripts/relayout_probe.py [out.json]"""
import argparse, importlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench

PKG = bench.PKG
mods = {"NR": importlib.import_module(PKG + ".neural_rendering"), "U": importlib.import_module(PKG + ".utils"),
        "syn": importlib.import_module(PKG + ".synthetic"), "par": importlib.import_module(PKG + ".parallel")}
lib = importlib.import_module(PKG + "._lib")
args = argparse.Namespace(scatter="sorted", volume_layout="contiguous")
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
wl = mods["syn"].CONFIGS["config5"]
res = {}
for n in (8, 4, 2):
    c = bench.Case(mods, wl, "bf16", dev, 0, 1, args, SB=1, n_rays=wl.rays_per_scene // n)
    row = {}
    for rep in range(2):
        for name, ratio in (("dense", 0.0), ("touched voxels", 0.25)):
            c.ren.sparse_relayout_ratio = ratio
            for _ in range(3):
                c.step()
            ms, _ = bench.timed_region(lambda: c.step(), 10, dev, 1)
            lib.timing_begin(); bench.timed_region(lambda: c.step(), 5, dev, 1); k = lib.timing_end()
            row.setdefault(name, []).append({"ms_per_step": round(ms / 10, 3), "relayout_ms": round(k["transpose"][0] / 5, 3),
                                             "relayout_launches": k["transpose"][1] // 5})
    res[f"16384/{n} = {wl.rays_per_scene // n} rays"] = row
    print(n, json.dumps(row))
    del c
    bench.release()
if len(sys.argv) > 1:
    json.dump(res, open(sys.argv[1], "w"), indent=1)
