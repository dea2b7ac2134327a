"""Which rounding pattern of utils.py:444-506 does CUDA-eager PyTorch execute?  Runs the oracle's gen_rays ON THE DEVICE
(ATen's CUDA kernels + cuBLAS' batched K = 3 product: what the reference runs on a GPU) against every flag combination
of nrf_raygen_ex and prints the bit-identical fraction of each.  Identity poses isolate the unprojection map (pixel
coordinate + norm), the arc poses add the direction product.  python scripts/raygen_probe.py [out.json]"""
import importlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
PKG = "real-robot-nerf-actor_b200"
ops = importlib.import_module(PKG + ".ops"); syn = importlib.import_module(PKG + ".synthetic")
from oracle import nerf_oracle as O

dev = torch.device("cuda", 0)


def frac(a, b):
    return float((a.view(torch.int32) == b.view(torch.int32)).float().mean())


cases = {
    "config2 128x128 f=153": (syn.arc_poses(2), 128, 128, torch.tensor(153.0), None),
    "160x120 f=(201.3,199.1) c=(77.2,61.9)": (syn.arc_poses(5), 160, 120, torch.tensor([201.3, 199.1]),
                                              torch.tensor([77.2, 61.9])),
    "80x60 f=76.18187": (syn.arc_poses(3), 80, 60, torch.tensor(76.18187), None),
}
res = {}
for name, (poses, W, H, f, c) in cases.items():
    eye = torch.eye(4)[None].repeat(poses.shape[0], 1, 1)
    per = {}
    for tag, P in (("identity", eye), ("arc", poses)):
        P = P.to(dev)
        ref = O.gen_rays(P, W, H, f.to(dev), 1.2, 4.0, c=None if c is None else c.to(dev))
        tab = {}
        for norm in (0, 8, 16, 24):
            for recip in (0, 4):
                for d in range(4):
                    fl = norm | recip | d
                    r = ops.raygen(P, W, H, f, 1.2, 4.0, c=c, flags=fl)
                    tab[fl] = (frac(r[..., 3:6], ref[..., 3:6]), float((r - ref).abs().max()))
        per[tag] = tab
    res[name] = per
    for tag, P in (("identity", eye), ("arc", poses)):      # a few mismatching elements of the best unprojection
        P = P.to(dev)
        ref = O.gen_rays(P, W, H, f.to(dev), 1.2, 4.0, c=None if c is None else c.to(dev))
        r = ops.raygen(P, W, H, f, 1.2, 4.0, c=c, flags=ops.NRF_RAYGEN_CUDA_EAGER)
        bad = (r[..., 3:6].view(torch.int32) != ref[..., 3:6].view(torch.int32)).nonzero()[:6]
        for b_, i_, j_, k_ in bad.tolist():
            print(f"  {tag} img {b_} px ({i_},{j_}) comp {k_}: ours {float(r[b_, i_, j_, 3 + k_])!r} ref "
                  f"{float(ref[b_, i_, j_, 3 + k_])!r}; pose row {P[b_, k_, :3].tolist()}")
    best = max(per["arc"], key=lambda k: per["arc"][k][0])
    print(f"{name}: best flags {best} -> arc {per['arc'][best]}, identity {per['identity'][best]}; "
          f"flags 0 -> arc {per['arc'][0]}, identity {per['identity'][0]}")
exact = [fl for fl in res[next(iter(res))]["arc"]
         if all(res[n][t][fl][0] == 1.0 for n in res for t in ("identity", "arc"))]
print("flag combinations bit-identical to CUDA-eager in every case:", exact)
out = {"exact_everywhere": exact, "cases": {n: {t: {str(k): v for k, v in tab.items()} for t, tab in per.items()}
                                           for n, per in res.items()}}
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], "w"), indent=1)
