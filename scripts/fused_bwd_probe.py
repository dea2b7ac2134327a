"""GPU probe: fused backward (one persistent data-gradient kernel + wgrads) vs the layer-by-layer backward.

    python scripts/fused_bwd_probe.py [N ...]
"""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "real-robot-nerf-actor_b200"
NR = importlib.import_module(PKG + ".neural_rendering")
ops = importlib.import_module(PKG + ".ops")
syn = importlib.import_module(PKG + ".synthetic")


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [256, 1000, 524288]
    C, H, D = 128, 512, 384
    mlp = NR.ResnetFC(d_in=42, d_out=4 + D, n_blocks=5, d_latent=C, d_hidden=H, combine_layer=3)
    syn.init_mlp_(mlp, seed=0)
    mlp = mlp.cuda()
    h = mlp.handle(ops.NRF_PREC_BF16)
    names = h.names()
    for N in sizes:
        gx = torch.Generator(device="cuda").manual_seed(N)
        fin = torch.zeros(N, h.sizes.kin_pad, device="cuda", dtype=torch.bfloat16)
        fin[:, :C + 42] = (torch.randn(N, C + 42, device="cuda", generator=gx) * 0.5).to(torch.bfloat16)
        dfield = torch.zeros(N, h.sizes.dout_pad, device="cuda", dtype=torch.bfloat16)
        dfield[:, :4 + D] = (torch.randn(N, 4 + D, device="cuda", generator=gx) * 0.1).to(torch.bfloat16)
        out, acts = h.forward(fin)

        def run(layered):
            grads = NR._zero_grads(h)
            dlat = h.backward(fin, acts, dfield, grads, deterministic=True, layered=layered)
            return dlat, grads

        dl_l, g_l = run(True)
        dl_f, g_f = run(False)
        torch.cuda.synchronize()
        dd = (dl_l - dl_f).abs().max().item()
        worst = max(((g_l[n] - g_f[n]).abs().max().item() / (g_l[n].abs().max().item() + 1e-30), n) for n in names)
        print(f"N={N}: dlatent max|diff| {dd:.3e} (max {dl_l.abs().max().item():.3e}); worst param-grad rel diff "
              f"{worst[0]:.3e} ({worst[1]})", flush=True)
        if N >= 65536:
            variants = (("layered", lambda: h.backward(fin, acts, dfield, g_l, layered=True)),
                        ("fused", lambda: h.backward(fin, acts, dfield, g_f)))
            best = {k: float("inf") for k, _ in variants}
            lib = importlib.import_module(PKG + "._lib")
            for rnd in range(3):
                for name, fn in variants:
                    fn()
                    torch.cuda.synchronize()
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record()
                    for _ in range(5):
                        fn()
                    b.record()
                    torch.cuda.synchronize()
                    best[name] = min(best[name], a.elapsed_time(b) / 5)
            for name, fn in variants:
                lib.timing_begin()
                for _ in range(5):
                    fn()
                k = {c: round(v[0] / 5, 3) for c, v in lib.timing_end().items() if v[1] > 0}
                print(f"   {name:8s} {best[name]:8.3f} ms   per-kernel ms {k}", flush=True)


if __name__ == "__main__":
    main()
