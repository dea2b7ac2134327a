"""GPU probe: fused whole-MLP forward kernel vs the layer-by-layer GEMM chain (same inputs, same packed weights).

    python scripts/fused_probe.py [N ...]
Prints max |diff| of the raw field outputs and of every saved operand, and the time of both paths.
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
    only = None
    argv = list(sys.argv[1:])
    if "--only" in argv:                         # --only infer|train: just that path, 3 launches (for ncu)
        i = argv.index("--only")
        only = argv[i + 1]
        del argv[i:i + 2]
    sizes = [int(a) for a in argv] or [256, 1000, 65536, 524288]
    C, H, D = 128, 512, 384
    mlp = NR.ResnetFC(d_in=42, d_out=4 + D, n_blocks=5, d_latent=C, d_hidden=H, combine_layer=3)
    syn.init_mlp_(mlp, seed=0)
    g = torch.Generator().manual_seed(5)
    with torch.no_grad():
        for n, p in mlp.named_parameters():          # non-zero biases so the bias path is exercised
            if n.endswith(".bias"):
                p.copy_(torch.randn(p.shape, generator=g) * 0.1)
    mlp = mlp.cuda()
    h = mlp.handle(ops.NRF_PREC_BF16)
    print("fused supported:", h.fused)
    for N in sizes:
        gx = torch.Generator(device="cuda").manual_seed(N)
        fin = torch.zeros(N, h.sizes.kin_pad, device="cuda", dtype=torch.bfloat16)
        fin[:, :C + 42] = (torch.randn(N, C + 42, device="cuda", generator=gx) * 0.5).to(torch.bfloat16)
        if only:
            acts = None
            for _ in range(3):
                out, acts = h.forward(fin, acts=acts, keep_acts=(only == "train"))
            torch.cuda.synchronize()
            print("ran", only, N, float(out.abs().max()))
            continue
        out_l, acts_l = h.forward(fin, layered=True)
        torch.cuda.synchronize()
        out_f, acts_f = h.forward(fin)
        torch.cuda.synchronize()
        out_i, acts_i = h.forward(fin, keep_acts=False)
        torch.cuda.synchronize()
        assert acts_i is None
        d = (out_f - out_l).abs().max().item()
        di = (out_i - out_f).abs().max().item()
        ref = out_l.abs().max().item()
        slots = 11
        al = acts_l.view(torch.bfloat16)[: slots * N * H].view(slots, N, H).float()
        af = acts_f.view(torch.bfloat16)[: slots * N * H].view(slots, N, H).float()
        ds = [(al[s] - af[s]).abs().max().item() for s in range(slots)]
        print(f"N={N}: out max|diff| fused-vs-layered {d:.3e} (|out| max {ref:.3f}), inference-vs-train {di:.3e}; "
              f"acts max|diff| per slot {['%.2e' % v for v in ds]}", flush=True)
        if N >= 65536:
            variants = (("layered", lambda: h.forward(fin, acts=acts_l, layered=True)),
                        ("fused train", lambda: h.forward(fin, acts=acts_f)),
                        ("fused infer", lambda: h.forward(fin, keep_acts=False)))
            best = {name: float("inf") for name, _ in variants}
            for rnd in range(3):                  # interleaved rounds, best of 3: clocks drift between boxes / runs
                for name, fn in variants:
                    for _ in range(2):
                        fn()
                    torch.cuda.synchronize()
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record()
                    for _ in range(10):
                        fn()
                    b.record()
                    torch.cuda.synchronize()
                    best[name] = min(best[name], a.elapsed_time(b) / 10)
            for name, ms in best.items():
                print(f"   {name:12s} {ms:8.3f} ms  {N * 6_076_416 / ms / 1e9:8.1f} TFLOP/s (algorithmic)  "
                      f"x{best['layered'] / ms:.2f} vs layered", flush=True)


if __name__ == "__main__":
    main()
