"""GPU diagnostics: where the warp roles of the fused MLP kernel spend their cycles (per-CTA clock64 counters).
    python scripts/fused_roles.py [N]"""
import ctypes as C
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
lib = importlib.import_module(PKG + "._lib").load()

N = int(sys.argv[1]) if len(sys.argv) > 1 else 524288
C_, H, D = 128, 512, 384
mlp = NR.ResnetFC(d_in=42, d_out=4 + D, n_blocks=5, d_latent=C_, d_hidden=H, combine_layer=3)
syn.init_mlp_(mlp, seed=0)
mlp = mlp.cuda()
h = mlp.handle(ops.NRF_PREC_BF16)
fin = torch.zeros(N, h.sizes.kin_pad, device="cuda", dtype=torch.bfloat16)
fin[:, :C_ + 42] = (torch.randn(N, C_ + 42, device="cuda") * 0.5).to(torch.bfloat16)
dfield = torch.zeros(N, h.sizes.dout_pad, device="cuda", dtype=torch.bfloat16)
dfield[:, :4 + D] = (torch.randn(N, 4 + D, device="cuda") * 0.1).to(torch.bfloat16)
out, acts = h.forward(fin)
grads = NR._zero_grads(h)
prof = torch.zeros(148 * 32, device="cuda", dtype=torch.int64)
lib.nrf_debug_set_fused_profile.argtypes = [C.c_void_p]
lib.nrf_debug_set_fused_profile.restype = None


def report(name, fn):
    fn()
    torch.cuda.synchronize()
    prof.zero_()
    lib.nrf_debug_set_fused_profile(C.c_void_p(prof.data_ptr()))
    fn()
    torch.cuda.synchronize()
    lib.nrf_debug_set_fused_profile(None)
    p = prof.view(148, 32).double().cpu()
    lead = p[0::2]
    f = lambda t: f"{t.mean().item() / 1e3:9.1f}k"
    print(f"{name}: (cycles, mean over CTAs)")
    print(f"  producer  total {f(p[:, 0])}  waiting for a free stage {f(p[:, 1])}")
    print(f"  issuer    total {f(lead[:, 2])}  waiting: accumulator free {f(lead[:, 3])}  A operand ready {f(lead[:, 4])}"
          f"  ring stage full {f(lead[:, 5])}")
    ep_tot = p[:, 8:24:2]
    ep_wait = p[:, 9:24:2]
    print(f"  epilogue  total {f(ep_tot)}  waiting for an accumulator {f(ep_wait)}  -> busy {f(ep_tot - ep_wait)}")


report("forward, inference", lambda: h.forward(fin, keep_acts=False))
report("forward, training", lambda: h.forward(fin, acts=acts))
report("backward", lambda: h.backward(fin, acts, dfield, grads))
