"""Stand-alone timing of the MLP GEMM shapes (for ncu and quick A/B runs).
   python scripts/gemm_probe.py [M]"""
import importlib, os, sys, math
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
ops = importlib.import_module("real-robot-nerf-actor_b200.ops")
M = int(sys.argv[1]) if len(sys.argv) > 1 else 524288
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
A = torch.randn(M, 512, device=dev, generator=g).to(torch.bfloat16)
Z = torch.randn(M, 192, device=dev, generator=g).to(torch.bfloat16)
W = (torch.randn(512, 512, device=dev, generator=g) / math.sqrt(512)).to(torch.bfloat16)
Wc = (torch.randn(512, 640, device=dev, generator=g) / math.sqrt(640)).to(torch.bfloat16)
bias = torch.randn(512, device=dev, generator=g)
xres = torch.randn(M, 512, device=dev, generator=g)
act = torch.empty(M, 512, device=dev, dtype=torch.bfloat16)
mask = torch.randn(M, 512, device=dev, generator=g).to(torch.bfloat16)

def bench(name, fn, flops, bytes_):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): fn()
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    print(f"{name:28s} {ms*1e3:8.1f} us  {flops/ms/1e9:7.1f} TFLOP/s  {bytes_/ms/1e6:7.1f} GB/s", flush=True)

xb = xres.to(torch.bfloat16)
gb = torch.empty_like(xb)
F = 2.0 * M * 512 * 512
bench("fc_0 (bias, relu act out)", lambda: ops.gemm(A, W, bias=bias, out_act=act, relu_act=True), F, M * 2048)
bench("plain f32 out", lambda: ops.gemm(A, W, out_f32=xres), F, M * 3072)
bench("fc_1 (K=640, resid, 2 outs)", lambda: ops.gemm(A, Wc, A2=Z[:, :128], bias=bias, resid=xb, out_act=xb, out_act2=act, relu_act2=True), F * 1.25, M * (1024 + 256 + 1024 + 1024 + 1024))
bench("dnet (mask, act out)", lambda: ops.gemm(A, W, mask_src=mask, out_act=act), F, M * 3072)
bench("gx (mask, resid, act out)", lambda: ops.gemm(A, W, mask_src=mask, resid=xb, out_act=gb), F, M * 4096)
G = torch.randn(M, 512, device=dev, generator=g).to(torch.bfloat16)
dW = torch.zeros(512, 512, device=dev)
db = torch.zeros(512, device=dev)
bench("wgrad 512x512 (no bias)", lambda: ops.wgrad(G, A, dW), F, M * 2048)
bench("wgrad 512x512 (+bias)", lambda: ops.wgrad(G, A, dW, db), F, M * 2048)
bench("wgrad 512x128 (+bias)", lambda: ops.wgrad(G, Z[:, :128], torch.zeros(512, 128, device=dev), db), F / 4, M * 1280)
