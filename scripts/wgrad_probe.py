"""GPU probe: one 512x512 weight-gradient GEMM over N samples (timing, or 3 launches for ncu with --ncu)."""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ops = importlib.import_module("real-robot-nerf-actor_b200.ops")
N = int([a for a in sys.argv[1:] if a.isdigit()][0]) if any(a.isdigit() for a in sys.argv[1:]) else 524288
G = (torch.randn(N, 512, device="cuda") * 0.1).to(torch.bfloat16)
A = torch.relu(torch.randn(N, 512, device="cuda")).to(torch.bfloat16)
dW = torch.zeros(512, 512, device="cuda")
db = torch.zeros(512, device="cuda")
for _ in range(3):
    ops.wgrad(G, A, dW, db)
torch.cuda.synchronize()
if "--ncu" not in sys.argv:
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        ops.wgrad(G, A, dW, db)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 20
    print(f"wgrad 512x512 over {N} samples: {ms:.3f} ms  {2 * N * 512 * 512 / ms / 1e9:.1f} TFLOP/s  "
          f"{N * 2048 / ms / 1e6:.1f} GB/s of operands")
