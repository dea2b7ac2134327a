#!/usr/bin/env python
"""bench.py -- render fwd+bwd throughput of the feature-NeRF hot path on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload config2]

One "step" = one NeuralRenderer training step on BASELINE.json's config 2 (per GPU: 2 scenes x 2048 rays,
64 coarse + 64 importance samples, 100^3 x 128-channel volume, ResnetFC 512 x 5 blocks, RGB + 384-d
feature heads): gen_rays -> ray subsample -> forward_nerf (coarse + fine) -> losses -> backward into the
voxel volume and all MLP parameters (+ one NCCL all-reduce of the MLP gradients when N > 1).
Metric: ray-samples/s = field evaluations R*(Kc + Kc+Kf) per step / step time.

`--impl reference` times the reference's algorithm on the host CPU (the oracle port, oracle/nerf_oracle.py:
torch CPU fp32, all host threads) on a bounded sample of the same workload.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
PKG = "real-robot-nerf-actor_b200"

FLOP_FWD = 6_076_416          # per field evaluation
FLOP_DGRAD = 6_033_408        # fwd minus the 42->512 input layer (no dgrad into PE / viewdirs)
FLOP_WGRAD = 6_076_416
FLOP_STEP = FLOP_FWD + FLOP_DGRAD + FLOP_WGRAD     # 18 186 240
# measured DRAM traffic of mlp_fused_kernel per launch at config 2 (ncu, profiles/r01c_fused_ncu.md):
# (3.607 + 7.263 + 6.700 + 3.326) GB over the 4 launches of a step
FUSED_DRAM_BYTES_PER_LAUNCH = 5.22e9


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d["hbm_gbs"], bf16_tflops=d["bf16_tflops"],
                    bf16_tflops_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]), source="measured")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "200", "-i", str(gpu_index)], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, smax, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1])); smax.append(float(r[2])); power.append(float(r[3]))
            except ValueError:
                continue
            for n, v in zip(names, r[5:9]):
                if v.strip().lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(smax), "reasons": sorted(reasons),
                "power_w_max": max(power), "samples": len(sm)}


# --------------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    NR = importlib.import_module(PKG + ".neural_rendering")
    U = importlib.import_module(PKG + ".utils")
    syn = importlib.import_module(PKG + ".synthetic")
    par = importlib.import_module(PKG + ".parallel")
    lib = importlib.import_module(PKG + "._lib")

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world} (launch with torch.distributed.run)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    wl = syn.CONFIGS[args.workload]
    SB, n_rays = wl.SB, wl.rays_per_scene

    cfg = U.default_config(voxel_shape=wl.S, d_latent=wl.C, d_embed=wl.D, n_coarse=wl.n_coarse, n_fine=wl.n_fine,
                           ray_chunk_size=n_rays, image_width=wl.W, image_height=wl.H)
    ren = NR.NeuralRenderer(cfg, torch.tensor(syn.BOUNDS), precision=args.precision)
    syn.init_mlp_(ren.nerf_model.mlp_coarse, seed=0)
    ren = ren.to(dev).train()
    ren.scatter = args.scatter
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    vol = torch.randn(SB, wl.C, wl.S, wl.S, wl.S, device=dev, generator=g) * 0.1
    if args.volume_layout == "channels_last_3d":       # what a conv3d producer run in that memory format hands over
        vol = vol.contiguous(memory_format=torch.channels_last_3d)
    vol.requires_grad_(True)
    # host-side inputs of a training step (what the data loader hands over), pinned
    poses_h = syn.arc_poses(SB).pin_memory()
    focal_h = torch.tensor(wl.focal, dtype=torch.float32).pin_memory()
    gh = torch.Generator().manual_seed(99 + rank)
    gt_rgb_h = torch.rand(SB, wl.H, wl.W, 3, generator=gh).pin_memory()
    gt_emb_h = torch.randn(SB, wl.H, wl.W, wl.D, generator=gh).pin_memory()
    h2d_bytes = sum(t.numel() * t.element_size() for t in (poses_h, focal_h, gt_rgb_h, gt_emb_h))
    poses_d, focal_d = poses_h.to(dev), focal_h.to(dev)
    gt_rgb_d, gt_emb_d = gt_rgb_h.to(dev), gt_emb_h.to(dev)
    params = [p for p in ren.parameters()]
    if world > 1 and args.allreduce == "overlap":     # ONE NCCL all-reduce of the flat MLP gradient, started inside the
        par.overlap_mlp_grad_allreduce(ren)           # backward: it runs under the volume-gradient scatter

    copy_stream = torch.cuda.Stream(device=dev)

    def step(host_inputs: bool):
        vol.grad = None
        for p in params:
            p.grad = None
        if host_inputs:
            # what a training loop does: the step's inputs come from pinned host memory; the copies run on a side
            # stream so that the 50 MB of target features (needed only by the losses) overlap the coarse pass
            main = torch.cuda.current_stream(dev)
            with torch.cuda.stream(copy_stream):
                poses, focal = poses_h.to(dev, non_blocking=True), focal_h.to(dev, non_blocking=True)
                ev_small = torch.cuda.Event()
                ev_small.record(copy_stream)
                gt_rgb, gt_emb = gt_rgb_h.to(dev, non_blocking=True), gt_emb_h.to(dev, non_blocking=True)
                ev_big = torch.cuda.Event()
                ev_big.record(copy_stream)
            main.wait_event(ev_small)
            for t in (poses, focal, gt_rgb, gt_emb):
                t.record_stream(main)
            ren.target_ready_event = ev_big          # compute_rendering_loss waits for it right before the losses
        else:
            poses, focal, gt_rgb, gt_emb = poses_d, focal_d, gt_rgb_d, gt_emb_d
            ren.target_ready_event = None
        out = ren(multi_scale_voxel_list=None, voxel_density=None, language=None, voxel_feat=vol,
                  voxel_poses=poses, focal=focal, gt_rgb=gt_rgb, gt_depth=None, gt_pose=poses, c=None,
                  lang_goal=None, gt_embed=gt_emb)
        out["loss"].backward()
        if world > 1 and args.allreduce == "post":
            par.allreduce_mlp_grads(ren)
        return out

    def timed(n_steps, host_inputs):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        prev, seen = None, 0.0
        for _ in range(n_steps):
            out = step(host_inputs)
            if host_inputs:
                # the step's result is read on the host EVERY step: the loss dictionary copies its scalars to pinned
                # memory asynchronously (LossDict), so a training loop logs step i while step i+1 is queued; the
                # last step's values are read before the timed region closes
                if prev is not None:
                    seen += prev["loss_rgb"] + prev["loss_embed"] + prev["loss_depth"]
                prev = out
        if host_inputs and prev is not None:
            seen += prev["loss_rgb"] + prev["loss_embed"] + prev["loss_depth"]
            assert seen == seen, "loss is NaN"
        b.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = torch.tensor([a.elapsed_time(b)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for _ in range(args.warmup):
        step(False)
    torch.cuda.synchronize()
    sampler = ClockSampler(local) if rank == 0 else None
    time.sleep(0.3)
    launches0 = lib.launch_count()
    ms_total = timed(args.steps, False)                 # the timed region behind `value`
    launches = lib.launch_count() - launches0
    # same K steps again with a CUDA-event pair around every kernel launch (recorded by the library on the
    # launching stream): per-kernel durations for the roofline; kept out of `value` because ~500 event
    # records per step add launch gaps
    lib.timing_begin()
    ms_total_ev = timed(args.steps, False)
    kern = lib.timing_end()
    clocks = sampler.stop() if sampler else None
    # end-to-end: host (pinned) inputs copied in, loss read back, every step
    step(True)
    ms_e2e = timed(args.steps, True)
    # optional mode, reported beside the headline and never mixed into it: the fine pass reuses the coarse pass's
    # field evaluations (bit-identical rendering; the MLP runs on Kc + Kf instead of Kc + (Kc + Kf) samples per ray)
    ms_reuse = None
    if not args.no_reuse_line:
        ren.reuse_coarse_evals = True
        for _ in range(2):
            step(False)
        ms_reuse = timed(args.steps, False)
        ren.reuse_coarse_evals = False

    evals_step = wl.evals * world
    ms_step = ms_total / args.steps
    value = evals_step / (ms_step * 1e-3)
    e2e_value = evals_step / (ms_e2e / args.steps * 1e-3)
    pk = peaks()
    # dominant kernel: mlp_fused_kernel (csrc/mlp_fused.cu), the whole forward MLP and the whole data-gradient chain
    # as one persistent tcgen05 kernel each (2 + 2 launches per step: coarse and fine pass).  Algorithmic FLOPs it
    # executes per step = (forward + dgrad) FLOP per evaluation x this rank's evaluations, minus the small dL/dz GEMM
    # that stays a separate launch; duration = sum over its launches (CUDA events on the launching stream).
    # (--precision fp32 or NRF_MLP_LAYERED=1: the per-layer gemm kernels are the dominant ones instead.)
    wg_ms, wg_n = kern["wgrad_tc"]
    fused_ms = kern["fused_fwd"][0] + kern["fused_bwd"][0]
    fused_n = kern["fused_fwd"][1] + kern["fused_bwd"][1]
    flop_dz = 2 * 3 * 512 * wl.C                      # dL/dz = [g_0|g_1|g_2] . W_z: separate GEMM
    if fused_n > 0:
        dom_name = "mlp_fused_kernel (whole-MLP forward + whole data-gradient chain, 4 launches/step)"
        dom_ms, dom_n = fused_ms, fused_n
        dom_flops = (FLOP_FWD + FLOP_DGRAD - flop_dz) * wl.evals * args.steps
    else:
        dom_name = "gemm_tc_kernel / gemm_simt_kernel (per-layer forward + dgrad GEMMs)"
        dom_ms = kern["gemm_tc"][0] + kern["simt"][0]
        dom_n = kern["gemm_tc"][1] + kern["simt"][1]
        dom_flops = (FLOP_FWD + FLOP_DGRAD) * wl.evals * args.steps
    achieved = dom_flops / (dom_ms * 1e-3) / 1e12 if dom_ms > 0 else 0.0
    roof = {"bound": "tensor", "kernel": dom_name, "achieved": round(achieved, 1),
            "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s", "frac": round(achieved / pk["bf16_tflops_sustained"], 4),
            "traffic": FUSED_DRAM_BYTES_PER_LAUNCH if fused_n > 0 and args.workload == "config2" else None,
            "traffic_source": "profiles/r01c_fused_ncu.md: dram__bytes_read.sum + dram__bytes_write.sum of the 4 "
                              "mlp_fused_kernel launches of one config-2 step (ncu --set full), averaged per launch; "
                              "algorithmic bytes per launch = 5.13e9 (13.5 KB/eval forward, 12.7 KB/eval backward)",
            "peak_source": pk["source"] + " (sustained: kernel timed inside a long step)",
            "launches": dom_n, "avg_launch_ms": round(dom_ms / max(dom_n, 1), 4),
            "algorithmic_flop_per_eval": {"fwd": FLOP_FWD, "dgrad": FLOP_DGRAD, "wgrad": FLOP_WGRAD},
            "wgrad_tc_tflops": round(FLOP_WGRAD * wl.evals * args.steps / (wg_ms * 1e-3) / 1e12, 1) if wg_ms > 0 else None,
            "step_frac_of_tensor_peak": round(FLOP_STEP * wl.evals / (ms_step * 1e-3) / 1e12 / pk["bf16_tflops_sustained"], 4)}
    kernel_ms = {k: round(v[0] / args.steps, 3) for k, v in kern.items() if v[1] > 0}
    line = {"metric": "render fwd+bwd ray-samples/s", "value": round(value, 1), "unit": "ray-samples/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_step, 3),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.precision, "data": "synthetic",
            "config": {"workload": f"{wl.name}: per GPU {SB} scenes x {n_rays} rays, {wl.n_coarse}+{wl.n_fine} samples, "
                                   f"{wl.S}^3 x {wl.C}ch volume, ResnetFC 512x5, RGB+{wl.D}d heads, fwd+bwd",
                       "evals_per_step": evals_step, "precision": args.precision, "scatter": args.scatter, "volume_layout": args.volume_layout,
                       "l2": "working set (1 GiB volume + ~20 GiB activations per step) >> 126 MB L2; no flush needed",
                       "parallelism": f"dp{world} over scenes; one NCCL all-reduce of the MLP grads ({args.allreduce})" if world > 1 else "single GPU"},
            "e2e": {"value": round(e2e_value, 1), "unit": "ray-samples/s", "ms_per_step": round(ms_e2e / args.steps, 3),
                    "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4 * 7,
                    "note": "poses, focal, gt_rgb, gt_embed copied from pinned host memory each step (side stream; the targets "
                            "are awaited right before the losses); the 7 loss scalars of every step are copied to pinned host memory "
                            "and read by the host one step later (LossDict), the last step's inside the timed region; "
                            "the voxel volume is device-resident "
                            "as in the reference (PerAct encoder output)"},
            "reuse_coarse_evals": None if ms_reuse is None else {
                "ms_per_step": round(ms_reuse / args.steps, 3),
                "value": round(evals_step / (ms_reuse / args.steps * 1e-3), 1), "unit": "ray-samples/s",
                "mlp_evals_per_step": world * SB * n_rays * (wl.n_coarse + wl.n_fine),
                "note": "opt-in NeuralRenderer.reuse_coarse_evals=True: same rendered samples per step (value counts "
                        "them as the headline does), the MLP evaluates each distinct sample once; NOT the headline"},
            "gpu_launches": launches, "clocks": clocks, "roofline": roof, "kernel_ms_per_step": kernel_ms,
            "ms_per_step_with_kernel_events": round(ms_total_ev / args.steps, 3)}
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(rays=args.cpu_rays, reps=1)
            if wl.train:                     # second half of the baseline leg: the same oracle port on the same GPU
                line["cpu_baseline"]["eager_pytorch_same_gpu"] = eager_gpu_baseline(dev, wl)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def eager_gpu_baseline(dev, wl, steps=2):
    """Part of the baseline leg (`cpu_baseline.eager_pytorch_same_gpu`), the only other place besides cpu_baseline()
    where bench.py executes oracle/ -- as the thing compared against, never on the product path.
    The reference's algorithm as eager PyTorch on the SAME GPU: the oracle restatement (oracle/nerf_oracle.py --
    the reference's own ATen ops: F.grid_sample, F.linear, cumprod, sort, searchsorted, in the reference's chunks of
    eval_batch_size points) at the FULL config-2 step, fp32 (TF32 off, as in the reference).  Reported only; the
    reference itself cannot travel to the GPU box (/root/reference is absent there)."""
    import torch
    from oracle import nerf_oracle as O
    syn = importlib.import_module(PKG + ".synthetic")
    try:
        SB, n_rays = wl.SB, wl.rays_per_scene
        params = {k: v.requires_grad_(True) for k, v in
                  O.init_params(d_in=42, d_latent=wl.C, d_hidden=512, d_out=4 + wl.D, seed=0, device=dev).items()}
        g = torch.Generator(device=dev).manual_seed(7)
        vol = (torch.randn(SB, wl.C, wl.S, wl.S, wl.S, device=dev, generator=g) * 0.1).requires_grad_(True)
        rays_all = O.gen_rays(syn.arc_poses(SB).to(dev), wl.W, wl.H, torch.tensor(wl.focal), 1.2, 4.0).reshape(SB, -1, 8)
        gt_rgb, gt_emb = (t.to(dev) for t in syn.make_targets(SB, n_rays, wl.D))

        def step(i):
            vol.grad = None
            for p in params.values():
                p.grad = None
            idx = syn.pick_ray_indices(wl.H * wl.W, n_rays, seed=i).to(dev)
            noise = {k: v.to(dev) for k, v in syn.make_noise(SB * n_rays, wl.n_coarse, wl.n_fine, seed=i).items()}
            out = O.forward_nerf(params, vol, rays_all[:, idx], syn.BOUNDS, wl.n_coarse, wl.n_fine, noise=noise)
            O.rendering_loss(out, gt_rgb, gt_emb)["loss"].backward()

        step(0)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(steps):
            step(i + 1)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / steps
        peak = torch.cuda.max_memory_allocated(dev) / 2 ** 30
        del params, vol
        torch.cuda.empty_cache()
        return {"value": round(wl.evals / (ms * 1e-3), 1), "unit": "ray-samples/s", "ms_per_step": round(ms, 1),
                "kind": "port", "peak_mem_gib": round(peak, 1),
                "sample": f"oracle port in eager PyTorch (fp32, reference chunking) on the same B200: the full "
                          f"{wl.name} step, {SB} x {n_rays} rays x ({wl.n_coarse} + {wl.n_coarse + wl.n_fine}) samples, "
                          f"mean of {steps} steps"}
    except Exception as e:                      # reported, never fatal for the bench line
        return {"unavailable": f"{type(e).__name__}: {e}"[:200]}


# ------------------------------------------------------------------ CPU baseline / reference arm
def _cpu_step_factory(rays_per_scene, S=100, C=128, D=384, Kc=64, Kf=64):
    """One CPU training step of the oracle port on a bounded sample of config 2 (1 scene)."""
    import torch
    from oracle import nerf_oracle as O
    syn = importlib.import_module(PKG + ".synthetic")
    torch.set_num_threads(os.cpu_count() or 1)
    params = {k: v.requires_grad_(True) for k, v in
              O.init_params(d_in=42, d_latent=C, d_hidden=512, d_out=4 + D, seed=0).items()}
    g = torch.Generator().manual_seed(7)
    vol = (torch.randn(1, C, S, S, S, generator=g) * 0.1).requires_grad_(True)
    poses = syn.arc_poses(1)
    rays_all = O.gen_rays(poses, 128, 128, torch.tensor(153.0), 1.2, 4.0).reshape(1, -1, 8)
    gt_rgb, gt_emb = syn.make_targets(1, rays_per_scene, D)

    def step(i):
        vol.grad = None
        for p in params.values():
            p.grad = None
        idx = syn.pick_ray_indices(128 * 128, rays_per_scene, seed=i)
        noise = syn.make_noise(rays_per_scene, Kc, Kf, seed=i)
        out = O.forward_nerf(params, vol, rays_all[:, idx], syn.BOUNDS, Kc, Kf, noise=noise)
        O.rendering_loss(out, gt_rgb, gt_emb)["loss"].backward()
    evals = rays_per_scene * (Kc + Kc + Kf)
    return step, evals


def cpu_baseline(rays=256, reps=1):
    step, evals = _cpu_step_factory(rays)
    step(0) if rays <= 64 else _cpu_step_factory(32)[0](0)       # warm the thread pool / allocator
    best = float("inf")
    for i in range(reps):
        t0 = time.perf_counter()
        step(i + 1)
        best = min(best, time.perf_counter() - t0)
    return {"value": round(evals / best, 1), "unit": "ray-samples/s", "cores": os.cpu_count(), "kind": "port",
            "sample": f"oracle port (torch CPU fp32, reference chunking): 1 scene x {rays} rays x (64 + 128) samples, "
                      f"100^3 x 128ch volume, fwd+bwd, best of {reps}: {best:.2f} s"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    step, evals = _cpu_step_factory(args.cpu_rays)
    for i in range(args.warmup):
        step(1000 + i)
    t0 = time.perf_counter()
    for i in range(args.steps):
        step(i)
    dt = time.perf_counter() - t0
    value = evals * args.steps / dt
    wl = importlib.import_module(PKG + ".synthetic").CONFIGS[args.workload]
    line = {"impl": "reference", "metric": "render fwd+bwd ray-samples/s", "value": round(value, 1),
            "unit": "ray-samples/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(dt / args.steps * 1e3, 1), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            "config": {"workload": f"{wl.name}: bounded CPU sample, 1 scene x {args.cpu_rays} rays, 64+64 samples, "
                                   "100^3 x 128ch volume, ResnetFC 512x5, RGB+384d heads, fwd+bwd",
                       "evals_per_step": evals},
            "cpu_baseline": {"value": round(value, 1), "unit": "ray-samples/s", "cores": os.cpu_count(), "kind": "port",
                             "sample": f"{args.steps} steps of 1 scene x {args.cpu_rays} rays x (64+128) samples"},
            "e2e": {"value": round(value, 1), "unit": "ray-samples/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="config2")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp16", "fp32"])
    ap.add_argument("--cpu-rays", type=int, default=128, dest="cpu_rays")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-reuse-line", action="store_true", dest="no_reuse_line")
    ap.add_argument("--allreduce", default="overlap", choices=["overlap", "post"],
                    help="N > 1: MLP-gradient all-reduce started inside the backward (default) or after it")
    ap.add_argument("--scatter", default="sorted", choices=["atomic", "sorted"])
    ap.add_argument("--volume-layout", default="contiguous", choices=["contiguous", "channels_last_3d"],
                    dest="volume_layout", help="memory format of the voxel volume handed to the renderer (default: "
                    "the reference's contiguous (SB,C,S,S,S); channels_last_3d skips both re-layout passes)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
