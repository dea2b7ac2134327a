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
# measured DRAM traffic of mlp_fused_kernel per launch at config 2 (ncu --set full, one step, 4 launches): read from
# profiles/fused_traffic.json (written from the committed capture by scripts/summarize_ncu.py traffic)
def _fused_traffic():
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "fused_traffic.json")))
        return float(d["dram_bytes_per_launch"]), d["source"]
    except Exception:
        return 5.22e9, "profiles/r02b_fused_ncu.md"


FUSED_DRAM_BYTES_PER_LAUNCH, FUSED_DRAM_SOURCE = _fused_traffic()


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
class Case:
    """One workload on this rank: renderer + device-resident volume + the host (pinned) and device copies of a training
    step's inputs."""

    def __init__(self, mods, wl, precision, dev, rank, world, args, SB=None, n_rays=None, focal=None, near_far=None,
                 vol_seed=None, share_from=None):
        import torch
        NR, U, syn = mods["NR"], mods["U"], mods["syn"]
        self.wl, self.dev, self.world, self.rank = wl, dev, world, rank
        self.SB = SB = wl.SB if SB is None else SB
        self.n_rays = n_rays = wl.rays_per_scene if n_rays is None else n_rays
        near, far = near_far if near_far is not None else (1.2, 4.0)
        cfg = U.default_config(voxel_shape=wl.S, d_latent=wl.C, d_embed=wl.D, n_coarse=wl.n_coarse, n_fine=wl.n_fine,
                               ray_chunk_size=n_rays, image_width=wl.W, image_height=wl.H, z_near=near, z_far=far)
        ren = NR.NeuralRenderer(cfg, torch.tensor(syn.BOUNDS), precision=precision)
        syn.init_mlp_(ren.nerf_model.mlp_coarse, seed=0)
        self.ren = ren = ren.to(dev).train()
        ren.scatter = args.scatter
        self.params = [p for p in ren.parameters()]
        if share_from is not None:                       # same volume / inputs, another precision mode
            for k in ("vol", "poses_h", "focal_h", "gt_rgb_h", "gt_emb_h", "poses_d", "focal_d", "gt_rgb_d", "gt_emb_d",
                      "h2d_bytes"):
                setattr(self, k, getattr(share_from, k))
            self.copy_stream = share_from.copy_stream
            return
        g = torch.Generator(device=dev).manual_seed(1234 + (rank if vol_seed is None else vol_seed))
        vol = torch.randn(SB, wl.C, wl.S, wl.S, wl.S, device=dev, generator=g) * 0.1
        if args.volume_layout == "channels_last_3d":     # what a conv3d producer run in that memory format hands over
            vol = vol.contiguous(memory_format=torch.channels_last_3d)
        self.vol = vol.requires_grad_(True)
        # host-side inputs of a training step (what the data loader hands over), pinned
        self.poses_h = syn.arc_poses(SB).pin_memory()
        self.focal_h = torch.tensor(wl.focal if focal is None else focal, dtype=torch.float32).pin_memory()
        gh = torch.Generator().manual_seed(99 + rank)
        self.gt_rgb_h = torch.rand(SB, wl.H, wl.W, 3, generator=gh).pin_memory()
        self.gt_emb_h = torch.randn(SB, wl.H, wl.W, wl.D, generator=gh).pin_memory()
        self.h2d_bytes = sum(t.numel() * t.element_size() for t in (self.poses_h, self.focal_h, self.gt_rgb_h, self.gt_emb_h))
        self.poses_d, self.focal_d = self.poses_h.to(dev), self.focal_h.to(dev)
        self.gt_rgb_d, self.gt_emb_d = self.gt_rgb_h.to(dev), self.gt_emb_h.to(dev)
        self.copy_stream = torch.cuda.Stream(device=dev)

    @property
    def evals(self):
        fine = (self.wl.n_coarse + self.wl.n_fine) if self.wl.n_fine > 0 else 0
        return self.SB * self.n_rays * (self.wl.n_coarse + fine)

    def step(self, host_inputs=False, post_allreduce=None, volume_allreduce=None):
        import torch
        ren, dev = self.ren, self.dev
        self.vol.grad = None
        for p in self.params:
            p.grad = None
        if host_inputs:
            # what a training loop does: the step's inputs come from pinned host memory; the copies run on a side
            # stream so that the 50 MB of target features (needed only by the losses) overlap the coarse pass
            main = torch.cuda.current_stream(dev)
            with torch.cuda.stream(self.copy_stream):
                poses, focal = self.poses_h.to(dev, non_blocking=True), self.focal_h.to(dev, non_blocking=True)
                ev_small = torch.cuda.Event()
                ev_small.record(self.copy_stream)
                gt_rgb, gt_emb = self.gt_rgb_h.to(dev, non_blocking=True), self.gt_emb_h.to(dev, non_blocking=True)
                ev_big = torch.cuda.Event()
                ev_big.record(self.copy_stream)
            main.wait_event(ev_small)
            for t in (poses, focal, gt_rgb, gt_emb):
                t.record_stream(main)
            ren.target_ready_event = ev_big          # compute_rendering_loss waits for it right before the losses
        else:
            poses, focal, gt_rgb, gt_emb = self.poses_d, self.focal_d, self.gt_rgb_d, self.gt_emb_d
            ren.target_ready_event = None
        out = ren(multi_scale_voxel_list=None, voxel_density=None, language=None, voxel_feat=self.vol,
                  voxel_poses=poses, focal=focal, gt_rgb=gt_rgb, gt_depth=None, gt_pose=poses, c=None,
                  lang_goal=None, gt_embed=gt_emb)
        out["loss"].backward()
        if post_allreduce is not None:
            post_allreduce(ren)
        if volume_allreduce is not None:
            volume_allreduce(self.vol.grad)
        return out


def release():
    """Between cases: the last step's buffers hang off autograd nodes that are only reclaimed by the cycle collector."""
    import gc
    import torch
    gc.collect()
    torch.cuda.empty_cache()
    if os.environ.get("NRF_BENCH_DEBUG"):
        sys.stderr.write(f"[bench] live after release: {torch.cuda.memory_allocated() / 2 ** 30:.2f} GiB\n")


def timed_region(fn, n_steps, dev, world, read_losses=False):
    """EXACTLY n_steps calls of fn() between a barrier + synchronize on both sides, CUDA events on the current stream.
    -> (max over ranks, [per-rank ms])."""
    import torch
    import torch.distributed as dist
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    prev, seen, out = None, 0.0, None
    for _ in range(n_steps):
        out = None
        out = fn()
        if read_losses:
            # the step's result is read on the host EVERY step: the loss dictionary copies its scalars to pinned
            # memory asynchronously (LossDict), so a training loop logs step i while step i+1 is queued; the last
            # step's values are read before the timed region closes
            if prev is not None:
                seen += prev["loss_rgb"] + prev["loss_embed"] + prev["loss_depth"]
            prev = out
    if read_losses and prev is not None:
        seen += prev["loss_rgb"] + prev["loss_embed"] + prev["loss_depth"]
        assert seen == seen, "loss is NaN"
    b.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = torch.tensor([a.elapsed_time(b)], device=dev)
    per_rank = [float(ms.item())]
    if world > 1:
        parts = [torch.empty_like(ms) for _ in range(world)]
        dist.all_gather(parts, ms)
        per_rank = [float(p.item()) for p in parts]
    return max(per_rank), per_rank


def mode_rel_err(mods, base: Case, args, precisions, n_rays=256):
    """Relative-L2 error of every tensor-core precision mode against this repo's fp32 parity mode (itself held to the
    reference's outputs at 1e-4 by the tests) on the SAME inputs: n_rays rays per scene of the workload, identical
    injected sampling noise, forward_nerf outputs of the fine pass."""
    import torch
    syn = mods["syn"]
    dev, wl = base.dev, base.wl
    R = base.SB * n_rays
    from_cpu = lambda d: {k: v.to(dev) for k, v in d.items()}
    noise = from_cpu(syn.make_noise(R, wl.n_coarse, wl.n_fine, seed=17))
    rays_all = mods["U"].gen_rays(base.poses_d, wl.W, wl.H, base.focal_d, base.ren.z_near, base.ren.z_far).reshape(base.SB, -1, 8)
    rays = rays_all[:, syn.pick_ray_indices(wl.W * wl.H, n_rays, seed=3).to(dev)].contiguous()
    outs = {}
    with torch.no_grad():
        for prec in ["fp32"] + list(precisions):
            c = Case(mods, wl, prec, dev, base.rank, base.world, args, SB=base.SB, n_rays=n_rays, share_from=base)
            c.ren.eval()
            c.ren.encode(None, None, None, base.vol.detach(), None, None, None)
            o = c.ren.forward_nerf(rays, noise=noise)
            outs[prec] = {k: o.fine[k].double() for k in ("rgb", "embed", "depth")}
            del c
    rel = lambda a, b: float((a - b).norm() / b.norm())
    return {p: {k: float(f"{rel(outs[p][k], outs['fp32'][k]):.3e}") for k in ("rgb", "embed", "depth")} for p in precisions}


def run_ours(args):
    import torch
    import torch.distributed as dist
    mods = {"NR": importlib.import_module(PKG + ".neural_rendering"), "U": importlib.import_module(PKG + ".utils"),
            "syn": importlib.import_module(PKG + ".synthetic"), "par": importlib.import_module(PKG + ".parallel"),
            "lib": importlib.import_module(PKG + "._lib")}
    syn, par, lib = mods["syn"], mods["par"], mods["lib"]

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world} (launch with torch.distributed.run)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    wl = syn.CONFIGS[args.workload]
    K, W = args.steps, args.warmup
    pk = peaks()

    main = Case(mods, wl, args.precision, dev, rank, world, args)
    ren = main.ren
    post = (lambda r: par.allreduce_mlp_grads(r)) if world > 1 and args.allreduce == "post" else None
    if world > 1 and args.allreduce == "overlap":     # ONE NCCL all-reduce of the flat MLP gradient, started inside the
        par.overlap_mlp_grad_allreduce(ren)           # backward: it runs under the volume-gradient scatter
    multi_gpu_check = multi_gpu_numerical_check(mods, main, world) if world > 1 else None

    for _ in range(W):
        main.step(post_allreduce=post)
    torch.cuda.synchronize()
    sampler = ClockSampler(local) if rank == 0 else None
    time.sleep(0.3)
    launches0 = lib.launch_count()
    ms_total, ms_rank = timed_region(lambda: main.step(post_allreduce=post), K, dev, world)   # behind `value`
    launches = lib.launch_count() - launches0
    # end-to-end: host (pinned) inputs copied in, loss read back, every step (right after `value`: the step slows by
    # ~5 % over the first seconds under the power cap - see `sustained` - and the two should see the same clocks)
    for _ in range(max(W, 3)):          # its own warm-up: the copy stream's allocator pool, the pinned staging ring
        main.step(host_inputs=True, post_allreduce=post)
    e2e_regions = [timed_region(lambda: main.step(host_inputs=True, post_allreduce=post), K, dev, world, read_losses=True)[0]
                   for _ in range(2)]
    ms_e2e = min(e2e_regions)      # two regions of K steps, the faster one: a single host hiccup (one 100 ms stall was
                                   # seen in one of four runs on one box) would otherwise decide the headline ratio
    # same K steps again with a CUDA-event pair around every kernel launch (recorded by the library on the
    # launching stream): per-kernel durations for the roofline; kept out of `value` because ~300 event
    # records per step add launch gaps
    lib.timing_begin()
    ms_total_ev, _ = timed_region(lambda: main.step(post_allreduce=post), K, dev, world)
    kern = lib.timing_end()
    # sample tiles (256 evaluations) that lie outside the grid altogether: the fused forward skips their latent
    # k-panels (exact zeros), 2 * 3 * C * 512 FLOP per evaluation that `achieved` below still counts as algorithmic work
    ren._touch_log = []
    main.step(post_allreduce=post)
    torch.cuda.synchronize()
    dead_evals = tiles_total = 0
    for t in ren._touch_log:
        n8 = (t.numel() + 7) // 8 * 8
        tt = torch.zeros(n8, dtype=torch.uint8, device=t.device)
        tt[:t.numel()] = t
        live = tt.view(-1, 8).sum(1) > 0              # (uint8 .any() stays uint8: its ~ is not a logical not)
        dead_evals += int((~live).sum()) * 256
        tiles_total += live.numel()
    ren._touch_log = None
    # a long timed region as well (20 steps are 0.3 s): >= 200 steps and >= 3 s, same step, same clock
    n_sus = 0 if args.sustain_steps <= 0 else max(args.sustain_steps, int(3000.0 / (ms_total / K)) + 1)
    sustained = None
    if n_sus:
        ms_sus, ms_sus_rank = timed_region(lambda: main.step(post_allreduce=post), n_sus, dev, world)
        sustained = {"steps": n_sus, "seconds": round(ms_sus / 1e3, 2), "ms_per_step": round(ms_sus / n_sus, 3),
                     "value": round(main.evals * world / (ms_sus / n_sus * 1e-3), 1), "unit": "ray-samples/s",
                     "ms_per_step_per_rank": [round(m / n_sus, 3) for m in ms_sus_rank]}
    clocks = sampler.stop() if sampler else None
    # optional mode, reported beside the headline and never mixed into it: the fine pass reuses the coarse pass's
    # field evaluations (bit-identical rendering; the MLP runs on Kc + Kf instead of Kc + (Kc + Kf) samples per ray)
    ms_reuse = None
    if not args.no_reuse_line:
        ren.reuse_coarse_evals = True
        for _ in range(2):
            main.step(post_allreduce=post)
        ms_reuse, _ = timed_region(lambda: main.step(post_allreduce=post), K, dev, world)
        ren.reuse_coarse_evals = False

    SB, n_rays = main.SB, main.n_rays
    evals_step = main.evals * world
    ms_step = ms_total / K
    value = evals_step / (ms_step * 1e-3)
    e2e_value = evals_step / (ms_e2e / K * 1e-3)
    # dominant kernel: mlp_fused_kernel (csrc/mlp_fused.cu), the whole forward MLP and the whole data-gradient chain
    # as one persistent tcgen05 kernel each (2 + 2 launches per step: coarse and fine pass).  Algorithmic FLOPs it
    # executes per step = (forward + dgrad) FLOP per evaluation x this rank's evaluations, minus the small dL/dz GEMM
    # that stays a separate launch; duration = sum over its launches (CUDA events on the launching stream).
    # (--precision fp32 / bf16x3 or NRF_MLP_LAYERED=1: the per-layer gemm kernels are the dominant ones instead.)
    wg_ms, wg_n = kern["wgrad_tc"]
    fused_ms = kern["fused_fwd"][0] + kern["fused_bwd"][0]
    fused_n = kern["fused_fwd"][1] + kern["fused_bwd"][1]
    flop_dz = 2 * 3 * 512 * wl.C                      # dL/dz = [g_0|g_1|g_2] . W_z: separate GEMM
    if fused_n > 0:
        dom_name = "mlp_fused_kernel (whole-MLP forward + whole data-gradient chain, 4 launches/step)"
        dom_ms, dom_n = fused_ms, fused_n
        dom_flops = (FLOP_FWD + FLOP_DGRAD - flop_dz) * main.evals * K
    else:
        dom_name = "gemm_tc_kernel / gemm_simt_kernel (per-layer forward + dgrad GEMMs)"
        dom_ms = kern["gemm_tc"][0] + kern["simt"][0]
        dom_n = kern["gemm_tc"][1] + kern["simt"][1]
        dom_flops = (FLOP_FWD + FLOP_DGRAD) * main.evals * K
    achieved = dom_flops / (dom_ms * 1e-3) / 1e12 if dom_ms > 0 else 0.0
    skipped_flops = dead_evals * flop_dz * K if fused_n > 0 else 0        # (2 * 3 * C * 512 per evaluation, as flop_dz)
    executed = (dom_flops - skipped_flops) / (dom_ms * 1e-3) / 1e12 if dom_ms > 0 else 0.0
    roof = {"bound": "tensor", "kernel": dom_name, "achieved": round(achieved, 1),
            "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s", "frac": round(achieved / pk["bf16_tflops_sustained"], 4),
            "executed": round(executed, 1), "frac_executed": round(executed / pk["bf16_tflops_sustained"], 4),
            "executed_note": f"`achieved` counts the algorithmic FLOPs of the reference's MLP; {dead_evals} of "
                             f"{main.evals} evaluations per step sit in 256-sample tiles without any sample inside the "
                             "grid, whose all-zero latent the fused forward does not multiply (bit-identical results): "
                             "`executed` leaves those products out",
            "traffic": FUSED_DRAM_BYTES_PER_LAUNCH if fused_n > 0 and args.workload == "config2" else None,
            "traffic_source": FUSED_DRAM_SOURCE + ": dram__bytes_read.sum + dram__bytes_write.sum of the 4 "
                              "mlp_fused_kernel launches of one config-2 step (ncu --set full), averaged per launch; "
                              "algorithmic bytes per launch = 5.13e9 (13.5 KB/eval forward, 12.7 KB/eval backward)",
            "peak_source": pk["source"] + " (sustained: kernel timed inside a long step)",
            "launches": dom_n, "avg_launch_ms": round(dom_ms / max(dom_n, 1), 4),
            "algorithmic_flop_per_eval": {"fwd": FLOP_FWD, "dgrad": FLOP_DGRAD, "wgrad": FLOP_WGRAD},
            "wgrad_tc_tflops": round(FLOP_WGRAD * main.evals * K / (wg_ms * 1e-3) / 1e12, 1) if wg_ms > 0 else None,
            "step_frac_of_tensor_peak": round(FLOP_STEP * main.evals / (ms_step * 1e-3) / 1e12 / pk["bf16_tflops_sustained"], 4)}
    kernel_ms = {k: round(v[0] / K, 3) for k, v in kern.items() if v[1] > 0}
    line = {"metric": "render fwd+bwd ray-samples/s", "value": round(value, 1), "unit": "ray-samples/s",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": round(ms_step, 3),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.precision, "data": "synthetic",
            "config": {"workload": f"{wl.name}: per GPU {SB} scenes x {n_rays} rays, {wl.n_coarse}+{wl.n_fine} samples, "
                                   f"{wl.S}^3 x {wl.C}ch volume, ResnetFC 512x5, RGB+{wl.D}d heads, fwd+bwd",
                       "evals_per_step": evals_step, "precision": args.precision, "scatter": args.scatter, "volume_layout": args.volume_layout,
                       "l2": "working set (1 GiB volume + ~20 GiB activations per step) >> 126 MB L2; no flush needed",
                       "parallelism": f"dp{world} over scenes; one NCCL all-reduce of the MLP grads ({args.allreduce})" if world > 1 else "single GPU"},
            "ms_per_step_per_rank": [round(m / K, 3) for m in ms_rank],
            "e2e": {"value": round(e2e_value, 1), "unit": "ray-samples/s", "ms_per_step": round(ms_e2e / K, 3),
                    "h2d_bytes_per_step": main.h2d_bytes, "d2h_bytes_per_step": 4 * 7,
                    "ms_per_step_regions": [round(m / K, 3) for m in e2e_regions],
                    "note": "the faster of two timed regions of K steps each; poses, focal, gt_rgb, gt_embed copied from pinned host memory each step (side stream; the targets "
                            "are awaited right before the losses); the 7 loss scalars of every step are copied to pinned host memory "
                            "and read by the host one step later (LossDict), the last step's inside the timed region; "
                            "the voxel volume is device-resident "
                            "as in the reference (PerAct encoder output)"},
            "sustained": sustained,
            "reuse_coarse_evals": None if ms_reuse is None else {
                "ms_per_step": round(ms_reuse / K, 3),
                "value": round(evals_step / (ms_reuse / K * 1e-3), 1), "unit": "ray-samples/s",
                "mlp_evals_per_step": world * SB * n_rays * (wl.n_coarse + wl.n_fine),
                "note": "opt-in NeuralRenderer.reuse_coarse_evals=True: same rendered samples per step (value counts "
                        "them as the headline does), the MLP evaluates each distinct sample once; NOT the headline"},
            "gpu_launches": launches, "clocks": clocks, "roofline": roof, "kernel_ms_per_step": kernel_ms,
            "ms_per_step_with_kernel_events": round(ms_total_ev / K, 3)}
    if multi_gpu_check is not None:
        line["multi_gpu_check"] = multi_gpu_check

    # ---- the precision modes on the same step: ms per step + error against the fp32 parity mode
    if not args.no_modes and wl.train:
        modes = {args.precision: {"ms_per_step": round(ms_step, 3)}}
        for prec in [p for p in ("bf16", "fp16", "bf16x3") if p != args.precision]:
            c = Case(mods, wl, prec, dev, rank, world, args, share_from=main)
            if world > 1 and args.allreduce == "overlap":
                par.overlap_mlp_grad_allreduce(c.ren)
            n = K if prec != "bf16x3" else max(3, K // 4)
            for _ in range(2):
                c.step(post_allreduce=post)
            ms_m, _ = timed_region(lambda: c.step(post_allreduce=post), n, dev, world)
            modes[prec] = {"ms_per_step": round(ms_m / n, 3), "steps": n}
            del c
            release()
        errs = mode_rel_err(mods, main, args, list(modes))
        for prec in modes:
            modes[prec]["rel_err_vs_fp32_mode"] = errs[prec]
        line["modes"] = modes
        line["modes_note"] = ("same config-2 step per mode; rel_err = relative L2 of the fine pass's rgb / embed / depth "
                              "against this repo's fp32 parity mode (held to the reference at 1e-4 by tests/) on 256 "
                              "rays per scene with identical injected sampling noise; fp16 = fp16 forward operands + bf16 "
                              "backward; bf16x3 = split-bf16 operands, three MMAs per product, layer by layer")

    # ---- in-box variant: near / far clipped to the volume and a longer lens, so that ~95 % of the samples gather
    # (in the BASELINE camera set-up only ~5 % of the samples fall inside the 1 m box): the HBM-side kernels under load
    if not args.no_extra and wl.train:
        ib = Case(mods, wl, args.precision, dev, rank, world, args, focal=500.0, near_far=(2.4, 3.2))
        if world > 1 and args.allreduce == "overlap":
            par.overlap_mlp_grad_allreduce(ib.ren)
        for _ in range(2):
            ib.step(post_allreduce=post)
        ms_ib, _ = timed_region(lambda: ib.step(post_allreduce=post), K, dev, world)
        lib.timing_begin()
        timed_region(lambda: ib.step(post_allreduce=post), K, dev, world)
        kib = lib.timing_end()
        gbs = lambda ms, bytes_per_eval: round(ib.evals * K * bytes_per_eval / (ms * 1e-3) / 1e9, 1) if ms > 0 else None
        enc, sca = gbs(kib["encode"][0], 4096), gbs(kib["scatter"][0], 4096)
        line["in_box"] = {"ms_per_step": round(ms_ib / K, 3), "value": round(ib.evals * world / (ms_ib / K * 1e-3), 1),
                          "camera": "focal 500, z_near 2.4, z_far 3.2: ~95 % of the samples inside the volume",
                          "kernel_ms_per_step": {k: round(v[0] / K, 3) for k, v in kib.items() if v[1] > 0},
                          "encode_gbs": enc, "encode_frac_of_hbm": round(enc / pk["hbm_gbs"], 3) if enc else None,
                          "scatter_gbs": sca, "scatter_frac_of_hbm": round(sca / pk["hbm_gbs"], 3) if sca else None,
                          "note": "algorithmic bytes: gather 8 corners x 128 ch x 4 B = 4096 B per evaluation (SURVEY 8d); "
                                  "scatter 4096 B per evaluation (the dense 0.97 GB gradient volume it also writes is not counted). "
                                  "A fraction above 1 is not a DRAM rate: neighbouring samples share corner rows, which then "
                                  "come from L2 (ncu, profiles/r02c_encode_tma_ncu.md: 0.66 GB of DRAM reads for 2.0 GB of "
                                  "corner rows in the fine pass; the kernel moves 9.2 TB/s from L2 to the SMs)"}
        del ib
        release()

    # ---- the other BASELINE configs (each a short run; N > 1: strong scaling - total work fixed)
    if not args.no_extra and not args.no_configs and args.workload == "config2":
        del main, ren
        release()
        line["configs"] = other_configs(mods, args, dev, rank, world)

    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(rays=args.cpu_rays, reps=1)
            if wl.train:                     # second half of the baseline leg: the same oracle port on the same GPU
                line["cpu_baseline"]["eager_pytorch_same_gpu"] = eager_gpu_baseline(dev, wl)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def multi_gpu_numerical_check(mods, case, world):
    """N > 1, before anything is timed: (a) the all-reduce started inside the backward gives bit-identical MLP gradients
    to the separate all-reduce after it; (b) the reduced gradient is the sum of the ranks' local gradients (gathered
    and added in rank order).  Asserted; the numbers go into the bench line."""
    import torch
    import torch.distributed as dist
    par = mods["par"]
    ren = case.ren
    det0, hook0 = ren.deterministic, ren._grad_allreduce
    ren.deterministic = True                         # ordered split reductions: run-to-run identical local gradients
    flat = lambda: torch.cat([p.grad.reshape(-1) for p in case.params])

    def run(mode):
        par.overlap_mlp_grad_allreduce(ren, enabled=(mode == "overlap"))
        torch.manual_seed(4321 + case.rank)
        case.step(post_allreduce=(lambda r: par.allreduce_mlp_grads(r)) if mode == "post" else None)
        return flat().clone()
    g_local, g_post, g_over = run("local"), run("post"), run("overlap")
    parts = [torch.empty_like(g_local) for _ in range(world)]
    dist.all_gather(parts, g_local)
    total = parts[0].double()
    for p in parts[1:]:
        total += p.double()
    rel = float((g_over.double() - total).norm() / total.norm())
    same = bool(torch.equal(g_post, g_over))
    ren.deterministic, ren._grad_allreduce = det0, hook0
    assert same, "overlapped MLP-gradient all-reduce differs from the separate one"
    assert rel < 1e-5, f"all-reduced MLP gradient is not the sum over ranks (rel {rel:.2e})"
    return {"overlap_equals_post_bitwise": same, "allreduce_vs_sum_of_local_grads_rel": float(f"{rel:.3e}"),
            "ranks": world, "mlp_grad_bytes": int(g_local.numel() * 4)}


def other_configs(mods, args, dev, rank, world):
    """BASELINE configs 3, 4, 5.  N > 1 is STRONG scaling here: config 3 shards the rays of the 5
    images, config 4 its 8 scenes, config 5 the rays of its one scene (MLP-gradient all-reduce + a dense all-reduce of
    the 4.1 GB volume gradient).  Short runs: 2 warm-up + a few timed iterations each."""
    import torch
    syn, par = mods["syn"], mods["par"]
    out = {}
    # config 3: inference, 5 cameras x 128 x 128 px, rays sharded contiguously, image rows all-gathered
    wl = syn.CONFIGS["config3"]
    c = Case(mods, wl, args.precision, dev, rank, world, args, SB=1, n_rays=4096, vol_seed=0)
    c.ren.eval()
    poses = syn.arc_poses(wl.n_cams).to(dev)
    vol = c.vol.detach()
    render = lambda: par.render_sharded(c.ren, vol, c.focal_d, poses, gather=world > 1)
    for _ in range(2):
        render()
    n = 5
    ms, per_rank = timed_region(render, n, dev, world)
    evals3 = wl.n_cams * wl.H * wl.W * (wl.n_coarse + wl.n_coarse + wl.n_fine)
    out["config3"] = {"workload": "inference: 5 cameras x 128x128 px, 64 + 128 samples/ray, 100^3 x 128ch volume; rays "
                                  f"sharded over {world} GPU(s), images all-gathered", "scaling": "strong",
                      "ms_per_render": round(ms / n, 3), "value": round(evals3 / (ms / n * 1e-3), 1), "unit": "ray-samples/s",
                      "evals_per_render": evals3, "ms_per_rank": [round(m / n, 3) for m in per_rank],
                      "frac_of_tensor_peak_fwd": round(FLOP_FWD * evals3 / world / (ms / n * 1e-3) / 1e12 / peaks()["bf16_tflops_sustained"], 4)}
    del c, vol, render
    release()
    # config 4: training, 8 scenes x 4096 rays, scenes data-parallel
    wl = syn.CONFIGS["config4"]
    if wl.SB % world == 0:
        c = Case(mods, wl, args.precision, dev, rank, world, args, SB=wl.SB // world)
        if world > 1:
            par.overlap_mlp_grad_allreduce(c.ren)
        for _ in range(2):
            c.step()
        n = 5
        ms, per_rank = timed_region(c.step, n, dev, world)
        out["config4"] = {"workload": f"training: 8 scenes x 4096 rays, 64+64 samples, {wl.SB // world} scene(s) per GPU, "
                                      "one NCCL all-reduce of the MLP grads", "scaling": "strong",
                          "ms_per_step": round(ms / n, 3), "value": round(c.evals * world / (ms / n * 1e-3), 1),
                          "unit": "ray-samples/s", "evals_per_step": c.evals * world,
                          "ms_per_rank": [round(m / n, 3) for m in per_rank]}
        del c
        release()
    # config 5: stress shape, one scene; N > 1: its rays split over the ranks, the volume gradient summed
    wl = syn.CONFIGS["config5"]
    if wl.rays_per_scene % world == 0:
        c = Case(mods, wl, args.precision, dev, rank, world, args, n_rays=wl.rays_per_scene // world, vol_seed=0)
        vred = None
        if world > 1:
            par.overlap_mlp_grad_allreduce(c.ren)
            c.ren.keep_voxel_counts = True                # the scatter's own per-voxel counts name the touched rows
            vred = lambda g: par.sparse_allreduce_volume_grad(g, counts=c.ren.last_voxel_counts)   # (dense form: below)
        for _ in range(2):
            c.step(volume_allreduce=vred)
        n = 5
        ms, per_rank = timed_region(lambda: c.step(volume_allreduce=vred), n, dev, world)
        rec = {"workload": f"training: 200^3 x 128ch volume (4.1 GB), 16384 rays x (128 + 256) samples, "
                           f"{wl.rays_per_scene // world} rays per GPU" +
                           ("; MLP-grad all-reduce + dense all-reduce of the 4.1 GB volume gradient" if world > 1 else ""),
               "scaling": "strong", "ms_per_step": round(ms / n, 3), "value": round(c.evals * world / (ms / n * 1e-3), 1),
               "unit": "ray-samples/s", "evals_per_step": c.evals * world, "ms_per_rank": [round(m / n, 3) for m in per_rank]}
        if world > 1:                               # the exchange alone, both forms, for the efficiency statement
            c.step(volume_allreduce=None)           # this rank's own gradient and the touched set of its scatter
            g_local, cnts = c.vol.grad.detach().clone(), c.ren.last_voxel_counts
            g = g_local.clone()
            sparse = lambda: par.sparse_allreduce_volume_grad(g.copy_(g_local), counts=cnts)
            stats = sparse()
            ms_s, _ = timed_region(sparse, 3, dev, world)
            ms_cp, _ = timed_region(lambda: g.copy_(g_local), 3, dev, world)
            ms_c, _ = timed_region(lambda: par.allreduce_volume_grad(g), 3, dev, world)
            ms_dense_step, _ = timed_region(lambda: c.step(volume_allreduce=par.allreduce_volume_grad), 3, dev, world)
            rec["volume_grad_exchange"] = {
                "sparse_bytes_per_rank_received": stats["bytes"], "sparse_rows_per_rank": stats["rows"],
                "sparse_exchange_ms": round((ms_s - ms_cp) / 3, 3),
                "dense_allreduce_ms": round(ms_c / 3, 3), "dense_bytes": int(g.numel() * 4),
                "ms_per_step_with_dense_allreduce": round(ms_dense_step / 3, 3),
                "note": "sparse: nrf_rows_gather -> all_gather of (voxel index, 128-vector) rows -> nrf_rows_merge (rank-order "
                        "sums, every touched 32-voxel tile written once); timed alone on this step's per-rank gradients"}
            del g_local
            rec["workload"] = rec["workload"].replace("dense all-reduce of the 4.1 GB volume gradient",
                                                      "sparse exchange of the touched voxel rows of the volume gradient")
            del g
        out["config5"] = rec
        del c
        release()
    return out


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
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp16", "bf16x3", "fp32"])
    ap.add_argument("--cpu-rays", type=int, default=128, dest="cpu_rays")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-reuse-line", action="store_true", dest="no_reuse_line")
    ap.add_argument("--no-modes", action="store_true", dest="no_modes", help="skip the per-precision-mode lines")
    ap.add_argument("--no-extra", action="store_true", dest="no_extra",
                    help="skip the in-box variant and BASELINE configs 3 / 4 / 5")
    ap.add_argument("--no-configs", action="store_true", dest="no_configs",
                    help="skip BASELINE configs 3 / 4 / 5 (the in-box variant still runs)")
    ap.add_argument("--sustain-steps", type=int, default=200, dest="sustain_steps",
                    help="length of the additional long timed region (0: skip)")
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
