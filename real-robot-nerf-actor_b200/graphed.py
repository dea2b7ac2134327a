"""CUDA-graph capture of the renderer's training step for fixed shapes.

At the reference's own training shape (nerfact.conf: ONE scene, 512-ray chunks, 64 latent channels: 98 304 field
evaluations) a step is ~55 kernel launches of a few tens of microseconds each, and the host (ctypes calls, autograd
bookkeeping, allocator) cannot enqueue them as fast as the GPU retires them.  `GraphedRenderLoss` captures the forward
(gen_rays -> ray subsample -> forward_nerf -> losses) and the backward (into the voxel volume and every MLP parameter)
into two CUDA graphs - the scheme of torch.cuda.make_graphed_callables - and exposes them as ONE autograd node, so the
renderer still sits in the middle of the caller's graph (PerAct encoder before it, the optimizer after it):

    step = GraphedRenderLoss(renderer, voxel_feat, gt_pose, focal, gt_rgb, gt_embed)     # example tensors: shapes only
    out = step(voxel_feat, gt_pose, focal, gt_rgb, gt_embed)     # LossDict, as renderer.forward(...) returns
    (bc_loss + 10.0 * out["loss"]).backward()

Ray indices and sampling noise are drawn inside the graphs by torch's graph-safe generator (fresh every replay).
Everything the kernels read is copied into static buffers first (the volume: one device-to-device copy, 0.08 ms for
100^3 x 64 channels); gradients come back as views of static buffers that the next replay overwrites - consume them
(optimizer step / accumulation by autograd, which copies) before the next call.  Shapes, dtypes and the renderer's
configuration are frozen at capture; `gt_depth`, a feature extractor and the NCCL hook are not captured (use the
eager path for those).
"""
from __future__ import annotations

import torch

from . import ops
from .neural_rendering import LossDict


class GraphedRenderLoss:
    def __init__(self, renderer, voxel_feat, gt_pose, focal, gt_rgb, gt_embed, c=None, warmup: int = 3):
        if renderer._grad_allreduce is not None:
            raise RuntimeError("GraphedRenderLoss: the overlapped NCCL all-reduce hook is not captured; all-reduce the "
                               "MLP gradients after the backward (parallel.allreduce_mlp_grads)")
        self.ren = ren = renderer
        dev = voxel_feat.device
        self.params = [p for p in ren.parameters() if p.requires_grad]
        self.s_vol = voxel_feat.detach().clone().requires_grad_(True)
        self.s_in = [t.detach().clone() for t in (gt_pose, focal, gt_rgb, gt_embed)]
        self.c = c
        self.s_gloss = torch.ones((), device=dev)
        trace0 = ren.trace_ranges
        ren.trace_ranges = False                           # profiler ranges are host-side objects: not capturable

        def fwd():
            pose, foc, rgb, emb = self.s_in
            return ren._loss_tensors(None, None, None, self.s_vol, pose, foc, rgb, None, pose, c=self.c, lang_goal=None,
                                     gt_embed=emb)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):                      # warm-up off the capture stream: caches, workspaces, cuBLAS-free
            for _ in range(warmup):
                loss, _ = fwd()
                torch.autograd.grad(loss, [self.s_vol] + self.params, self.s_gloss, allow_unused=True)
        torch.cuda.current_stream(dev).wait_stream(side)
        self.pool = torch.cuda.graph_pool_handle()
        self.g_fwd, self.g_bwd = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        ops._INTR_CACHE.clear()                            # the [fx, fy, cx, cy] vector must be rebuilt INSIDE the graph
        with torch.cuda.graph(self.g_fwd, pool=self.pool):
            self.s_loss, self.s_scalars = fwd()
        with torch.cuda.graph(self.g_bwd, pool=self.pool):
            grads = torch.autograd.grad(self.s_loss, [self.s_vol] + self.params, self.s_gloss, allow_unused=True)
        ops._INTR_CACHE.clear()                            # ... and must not leak into later eager calls
        self.s_vol_grad = grads[0]
        # the MLP gradients are views of ONE flat buffer (neural_rendering._zero_grads): a replay's results are handed
        # to autograd as views of one clone of it (autograd may keep what it is given as `.grad`)
        pg = list(grads[1:])
        bases = {id(x._base): x._base for x in pg if x is not None and x._base is not None}
        self.s_flat = list(bases.values())
        self.s_pviews = [None if x is None else (list(bases).index(id(x._base)), x.storage_offset(), tuple(x.shape))
                         if x._base is not None else x for x in pg]
        ren.trace_ranges = trace0
        self._fn = _make_function(self)

    def __call__(self, voxel_feat, gt_pose, focal, gt_rgb, gt_embed):
        loss, scalars = self._fn(voxel_feat, gt_pose, focal, gt_rgb, gt_embed, *self.params)
        return LossDict(loss, scalars)


def _make_function(g: GraphedRenderLoss):
    class _Graphed(torch.autograd.Function):
        @staticmethod
        def forward(ctx, voxel_feat, gt_pose, focal, gt_rgb, gt_embed, *params):
            with torch.no_grad():
                g.s_vol.copy_(voxel_feat)
                for dst, src in zip(g.s_in, (gt_pose, focal, gt_rgb, gt_embed)):
                    if dst.data_ptr() != src.data_ptr():
                        dst.copy_(src)
            g.g_fwd.replay()
            ctx.mark_non_differentiable(g.s_scalars)
            return g.s_loss.detach().clone(), g.s_scalars.detach().clone()

        @staticmethod
        def backward(ctx, d_loss, _d_scalars):
            g.s_gloss.copy_(d_loss)
            g.g_bwd.replay()
            flats = [f.clone() for f in g.s_flat]
            pgrads = []
            for v in g.s_pviews:
                if v is None:
                    pgrads.append(None)
                elif isinstance(v, tuple):
                    which, off, shape = v
                    n = 1
                    for d in shape:
                        n *= d
                    pgrads.append(flats[which][off:off + n].view(shape))
                else:
                    pgrads.append(v.clone())
            return (g.s_vol_grad, None, None, None, None, *pgrads)
    return _Graphed.apply
