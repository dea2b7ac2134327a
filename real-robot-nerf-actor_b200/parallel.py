"""Multi-GPU plumbing: one process per GPU, rays / scenes sharded, no collective inside the render.

The reference's hot path is single-GPU; its ancestor shards the ray axis with nn.DataParallel
(featurenerf_robo/featurenerf/src/render/nerf_embed.py:412-429).  Here:
  * inference (`render_sharded`): each rank renders a contiguous slice of the flattened ray list;
    the only exchange is the optional all_gather of the finished image rows;
  * training (`overlap_mlp_grad_allreduce`, or `allreduce_mlp_grads` after the backward): scenes (or ray
    slices) are data-parallel; the field MLP is replicated, so its gradients are summed with ONE all_reduce
    over a flat fp32 buffer (12.2 MB at the BASELINE dims), started inside the backward so that it runs
    under the volume-gradient scatter.  The volume gradient stays local when each rank owns whole scenes;
    `allreduce_volume_grad` covers the case of one scene split over ranks.
"""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced [lo, hi) slice of n items for `rank` (first n % world ranks get one more)."""
    base, extra = divmod(n, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_rays(rays: torch.Tensor, world_size: int, rank: int) -> torch.Tensor:
    """rays (N,8) -> this rank's contiguous slice."""
    lo, hi = shard_bounds(rays.shape[0], world_size, rank)
    return rays[lo:hi]


def _flatten(tensors: List[torch.Tensor]) -> torch.Tensor:
    return torch.cat([t.reshape(-1) for t in tensors]) if tensors else torch.empty(0)


def allreduce_mlp_grads(module: torch.nn.Module, group=None, average: bool = False) -> int:
    """Sums (or averages) every parameter gradient of `module` across ranks with one all_reduce.

    Aliased parameters (mlp_fine is mlp_coarse when share_mlp) are visited once.  Returns the
    number of bytes reduced.  Parameters without a gradient contribute zeros so that every rank
    issues an identically-shaped collective.
    """
    params = [p for p in module.parameters() if p.requires_grad]
    if not params or not dist.is_available() or not dist.is_initialized():
        return 0
    grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in params]
    flat = _flatten(grads)
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if average:
        flat /= dist.get_world_size(group)
    off = 0
    for p, g in zip(params, grads):
        n = g.numel()
        if p.grad is None:
            p.grad = flat[off:off + n].view_as(p).clone()
        else:
            p.grad.copy_(flat[off:off + n].view_as(p))
        off += n
    return flat.numel() * flat.element_size()


def overlap_mlp_grad_allreduce(renderer, group=None, average: bool = False, enabled: bool = True) -> None:
    """Data-parallel training without a separate reduction step: the renderer's backward all-reduces the flat MLP
    gradient buffer (ONE collective, 12.2 MB at the BASELINE dims) as soon as the last weight-gradient kernel of the
    step is enqueued, so the collective runs under the volume-gradient scatter instead of after the backward
    (SURVEY 8e).  The `.grad`s autograd hands out are already summed (or averaged) over the ranks: do NOT call
    `allreduce_mlp_grads` as well.  Only the gradients produced by `forward_nerf` / `forward` are covered."""
    if not enabled:
        renderer._grad_allreduce = None
        return

    def start(flats):
        if not (dist.is_available() and dist.is_initialized()):
            return None
        works = [dist.all_reduce(f, op=dist.ReduceOp.SUM, group=group, async_op=True) for f in flats]

        def finish():
            for w in works:
                w.wait()                       # the current stream waits; the host does not
            if average:
                for f in flats:
                    f /= dist.get_world_size(group)
        return finish

    renderer._grad_allreduce = start


def allreduce_volume_grad(grad: torch.Tensor, group=None) -> torch.Tensor:
    """Dense sum of a volume gradient across ranks (one scene's rays split over GPUs)."""
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(grad, op=dist.ReduceOp.SUM, group=group)
    return grad


@torch.no_grad()
def render_sharded(renderer, voxel_feat, focal, tgt_pose, c=None, group=None, gather: bool = True):
    """`NeuralRenderer.rendering` with the flattened ray list split contiguously over the ranks.

    Every rank holds the volume and the MLP; rank r renders rays [lo_r, hi_r).  With `gather` the
    per-rank pieces are all_gather'ed and every rank returns full (B,H,W,.) images; otherwise each
    rank returns its own slice (n_r, .) plus its bounds.
    """
    from .utils import gen_rays
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    rays = gen_rays(tgt_pose, renderer.W, renderer.H, focal, renderer.z_near, renderer.z_far, c=c)
    B, H, W, _ = rays.shape
    flat = rays.reshape(B * H * W, 8)
    lo, hi = shard_bounds(flat.shape[0], world, rank)
    renderer.encode(None, None, None, voxel_feat, None, focal, c)
    rgbs, embs, deps = [], [], []
    from . import ops
    from .neural_rendering import _is_channels_last_3d
    renderer._vol_cl_held = None if _is_channels_last_3d(voxel_feat) else \
        (voxel_feat, ops.volume_to_channels_last(voxel_feat))                       # one re-layout for all chunks
    try:
        for i in range(lo, hi, renderer.render_chunk_rays):
            out = renderer.forward_nerf(flat[i:min(i + renderer.render_chunk_rays, hi)].unsqueeze(0)).fine
            rgbs.append(out.rgb.squeeze(0)); embs.append(out.embed.squeeze(0)); deps.append(out.depth.squeeze(0))
    finally:
        renderer._vol_cl_held = None
    rgb, emb, dep = torch.cat(rgbs), torch.cat(embs), torch.cat(deps)
    if not gather or world == 1:
        if world == 1:
            return rgb.reshape(B, H, W, 3), emb.reshape(B, H, W, -1), dep.reshape(B, H, W)
        return (rgb, emb, dep), (lo, hi)
    sizes = [shard_bounds(flat.shape[0], world, r) for r in range(world)]
    nmax = max(h - l for l, h in sizes)

    def gather_rows(x):
        pad = torch.zeros(nmax, *x.shape[1:], device=x.device, dtype=x.dtype)
        pad[:x.shape[0]] = x
        parts = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(parts, pad, group=group)
        return torch.cat([p[:h - l] for p, (l, h) in zip(parts, sizes)])
    return (gather_rows(rgb).reshape(B, H, W, 3), gather_rows(emb).reshape(B, H, W, -1),
            gather_rows(dep).reshape(B, H, W))
