"""Multi-GPU plumbing: one process per GPU, rays / scenes sharded, no collective inside the render.

The reference's hot path is single-GPU; its ancestor shards the ray axis with nn.DataParallel
(featurenerf_robo/featurenerf/src/render/nerf_embed.py:412-429).  Here:
  * inference (`render_sharded`): each rank renders a contiguous slice of the flattened ray list;
    the only exchange is the optional all_gather of the finished image rows;
  * training (`overlap_mlp_grad_allreduce`, or `allreduce_mlp_grads` after the backward): scenes (or ray
    slices) are data-parallel; the field MLP is replicated, so its gradients are summed with ONE all_reduce
    over a flat fp32 buffer (12.2 MB at the BASELINE dims), started inside the backward so that it runs
    under the volume-gradient scatter.  The volume gradient stays local when each rank owns whole scenes;
    one scene split over ranks (config 5) sums it with `sparse_allreduce_volume_grad` (touched voxel rows only) or the
    dense `allreduce_volume_grad`.
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
def sparse_allreduce_volume_grad(grad: torch.Tensor, group=None, counts: torch.Tensor = None) -> dict:
    """Sum of a volume gradient across ranks by exchanging only the voxels a rank's rays touched (SURVEY 8e, config 5:
    ONE scene whose rays are split over the GPUs).  grad: (SB,C,S0,S1,S2), contiguous or channels_last_3d, summed IN
    PLACE; every rank ends with bit-identical values (the contributions are added in rank order on every rank).

    A rank's rays touch a small part of the grid (the 200^3 x 128-channel gradient is 4.1 GB, the rows written by
    16 384 / N rays a few hundred MB), so instead of a dense all-reduce each rank all-gathers (voxel index, C-vector)
    rows: bytes on the wire = sum of the touched rows x (4 C + 8) instead of 2 (N-1)/N x the whole volume.
    Costs one host synchronisation (the row counts size the exchange buffers).  Returns the exchange statistics.
    counts: optional (SB*V,) per-voxel entry counts of the scatter that produced `grad`
    (`renderer.keep_voxel_counts = True` -> `renderer.last_voxel_counts`): the touched set is then `counts > 0`
    (32 MB at 200^3) instead of a scan of the dense gradient (4.1 GB read + a 1 GB boolean temporary)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return {"rows": 0, "bytes": 0}
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    SB, C = grad.shape[:2]
    V = grad[0, 0].numel()
    # per scene a 2-D view with the voxel axis `vdim`: (C, V) for a contiguous gradient, (V, C) for channels_last_3d; every
    # gather / scatter below then moves along the unit-stride axis of the neighbouring (ascending) voxel indices
    if grad.is_contiguous():
        views, vdim = [grad[b].reshape(C, V) for b in range(SB)], 1
    elif grad.is_contiguous(memory_format=torch.channels_last_3d):
        views, vdim = [grad[b].permute(1, 2, 3, 0).reshape(V, C) for b in range(SB)], 0
    else:
        raise ValueError("sparse_allreduce_volume_grad: gradient must be contiguous or channels_last_3d")
    assert all(v.data_ptr() == grad[b].data_ptr() for b, v in enumerate(views)), "expected views, got copies"
    if counts is not None:
        assert counts.numel() == SB * V, "counts must have one entry per voxel"
        touched = counts.reshape(-1) > 0
    else:
        touched = torch.stack([(v != 0).any(1 - vdim) for v in views]).reshape(-1)   # (SB * V)
    idx = touched.nonzero().squeeze(1)                                           # flat (scene * V + voxel), ascending
    if grad.is_cuda and C % 4 == 0:
        return _sparse_allreduce_rows_cuda(grad, idx, group, exact_touched=True)
    counts = torch.zeros(world, device=grad.device, dtype=torch.int64)
    counts[rank] = idx.numel()
    dist.all_reduce(counts, group=group)
    counts = counts.tolist()                                                     # the one host sync
    cap = max(max(counts), 1)
    shape = (C, cap) if vdim == 1 else (cap, C)
    my_rows = torch.zeros(shape, device=grad.device, dtype=grad.dtype)
    my_idx = torch.zeros(cap, device=grad.device, dtype=torch.int64)
    n = idx.numel()
    my_idx[:n] = idx
    bounds = torch.searchsorted(idx, torch.arange(SB + 1, device=idx.device) * V).tolist() if SB > 1 else [0, n]

    def rows_of(buf, lo, hi):
        return buf[:, lo:hi] if vdim == 1 else buf[lo:hi]
    for b in range(SB):
        lo, hi = bounds[b], bounds[b + 1]
        if hi > lo:
            rows_of(my_rows, lo, hi).copy_(views[b].index_select(vdim, idx[lo:hi] - b * V))
    all_rows = [torch.empty_like(my_rows) for _ in range(world)]
    all_idx = [torch.empty_like(my_idx) for _ in range(world)]
    dist.all_gather(all_rows, my_rows, group=group)
    dist.all_gather(all_idx, my_idx, group=group)
    # own touched rows are cleared, then every rank's rows (own included) are added in rank order: identical bits on
    # every rank; within one rank's list the voxel indices are unique, so index_add_ has no collisions
    for b in range(SB):
        lo, hi = bounds[b], bounds[b + 1]
        if hi > lo:
            views[b].index_fill_(vdim, idx[lo:hi] - b * V, 0.0)
    for r in range(world):
        k = counts[r]
        if k == 0:
            continue
        i_r = all_idx[r][:k]
        bnd = torch.searchsorted(i_r, torch.arange(SB + 1, device=i_r.device) * V).tolist() if SB > 1 else [0, k]
        for b in range(SB):
            lo, hi = bnd[b], bnd[b + 1]
            if hi > lo:
                views[b].index_add_(vdim, i_r[lo:hi] - b * V, rows_of(all_rows[r], lo, hi))
    return {"rows": counts, "bytes": int(cap * world * (C * grad.element_size() + 8))}


def _sparse_allreduce_rows_cuda(grad, idx, group, exact_touched=False):
    """The device path of sparse_allreduce_volume_grad: nrf_rows_gather -> all_gather of (rows, indices) ->
    nrf_rows_merge (every voxel some rank lists = the sum of its rows in rank order, own rows included: the same bits
    on every rank; the volume is written once per touched tile and never read).  One host synchronisation (the row
    counts size the buffers)."""
    from . import ops
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    C = grad.shape[1]
    n = idx.numel()
    counts = torch.zeros(world, device=grad.device, dtype=torch.int64)
    counts[rank] = n
    dist.all_reduce(counts, group=group)
    counts = counts.tolist()                                                     # the one host sync
    cap = max(max(counts), 1)
    my_rows = torch.empty((cap, C), device=grad.device, dtype=torch.float32)
    my_idx = torch.empty(cap, device=grad.device, dtype=torch.int64)
    my_idx[:n] = idx
    ops.rows_gather(grad, idx, out=my_rows)
    all_rows = torch.empty((world, cap, C), device=grad.device, dtype=torch.float32)
    all_idx = torch.empty((world, cap), device=grad.device, dtype=torch.int64)
    dist.all_gather_into_tensor(all_rows, my_rows, group=group)
    dist.all_gather_into_tensor(all_idx, my_idx, group=group)
    # every listed voxel = its rows summed in rank order, written once; with the touched set from the scatter's own
    # counts every other voxel of this rank's gradient is a zero the scatter wrote: whole tiles, full sectors
    ops.rows_merge(grad, all_rows, all_idx, counts, unlisted_are_zero=exact_touched)
    return {"rows": counts, "bytes": int(cap * world * (C * 4 + 8))}


@torch.no_grad()
def sparse_allreduce_phases(grad, counts, group=None) -> dict:
    """Diagnostics (scripts/exchange_probe.py): the device path of sparse_allreduce_volume_grad phase by phase, each
    bracketed by a device synchronisation (so the sum exceeds the pipelined call).  Milliseconds."""
    import time
    from . import ops
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    C = grad.shape[1]
    out = {}

    def phase(name, fn):
        torch.cuda.synchronize()
        t = time.perf_counter()
        r = fn()
        torch.cuda.synchronize()
        out[name] = round((time.perf_counter() - t) * 1e3, 3)
        return r
    idx = phase("nonzero", lambda: (counts.reshape(-1) > 0).nonzero().squeeze(1))
    n = idx.numel()

    def sizes():
        c = torch.zeros(world, device=grad.device, dtype=torch.int64)
        c[rank] = n
        dist.all_reduce(c, group=group)
        return c.tolist()
    cnt = phase("counts_allreduce", sizes)
    cap = max(max(cnt), 1)
    my_rows = torch.empty((cap, C), device=grad.device, dtype=torch.float32)
    my_idx = torch.empty(cap, device=grad.device, dtype=torch.int64)
    my_idx[:n] = idx
    phase("rows_gather", lambda: ops.rows_gather(grad, idx, out=my_rows))
    all_rows = torch.empty((world, cap, C), device=grad.device, dtype=torch.float32)
    all_idx = torch.empty((world, cap), device=grad.device, dtype=torch.int64)
    phase("all_gather_rows", lambda: dist.all_gather_into_tensor(all_rows, my_rows, group=group))
    phase("all_gather_idx", lambda: dist.all_gather_into_tensor(all_idx, my_idx, group=group))
    phase("rows_merge", lambda: ops.rows_merge(grad, all_rows, all_idx, cnt, unlisted_are_zero=True))
    out["rows"] = cnt
    return out


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
