"""Drop-in for the reference's voxelizer, `voxel_grid_real.VoxelGrid` (voxel_grid_real.py:15-233): the step that
turns the fused multi-camera point cloud into the (B, S, S, S, 3+F+3+1) grid the PerAct encoder consumes -- the
producer side of the volume this package renders from.

Same constructor and call signature; `coords_to_bounding_voxel_grid` forwards to `nrf_voxelize` (csrc/voxelize.cu):
counting sort of the points by voxel + one thread per voxel, no float atomics -> bit-reproducible, and bit-identical
to the reference run on the CPU (points of a voxel are added in ascending index, as torch's CPU scatter_add_ does).
There is no CPU path: CPU tensors raise.
"""
from __future__ import annotations


import torch
from torch import nn

from . import _lib
from ._lib import check, ptr, stream_ptr

MIN_DENOMINATOR = 1e-12      # voxel_grid_real.py:11


class VoxelGrid(nn.Module):
    """voxel_grid_real.py:15-99.  Buffers the reference precomputes for its index arithmetic (flat output, tiled batch
    indices, index grid: ~0.5 GB at batch 8, S=100) are not needed and not allocated."""

    def __init__(self, coord_bounds, voxel_size: int, device, batch_size, feature_size, max_num_coords: int):
        super().__init__()
        self._device = device
        self._voxel_size = int(voxel_size)
        self._voxel_shape = [self._voxel_size] * 3
        self._voxel_d = float(self._voxel_size)
        self._voxel_feature_size = 4 + feature_size
        self._batch_size = batch_size
        self._num_coords = max_num_coords
        self.register_buffer("_coord_bounds", torch.tensor(coord_bounds, dtype=torch.float).reshape(1, 6))
        self.register_buffer("_dims_orig", torch.tensor([self._voxel_shape], dtype=torch.int32))

    def _geometry(self, coord_bounds, B, device):
        """(B, 6) = [bb_min - res | res + 1e-12] with the reference's own fp32 tensor ops (:176-186)."""
        bounds = self._coord_bounds if coord_bounds is None else coord_bounds
        bounds = torch.as_tensor(bounds, dtype=torch.float32, device=device).reshape(-1, 6)
        bb_mins, bb_maxs = bounds[..., 0:3], bounds[..., 3:6]
        bb_ranges = bb_maxs - bb_mins
        res = bb_ranges / (self._dims_orig.to(device).float() + MIN_DENOMINATOR)
        denom = res + MIN_DENOMINATOR
        geom = torch.cat([bb_mins - res, denom], -1)
        return geom.expand(B, 6).contiguous()

    def coords_to_bounding_voxel_grid(self, coords, coord_features=None, coord_bounds=None, only_features=False):
        """coords (B, N, 3) world points, coord_features (B, N, F) -> (B, S, S, S, 3 + F + 3 + 1):
        [mean xyz, mean features, voxel index / S, occupancy] (voxel_grid_real.py:175-233)."""
        if not (isinstance(coords, torch.Tensor) and coords.is_cuda):
            raise _lib.NrfError("VoxelGrid: expected CUDA tensors (there is no CPU fallback)")
        coords = coords.to(torch.float32).contiguous()
        B, N, _ = coords.shape
        F = 0
        if coord_features is not None:
            coord_features = coord_features.to(torch.float32).contiguous()
            F = coord_features.shape[-1]
        S = self._voxel_size
        geom = self._geometry(coord_bounds, B, coords.device)
        out = torch.empty(B, S, S, S, 3 + F + 4, device=coords.device, dtype=torch.float32)
        lib = _lib.load()
        ws = torch.empty(lib.nrf_voxelize_workspace_bytes(B, N, S), device=coords.device, dtype=torch.uint8)
        with torch.cuda.device(coords.device):          # launched on the tensors' device, whatever is current
            check(lib.nrf_voxelize(ptr(coords), ptr(coord_features), B, N, F, ptr(geom), S, ptr(out), ptr(ws),
                                   stream_ptr()), "nrf_voxelize")
        return out if not only_features else out[..., :-7]

    forward = coords_to_bounding_voxel_grid
