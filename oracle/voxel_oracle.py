"""TEST INFRASTRUCTURE (not a product path): CPU restatement of the reference voxelizer,
voxel_grid_real.py:175-233 `VoxelGrid.coords_to_bounding_voxel_grid`, in plain torch.

Pinned by tests/golden/voxelize_small.npz, produced by the UNMODIFIED reference class (tests/golden/make_golden.py,
`run_voxelizer`; the reference module imports only torch and numpy).  Only tests/ may import this.
"""
import torch

MIN_DENOMINATOR = 1e-12          # voxel_grid_real.py:11


def voxelize(coords, feats, bounds, S):
    """coords (B,N,3), feats (B,N,F) or None, bounds 6-vector or (B,6) -> (B,S,S,S,3+F+3+1) fp32."""
    coords = coords.float()
    B, N, _ = coords.shape
    bounds = torch.as_tensor(bounds, dtype=torch.float32).reshape(-1, 6)
    bb_mins, bb_maxs = bounds[:, 0:3], bounds[:, 3:6]
    dims_orig = torch.tensor([[S, S, S]], dtype=torch.int32)
    res = (bb_maxs - bb_mins) / (dims_orig.float() + MIN_DENOMINATOR)                    # :74 / :183
    denom = res + MIN_DENOMINATOR                                                       # :78 / :184
    shifted = bb_mins - res                                                             # :186
    idx = torch.floor((coords - shifted.unsqueeze(1)) / denom.unsqueeze(1)).int()       # :187-188
    idx = torch.max(torch.min(idx, torch.tensor(S + 1, dtype=torch.int32)), torch.tensor(0, dtype=torch.int32))
    vals = coords if feats is None else torch.cat([coords, feats.float()], -1)          # :196-198
    vals = torch.cat([vals, torch.ones(B, N, 1)], -1)                                   # :206-207
    nch = vals.shape[-1]
    W = S + 2
    flat = ((torch.arange(B).view(B, 1) * W + idx[..., 0].long()) * W + idx[..., 1].long()) * W + idx[..., 2].long()
    total = torch.zeros(B * W * W * W, nch)
    total.index_add_(0, flat.reshape(-1), vals.reshape(-1, nch))                        # scatter_add_ (:118)
    cnt = torch.zeros(B * W * W * W)
    cnt.index_add_(0, flat.reshape(-1), torch.ones(B * N))
    total = total / cnt.clamp(min=1).unsqueeze(-1)                                      # :125-130
    vox = total.view(B, W, W, W, nch)[:, 1:-1, 1:-1, 1:-1]                               # :212
    occupied = (vox[..., -1:] > 0).float()                                              # :220
    ar = torch.arange(S, dtype=torch.float32)
    grid = torch.stack(torch.meshgrid(ar, ar, ar, indexing="ij"), -1).unsqueeze(0).expand(B, -1, -1, -1, -1)
    return torch.cat([vox[..., :-1], grid / float(S), occupied], -1)                    # :224-226
