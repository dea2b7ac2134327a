"""Synthetic workload factory for the BASELINE.json configs.

Everything is drawn from seeded CPU generators so that the build container (where the golden
fixtures are produced from the reference) and the GPU box regenerate bit-identical inputs.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch

BOUNDS = (-0.1, -0.3, -0.2, 0.8, 0.7, 0.7)      # train_nerfact_multi_kitchen.py:1131
BOX_CENTRE = (0.35, 0.2, 0.25)


@dataclass
class Workload:
    name: str
    S: int            # voxel grid side
    C: int            # latent channels
    D: int            # embed channels
    SB: int           # scenes
    rays_per_scene: int
    n_coarse: int
    n_fine: int
    H: int = 128
    W: int = 128
    focal: float = 153.0
    n_cams: int = 1
    train: bool = True

    @property
    def evals(self) -> int:
        """Field evaluations per forward_nerf: R*(Kc + (Kc+Kf)) (SURVEY 8d 'Metric')."""
        R = self.SB * self.rays_per_scene
        fine = (self.n_coarse + self.n_fine) if self.n_fine > 0 else 0
        return R * (self.n_coarse + fine)


CONFIGS = {
    "config1": Workload("config1", 100, 128, 384, 1, 512, 64, 0, train=False),
    "config2": Workload("config2", 100, 128, 384, 2, 2048, 64, 64),
    "config3": Workload("config3", 100, 128, 384, 1, 5 * 128 * 128, 64, 64, n_cams=5, train=False),
    "config4": Workload("config4", 100, 128, 384, 8, 4096, 64, 64),
    "config5": Workload("config5", 200, 128, 384, 1, 16384, 128, 128),
    # the reference's own training shape (nerfact.conf:22-28,:606): one scene, 512-ray chunks, 64 latent channels,
    # 512-d features -- a launch-bound step, kept for host-overhead work (scripts/profile_step.py nerfact)
    "nerfact": Workload("nerfact", 100, 64, 512, 1, 512, 64, 64),
}


def look_at_pose(eye, target=BOX_CENTRE, up=(0.0, 0.0, 1.0)):
    """OpenGL camera-to-world (camera looks down -z), 4x4 fp32."""
    eye = torch.tensor(eye, dtype=torch.float64)
    tgt = torch.tensor(target, dtype=torch.float64)
    back = eye - tgt
    back = back / back.norm()
    right = torch.linalg.cross(torch.tensor(up, dtype=torch.float64), back)
    right = right / right.norm()
    upv = torch.linalg.cross(back, right)
    m = torch.eye(4, dtype=torch.float64)
    m[:3, 0], m[:3, 1], m[:3, 2], m[:3, 3] = right, upv, back, eye
    return m.to(torch.float32)


def arc_poses(n, radius=2.8, elev_deg=30.0):
    """n cameras on an arc of `radius` metres around the box centre, looking at it."""
    poses = []
    for i in range(n):
        az = math.radians(-60.0 + 120.0 * (i + 0.5) / n)
        el = math.radians(elev_deg)
        eye = (BOX_CENTRE[0] + radius * math.cos(el) * math.cos(az),
               BOX_CENTRE[1] + radius * math.cos(el) * math.sin(az),
               BOX_CENTRE[2] + radius * math.sin(el))
        poses.append(look_at_pose(eye))
    return torch.stack(poses)


def make_volume(SB, C, S, seed=0, scale=0.1):
    """voxel_feat ~ scale*N(0,1), (SB,C,S,S,S) fp32 channel-first, CPU."""
    g = torch.Generator().manual_seed(1000 + seed)
    return torch.randn(SB, C, S, S, S, generator=g) * scale


def make_targets(SB, n, D, seed=0):
    g = torch.Generator().manual_seed(2000 + seed)
    return torch.rand(SB, n, 3, generator=g), torch.randn(SB, n, D, generator=g)


def make_noise(R, n_coarse, n_fine, seed=0, perturb=True):
    """Pre-drawn noise for one forward_nerf: dict(coarse, u, fine); zeros when perturb is off.

    `u` (the inverse-CDF draw) is always random in [0,1): a zero `u` would put every fine
    sample in bin 0.
    """
    g = torch.Generator().manual_seed(3000 + seed)
    n = {"u": torch.rand(R, n_fine, generator=g)} if n_fine > 0 else {}
    if perturb:
        n["coarse"] = torch.rand(R, n_coarse, generator=g)
        if n_fine > 0:
            n["fine"] = torch.rand(R, n_fine, generator=g)
    return n


def pick_ray_indices(n_pixels, n, seed=0):
    g = torch.Generator().manual_seed(4000 + seed)
    return torch.randint(n_pixels, (n,), generator=g)


def init_mlp_(mlp, seed=0):
    """Reference init (resnetfc.py:38-41,92-98,126-128) from a seeded CPU generator, with
    fc_1.weight ~ N(0, 2/d_hidden) instead of zeros so the residual blocks are not identities
    (SURVEY.md 9.8).  In-place on a ResnetFC parameter container; returns it."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in mlp.named_parameters():
            if name.endswith(".bias"):
                p.zero_()
            else:
                fan_in = p.shape[1]
                p.copy_((torch.randn(p.shape, generator=g) * math.sqrt(2.0 / fan_in)).to(p.device))
    return mlp


def voxelizer_points(B, N, F, seed):
    """Clustered synthetic point cloud (several points per voxel, some outside the bounds, exact face hits)."""
    g = torch.Generator().manual_seed(seed)
    b = torch.tensor(BOUNDS)
    centres = torch.rand(B, 12, 3, generator=g) * (b[3:] - b[:3]) + b[:3]
    which = torch.randint(12, (B, N), generator=g)
    coords = torch.gather(centres, 1, which.unsqueeze(-1).expand(-1, -1, 3)) + 0.03 * torch.randn(B, N, 3, generator=g)
    coords[:, :8] = torch.rand(B, 8, 3, generator=g) * 3.0 - 1.0            # far outside
    coords[:, 8] = b[:3]                                                    # exactly on the lower corner
    coords[:, 9] = b[3:]                                                    # exactly on the upper corner
    feats = torch.rand(B, N, F, generator=g)
    return coords, feats
