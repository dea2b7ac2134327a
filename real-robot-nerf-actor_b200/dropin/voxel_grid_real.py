"""`from voxel_grid_real import VoxelGrid` shim: with this directory first on sys.path the reference's training
scripts pick up the B200 voxelizer unchanged.  See INTEGRATION.md."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _root not in sys.path:
    sys.path.insert(0, _root)
_impl = importlib.import_module("real-robot-nerf-actor_b200.voxel_grid")
VoxelGrid = _impl.VoxelGrid
MIN_DENOMINATOR = _impl.MIN_DENOMINATOR
