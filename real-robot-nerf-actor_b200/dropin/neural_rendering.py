"""`from neural_rendering import NeuralRenderer` shim: put this directory first on sys.path and the
reference's train_nerfact_*_kitchen.py import lines (train_nerfact_multi_kitchen.py:52) pick up the
B200 renderer unchanged.  See INTEGRATION.md."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _root not in sys.path:
    sys.path.insert(0, _root)
_impl = importlib.import_module("real-robot-nerf-actor_b200.neural_rendering")
NeuralRenderer = _impl.NeuralRenderer
PixelNeRFEmbedNet = _impl.PixelNeRFEmbedNet
PSNR_torch = _impl.PSNR_torch
