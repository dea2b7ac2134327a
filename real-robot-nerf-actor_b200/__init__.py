"""B200-native feature-NeRF renderer: drop-in for the reference's `NeuralRenderer` hot path.

The directory name carries a hyphen (it is fixed by the project layout); import it with
`importlib.import_module("real-robot-nerf-actor_b200")` or put `real-robot-nerf-actor_b200/dropin`
on `sys.path` and `from neural_rendering import NeuralRenderer` exactly as the reference's
train_nerfact_*_kitchen.py scripts do.
"""
__all__ = ["synthetic"]
