"""Stub loader that imports the UNMODIFIED reference renderer from /root/reference.

TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ may be imported by the product
package; only tests/, tests/golden/make_golden.py, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference leg use it.

The reference hot-path files do not import as shipped:
  * models_embed.py:6,8   imports agents.nerfodiseactv4_bc.{resnetfc,utils}
  * resnetfc.py:6,8,9     imports agents.nerfinact_bc.utils, .network_utils, .attention
  * attention.py:12       imports ldm.modules.diffusionmodules.util
  * utils.py:2-10         imports pyrender, trimesh, rlbench, pyrep at module top
  * neural_rendering.py:4,13,125 imports termcolor, dotmap, odise
This module registers stand-ins for the missing third-party packages and fake
`agents.*` packages whose __path__ is the reference root, then imports the four
reference files without touching them.  It only works where /root/reference
exists (the build container); the GPU box never calls it.
"""
from __future__ import annotations

import importlib
import os
import sys
import types
from unittest import mock

REFERENCE_ROOT = os.environ.get("NRF_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "neural_rendering.py"))


class _AttrDict(dict):
    """dict with attribute access: stands in for dotmap.DotMap on this path."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    def toDict(self):
        return dict(self)


class Cfg(dict):
    """Config object giving both cfg.key and cfg["key"] (pyhocon ConfigTree stand-in)."""

    def __init__(self, d):
        super().__init__()
        for k, v in d.items():
            self[k] = Cfg(v) if isinstance(v, dict) else v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def default_cfg(**over):
    """nerfact.conf:14-104 `neural_renderer{}` block at the BASELINE dims (C=128, D=384)."""
    d = dict(
        foundation_model_name="diffusion", d_embed=384, d_latent=128,
        use_multi_scale_voxel=False, d_multi_scale_latent=266, use_depth_supervision=False,
        lambda_embed=0.01, lambda_depth=0.0, threshold_depth_supervision=0.8,
        ray_chunk_size=512, d_lang=128, voxel_shape=100, share_mlp=True,
        image_width=128, image_height=128, z_near=1.2, z_far=4.0,
        regress_coord=False, regress_attention=False, ret_last_feat=False,
        use_code=True, use_code_viewdirs=False, use_freenerf=False, use_xyz=True,
        n_coarse=64, n_fine=64, n_fine_depth=0, white_bkgd=False, lindisp=False,
        normalize_z=False, canon_xyz=True, use_viewdirs=True, eval_batch_size=4096,
        noise_std=0.0, depth_std=0.001,
        mlp=dict(n_blocks=5, d_hidden=512, combine_layer=3, combine_type="average",
                 beta=0.0, use_spade=False, use_language=False),
        code=dict(num_freqs=6, freq_factor=1.5, include_input=True),
    )
    for k, v in over.items():
        if isinstance(v, dict) and isinstance(d.get(k), dict):
            d[k].update(v)
        else:
            d[k] = v
    return Cfg(d)


_loaded = None


def load_reference():
    """Returns the reference `neural_rendering` module (imported unmodified)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not reference_available():
        raise RuntimeError(f"reference not present at {REFERENCE_ROOT}")

    def stub(name, **attrs):
        m = sys.modules.get(name)
        if m is None:
            m = mock.MagicMock(name=name)
            m.__name__ = name
            m.__path__ = []
            sys.modules[name] = m
        for k, v in attrs.items():
            setattr(m, k, v)
        return m

    tc = types.ModuleType("termcolor")
    tc.colored = lambda s, *a, **k: s
    tc.cprint = lambda *a, **k: None
    sys.modules.setdefault("termcolor", tc)
    dm = types.ModuleType("dotmap")
    dm.DotMap = _AttrDict
    sys.modules.setdefault("dotmap", dm)
    for name in ("pyrender", "pyrender.trackball", "trimesh", "rlbench", "rlbench.backend",
                 "rlbench.backend.const", "rlbench.backend.observation", "pyrep", "pyrep.const",
                 "pyrep.objects", "odise", "odise.modeling", "odise.modeling.meta_arch",
                 "odise.modeling.meta_arch.ldm", "ldm", "ldm.modules",
                 "ldm.modules.diffusionmodules", "ldm.modules.diffusionmodules.util", "xformers",
                 "xformers.ops", "vlab"):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                stub(name)
    for pkg in ("agents", "agents.nerfodiseactv4_bc", "agents.nerfinact_bc"):
        if pkg not in sys.modules:
            m = types.ModuleType(pkg)
            m.__path__ = [REFERENCE_ROOT]
            sys.modules[pkg] = m
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        saved_utils = sys.modules.pop("utils", None)
        ref_utils = importlib.import_module("utils")
        sys.modules["agents.nerfodiseactv4_bc.utils"] = ref_utils
        sys.modules["agents.nerfinact_bc.utils"] = ref_utils
        nr = importlib.import_module("neural_rendering")
    finally:
        sys.path.remove(REFERENCE_ROOT)
    # keep the reference modules reachable under private names, and free the generic ones
    for generic in ("utils", "neural_rendering", "models_embed", "dino",
                    "vision_transformer_flexible"):
        mod = sys.modules.pop(generic, None)
        if mod is not None:
            sys.modules["_nrf_reference_" + generic] = mod
    if saved_utils is not None:
        sys.modules["utils"] = saved_utils
    _loaded = nr
    return nr


def build_reference_renderer(cfg=None, bounds=None):
    """Constructs the reference NeuralRenderer on CPU (models_embed.py:31 calls .cuda())."""
    import torch
    nr = load_reference()
    cfg = cfg if cfg is not None else default_cfg()
    if bounds is None:
        bounds = torch.tensor([-0.1, -0.3, -0.2, 0.8, 0.7, 0.7])
    if torch.cuda.is_available():
        return nr.NeuralRenderer(cfg, coordinate_bounds=bounds)
    with mock.patch.object(torch.Tensor, "cuda", lambda self, *a, **k: self):
        return nr.NeuralRenderer(cfg, coordinate_bounds=bounds)


class inject_noise:
    """Replaces torch.rand_like / rand / randn_like by a queue of pre-drawn tensors.

    The reference has no `perturb` flag (neural_rendering.py:172,194,200,218 always draw);
    'perturb off' is the injection of zeros.  `draws` is a list consumed in call order;
    an entry of None means 'zeros of the requested shape'.
    """

    def __init__(self, draws):
        self.draws = list(draws)

    def __enter__(self):
        import torch
        self._p = []
        q = self.draws

        def pop_like(t):
            v = q.pop(0)
            return torch.zeros_like(t) if v is None else v.to(t.device).reshape(t.shape)

        def pop_shape(*shape, **kw):
            v = q.pop(0)
            if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
                shape = tuple(shape[0])
            dev = kw.get("device", None)
            z = torch.zeros(*shape, dtype=kw.get("dtype", torch.float32), device=dev)
            return z if v is None else v.to(z.device).reshape(z.shape)

        for name, fn in (("rand_like", pop_like), ("randn_like", pop_like), ("rand", pop_shape)):
            p = mock.patch.object(torch, name, fn)
            p.start()
            self._p.append(p)
        return self

    def __exit__(self, *exc):
        for p in self._p:
            p.stop()
        return False
