"""Ray / encoding helpers with the reference's names (utils.py:434-567), on the C ABI where it matters."""
from __future__ import annotations

import numpy as np
import torch

from . import ops


class AttrDict(dict):
    """Attribute-access dict: the role dotmap.DotMap plays for forward_nerf's outputs
    (neural_rendering.py:419,447)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    def toDict(self):
        return {k: (v.toDict() if isinstance(v, AttrDict) else v) for k, v in self.items()}


class ConfigDict(dict):
    """Config object supporting both cfg.key and cfg["key"], like pyhocon's ConfigTree
    (nerfact.conf is read that way: neural_rendering.py:93-154, models_embed.py:26-120)."""

    def __init__(self, d=None, **kw):
        super().__init__()
        for k, v in {**(d or {}), **kw}.items():
            self[k] = ConfigDict(v) if isinstance(v, dict) and not isinstance(v, ConfigDict) else v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def default_config(**over) -> ConfigDict:
    """The `neural_renderer{}` block of nerfact.conf:14-104, at the BASELINE dims (C=128, D=384)."""
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
            d[k] = {**d[k], **v}
        else:
            d[k] = v
    return ConfigDict(d)


def _conf_tokens(text):
    """Tokens of the HOCON subset nerfact.conf uses: { } [ ] = : , newline, quoted / unquoted scalars; # and // comments."""
    i, n = 0, len(text)
    while i < n:
        ch = text[i]
        if ch == "#" or text.startswith("//", i):
            while i < n and text[i] != "\n":
                i += 1
        elif ch == "\n":
            yield ("nl", ch)
            i += 1
        elif ch.isspace():
            i += 1
        elif ch in "{}[]=:,":
            yield (ch, ch)
            i += 1
        elif ch in "\"'":
            j = text.find(ch, i + 1)
            if j < 0:
                raise ValueError("unterminated string in config")
            yield ("str", text[i + 1:j])
            i = j + 1
        else:
            j = i
            while j < n and text[j] not in "{}[]=:,\n#" and not text.startswith("//", j):
                j += 1
            yield ("raw", text[i:j].strip())
            i = j


def _conf_scalar(raw):
    """An unquoted value the way pyhocon reads it: int, float, true / false in any case, null; anything else - `None`
    and `average` included (nerfact.conf:43-44,:95) - stays a string."""
    low = raw.lower()
    if low in ("true", "false"):
        return low == "true"
    if low == "null":
        return None
    for cast in (int, float):
        try:
            return cast(raw)
        except ValueError:
            pass
    return raw


def parse_conf(text: str) -> ConfigDict:
    """Parses the HOCON subset of the reference's nerfact.conf (key = value / key : value / key { ... } blocks, dotted
    keys, [lists], quoted and bare scalars, # and // comments) into a ConfigDict, so that the config loads without
    pyhocon (`ConfigFactory.parse_file`, train_nerfact_multi_kitchen.py:1244-1245): the result answers both
    `conf['neural_renderer'].image_width = W` and `conf.mlp.d_hidden` like pyhocon's ConfigTree.  Substitutions
    (${...}), includes and multi-line strings are not part of the subset and raise."""
    toks = list(_conf_tokens(text))
    pos = 0

    def peek():
        return toks[pos] if pos < len(toks) else ("eof", "")

    def skip_seps():
        nonlocal pos
        while peek()[0] in ("nl", ","):
            pos += 1

    def value():
        nonlocal pos
        kind, tok = peek()
        if kind == "{":
            pos += 1
            return obj("}")
        if kind == "[":
            pos += 1
            out = []
            skip_seps()
            while peek()[0] != "]":
                if peek()[0] == "eof":
                    raise ValueError("unterminated list in config")
                out.append(value())
                skip_seps()
            pos += 1
            return out
        if kind == "str":
            pos += 1
            return tok
        if kind == "raw":
            if "${" in tok:
                raise ValueError("config substitutions are not supported")
            pos += 1
            return _conf_scalar(tok)
        raise ValueError(f"unexpected {tok!r} in config")

    def put(d, dotted, v):
        parts = dotted.split(".")
        for k in parts[:-1]:
            d = d.setdefault(k, {})
        if isinstance(v, dict) and isinstance(d.get(parts[-1]), dict):
            d[parts[-1]].update(v)            # HOCON merges repeated objects
        else:
            d[parts[-1]] = v

    def obj(end):
        nonlocal pos
        d = {}
        skip_seps()
        while peek()[0] != end:
            kind, key = peek()
            if kind not in ("raw", "str"):
                raise ValueError(f"expected a key, got {key!r}")
            if key == "include":
                raise ValueError("config includes are not supported")
            pos += 1
            if peek()[0] in ("=", ":"):
                pos += 1
            elif peek()[0] != "{":
                raise ValueError(f"expected '=', ':' or '{{' after key {key!r}")
            put(d, key, value())
            skip_seps()
        pos += 1
        return d

    toks.append(("eof", ""))
    skip_seps()
    if peek()[0] == "{":
        pos += 1
        d = obj("}")
    else:
        d = obj("eof")
    return ConfigDict(d)


def load_conf(path) -> ConfigDict:
    """`ConfigFactory.parse_file(path)` for nerfact.conf without pyhocon (see parse_conf)."""
    with open(path) as f:
        return parse_conf(f.read())


def repeat_interleave(input, repeats, dim=0):
    """utils.py:434-441."""
    output = input.unsqueeze(1).expand(-1, repeats, *input.shape[1:])
    return output.reshape(-1, *input.shape[1:])


def combine_interleaved(t, inner_dims=(1,), agg_type="average"):
    """utils.py:509-519."""
    if len(inner_dims) == 1 and inner_dims[0] == 1:
        return t
    t = t.reshape(-1, *inner_dims, *t.shape[1:])
    if agg_type == "average":
        return torch.mean(t, dim=1)
    if agg_type == "max":
        return torch.max(t, dim=1)[0]
    raise NotImplementedError("Unsupported combine type " + agg_type)


def gen_rays(poses, width, height, focal, z_near, z_far, c=None):
    """utils.py:477-506 -> (B,H,W,8), computed by nrf_raygen."""
    f = focal.squeeze() if torch.is_tensor(focal) else focal
    return ops.raygen(poses, width, height, f, z_near, z_far, c=c)


def unproj_map(width, height, f, c=None, device="cuda"):
    """utils.py:444-474 -> (H,W,3): the direction part of gen_rays for an identity pose."""
    eye = torch.eye(4, device=device, dtype=torch.float32).unsqueeze(0)
    return ops.raygen(eye, width, height, f, 0.0, 0.0, c=c)[0, :, :, 3:6].contiguous()


class PositionalEncoding(torch.nn.Module):
    """utils.py:521-567: carries the reference's persistent buffers `_freqs` / `_phases` (state_dict
    compatibility); the encoding itself is fused into nrf_encode_points."""

    def __init__(self, num_freqs=6, d_in=3, freq_factor=np.pi, include_input=True):
        super().__init__()
        self.num_freqs = num_freqs
        self.d_in = d_in
        self.freq_factor = freq_factor
        self.freqs = freq_factor * 2.0 ** torch.arange(0, num_freqs)
        self.d_out = self.num_freqs * 2 * d_in
        self.include_input = include_input
        if include_input:
            self.d_out += d_in
        self.register_buffer("_freqs", torch.repeat_interleave(self.freqs, 2).view(1, -1, 1))
        _phases = torch.zeros(2 * self.num_freqs)
        _phases[1::2] = np.pi * 0.5
        self.register_buffer("_phases", _phases.view(1, -1, 1))

    @classmethod
    def from_conf(cls, conf, d_in=3):
        g = lambda k: conf[k] if isinstance(conf, dict) else getattr(conf, k)
        return cls(g("num_freqs"), d_in, g("freq_factor"), g("include_input"))


def pca_fit_transform(x: torch.Tensor, n_components: int) -> torch.Tensor:
    """sklearn.decomposition.PCA(n_components).fit_transform(x) on the device x lives on (neural_rendering.py:640-646
    runs it on the CPU through numpy: a device -> host -> device round trip of the whole target feature map and a host
    stall in every training step).

    x (N, D) -> scores (N, n_components) = (x - mean) . V_k, V_k the top principal axes.  Exact PCA through the (D, D)
    covariance eigen-decomposition in fp64 (N >> D here: N = B*H*W pixels); signs follow sklearn's svd_flip
    (u_based_decision=False: the largest-magnitude entry of every axis is positive), so the result equals sklearn's
    `svd_solver="full"` / "covariance_eigh" output up to rounding.  (sklearn's "auto" may pick the randomized solver,
    whose output is itself only approximate and seed-dependent.)"""
    if x.dim() != 2:
        raise ValueError("pca_fit_transform expects (N, D)")
    n, d = x.shape
    if not 0 < n_components <= min(n, d):
        raise ValueError(f"n_components={n_components} must be in (0, min(N, D)={min(n, d)}]")
    xd = x.to(torch.float64)
    xc = xd - xd.mean(0, keepdim=True)
    cov = xc.t() @ xc                                   # (D, D); the 1/(N-1) factor does not change the axes
    evals, evecs = torch.linalg.eigh(cov)               # ascending
    v = evecs[:, -n_components:].flip(1)                # (D, k), descending variance
    idx = v.abs().argmax(0)
    signs = torch.sign(v[idx, torch.arange(n_components, device=v.device)])
    signs = torch.where(signs == 0, torch.ones_like(signs), signs)
    return (xc @ (v * signs)).to(x.dtype)
