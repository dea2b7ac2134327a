"""Host-side logic that needs no GPU: the device PCA against sklearn, the loss dictionary's dict behaviour."""
import numpy as np
import pytest
import torch

from tests.conftest import load_pkg


def test_pca_fit_transform_equals_sklearn():
    """neural_rendering.py:640-646 reduces the target feature map with sklearn's PCA on the CPU; the replacement runs on
    whatever device the features live on and must give the same scores (signs included)."""
    sk = pytest.importorskip("sklearn.decomposition")
    U = load_pkg("utils")
    g = torch.Generator().manual_seed(0)
    for n, d, k in ((2000, 96, 24), (513, 40, 7), (300, 64, 64)):
        x = torch.randn(n, d, generator=g) @ torch.randn(d, d, generator=g) + 3.0
        ref = sk.PCA(n_components=k, svd_solver="full").fit_transform(x.numpy())
        got = U.pca_fit_transform(x, k).numpy()
        assert got.shape == ref.shape
        assert np.abs(got - ref).max() <= 2e-4 * np.abs(ref).max(), (n, d, k)
    with pytest.raises(ValueError):
        U.pca_fit_transform(torch.randn(10, 4), 5)


def test_parse_conf_reads_the_hocon_subset():
    """utils.parse_conf: the pieces of HOCON nerfact.conf is written in (blocks, `=` / `:`, dotted keys, lists, quoted and
    bare scalars, comments), typed the way pyhocon types them."""
    U = load_pkg("utils")
    c = U.parse_conf("""
        # a comment
        top = 3   // trailing comment
        renderer {
          name = diffusion      # bare string
          path = None           # pyhocon keeps `None` as a string; only null is None
          missing = null
          on = True
          ratio : 1.5e-1
          sizes = [100, 2.5, "a b", false]
          mlp { d_hidden = 512 }
          mlp.beta = 0.0
        }
        renderer.mlp.n_blocks = 5
    """)
    r = c["renderer"]
    assert c.top == 3 and r.name == "diffusion" and r.path == "None" and r.missing is None and r.on is True
    assert r.ratio == 0.15 and r.sizes == [100, 2.5, "a b", False]
    assert dict(r.mlp) == {"d_hidden": 512, "beta": 0.0, "n_blocks": 5} and isinstance(r.mlp, U.ConfigDict)
    r.image_width = 80                                     # train_nerfact_multi_kitchen.py:1246 writes through attributes
    assert c["renderer"]["image_width"] == 80
    assert dict(U.parse_conf("{ a = 1 }")) == {"a": 1}
    for bad in ("a = ${b}", "include \"x.conf\"", "a = [1, 2", "a = \"open", "= 3"):
        with pytest.raises(ValueError):
            U.parse_conf(bad)


def test_the_reference_nerfact_conf_builds_the_renderer():
    """The reference's own nerfact.conf, read without pyhocon, through the script's own lines
    (train_nerfact_multi_kitchen.py:1244-1248) into the drop-in renderer: every key the renderer reads is there and
    the model has nerfact.conf's dims (64 latent channels, 512-d features, 64 + 32 + 16 samples)."""
    import os
    path = "/root/reference/nerfact.conf"
    if not os.path.exists(path):
        pytest.skip("the reference tree is not on this machine")
    U, NR = load_pkg("utils"), load_pkg("neural_rendering")
    conf = U.load_conf(path)
    assert conf.language_model == "CLIP" and conf.voxel_sizes == [100] and conf.bounds_offset == [0.15]
    conf["neural_renderer"].image_width = 80
    conf["neural_renderer"].image_height = 60
    nr = conf["neural_renderer"]
    assert (nr.d_latent, nr.d_embed, nr.n_coarse, nr.n_fine, nr.n_fine_depth) == (64, 512, 64, 32, 16)
    assert nr.dino_path == "None" and nr.mlp.combine_type == "average" and nr.code.freq_factor == 1.5
    ren = NR.NeuralRenderer(nr, torch.tensor([-0.1, -0.3, -0.2, 0.8, 0.7, 0.7]))
    assert (ren.W, ren.H, ren.n_coarse, ren.n_fine, ren.n_fine_depth) == (80, 60, 64, 32, 16)
    assert (ren.z_near, ren.z_far, ren.lambda_embed, ren.eval_batch_size) == (1.2, 4.0, 0.01, 4096)
    mlp = ren.nerf_model.mlp_coarse
    assert mlp.lin_in.weight.shape == (512, 42) and mlp.lin_out.weight.shape == (4 + 512, 512)
    assert len(mlp.lin_z) == 3 and mlp.lin_z[0].weight.shape == (512, 64) and len(mlp.blocks) == 5
    assert ren.nerf_model.mlp_fine is mlp                   # share_mlp = True
