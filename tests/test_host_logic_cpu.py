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
