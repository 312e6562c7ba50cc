"""oracle/pca_np.py against the installed scikit-learn (the dependency the reference calls, concat_pca_sn.py:33-36,56-61)."""
import numpy as np


def test_ensemble_tail_matches_sklearn():
    from sklearn.decomposition import PCA
    from sklearn.preprocessing import normalize
    from oracle import pca_np
    rng = np.random.default_rng(0)
    parts = [rng.standard_normal((300, d)).astype(np.float32) * s for d, s in ((64, 1.0), (48, 3.0), (32, 0.1))]
    parts[1][7] = 0.0                                                   # an all-zero descriptor stays zero
    x = np.concatenate([normalize(p) for p in parts], axis=1)
    pca = PCA(n_components=16, random_state=2023).fit(x)
    ref = pca.transform(x)
    got = pca_np.ensemble_pca(parts, pca.mean_, pca.components_)
    np.testing.assert_allclose(got, ref, rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(pca_np.normalize_rows(parts[1]), normalize(parts[1]), rtol=1e-6, atol=1e-7)
