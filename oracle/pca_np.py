"""CPU restatement of the reference's ensemble tail -- TEST INFRASTRUCTURE (see oracle/__init__.py).

The arithmetic lives in scikit-learn (reference pin: `pip install scikit-learn` without a version, dockerfile /
setup.sh; installed here: see sklearn.__version__), called at D/infer/concat_pca_sn.py:56-64 and
D/infer/extract_query_feats.py:169-204:
    vid_feats = [normalize(x[vid].feature) for x in features_list]      # sklearn.preprocessing.normalize, L2 rows
    vid_feats = np.concatenate(vid_feats, axis=1)
    vid_feats = pca_model.transform(vid_feats)                          # (X - mean_) @ components_.T, whiten=False
Pinned by tests/test_oracle_pca.py against the installed scikit-learn itself.
"""
import numpy as np


def normalize_rows(x: np.ndarray) -> np.ndarray:
    x = np.asarray(x, dtype=np.float32)
    norms = np.sqrt((x.astype(np.float32) ** 2).sum(axis=1, dtype=np.float32))
    norms = np.where(norms == 0.0, np.float32(1.0), norms)          # sklearn: zero rows are left untouched
    return x / norms[:, None]


def ensemble_pca(parts, mean, components) -> np.ndarray:
    x = np.concatenate([normalize_rows(p) for p in parts], axis=1)
    return (x - np.asarray(mean, np.float32)) @ np.asarray(components, np.float32).T
