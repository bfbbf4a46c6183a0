"""Shared test helpers: seeded inputs with cluster structure so mined extrema have gaps >> rounding noise."""
import numpy as np


def clustered(B, D, ids_per=4, noise=0.25, seed=0, shuffle=True, dtype=np.float32):
    rng = np.random.default_rng(seed)
    n_ids = max(2, (B + ids_per - 1) // ids_per)
    labels = np.repeat(np.arange(n_ids), ids_per)[:B].astype(np.int32)
    cent = rng.standard_normal((n_ids, D))
    E = cent[labels] + noise * rng.standard_normal((B, D))
    if shuffle:
        perm = rng.permutation(B)
        E, labels = E[perm], labels[perm]
    return E.astype(dtype), labels


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return np.abs(a - b).max() / max(1e-30, np.abs(b).max())
