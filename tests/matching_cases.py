"""Seeded query / reference / candidate sets shared by the matching-feature tests (CPU pin test and GPU parity test)."""
import numpy as np


def make_case(seed=0, n_q=10, n_r=12, d=32):
    rng = np.random.default_rng(seed)
    unit = lambda x: (x / np.linalg.norm(x, axis=1, keepdims=True)).astype(np.float32)
    ref = {f"R{i}": unit(rng.standard_normal((int(rng.integers(3, 260)), d))) for i in range(n_r)}
    query, len_map = {}, {}
    for i in range(n_q):
        base = int(rng.integers(2, 70))
        copies = int(rng.integers(1, 5))
        f = rng.standard_normal((base * copies, d))
        if copies > 1:                                       # one copy resembles reference i
            k = int(rng.integers(0, copies))
            src = ref[f"R{i % n_r}"]
            L = min(base, len(src))
            f[k * base:k * base + L] = src[:L] + 0.3 * rng.standard_normal((L, d)) / np.sqrt(d)
        if i == 3:
            f = f[: len(f) - 1] if len(f) > base else f      # ragged last copy (len not a multiple of num_data)
        query[f"Q{i}"] = unit(f)
        len_map[f"Q{i}"] = base
    query["Q_big"] = unit(rng.standard_normal((300, d)))      # more rows than the image holds
    len_map["Q_big"] = 300
    cands = [(f"Q{i}", f"R{(i + j) % n_r}", float(rng.uniform())) for i in range(n_q) for j in range(2)]
    cands.append(("Q_big", "R0", 0.5))
    return query, ref, cands, len_map
