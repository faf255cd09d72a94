"""Synthetic document-term corpora (SURVEY.md §8d recipe).

Token-level sampler: document weights ~ lognormal(0, 0.6), term distribution Zipf
p(r) ∝ 1/r over m ranks, T tokens drawn independently, duplicates summed into counts.
T is solved analytically (Poisson occupancy, quantile-binned document weights) so that the
number of distinct (doc, term) pairs lands on the requested nnz.  ``planted=True`` mixes in
topic-specific vocabularies (50/50 with the global Zipf, Dirichlet(0.1) doc-topic weights)
so that EM contracts to a well separated optimum; marginals stay Zipf.

Used by bench.py, the tests and __graft_entry__.smoke(); pure numpy/scipy host code.
"""
import numpy as np
import scipy.sparse as sp

# (n_docs, n_terms, target_nnz, n_components, n_iter, matrix_seed) per BASELINE.json config
CONFIGS = {
    "C1": dict(n=2_000, m=5_000, nnz=200_000, k=10, n_iter=50, seed=1),
    "C2": dict(n=100_000, m=50_000, nnz=10_000_000, k=20, n_iter=100, seed=0),
    "C3": dict(n=100_000, m=50_000, nnz=10_000_000, k=128, n_iter=100, seed=0),
    "C5": dict(n=1_000_000, m=200_000, nnz=200_000_000, k=20, n_iter=50, seed=2),
}


def _zipf(m, s=1.0):
    p = 1.0 / np.arange(1, m + 1, dtype=np.float64) ** s
    return p / p.sum()


def _expected_nnz(T, wq, wq_count, pw):
    # E[#distinct pairs] = sum_d sum_w 1 - exp(-T p_d p_w), docs binned into quantiles
    lam = T * wq[:, None] * pw[None, :]
    return float((wq_count[:, None] * -np.expm1(-lam)).sum())


def solve_tokens(doc_w, pw, target_nnz, bins=96):
    order = np.sort(doc_w)
    chunks = np.array_split(order, min(bins, order.shape[0]))
    wq = np.array([c.mean() for c in chunks])
    cnt = np.array([c.shape[0] for c in chunks], dtype=np.float64)
    lo, hi = float(target_nnz), float(target_nnz) * 64.0
    for _ in range(60):
        mid = 0.5 * (lo + hi)
        if _expected_nnz(mid, wq, cnt, pw) < target_nnz:
            lo = mid
        else:
            hi = mid
    return int(0.5 * (lo + hi))


def _draw(cdf, u):
    idx = np.searchsorted(cdf, u, side="right")
    np.minimum(idx, cdf.shape[0] - 1, out=idx)
    return idx


def make_corpus(n, m, nnz, seed=0, planted=False, k_true=10, dtype=np.int32,
                chunk=4_000_000, return_info=False):
    """Return a CSR matrix (sorted indices, no duplicates, integer counts) with about
    ``nnz`` stored entries."""
    rng = np.random.default_rng(seed)
    doc_w = rng.lognormal(0.0, 0.6, size=n)
    doc_w /= doc_w.sum()
    pw = _zipf(m)
    T = solve_tokens(doc_w, pw, nnz)
    doc_cdf = np.cumsum(doc_w)
    term_cdf = np.cumsum(pw)
    if planted:
        theta = rng.dirichlet(np.full(k_true, 0.1), size=n)
        theta_cdf = np.cumsum(theta, axis=1)
        perms = np.stack([rng.permutation(m) for _ in range(k_true)])

    keys = []
    done = 0
    while done < T:
        c = min(chunk, T - done)
        d = _draw(doc_cdf, rng.random(c))
        r = _draw(term_cdf, rng.random(c))
        if planted:
            u = rng.random(c)
            z = (u[:, None] > theta_cdf[d]).sum(axis=1)
            np.minimum(z, k_true - 1, out=z)
            use_topic = rng.random(c) < 0.5
            r = np.where(use_topic, perms[z, r], r)
        keys.append(d.astype(np.int64) * m + r)
        done += c
    keys = np.concatenate(keys)
    uniq, counts = np.unique(keys, return_counts=True)
    del keys
    rows = uniq // m
    cols = (uniq - rows * m).astype(np.int32)
    indptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(np.bincount(rows, minlength=n), out=indptr[1:])
    X = sp.csr_matrix((counts.astype(dtype), cols, indptr.astype(np.int32)), shape=(n, m))
    X.has_sorted_indices = True
    if return_info:
        info = dict(tokens=int(T), seed=int(seed), nnz=int(X.nnz), sum_counts=int(counts.sum()),
                    max_count=int(counts.max()), planted=bool(planted))
        return X, info
    return X


def make_config(name, planted=False, return_info=False):
    c = CONFIGS[name]
    return make_corpus(c["n"], c["m"], c["nnz"], seed=c["seed"], planted=planted,
                       return_info=return_info)
