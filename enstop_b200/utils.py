"""Host-side helpers that keep the semantics of enstop/utils.py (numpy, no numba).

``normalize`` (utils.py:8-41), ``log_lift`` / ``mean_log_lift`` (utils.py:44-146),
``coherence`` / ``mean_coherence`` (utils.py:149-273), ``standardize_input``
(utils.py:276-280).  These are not on the timed path; they exist so that
``PLSA.coherence()`` / ``PLSA.log_lift()`` keep working after the switch.
"""
import numpy as np
from scipy.sparse import csc_matrix, issparse
from sklearn.preprocessing import normalize as sklearn_normalize

try:  # sklearn's own validator, as the reference prefers (utils.py:282-284 is its fallback)
    from sklearn.utils.validation import _check_sample_weight
except ImportError:  # pragma: no cover
    def _check_sample_weight(sample_weight, X, dtype=None):
        n = X.shape[0]
        if sample_weight is None:
            return np.ones(n, dtype=dtype or np.float64)
        sample_weight = np.asarray(sample_weight, dtype=dtype or np.float64)
        if sample_weight.ndim == 0:
            return np.full(n, float(sample_weight), dtype=dtype or np.float64)
        if sample_weight.shape != (n,):
            raise ValueError("sample_weight.shape == {}, expected {}!".format(
                sample_weight.shape, (n,)))
        return np.ascontiguousarray(sample_weight)


def normalize(ndarray, axis=0):
    """In-place L1 normalisation along ``axis`` with float64 marginals; slices whose
    marginal is not > 0 are left untouched (utils.py:22-41)."""
    if axis not in (0, 1):
        raise ValueError("axis must be 0 or 1")
    marginal = ndarray.sum(axis=axis, dtype=np.float64)
    safe = np.where(marginal > 0.0, marginal, 1.0)
    if axis == 1:
        ndarray /= safe[:, None]
    else:
        ndarray /= safe[None, :]
    return ndarray


def standardize_input(input_matrix):
    """Float input is L1 row-normalised, integer counts are left alone (utils.py:276-280;
    the reference's ``np.float`` spelling fails on numpy >= 1.24, the intent is kept)."""
    if input_matrix.dtype in (np.float32, np.float64):
        return sklearn_normalize(input_matrix, norm="l1")
    return input_matrix


def _empirical_probs(data):
    p = np.array(data.sum(axis=0)).squeeze().astype(np.float64)
    return p / p.sum()


def _log_lift(topics, z, empirical_probs, n=-1):
    if n <= 0:
        sel = np.arange(topics.shape[1])
        denom = topics.shape[1]
    else:
        sel = np.argsort(topics[z])[-n:]
        denom = n
    ok = empirical_probs[sel] > 0
    total = float(np.sum(topics[z, sel][ok] / empirical_probs[sel][ok]))
    return np.log(total / denom)


def log_lift(topics, z, data, n_words=-1):
    normalized = np.array(topics, dtype=np.float64)
    normalize(normalized, axis=1)
    return _log_lift(normalized, z, _empirical_probs(data), n=n_words)


def mean_log_lift(topics, data, n_words=-1):
    # utils.py:141-146 scores the un-normalised `topics` in the mean; kept
    probs = _empirical_probs(data)
    return float(np.mean([_log_lift(topics, z, probs, n=n_words)
                          for z in range(topics.shape[0])]))


def _coherence(topics, z, n, indices, indptr, n_docs_per_word):
    top_words = np.argsort(topics[z])[-n:]
    docs = [indices[indptr[w]: indptr[w + 1]] for w in top_words]
    total = 0.0
    for i in range(n - 1):
        w = top_words[i]
        if n_docs_per_word[w] == 0:
            continue
        for j in range(i + 1, n):
            co = np.intersect1d(docs[i], docs[j], assume_unique=True).shape[0]
            total += np.log((co + 1.0) / n_docs_per_word[w])
    return total


def _csc(data):
    return data.tocsc() if issparse(data) else csc_matrix(data)


def coherence(topics, z, data, n_words=20):
    csc = _csc(data)
    csc.sort_indices()
    n_docs_per_word = np.array((data > 0).sum(axis=0)).squeeze()
    return _coherence(topics, z, n_words, csc.indices, csc.indptr, n_docs_per_word)


def mean_coherence(topics, data, n_words=20):
    csc = _csc(data)
    csc.sort_indices()
    n_docs_per_word = np.array((data > 0).sum(axis=0)).squeeze()
    return float(np.mean([_coherence(topics, z, n_words, csc.indices, csc.indptr,
                                     n_docs_per_word) for z in range(topics.shape[0])]))
