"""ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the shipped product path.

ctypes front end to ``libplsa_oracle.so`` (the C restatement of enstop/plsa.py, see
``plsa_oracle.c``) plus the host-side glue of the reference restated in numpy:

* ``plsa_init_random``  — plsa.py:451-456,510-511 (two ``rng.rand`` draws, P(w|z) first,
  then float64 L1 row normalisation, utils.py:22-41)
* ``plsa_fit``          — plsa.py:707-730
* ``plsa_refit``        — plsa.py:975-997
* ``all_pairs_kl_divergence`` / ``all_pairs_hellinger_distance`` — enstop_.py:234-263 (+ the
  body of umap.distances.hellinger, which the reference file quotes at :30-46), float64 numpy

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import this module.  ``enstop_b200`` never does.

Parity status: PINNED — checked against outputs of the reference itself
(tests/golden/make_golden.py ran /root/reference/enstop/plsa.py in the build container,
tests/golden/make_golden_distances.py compiled the reference's own all-pairs functions;
tests/test_oracle_golden.py compares).
"""
import ctypes
import os
import subprocess

import numpy as np
from sklearn.utils import check_random_state

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libplsa_oracle.so")
_lib = None

_i32p = ctypes.POINTER(ctypes.c_int32)
_f32p = ctypes.POINTER(ctypes.c_float)
_f64p = ctypes.POINTER(ctypes.c_double)
_i64 = ctypes.c_int64


def build(force=False):
    """Compile the C restatement with the system gcc (oracle/Makefile)."""
    src_mtime = max(
        os.path.getmtime(os.path.join(_HERE, f))
        for f in ("plsa_oracle.c", "plsa_oracle_impl.h", "Makefile")
    )
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < src_mtime:
        env = dict(os.environ)
        env.pop("CC", None)
        subprocess.run(["make", "-C", _HERE, "-B", "libplsa_oracle.so"], check=True,
                       env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    return _SO


def lib():
    global _lib
    if _lib is not None:
        return _lib
    try:
        build()
        L = ctypes.CDLL(_SO)
    except (OSError, subprocess.CalledProcessError):
        build(force=True)
        L = ctypes.CDLL(_SO)

    L.oracle_e_step_f32.argtypes = [_i32p, _i32p, _i64, _f32p, _f32p, _f32p, _i64, _i64,
                                    ctypes.c_float]
    L.oracle_e_step_f32.restype = None
    L.oracle_m_step_f32.argtypes = [_i32p, _i32p, _f32p, _i64, _f32p, _f32p, _f32p, _f32p,
                                    _f32p, _f32p, _i64, _i64, _i64]
    L.oracle_m_step_f32.restype = None
    L.oracle_log_likelihood_f32.argtypes = [_i32p, _i32p, _f32p, _i64, _f32p, _f32p, _f32p,
                                            _i64, _i64]
    L.oracle_log_likelihood_f32.restype = ctypes.c_float
    L.oracle_log_likelihood_f64.argtypes = [_i32p, _i32p, _f64p, _i64, _f64p, _f64p, _f64p,
                                            _i64, _i64]
    L.oracle_log_likelihood_f64.restype = ctypes.c_double
    L.oracle_fit_inner_f32.argtypes = [_i32p, _i32p, _f32p, _i64, _f32p, _f32p, _f32p, _i64,
                                       _i64, _i64, _i64, _i64, ctypes.c_double,
                                       ctypes.c_float, ctypes.c_int, _f64p, _i64,
                                       ctypes.POINTER(_i64)]
    L.oracle_fit_inner_f32.restype = _i64
    L.oracle_fit_inner_f64.argtypes = [_i32p, _i32p, _f64p, _i64, _f64p, _f64p, _f64p, _i64,
                                       _i64, _i64, _i64, _i64, ctypes.c_double,
                                       ctypes.c_double, ctypes.c_int, _f64p, _i64,
                                       ctypes.POINTER(_i64)]
    L.oracle_fit_inner_f64.restype = _i64
    L.oracle_refit_inner_f32.argtypes = [_i32p, _i32p, _f32p, _i64, _f32p, _f32p, _f32p,
                                         _i64, _i64, _i64, _i64, _i64, ctypes.c_double,
                                         ctypes.c_float]
    L.oracle_refit_inner_f32.restype = _i64
    L.oracle_refit_inner_f64.argtypes = [_i32p, _i32p, _f64p, _i64, _f64p, _f64p, _f64p,
                                         _i64, _i64, _i64, _i64, _i64, ctypes.c_double,
                                         ctypes.c_double]
    L.oracle_refit_inner_f64.restype = _i64
    L.oracle_normalize_rows_f64.argtypes = [_f64p, _i64, _i64]
    L.oracle_normalize_rows_f64.restype = None
    L.oracle_set_num_threads.argtypes = [ctypes.c_int]
    L.oracle_set_num_threads.restype = ctypes.c_int
    _lib = L
    return L


def _p(a, ct):
    return a.ctypes.data_as(ct)


def set_num_threads(n=0):
    """Threads used by the parallel (prange) loops; 0 = all cores.  Returns the count."""
    return int(lib().oracle_set_num_threads(int(n)))


def _coo(X, dtype):
    """plsa.py:714 — ``A = X.tocoo().astype(np.float32)`` (row-major when X is CSR)."""
    A = X.tocoo()
    rows = np.ascontiguousarray(A.row, dtype=np.int32)
    cols = np.ascontiguousarray(A.col, dtype=np.int32)
    vals = np.ascontiguousarray(A.data.astype(np.float32), dtype=dtype)
    return rows, cols, vals


def normalize_rows(a):
    """utils.py:22-41 with axis=1 (in place, float64)."""
    assert a.dtype == np.float64 and a.flags.c_contiguous
    lib().oracle_normalize_rows_f64(_p(a, _f64p), a.shape[0], a.shape[1])
    return a


def plsa_init_random(n, m, k, rng):
    """plsa.py:454-456 + 510-511: P(w|z) is drawn first, then P(z|d); both float64 and
    L1 row-normalised.  Returns (p_z_given_d, p_w_given_z) like plsa_init."""
    p_w_given_z = rng.rand(k, m)
    p_z_given_d = rng.rand(n, k)
    normalize_rows(p_w_given_z)
    normalize_rows(p_z_given_d)
    return p_z_given_d, p_w_given_z


def fit_inner(rows, cols, vals, pwz, pzd, sw, n_iter=100, n_iter_per_test=10,
              tolerance=0.001, e_step_thresh=1e-32, use_sample_weights=False,
              precision="f32"):
    """plsa.py:516-640.  Mutates pwz / pzd in place; returns (iters_run, ll_trace)."""
    L = lib()
    k, m = pwz.shape
    n = pzd.shape[0]
    cap = n_iter // max(1, n_iter_per_test) + 3
    trace = np.zeros(cap, dtype=np.float64)
    n_ll = _i64(0)
    if precision == "f32":
        fn, fp, dt = L.oracle_fit_inner_f32, _f32p, np.float32
    else:
        fn, fp, dt = L.oracle_fit_inner_f64, _f64p, np.float64
    for a in (vals, pwz, pzd, sw):
        assert a.dtype == dt and a.flags.c_contiguous
    iters = fn(_p(rows, _i32p), _p(cols, _i32p), _p(vals, fp), vals.shape[0], _p(pwz, fp),
               _p(pzd, fp), _p(sw, fp), n, m, k, n_iter, n_iter_per_test, float(tolerance),
               e_step_thresh, int(bool(use_sample_weights)), _p(trace, _f64p), cap,
               ctypes.byref(n_ll))
    if iters < 0:
        raise MemoryError("oracle: allocation of the nnz x k posterior failed")
    return int(iters), trace[: min(cap, n_ll.value)].copy()


def plsa_fit(X, k, sample_weight, init="random", n_iter=100, n_iter_per_test=10,
             tolerance=0.001, e_step_thresh=1e-32, random_state=None, precision="f32",
             return_info=False):
    """plsa.py:643-730 restated.  ``init`` is "random" or a (p_z_given_d, p_w_given_z)
    tuple (plsa.py:505-506; still normalised, plsa.py:510-511)."""
    rng = check_random_state(random_state)
    n, m = X.shape
    if isinstance(init, str):
        if init != "random":
            raise ValueError("oracle restates only init='random' and tuple init")
        pzd, pwz = plsa_init_random(n, m, k, rng)
    else:
        pzd = np.array(init[0], dtype=np.float64, order="C")
        pwz = np.array(init[1], dtype=np.float64, order="C")
        normalize_rows(pwz)
        normalize_rows(pzd)
    # plsa.py:709-710: the float32 cast happens before the EM loop in every precision
    pzd = pzd.astype(np.float32, order="C")
    pwz = pwz.astype(np.float32, order="C")
    sample_weight = np.asarray(sample_weight, dtype=np.float32)
    use_sw = bool(np.any(sample_weight != 1.0))  # plsa.py:712
    dt = np.float32 if precision == "f32" else np.float64
    rows, cols, vals = _coo(X, dt)
    pzd = np.ascontiguousarray(pzd, dtype=dt)
    pwz = np.ascontiguousarray(pwz, dtype=dt)
    sw = np.ascontiguousarray(sample_weight, dtype=dt)
    iters, trace = fit_inner(rows, cols, vals, pwz, pzd, sw, n_iter, n_iter_per_test,
                             tolerance, e_step_thresh, use_sw, precision)
    if return_info:
        return pzd, pwz, {"n_iter": iters, "ll_trace": trace}
    return pzd, pwz


def plsa_refit(X, topics, sample_weight, n_iter=50, n_iter_per_test=10, tolerance=0.005,
               e_step_thresh=1e-32, random_state=None, precision="f32"):
    """plsa.py:923-997 restated."""
    L = lib()
    dt = np.float32 if precision == "f32" else np.float64
    rows, cols, vals = _coo(X, dt)
    k = topics.shape[0]
    rng = check_random_state(random_state)
    pzd = rng.rand(X.shape[0], k)
    normalize_rows(pzd)
    pzd = np.ascontiguousarray(pzd.astype(np.float32), dtype=dt)
    tp = np.ascontiguousarray(np.asarray(topics).astype(np.float32), dtype=dt)
    sw = np.ascontiguousarray(np.asarray(sample_weight, dtype=np.float32), dtype=dt)
    if precision == "f32":
        fn, fp = L.oracle_refit_inner_f32, _f32p
    else:
        fn, fp = L.oracle_refit_inner_f64, _f64p
    iters = fn(_p(rows, _i32p), _p(cols, _i32p), _p(vals, fp), vals.shape[0], _p(tp, fp),
               _p(pzd, fp), _p(sw, fp), X.shape[0], X.shape[1], k, n_iter, n_iter_per_test,
               float(tolerance), e_step_thresh)
    if iters < 0:
        raise MemoryError("oracle: allocation of the nnz x k posterior failed")
    return pzd


def log_likelihood(X, pwz, pzd, sample_weight=None, precision="f64"):
    """plsa.py:329-386 on arbitrary factors (float64 reduction by default)."""
    L = lib()
    dt = np.float32 if precision == "f32" else np.float64
    rows, cols, vals = _coo(X, dt)
    if sample_weight is None:
        sample_weight = np.ones(X.shape[0])
    pwz = np.ascontiguousarray(pwz, dtype=dt)
    pzd = np.ascontiguousarray(pzd, dtype=dt)
    sw = np.ascontiguousarray(sample_weight, dtype=dt)
    if precision == "f32":
        return float(L.oracle_log_likelihood_f32(_p(rows, _i32p), _p(cols, _i32p),
                                                 _p(vals, _f32p), vals.shape[0],
                                                 _p(pwz, _f32p), _p(pzd, _f32p),
                                                 _p(sw, _f32p), pwz.shape[1], pwz.shape[0]))
    return float(L.oracle_log_likelihood_f64(_p(rows, _i32p), _p(cols, _i32p),
                                             _p(vals, _f64p), vals.shape[0], _p(pwz, _f64p),
                                             _p(pzd, _f64p), _p(sw, _f64p), pwz.shape[1],
                                             pwz.shape[0]))


def all_pairs_kl_divergence(distributions):
    """enstop_.py:234-250: result[i, j] = sum_w a log2(a / b) over entries where both are > 0."""
    P = np.asarray(distributions, dtype=np.float64)
    pos = P > 0
    L = np.zeros_like(P)
    L[pos] = np.log2(P[pos])
    n = P.shape[0]
    out = np.zeros((n, n))
    for i in range(n):
        both = pos & pos[i][None, :]
        out[i] = np.where(both, P[i][None, :] * (L[i][None, :] - L), 0.0).sum(axis=1)
    return out


def all_pairs_hellinger_distance(distributions):
    """enstop_.py:253-263 with umap.distances.hellinger (quoted at enstop_.py:30-46):
    sqrt(1 - sum_w sqrt(a b) / sqrt(|a|_1 |b|_1)); 0 for two all-zero rows, 1 when exactly
    one row is all-zero."""
    P = np.asarray(distributions, dtype=np.float64)
    R = np.sqrt(P)
    l1 = P.sum(axis=1)
    inner = R @ R.T
    denom = np.sqrt(np.outer(l1, l1))
    with np.errstate(divide="ignore", invalid="ignore"):
        d = np.sqrt(np.clip(1.0 - inner / denom, 0.0, None))
    zero = l1 == 0
    d[np.ix_(zero, ~zero)] = 1.0
    d[np.ix_(~zero, zero)] = 1.0
    d[np.ix_(zero, zero)] = 0.0
    return d
