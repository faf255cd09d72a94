"""pLSA on a B200: the host-side mirror of enstop/plsa.py.

Same names, arguments, defaults and error behaviour as the reference —
``plsa_init`` (plsa.py:412-513), ``plsa_fit`` (plsa.py:643-730), ``plsa_refit``
(plsa.py:923-997) and the sklearn-style ``PLSA`` estimator (plsa.py:1000-1285) — with the EM
loop (``plsa_fit_inner`` / ``plsa_refit_inner``, plsa.py:516-640 / :819-920) carried out by
hand-written sm_100a kernels behind the C ABI in ``include/plsa_b200.h``.  Host code is
plain numpy/scipy + ctypes; initialisation stays on the host so that a seed produces the
reference's exact ``RandomState`` stream (plsa.py:455-456).

Extra keyword-only arguments (``device``, ``context``) and fitted attributes (``n_iter_``,
``log_likelihood_trace_``) are additive; positional use is unchanged.
"""
import os
import threading

import numpy as np
from scipy.sparse import csr_matrix, issparse
from sklearn.base import BaseEstimator, TransformerMixin
from sklearn.decomposition import non_negative_factorization
from sklearn.utils import check_array, check_random_state
from sklearn.utils.extmath import randomized_svd

from . import _lib
from .utils import (_check_sample_weight, coherence, log_lift, mean_coherence, mean_log_lift,
                    normalize, standardize_input)


def default_device():
    return int(os.environ.get("ENSTOP_B200_DEVICE", "0"))


def _l2(x):
    return np.sqrt(np.sum(np.square(x)))


def plsa_init(X, k, init="random", rng=np.random):
    """Initial P(z|d) [n, k] and P(w|z) [k, m], float64, L1 row-normalised.

    ``"random"`` draws P(w|z) first and P(z|d) second from ``rng.rand`` (plsa.py:454-456);
    ``"nndsvd"`` is the non-negative double SVD start of sklearn's NMF built on
    ``randomized_svd`` (plsa.py:458-493); ``"nmf"`` runs sklearn's coordinate-descent NMF
    (plsa.py:495-504); a tuple/list is taken as given (plsa.py:505-506)."""
    n, m = X.shape
    if isinstance(init, str) and init == "random":
        p_w_given_z = rng.rand(k, m)
        p_z_given_d = rng.rand(n, k)
    elif isinstance(init, str) and init == "nndsvd":
        U, S, V = randomized_svd(X, k)
        p_z_given_d, p_w_given_z = np.zeros(U.shape), np.zeros(V.shape)
        p_z_given_d[:, 0] = np.sqrt(S[0]) * np.abs(U[:, 0])
        p_w_given_z[0, :] = np.sqrt(S[0]) * np.abs(V[0, :])
        for j in range(1, k):
            x, y = U[:, j], V[j, :]
            x_pos, y_pos = np.maximum(x, 0), np.maximum(y, 0)
            x_neg, y_neg = np.abs(np.minimum(x, 0)), np.abs(np.minimum(y, 0))
            pos = (_l2(x_pos), _l2(y_pos))
            neg = (_l2(x_neg), _l2(y_neg))
            if pos[0] * pos[1] > neg[0] * neg[1]:
                u, v, sigma = x_pos / pos[0], y_pos / pos[1], pos[0] * pos[1]
            else:
                u, v, sigma = x_neg / neg[0], y_neg / neg[1], neg[0] * neg[1]
            scale = np.sqrt(S[j] * sigma)
            p_z_given_d[:, j] = scale * u
            p_w_given_z[j, :] = scale * v
    elif isinstance(init, str) and init == "nmf":
        p_z_given_d, p_w_given_z, _ = non_negative_factorization(
            X, n_components=k, init="nndsvd", solver="cd", beta_loss=2, tol=1e-2, max_iter=100)
    elif isinstance(init, (tuple, list)):
        p_z_given_d, p_w_given_z = init
        p_z_given_d = np.array(p_z_given_d, dtype=np.float64)
        p_w_given_z = np.array(p_w_given_z, dtype=np.float64)
        if p_z_given_d.shape != (n, k) or p_w_given_z.shape != (k, m):
            raise ValueError("init arrays must have shapes ({}, {}) and ({}, {})".format(
                n, k, k, m))
    else:
        raise ValueError("Unrecognized init {}".format(init))
    normalize(p_w_given_z, axis=1)
    normalize(p_z_given_d, axis=1)
    return p_z_given_d, p_w_given_z


def _random_init_f32(n, m, k, rng, ctx=None):
    """The "random" start of plsa_init + the float32 cast of plsa_fit (plsa.py:454-456,
    510-511, 709-710) in one pass over the RandomState's own MT19937 stream
    (csrc/host_init.cpp): bit-identical factors, a quarter of the host time.  With ``ctx`` the
    factors are drawn straight into the context's page-locked staging (views, valid until the
    context's next fit).  Returns None when ``rng`` is not a legacy RandomState (then numpy
    does it)."""
    try:
        out_pzd = out_pwz = None
        if ctx is not None and n > 0 and m > 0:
            out_pzd, out_pwz = ctx.pinned_factors(n, m, k)
        p_w_given_z = _lib.random_rows(rng, k, m, out=out_pwz)
        if p_w_given_z is None:
            return None
        p_z_given_d = _lib.random_rows(rng, n, k, out=out_pzd)
    except _lib.PlsaError:
        return None
    return p_z_given_d, p_w_given_z


def _as_csr(X):
    if not issparse(X):
        X = csr_matrix(X)
    elif X.format != "csr":
        X = X.tocsr()
    return X


class _Staging:
    """Uploads the corpus and builds its device-side structures on a helper thread while the
    caller draws the seeded initial factors on the host (ctypes releases the GIL)."""

    def __init__(self, X, k, device, context, refit):
        self.owned = context is None
        self.error = None
        self.thread = None
        if self.owned:
            self.ctx = _lib.acquire_context(default_device() if device is None else device)
            self.thread = threading.Thread(target=self._run, args=(_as_csr(X), k, refit))
            self.thread.start()
        else:
            self.ctx = context

    def _run(self, X, k, refit):
        try:
            # a full fit needs the term-major copy: its sort starts while the values upload
            self.ctx.set_option("presort", 0 if refit else 1)
            self.ctx.upload_csr(X)
            self.ctx.prepare(k, refit)
        except BaseException as exc:  # re-raised by wait()
            self.error = exc

    def wait(self):
        if self.thread is not None:
            self.thread.join()
            self.thread = None
        if self.error is not None:
            raise self.error
        return self.ctx

    def close(self):
        if self.thread is not None:
            self.thread.join()
            self.thread = None
        if self.owned and self.ctx is not None:
            if self.error is None:
                _lib.release_context(self.ctx)   # buffers stay cached for the next fit
            else:
                self.ctx.close()
            self.ctx = None


def shard_rows(indptr, n_shards):
    """Row boundaries [b_0 = 0, ..., b_G = n] of G contiguous document shards with (nearly)
    equal numbers of stored entries and at least one document each."""
    indptr = np.asarray(indptr)
    n = len(indptr) - 1
    n_shards = int(n_shards)
    if n_shards < 1 or n < n_shards:
        raise ValueError("cannot cut {} documents into {} shards".format(n, n_shards))
    targets = indptr[-1] * np.arange(1, n_shards, dtype=np.float64) / n_shards
    cuts = np.searchsorted(indptr, targets, side="left")
    bounds = [0]
    for g, c in enumerate(cuts, start=1):     # strictly increasing, room left for the rest
        bounds.append(int(min(max(c, bounds[-1] + 1), n - (n_shards - g))))
    bounds.append(n)
    return bounds


_shard_comms = {}            # tuple(devices) -> [Comm per rank], reused by later sharded fits
_shard_comm_lock = threading.Lock()


def release_shard_communicators():
    """Destroy the NCCL communicators kept for sharded fits."""
    with _shard_comm_lock:
        held = list(_shard_comms.values())
        _shard_comms.clear()
    for comms in held:
        for c in comms:
            c.close()


class ThreadPeerExchange:
    """Peer-memory setup between the ranks of a sharded fit that live in ONE process (one
    host thread per GPU): every rank prepares its exchange block, the device addresses are
    swapped through this object, every rank attaches its peers.  If any rank cannot (no
    peer access between two of the GPUs), all ranks fall back to the NCCL all-reduce."""

    def __init__(self, devices, timeout=300.0):
        self.devices = list(devices)
        self.barrier = threading.Barrier(len(self.devices), timeout=timeout)
        self.base = [0] * len(self.devices)
        self.ok = [True] * len(self.devices)

    def __call__(self, ctx, rank):
        try:
            self.base[rank] = ctx.shard_p2p_prepare()[0]
        except _lib.PlsaError:
            self.ok[rank] = False
        self.barrier.wait()
        if all(self.ok):
            try:
                for p, dev in enumerate(self.devices):
                    if p != rank:
                        ctx.shard_p2p_attach(p, dev, base=self.base[p])
            except _lib.PlsaError:
                self.ok[rank] = False
        self.barrier.wait()
        if not all(self.ok):
            ctx.set_option("p2p", 0)
        return all(self.ok)

    def finish(self, ctx):
        """Every rank unmaps its peers' blocks before any rank frees its own."""
        ctx.shard_p2p_detach()
        self.barrier.wait()

    def abort(self):
        self.barrier.abort()


def plsa_fit_shard(X_rows, k, p_z_given_d_rows, p_w_given_z, sample_weight_rows, comm, device,
                   n_iter=100, n_iter_per_test=10, tolerance=0.001, e_step_thresh=1e-32,
                   use_sample_weights=False, profile=False, exchange=None):
    """One rank of a document-sharded fit: this rank's rows of X and of P(z|d), the full
    P(w|z), the shard communicator (``_lib.Comm``).  Every rank calls it with the same scalar
    arguments.  ``exchange(ctx, rank)`` (optional) sets up the peer-memory all-reduce.
    Returns (P(z|d) rows, full P(w|z), info)."""
    ctx = _lib.acquire_context(device)   # pooled: device buffers survive between fits
    ok = False
    try:
        ctx.set_option("p2p", 1)
        # exchange shape: -1 = two-shot (reduce a slice, gather the slices) from 4 ranks up
        ctx.set_option("p2p_two_shot", int(os.environ.get("ENSTOP_B200_TWO_SHOT", "-1")))
        ctx.set_option("presort", 1)
        ctx.upload_csr(_as_csr(X_rows))
        ctx.set_shard(comm)
        ctx.set_factors(np.ascontiguousarray(p_z_given_d_rows, dtype=np.float32),
                        np.ascontiguousarray(p_w_given_z, dtype=np.float32))
        ctx.set_sample_weight(sample_weight_rows if use_sample_weights else None)
        ctx.prepare(k, False)    # sort, work items, module load: before the ranks line up
        p2p = bool(exchange(ctx, comm.rank)) if exchange is not None else False
        if profile:
            ctx.set_profiling(True)
        iters, trace = ctx.em(n_iter, n_iter_per_test, tolerance, e_step_thresh, refit=False,
                              use_sample_weights=use_sample_weights)
        pzd, pwz = ctx.get_factors()
        info = {"n_iter": iters, "ll_trace": trace, "em_ms": ctx.last_em_ms,
                "launches": ctx.launches, "profile": ctx.profile() if profile else None,
                "p2p": p2p}
        if profile:
            ctx.set_profiling(False)
        if p2p and hasattr(exchange, "finish"):
            exchange.finish(ctx)
        ctx.set_shard(None)
        ok = True
    finally:
        if ok:
            _lib.release_context(ctx)
        else:
            ctx.close()
    return pzd, pwz, info


def _plsa_fit_sharded(X, k, sample_weight, init, n_iter, n_iter_per_test, tolerance,
                      e_step_thresh, random_state, devices, p2p=True):
    """plsa_fit over several GPUs of one box: one host thread, context and NCCL rank per
    device, documents cut into contiguous shards of equal stored entries."""
    X = _as_csr(X)
    n, m = X.shape
    rng = check_random_state(random_state)
    fast = _random_init_f32(n, m, k, rng) if isinstance(init, str) and init == "random" else None
    if fast is not None:
        p_z_given_d, p_w_given_z = fast
    else:
        p_z_given_d, p_w_given_z = plsa_init(X, k, init=init, rng=rng)
        p_z_given_d = p_z_given_d.astype(np.float32, order="C")
        p_w_given_z = p_w_given_z.astype(np.float32, order="C")
    sample_weight = np.asarray(sample_weight, dtype=np.float32)
    use_sw = bool(np.any(sample_weight != 1.0))
    G = len(devices)
    bounds = shard_rows(X.indptr, G)
    key = tuple(devices)
    with _shard_comm_lock:
        comms = _shard_comms.pop(key, None)   # communicators are kept between fits
    uid = _lib.Comm.unique_id() if comms is None else None
    comms = comms or [None] * G
    results, errors = [None] * G, [None] * G
    exchange = ThreadPeerExchange(devices) if p2p else None
    # Every rank's communicator exists before any rank touches its shard: a rank that fails later
    # (upload, allocation) can then abort the others instead of leaving them in ncclCommInitRank.
    n_dev = _lib.device_count()
    bad = [d for d in devices if not 0 <= d < n_dev]
    if bad or len(set(devices)) != G:
        raise ValueError("devices must be distinct CUDA ordinals below {}: {}".format(n_dev, devices))
    if any(c is None for c in comms):
        def make_comm(r):
            try:
                comms[r] = _lib.Comm(devices[r], G, r, uid)
            except BaseException as exc:
                errors[r] = exc
        makers = [threading.Thread(target=make_comm, args=(r,)) for r in range(G)]
        for t in makers:
            t.start()
        for t in makers:
            t.join()
        if any(e is not None for e in errors):
            for c in comms:
                if c is not None:
                    c.abort()
                    c.close()
            raise next(e for e in errors if e is not None)

    def worker(r):
        try:
            comm = comms[r]
            lo, hi = bounds[r], bounds[r + 1]
            results[r] = plsa_fit_shard(X[lo:hi], k, p_z_given_d[lo:hi], p_w_given_z,
                                        sample_weight[lo:hi], comm, devices[r], n_iter,
                                        n_iter_per_test, tolerance, e_step_thresh, use_sw,
                                        exchange=exchange)
        except BaseException as exc:
            errors[r] = exc
            if exchange is not None:
                exchange.abort()
            for c in comms:          # peers blocked in a collective with this rank return
                if c is not None:
                    c.abort()

    threads = [threading.Thread(target=worker, args=(r,)) for r in range(G)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if any(e is not None for e in errors):
        for c in comms:
            if c is not None:
                c.close()
        raise next(e for e in errors if e is not None)
    with _shard_comm_lock:
        _shard_comms[key] = comms
    iters = {res[2]["n_iter"] for res in results}
    if len(iters) != 1:
        raise _lib.PlsaError("sharded fit: ranks disagree on the iteration count {}".format(iters))
    info = dict(results[0][2])
    info["em_ms"] = max(res[2]["em_ms"] for res in results)
    info["launches"] = sum(res[2]["launches"] for res in results)
    info["shard_bounds"] = bounds
    return np.concatenate([res[0] for res in results]), results[0][1], info


def plsa_fit(X, k, sample_weight, init="random", n_iter=100, n_iter_per_test=10,
             tolerance=0.001, e_step_thresh=1e-32, random_state=None, *, device=None,
             context=None, return_info=False, devices=None, download=True, precheck=None):
    """Fit pLSA with ``k`` topics; returns ``(p_z_given_d [n,k], p_w_given_z [k,m])`` float32.

    Drop-in for enstop.plsa.plsa_fit (plsa.py:643-730).  ``context`` (an
    ``enstop_b200._lib.Context`` whose resident corpus is X) skips the upload — used by the
    ensemble for its bootstrapped members.  ``devices`` (a list of two or more CUDA ordinals)
    shards the documents of this one fit over those GPUs.  ``download=False`` leaves the
    factors on the device (the ensemble stashes P(w|z) there) and returns (None, None)."""
    if devices is not None and len(devices) > 1:
        if precheck is not None:
            precheck()
        out = _plsa_fit_sharded(X, k, sample_weight, init, n_iter, n_iter_per_test, tolerance,
                                e_step_thresh, random_state, [int(d) for d in devices],
                                p2p=os.environ.get("ENSTOP_B200_P2P", "1") != "0")
        return out if return_info else out[:2]
    if devices is not None and len(devices) == 1 and device is None:
        device = int(devices[0])
    staging = _Staging(X, k, device, context, refit=False)
    try:
        rng = check_random_state(random_state)
        fast = _random_init_f32(X.shape[0], X.shape[1], k, rng, ctx=staging.ctx) \
            if isinstance(init, str) and init == "random" else None
        if fast is not None:
            p_z_given_d, p_w_given_z = fast
        else:
            p_z_given_d, p_w_given_z = plsa_init(X, k, init=init, rng=rng)
            p_z_given_d = p_z_given_d.astype(np.float32, order="C")
            p_w_given_z = p_w_given_z.astype(np.float32, order="C")
        sample_weight = np.asarray(sample_weight, dtype=np.float32)
        use_sample_weights = bool(np.any(sample_weight != 1.0))  # plsa.py:712
        if precheck is not None:
            precheck()      # input validation that ran beside the staging; may raise
        ctx = staging.wait()
        ctx.set_factors(p_z_given_d, p_w_given_z)
        ctx.set_sample_weight(sample_weight if use_sample_weights else None)
        iters, trace = ctx.em(n_iter, n_iter_per_test, tolerance, e_step_thresh, refit=False,
                              use_sample_weights=use_sample_weights)
        p_z_given_d, p_w_given_z = ctx.get_factors() if download else (None, None)
        info = {"n_iter": iters, "ll_trace": trace, "em_ms": ctx.last_em_ms,
                "launches": ctx.launches}
    finally:
        staging.close()
    if return_info:
        return p_z_given_d, p_w_given_z, info
    return p_z_given_d, p_w_given_z


def plsa_refit(X, topics, sample_weight, n_iter=50, n_iter_per_test=10, tolerance=0.005,
               e_step_thresh=1e-32, random_state=None, *, device=None, context=None,
               return_info=False):
    """Estimate P(z|d) for documents X against frozen ``topics`` [k, m].

    Drop-in for enstop.plsa.plsa_refit (plsa.py:923-997): fresh ``rng.rand(n, k)`` start,
    E-step + P(z|d)-only M-step; as in the reference the loop always runs ``n_iter``
    iterations (its early stop is guarded by ``LL > 0``, plsa.py:913)."""
    topics = np.ascontiguousarray(topics, dtype=np.float32)
    k = topics.shape[0]
    staging = _Staging(X, k, device, context, refit=True)
    try:
        rng = check_random_state(random_state)
        p_z_given_d = _lib.random_rows(rng, X.shape[0], k)  # plsa.py:979-981, fast path
        if p_z_given_d is None:
            p_z_given_d = rng.rand(X.shape[0], k)
            normalize(p_z_given_d, axis=1)
            p_z_given_d = p_z_given_d.astype(np.float32)
        sample_weight = np.asarray(sample_weight, dtype=np.float32)
        ctx = staging.wait()
        ctx.set_factors(p_z_given_d, topics)
        ctx.set_sample_weight(None if not np.any(sample_weight != 1.0) else sample_weight)
        iters, _ = ctx.em(n_iter, n_iter_per_test, tolerance, e_step_thresh, refit=True)
        p_z_given_d, _ = ctx.get_factors(want_pwz=False)
        info = {"n_iter": iters, "em_ms": ctx.last_em_ms, "launches": ctx.launches}
    finally:
        staging.close()
    if return_info:
        return p_z_given_d, info
    return p_z_given_d


class _StoredZeros(Exception):
    """The matrix stores explicit zeros: empty rows must be found from the row sums."""


class _ValueCheck:
    """min over the stored values on a helper thread (numpy releases the GIL in reductions)."""

    ASYNC_FROM = 2_000_000   # stored entries; below this the check costs less than a thread

    def __init__(self, X):
        self.min = 0
        self.thread = None
        if X.nnz >= self.ASYNC_FROM:
            self.thread = threading.Thread(target=self._run, args=(X.data,))
            self.thread.start()
        elif X.nnz:
            self.min = X.data.min()

    def _run(self, data):
        self.min = data.min()

    def result(self):
        if self.thread is not None:
            self.thread.join()
            self.thread = None
        if self.min < 0:
            raise ValueError("PLSA is only valid for matrices with non-negative entries")
        if self.min == 0:
            raise _StoredZeros()


class PLSA(BaseEstimator, TransformerMixin):
    """Probabilistic Latent Semantic Analysis, sklearn-style (mirrors plsa.py:1000-1285).

    Parameters are the reference's (plsa.py:1074-1084): ``n_components=10``,
    ``init="random"`` (``"random"``, ``"nndsvd"``, ``"nmf"`` or a tuple of arrays),
    ``n_iter=100``, ``n_iter_per_test=10``, ``tolerance=0.001``, ``e_step_thresh=1e-32``,
    ``transform_random_seed=42``, ``random_state=None``; plus ``device`` (CUDA ordinal,
    default ``$ENSTOP_B200_DEVICE`` or 0) and ``devices`` (two or more ordinals: the
    documents of the fit are sharded over those GPUs, P(w|z) summed over NVLink once per EM
    iteration).

    Attributes: ``components_`` (P(w|z), [n_topics, n_words] float32), ``embedding_``
    (P(z|d), [n_docs, n_topics]), ``training_data_``; additionally ``n_iter_`` and
    ``log_likelihood_trace_`` (the values the early-stop test saw).
    """

    def __init__(self, n_components=10, init="random", n_iter=100, n_iter_per_test=10,
                 tolerance=0.001, e_step_thresh=1e-32, transform_random_seed=42,
                 random_state=None, device=None, devices=None):
        self.n_components = n_components
        self.init = init
        self.n_iter = n_iter
        self.n_iter_per_test = n_iter_per_test
        self.tolerance = tolerance
        self.e_step_thresh = e_step_thresh
        self.transform_random_seed = transform_random_seed
        self.random_state = random_state
        self.device = device
        self.devices = devices

    def fit(self, X, y=None, sample_weight=None):
        self.fit_transform(X, sample_weight=sample_weight)
        return self

    def fit_transform(self, X, y=None, sample_weight=None):
        X = check_array(X, accept_sparse="csr")
        X = standardize_input(X)
        if not issparse(X):
            X = csr_matrix(X)
        sample_weight = _check_sample_weight(sample_weight, X, dtype=np.float32)
        # The sign check of plsa.py:1146-1149 reads every stored value (2 ms at 10 M entries): it
        # runs on a helper thread while the fit stages the corpus and draws its start, and is
        # joined before the first EM iteration.  Rows are taken as empty iff they store
        # nothing, which is exact unless explicit zeros are stored; then the row sums of
        # plsa.py:1151-1153 decide and the fit is redone on the stripped matrix.
        check = _ValueCheck(X)
        good_rows = np.diff(X.indptr) != 0
        try:
            if check.thread is None:      # small input: checked on the spot
                check.result()
                check = None
            U, V, info, good_rows = self._fit_rows(X, sample_weight, good_rows, check)
        except _StoredZeros:
            row_sums = np.array(X.sum(axis=1).T)[0]
            U, V, info, good_rows = self._fit_rows(X, sample_weight, row_sums != 0, None)
        zero_rows_found = not np.all(good_rows)
        if zero_rows_found:
            self.embedding_ = np.zeros((X.shape[0], self.n_components))
            self.embedding_[good_rows] = U
        else:
            self.embedding_ = U
        self.components_ = V
        self.training_data_ = X
        self.n_iter_ = info["n_iter"]
        self.log_likelihood_trace_ = info["ll_trace"]
        return self.embedding_

    def _fit_rows(self, X, sample_weight, good_rows, check):
        if not np.all(good_rows):
            data_for_fitting = X[good_rows]
            # plsa.py:1144 vs :1156-1164 leaves the weights unaligned with the stripped
            # matrix; the weights of the kept rows are what is meant
            sample_weight = sample_weight[good_rows]
        else:
            data_for_fitting = X
        U, V, info = plsa_fit(data_for_fitting, self.n_components, sample_weight, self.init,
                              self.n_iter, self.n_iter_per_test, self.tolerance,
                              self.e_step_thresh, self.random_state, device=self.device,
                              return_info=True, devices=self.devices,
                              precheck=check.result if check is not None else None)
        return U, V, info, good_rows

    def transform(self, X, y=None):
        X = check_array(X, accept_sparse="csr")
        random_state = check_random_state(self.transform_random_seed)
        sample_weight = _check_sample_weight(None, X, dtype=np.float32)
        if not issparse(X):
            X = csr_matrix(X)
        return plsa_refit(X, self.components_, sample_weight, n_iter=50, n_iter_per_test=5,
                          tolerance=0.001, random_state=random_state, device=self.device)

    def coherence(self, topic_num=None, n_words=20):
        if not isinstance(topic_num, int) and topic_num is not None:
            raise ValueError("Topic number must be an integer or None.")
        if topic_num is None:
            return mean_coherence(self.components_, self.training_data_, n_words)
        if 0 <= topic_num < self.n_components:
            return coherence(self.components_, topic_num, self.training_data_, n_words)
        raise ValueError("Topic number must be in range 0 to {}".format(self.n_components))

    def log_lift(self, topic_num=None, n_words=20):
        if not isinstance(topic_num, int) and topic_num is not None:
            raise ValueError("Topic number must be an integer or None.")
        if topic_num is None:
            return mean_log_lift(self.components_, self.training_data_, n_words)
        if 0 <= topic_num < self.n_components:
            return log_lift(self.components_, topic_num, self.training_data_, n_words)
        raise ValueError("Topic number must be in range 0 to {}".format(self.n_components))
