"""Ensemble topic modelling on B200s: the host-side mirror of enstop/enstop_.py.

``plsa_topics`` (enstop_.py:56-115), ``ensemble_of_topics`` (:164-231), the topic combiners
(:266-414), ``ensemble_fit`` (:417-584) and ``EnsembleTopics`` (:587-927) keep their names,
arguments and defaults.  What changes is where the work runs:

* every ensemble member (bootstrap resample + pLSA fit) runs on a GPU; members are sharded
  round-robin over the visible devices, one host thread and one resident copy of the corpus
  per device (the reference fans the same members out over dask/joblib *threads*,
  enstop_.py:209-217);
* the bootstrap resample is a device-side row gather of the resident CSR (the index vector
  is still drawn on the host from the member's RandomState, enstop_.py:86-87);
* members' P(w|z) stay on their device until one NCCL gather stacks them on device 0
  (``np.vstack(topics)``, enstop_.py:231); clustering stays on the host; the final
  document-vector refit (enstop_.py:565-570) runs on device 0.

Documented deviations:
* the reference hands the *same* ``random_state`` to every member, so an int seed makes all
  n_starts runs identical; here member r gets seed ``RandomState(random_state).randint(...)[r]``
  — member r equals the reference's ``plsa_topics(X, k, random_state=seed_r)``;
* ``hdbscan`` / ``umap`` are the reference's un-pinned third-party dependencies and are not
  required: clustering uses ``sklearn.cluster.HDBSCAN`` (same algorithm, ``leaf`` selection);
  ``"hellinger_umap"`` uses ``umap`` when importable and otherwise falls back, with a
  warning, to ``"hellinger"`` on the exact distance matrix;
* ``model="nmf"`` (sklearn's solver, enstop_.py:118-161, broken on current sklearn) is not
  offered;
* ``EnsembleTopics.transform`` passes unit sample weights (the reference call at
  enstop_.py:847-854 omits the argument and raises TypeError).
"""
import threading
from warnings import warn

import numpy as np
from scipy.sparse import csr_matrix, issparse
from sklearn.base import BaseEstimator, TransformerMixin
from sklearn.utils import check_array, check_random_state

from . import _lib
from .plsa import default_device, plsa_fit, plsa_refit
from .utils import (_check_sample_weight, coherence, log_lift, mean_coherence, mean_log_lift,
                    normalize)

_MAX_SEED = 2 ** 31 - 1


def member_seeds(random_state, n_runs):
    """One int seed per ensemble member, deterministic in ``random_state``."""
    rng = check_random_state(random_state)
    return [int(s) for s in rng.randint(0, _MAX_SEED, size=n_runs)]


def bootstrap_indices(n_docs, random_state):
    """enstop_.py:85-87: ``rng.randint(0, n, size=n)`` from the member's RandomState."""
    rng = check_random_state(random_state)
    return rng.randint(0, n_docs, size=n_docs)


def plsa_topics(X, k, **kwargs):
    """Bootstrap-resample the corpus and fit pLSA to it; returns P(w|z) [k, n_words]
    (enstop_.py:56-115).  Extra kwargs: ``device``, ``context`` (a resident corpus)."""
    context = kwargs.get("context", None)
    owned = context is None
    if owned:
        device = kwargs.get("device", None)
        context = _lib.Context(default_device() if device is None else device)
        context.upload_csr(X.tocsr())
    try:
        n, m = X.shape
        init = kwargs.get("init", "random")
        fit_matrix = None
        if kwargs.get("bootstrap", True):
            idx = bootstrap_indices(n, kwargs.get("random_state", None))
            context.bootstrap(idx)
            if isinstance(init, str) and init != "random":
                fit_matrix = X.tocsr()[idx]      # SVD / NMF starts need the resampled matrix
        else:
            context.bootstrap(None)
            fit_matrix = X
        n_fit = context.shape[0]
        if fit_matrix is None:
            fit_matrix = _Shape((n_fit, m))       # a random start only needs the shape
        sample_weight = np.ones(n_fit, dtype=np.float32)
        _, topic_vocab = plsa_fit(
            fit_matrix, k, sample_weight,
            init=init,
            n_iter=kwargs.get("n_iter", 100),
            n_iter_per_test=kwargs.get("n_iter_per_test", 10),
            tolerance=kwargs.get("tolerance", 0.001),
            e_step_thresh=kwargs.get("e_step_thresh", 1e-16),
            random_state=kwargs.get("random_state", None),
            context=context, download=kwargs.get("download", True))
    finally:
        if owned:
            context.close()
    return topic_vocab


class _Shape:
    """Stands in for the bootstrapped matrix where only its shape is needed (random init)."""

    def __init__(self, shape):
        self.shape = shape


def shard_members(n_runs, n_shards):
    """Member r runs on shard r mod n_shards (a shard = one GPU / one rank)."""
    return [[r for r in range(n_runs) if r % n_shards == i] for i in range(n_shards)]


def stack_in_member_order(stacked, shards, k):
    """Rows of a shard-major gather ([shard 0's members..., shard 1's members..., ...], k
    rows each) re-ordered to member order — the np.vstack order of enstop_.py:231."""
    order = [r for members in shards for r in members]
    m = stacked.shape[1]
    blocks = stacked.reshape(len(order), k, m)
    out = np.empty_like(stacked)
    for pos, r in enumerate(order):
        out[r * k:(r + 1) * k] = blocks[pos]
    return out


def distances_in_member_order(dist, shards, k):
    """The same re-ordering for an all-pairs matrix computed on the shard-major stack."""
    order = [r for members in shards for r in members]
    perm = np.empty(len(order) * k, dtype=np.int64)       # perm[new row] = old row
    for pos, r in enumerate(order):
        perm[r * k:(r + 1) * k] = np.arange(pos * k, (pos + 1) * k)
    return dist[np.ix_(perm, perm)]


def resolve_devices(devices=None, n_jobs=None):
    count = _lib.device_count()
    if count < 1:
        raise _lib.PlsaError("no CUDA device visible; enstop_b200 has no CPU fallback")
    if devices is None:
        devices = list(range(count))
    devices = [int(d) for d in devices]
    if n_jobs is not None and n_jobs > 0:
        devices = devices[: max(1, min(len(devices), int(n_jobs)))]
    return devices


def fit_members_on_device(X, k, device, members, seeds, **kwargs):
    """Fit the ensemble members ``members`` (indices into ``seeds``) on one GPU and leave their
    P(w|z) stashed on it.  Two contexts ("lanes") per device, each with its own host thread and
    stream: while one lane's member is in its EM loop the other draws its bootstrap and seeded
    start on the host and rebuilds its device structures, so the GPU does not idle between
    members.  Returns (context holding the device's stash, members in stash order, every
    context used); the caller gathers, then hands the contexts to ``release_member_contexts``.
    This is the per-GPU unit of enstop_.py:209-217 — one call per device in a one-process run
    (``ensemble_of_topics``), one call per rank in a one-process-per-GPU run (bench.py)."""
    members = list(members)
    lanes = [members[0::2], members[1::2]] if len(members) > 1 else [members]
    contexts, errors = [None] * len(lanes), []

    def worker(lane):
        try:
            ctx = _lib.acquire_context(device)   # pooled: creating and destroying a context
            contexts[lane] = ctx                 # (pinned staging, ~20 device buffers) costs more
                                                 # than several members
            ctx.set_option("presort", 0)         # the members sort their bootstrap samples,
            ctx.upload_csr(X)                    # not this base corpus
            for slot, r in enumerate(lanes[lane]):
                kw = dict(kwargs)
                kw["random_state"] = seeds[r]
                kw["context"] = ctx
                kw["download"] = False           # the topics stay on the device until the gather
                plsa_topics(X, k, **kw)
                ctx.stash_topics(slot, len(lanes[lane]))
        except Exception as exc:  # surfaced after join
            errors.append(exc)

    threads = [threading.Thread(target=worker, args=(lane,)) for lane in range(len(lanes))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    used = [c for c in contexts if c is not None]
    if errors:
        for c in used:
            c.close()
        raise errors[0]
    if len(lanes) > 1:      # one stash per device: lane 0's members, then lane 1's
        _lib.stash_append(contexts[0], contexts[1], len(lanes[0]), len(lanes[1]))
    return contexts[0], [r for lane in lanes for r in lane], used


def release_member_contexts(contexts, failed=False):
    for ctx in contexts:
        if failed:
            ctx.close()
        else:
            ctx.bootstrap(None)
            _lib.release_context(ctx)


def ensemble_of_topics(X, k, model="plsa", n_jobs=4, n_runs=16, parallelism="threads",
                       devices=None, return_seeds=False, return_distances=None, **kwargs):
    """Topics of ``n_runs`` bootstrapped pLSA fits stacked as [n_runs * k, n_words]
    (enstop_.py:164-231).  Members are sharded over ``devices`` (default: all visible GPUs,
    at most ``n_jobs`` of them), member r on device r mod G.  ``return_distances``
    ("hellinger" / "kl"): also the all-pairs matrix of enstop_.py:234-263, computed on the
    gathered stack while it is still on the root GPU; the result is then a tuple."""
    if model != "plsa":
        raise ValueError('Model must be "plsa" (the sklearn NMF alternative is not offered)')
    if parallelism not in ("dask", "joblib", "threads", "none"):
        raise ValueError("Unrecognized parallelism {}; should be one of {}".format(
            parallelism, ("dask", "joblib")))
    X = X.tocsr() if issparse(X) else csr_matrix(X)
    devices = resolve_devices(devices, None if parallelism == "none" else n_jobs)
    if parallelism == "none":
        devices = devices[:1]
    seeds = member_seeds(kwargs.get("random_state", None), n_runs)
    assign = dict(zip(devices, shard_members(n_runs, len(devices))))
    used = [d for d in devices if assign[d]]
    results, errors = {}, []

    def worker(dev):
        try:
            results[dev] = fit_members_on_device(X, k, dev, assign[dev], seeds, **kwargs)
        except Exception as exc:  # surfaced after join
            errors.append(exc)

    threads = [threading.Thread(target=worker, args=(d,)) for d in used]
    if len(used) > 1:   # NCCL communicators of the final gather, created meanwhile
        threads.append(threading.Thread(target=lambda: _lib.gather_warmup(used)))
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    all_contexts = [c for res in results.values() for c in res[2]]
    try:
        if errors:
            raise errors[0]
        stacked = _lib.gather_topics([results[d][0] for d in used],
                                     [len(assign[d]) for d in used])
        dist = None
        if return_distances is not None:
            dist = _lib.gathered_distances(results[used[0]][0], return_distances)
    except BaseException:
        release_member_contexts(all_contexts, failed=True)
        raise
    release_member_contexts(all_contexts)
    orders = [results[d][1] for d in used]
    out = stack_in_member_order(stacked, orders, k)
    extra = []
    if return_seeds:
        extra.append(seeds)
    if return_distances is not None:
        extra.append(distances_in_member_order(dist, orders, k))
    return (out, *extra) if extra else out


# ---- distances between topics (enstop_.py:234-263) ------------------------------------------
def all_pairs_kl_divergence(distributions, device=None):
    """result[i, j] = sum_w a log2(a / b) over entries where both are > 0 (enstop_.py:234-250),
    evaluated on the GPU (csrc: topic_pairs_kernel<1>); float64 [N, N]."""
    return _lib.topic_distances(distributions, "kl", default_device() if device is None else device)


def all_pairs_hellinger_distance(distributions, device=None):
    """sqrt(1 - sum_w sqrt(a b) / sqrt(|a|_1 |b|_1)) (umap.distances.hellinger, as used at
    enstop_.py:253-263), evaluated on the GPU in the cancellation-free form
    sqrt(1/2 sum_w (sqrt(a/|a|) - sqrt(b/|b|))^2) (csrc: topic_pairs_kernel<0>)."""
    return _lib.topic_distances(distributions, "hellinger",
                                default_device() if device is None else device)


def _combine(all_topics, labels, weights=None):
    """Cluster representative: mean of sqrt(topic), squared, renormalised
    (enstop_.py:311-312, 348-349, 399-405)."""
    n_clusters = int(labels.max()) + 1 if labels.size else 0
    result = np.empty((max(n_clusters, 0), all_topics.shape[1]), dtype=np.float32)
    for i in range(n_clusters):
        mask = labels == i
        w = None if weights is None else weights[mask]
        if w is not None and not np.any(w > 0):
            w = None
        result[i] = np.average(np.sqrt(all_topics[mask]), axis=0, weights=w) ** 2
        result[i] /= result[i].sum()
    return result


def _hdbscan(**kw):
    from sklearn.cluster import HDBSCAN
    return HDBSCAN(cluster_selection_method="leaf", copy=True, **kw)


def generate_combined_topics_kl(all_topics, min_samples=5, min_cluster_size=5, distances=None):
    """enstop_.py:266-314: mutual reachability from the asymmetric KL matrix with the
    min_samples-th neighbour as core divergence, single linkage, leaf clusters.
    ``distances``: a precomputed all-pairs matrix (otherwise computed on the GPU)."""
    div = all_pairs_kl_divergence(all_topics) if distances is None else np.asarray(distances)
    core = np.sort(div, axis=1)[:, min(min_samples, div.shape[0] - 1)]
    tiled = np.tile(core, (core.shape[0], 1))
    mreach = np.dstack([div, div.T, tiled, tiled.T]).max(axis=-1)
    np.fill_diagonal(mreach, 0.0)
    # min_samples=1: core distance 0, so the supplied matrix is used as it stands
    labels = _hdbscan(min_samples=1, min_cluster_size=min_cluster_size,
                      metric="precomputed").fit_predict(mreach)
    return _combine(all_topics, labels)


def generate_combined_topics_hellinger(all_topics, min_samples=5, min_cluster_size=5,
                                       distances=None):
    """enstop_.py:317-351.  ``distances``: a precomputed all-pairs matrix (otherwise computed
    on the GPU)."""
    dist = all_pairs_hellinger_distance(all_topics) if distances is None else np.asarray(distances)
    labels = _hdbscan(min_samples=min_samples, min_cluster_size=min_cluster_size,
                      metric="precomputed").fit_predict(dist)
    return _combine(all_topics, labels)


def generate_combined_topics_hellinger_umap(all_topics, min_samples=5, min_cluster_size=5,
                                            n_neighbors=15, reduced_dim=5):
    """enstop_.py:354-407: UMAP(hellinger) to 5-d, HDBSCAN(leaf, allow_single_cluster),
    membership-weighted mean of sqrt(topic).  Needs ``umap``; without it the exact Hellinger
    matrix is clustered directly (a documented deviation)."""
    try:
        import umap
    except ImportError:
        warn("umap is not installed: topic_combination='hellinger_umap' falls back to "
             "'hellinger' (HDBSCAN on the exact Hellinger distance matrix)")
        return generate_combined_topics_hellinger(all_topics, min_samples, min_cluster_size)
    embedding = umap.UMAP(n_neighbors=n_neighbors, n_components=reduced_dim,
                          metric="hellinger").fit_transform(all_topics)
    clusterer = _hdbscan(min_samples=min_samples, min_cluster_size=min_cluster_size,
                         allow_single_cluster=True).fit(embedding)
    return _combine(all_topics, clusterer.labels_, clusterer.probabilities_)


_topic_combiner = {
    "kl_divergence": generate_combined_topics_kl,
    "hellinger": generate_combined_topics_hellinger,
    "hellinger_umap": generate_combined_topics_hellinger_umap,
}


def ensemble_fit(X, estimated_n_topics=10, model="plsa", init="random", min_samples=3,
                 min_cluster_size=4, n_starts=16, n_jobs=1, parallelism="threads",
                 topic_combination="hellinger_umap", bootstrap=True, n_iter=100,
                 n_iter_per_test=10, tolerance=0.001, e_step_thresh=1e-16, lift_factor=1,
                 beta_loss=1, alpha=0.0, solver="mu", random_state=None, devices=None,
                 return_all_topics=False):
    """Stable topics from an ensemble of bootstrapped pLSA fits and the documents'
    P(z|d) against them (enstop_.py:417-584).  Returns (doc_vectors, stable_topics)."""
    if topic_combination not in _topic_combiner:
        raise ValueError("topic_combination must be one of {}".format(
            tuple(_topic_combiner.keys())))
    X = check_array(X, accept_sparse="csr", dtype=np.float32)
    if not issparse(X):
        X = csr_matrix(X, dtype=np.float32)
    # the combiner's distance matrix comes from the stack while it is resident on the root GPU
    have_umap = True
    if topic_combination == "hellinger_umap":
        try:
            import umap  # noqa: F401
        except ImportError:
            have_umap = False
    kind = {"kl_divergence": "kl", "hellinger": "hellinger",
            "hellinger_umap": None if have_umap else "hellinger"}[topic_combination]
    res = ensemble_of_topics(
        X, estimated_n_topics, model, n_jobs, n_starts, parallelism, devices=devices,
        init=init, n_iter=n_iter, n_iter_per_test=n_iter_per_test, tolerance=tolerance,
        e_step_thresh=e_step_thresh, bootstrap=bootstrap, random_state=random_state,
        return_distances=kind)
    all_topics, dist = res if kind is not None else (res, None)
    if topic_combination == "hellinger_umap" and have_umap:
        stable_topics = generate_combined_topics_hellinger_umap(all_topics, min_samples, min_cluster_size)
    elif topic_combination == "hellinger_umap":
        warn("umap is not installed: topic_combination='hellinger_umap' falls back to "
             "'hellinger' (HDBSCAN on the exact Hellinger distance matrix)")
        stable_topics = generate_combined_topics_hellinger(all_topics, min_samples, min_cluster_size,
                                                           distances=dist)
    else:
        stable_topics = _topic_combiner[topic_combination](all_topics, min_samples, min_cluster_size,
                                                           distances=dist)
    if stable_topics.shape[0] == 0:
        raise ValueError("no stable topic cluster was found; lower min_cluster_size or "
                         "min_samples, or raise n_starts")
    if lift_factor != 1:
        stable_topics **= lift_factor
        normalize(stable_topics, axis=1)
    sample_weight = _check_sample_weight(None, X, dtype=np.float32)
    dev = resolve_devices(devices)[0]
    doc_vectors = plsa_refit(X, stable_topics, sample_weight, e_step_thresh=e_step_thresh,
                             random_state=random_state, device=dev)
    if return_all_topics:
        return doc_vectors, stable_topics, all_topics
    return doc_vectors, stable_topics


class EnsembleTopics(BaseEstimator, TransformerMixin):
    """Ensemble Topic Modelling (EnsTop), sklearn-style (mirrors enstop_.py:587-927).

    Constructor arguments and defaults are the reference's (enstop_.py:709-730):
    ``n_components=10, model="plsa", init="random", n_starts=16, min_samples=3,
    min_cluster_size=5, n_jobs=8, parallelism="dask", topic_combination="hellinger_umap",
    bootstrap=True, n_iter=80, n_iter_per_test=10, tolerance=0.001, e_step_thresh=1e-32,
    lift_factor=1, beta_loss=1, alpha=0.0, solver="mu", transform_random_seed=42,
    random_state=None``; ``n_jobs`` bounds the number of GPUs used, ``parallelism`` is
    accepted for compatibility (members always run as one host thread per GPU), and
    ``devices`` (extra) pins the CUDA ordinals.

    Attributes: ``components_`` (stable topics, [n_components_, n_words]), ``embedding_``
    (P(z|d) [n_docs, n_components_]), ``training_data_``, ``n_components_``.

    Clustering stage: the reference clusters with ``hdbscan`` and, for the default
    ``topic_combination="hellinger_umap"``, embeds the topics with ``umap`` first.  Here
    HDBSCAN is ``sklearn.cluster.HDBSCAN`` and, when ``umap`` cannot be imported, the default
    falls back — with a warning — to ``"hellinger"`` (HDBSCAN on the exact all-pairs Hellinger
    matrix, computed on the GPU): ``n_components_`` and ``components_`` can then differ from
    what the reference's UMAP route would select.  The members' topics and the
    membership-weighted cluster representatives are the reference's.
    """

    def __init__(self, n_components=10, model="plsa", init="random", n_starts=16, min_samples=3,
                 min_cluster_size=5, n_jobs=8, parallelism="dask",
                 topic_combination="hellinger_umap", bootstrap=True, n_iter=80,
                 n_iter_per_test=10, tolerance=0.001, e_step_thresh=1e-32, lift_factor=1,
                 beta_loss=1, alpha=0.0, solver="mu", transform_random_seed=42,
                 random_state=None, devices=None):
        self.n_components = n_components
        self.model = model
        self.init = init
        self.n_starts = n_starts
        self.min_samples = min_samples
        self.min_cluster_size = min_cluster_size
        self.n_jobs = n_jobs
        self.parallelism = parallelism
        self.topic_combination = topic_combination
        self.bootstrap = bootstrap
        self.n_iter = n_iter
        self.n_iter_per_test = n_iter_per_test
        self.tolerance = tolerance
        self.e_step_thresh = e_step_thresh
        self.lift_factor = lift_factor
        self.beta_loss = beta_loss
        self.alpha = alpha
        self.solver = solver
        self.transform_random_seed = transform_random_seed
        self.random_state = random_state
        self.devices = devices

    def fit(self, X, y=None):
        self.fit_transform(X)
        return self

    def fit_transform(self, X, y=None, **fit_params):
        X = check_array(X, accept_sparse="csr")
        if not issparse(X):
            X = csr_matrix(X)
        U, V = ensemble_fit(
            X, self.n_components, self.model, self.init, self.min_samples,
            self.min_cluster_size, self.n_starts, self.n_jobs, self.parallelism,
            self.topic_combination, self.bootstrap, self.n_iter, self.n_iter_per_test,
            self.tolerance, self.e_step_thresh, self.lift_factor, self.beta_loss, self.alpha,
            self.solver, self.random_state, devices=self.devices)
        self.components_ = V
        self.embedding_ = U
        self.training_data_ = X
        self.n_components_ = self.components_.shape[0]
        return U

    def transform(self, X, y=None):
        X = check_array(X, accept_sparse="csr")
        random_state = check_random_state(self.transform_random_seed)
        if not issparse(X):
            X = csr_matrix(X)
        sample_weight = _check_sample_weight(None, X, dtype=np.float32)
        dev = resolve_devices(self.devices)[0]
        return plsa_refit(X, self.components_, sample_weight, n_iter=50, n_iter_per_test=5,
                          tolerance=0.001, random_state=random_state, device=dev)

    def _score(self, one, mean, topic_num, n_words):
        if not isinstance(topic_num, int) and topic_num is not None:
            raise ValueError("Topic number must be an integer or None.")
        if topic_num is None:
            return mean(self.components_, self.training_data_, n_words=n_words)
        if 0 <= topic_num < self.n_components:
            return one(self.components_, topic_num, self.training_data_, n_words=n_words)
        raise ValueError("Topic number must be in range 0 to {}".format(self.n_components))

    def coherence(self, topic_num=None, n_words=20):
        return self._score(coherence, mean_coherence, topic_num, n_words)

    def log_lift(self, topic_num=None, n_words=20):
        return self._score(log_lift, mean_log_lift, topic_num, n_words)
