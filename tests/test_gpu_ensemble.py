"""GPU: the ensemble path — members sharded over the visible devices (one host thread per
GPU), device-side bootstrap, NCCL gather of the stacked topics, host clustering, refit."""
import numpy as np
import pytest
from conftest import rel_l2

from enstop_b200 import EnsembleTopics, _lib, enstop_, plsa, synth
from oracle import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def corpus():
    return synth.make_corpus(1500, 2000, 90_000, seed=11, planted=True, k_true=6)


def test_member_equals_reference_plsa_topics(corpus):
    """Member r of the ensemble == the reference's plsa_topics(X, k, random_state=seed_r):
    bootstrap rows rng.randint(0, n, n) (enstop_.py:85-88), then plsa_fit with the same seed."""
    X = corpus
    k, n_runs = 6, 4
    kw = dict(n_iter=15, n_iter_per_test=5, tolerance=0.0, e_step_thresh=1e-32, random_state=7)
    stacked, seeds = enstop_.ensemble_of_topics(X, k, n_runs=n_runs, n_jobs=8, return_seeds=True,
                                                **kw)
    assert stacked.shape == (n_runs * k, X.shape[1]) and stacked.dtype == np.float32
    for r in (0, n_runs - 1):
        idx = np.random.RandomState(seeds[r]).randint(0, X.shape[0], size=X.shape[0])
        B = X[idx]
        _, ref = oracle.plsa_fit(B, k, np.ones(B.shape[0], dtype=np.float32), n_iter=15,
                                 n_iter_per_test=5, tolerance=0.0, random_state=seeds[r],
                                 precision="f64")
        assert rel_l2(stacked[r * k:(r + 1) * k], ref) < 1e-5
    # same seed -> same stack, whatever the number of devices used
    again = enstop_.ensemble_of_topics(X, k, n_runs=n_runs, n_jobs=1, **kw)
    assert np.array_equal(stacked, again)


@pytest.mark.skipif(_lib.device_count() < 2, reason="needs 2 GPUs")
def test_nccl_gather_across_devices(corpus):
    X = corpus
    k = 5
    sw = np.ones(X.shape[0], dtype=np.float32)
    ctxs = [_lib.Context(d) for d in (0, 1)]
    topics = []
    for i, ctx in enumerate(ctxs):
        ctx.upload_csr(X)
        _, t = plsa.plsa_fit(X, k, sw, n_iter=5, tolerance=0.0, random_state=i, context=ctx)
        ctx.stash_topics(0, 1)
        topics.append(t)
    stacked = _lib.gather_topics(ctxs, [1, 1])
    for ctx in ctxs:
        ctx.close()
    assert np.array_equal(stacked, np.vstack(topics))


def test_distances_of_the_gathered_stack(corpus):
    """The all-pairs matrix computed on the stack the gather left on the root GPU equals the one
    computed from the host copy (and the float64 oracle), re-ordered to member order."""
    X = corpus
    k, n_runs = 5, 5
    kw = dict(n_iter=8, n_iter_per_test=5, tolerance=0.0, e_step_thresh=1e-32, random_state=3)
    for kind, ref in (("hellinger", oracle.all_pairs_hellinger_distance),
                      ("kl", oracle.all_pairs_kl_divergence)):
        stacked, dist = enstop_.ensemble_of_topics(X, k, n_runs=n_runs, n_jobs=8,
                                                   return_distances=kind, **kw)
        assert dist.shape == (n_runs * k, n_runs * k)
        assert np.array_equal(dist, _lib.topic_distances(stacked, kind, 0))
        assert np.allclose(dist, ref(stacked), rtol=1e-5, atol=2e-6)


def test_ensemble_topics_estimator(corpus):
    X = corpus
    model = EnsembleTopics(n_components=6, n_starts=6, n_iter=30, topic_combination="hellinger",
                           min_samples=2, min_cluster_size=3, random_state=3)
    emb = model.fit_transform(X)
    assert model.components_.shape[1] == X.shape[1]
    assert model.n_components_ == model.components_.shape[0] >= 2
    assert emb.shape == (X.shape[0], model.n_components_)
    assert np.allclose(model.components_.sum(axis=1), 1.0, atol=1e-4)
    good = np.asarray(X.sum(axis=1)).ravel() != 0
    assert np.allclose(emb[good].sum(axis=1), 1.0, atol=1e-4)
    # the planted topics are recovered: every stable topic is close to one planted vocabulary
    t = model.transform(X[:25])
    assert t.shape == (25, model.n_components_)
    assert np.isfinite(model.coherence(n_words=10)) and np.isfinite(model.log_lift(n_words=10))
    with pytest.raises(ValueError, match="topic_combination"):
        EnsembleTopics(topic_combination="bogus").fit(X)


def test_topic_distances_kernel():
    """plsa_topic_distances (GPU) against the reference's golden values and the float64
    oracle, including exact zeros, an all-zero row, a duplicate row and un-normalised rows."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden",
                             "topic_distances.npz"))
    topics = g["topics"]
    H = enstop_.all_pairs_hellinger_distance(topics)
    K = enstop_.all_pairs_kl_divergence(topics)
    assert H.shape == K.shape == (25, 25) and H.dtype == np.float64
    assert np.allclose(H ** 2, g["hellinger"] ** 2, atol=1e-6)
    assert np.allclose(K, g["kl"], rtol=1e-5, atol=1e-5)
    assert H[11, 0] == 1.0 and H[0, 11] == 1.0 and H[11, 11] == 0.0 and H[12, 13] == 0.0
    # a larger, ragged case (sizes not multiples of the 32 x 32 tile or the 32-term block)
    rng = np.random.RandomState(5)
    big = rng.dirichlet(np.full(1013, 0.05), size=77).astype(np.float32)
    big[big < 1e-5] = 0.0
    Ho, Ko = oracle.all_pairs_hellinger_distance(big), oracle.all_pairs_kl_divergence(big)
    assert np.allclose(enstop_.all_pairs_hellinger_distance(big), Ho, atol=2e-6)
    assert np.allclose(enstop_.all_pairs_kl_divergence(big), Ko, rtol=1e-5, atol=1e-5)
