"""GPU: one fit with its documents sharded over several GPUs (include/plsa_b200.h
plsa_set_shard).  The one-rank tests run the sharded code path (separate column-sum kernels,
packed log-likelihood mailbox, collectives as no-ops) on a single device; the two-rank test
needs two GPUs and goes over NCCL."""
import numpy as np
import pytest
from conftest import rel_l2

from enstop_b200 import PLSA, _lib, plsa, synth
from oracle import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def corpus():
    return synth.make_corpus(3000, 2500, 160_000, seed=21, planted=True, k_true=8)


def _fit_ctx(X, k, n_iter, seed, comm=None, tolerance=0.0, per_test=10):
    rng = np.random.RandomState(seed)
    pzd0, pwz0 = plsa.plsa_init(X, k, "random", rng)
    with _lib.Context(0) as ctx:
        ctx.upload_csr(X)
        if comm is not None:
            ctx.set_shard(comm)
        ctx.set_factors(pzd0.astype(np.float32), pwz0.astype(np.float32))
        ctx.set_sample_weight(None)
        iters, trace = ctx.em(n_iter, per_test, tolerance)
        pzd, pwz = ctx.get_factors()
        if comm is not None:
            ctx.set_shard(None)
    return pzd, pwz, iters, trace


@pytest.mark.parametrize("k", [8, 20, 40])
def test_one_rank_shard_path_matches_plain_fit(corpus, k):
    X = corpus
    comm = _lib.Comm(0, 1, 0, None)
    a = _fit_ctx(X, k, 25, 5)
    b = _fit_ctx(X, k, 25, 5, comm=comm)
    comm.close()
    assert a[2] == b[2] == 25
    assert rel_l2(b[0], a[0]) < 2e-6 and rel_l2(b[1], a[1]) < 2e-6
    assert np.allclose(a[3], b[3], rtol=1e-7)
    sw = np.ones(X.shape[0], dtype=np.float32)
    ref_pzd, ref_pwz = oracle.plsa_fit(X, k, sw, n_iter=25, tolerance=0.0, random_state=5,
                                       precision="f64")
    assert rel_l2(b[1], ref_pwz) < 1e-5 and rel_l2(b[0], ref_pzd) < 1e-5


def test_one_rank_shard_path_early_stop(corpus):
    """Same stop decision and trace through the packed {log-likelihood, flag} mailbox."""
    X = corpus
    comm = _lib.Comm(0, 1, 0, None)
    a = _fit_ctx(X, 8, 200, 9, tolerance=1e-4, per_test=5)
    b = _fit_ctx(X, 8, 200, 9, comm=comm, tolerance=1e-4, per_test=5)
    comm.close()
    assert a[2] == b[2] and a[2] < 200
    assert len(a[3]) == len(b[3]) and np.allclose(a[3], b[3], rtol=1e-7)


@pytest.mark.skipif(_lib.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("k,weighted,p2p", [(10, False, True), (20, True, True), (20, False, False),
                                            (128, False, True)])
def test_two_gpu_sharded_fit_matches_oracle(corpus, k, weighted, p2p, monkeypatch):
    """p2p: the ranks' P(w|z) sums are added by shard_reduce_kernel over NVLink peer memory;
    otherwise by ncclAllReduce."""
    monkeypatch.setenv("ENSTOP_B200_P2P", "1" if p2p else "0")
    X = corpus
    n = X.shape[0]
    sw = np.ones(n, dtype=np.float32)
    if weighted:
        sw = (0.5 + np.random.RandomState(3).rand(n)).astype(np.float32)
    kw = dict(n_iter=30, n_iter_per_test=10, tolerance=0.0, random_state=13)
    pzd, pwz, info = plsa.plsa_fit(X, k, sw, devices=[0, 1], return_info=True, **kw)
    one_pzd, one_pwz = plsa.plsa_fit(X, k, sw, device=0, **kw)
    ref_pzd, ref_pwz = oracle.plsa_fit(X, k, sw, precision="f64", **kw)
    assert info["n_iter"] == 30 and len(info["shard_bounds"]) == 3 and info["p2p"] == p2p
    assert pzd.shape == (n, k) and pwz.shape == (k, X.shape[1])
    assert rel_l2(pwz, ref_pwz) < 1e-5 and rel_l2(pzd, ref_pzd) < 1e-5
    assert rel_l2(pwz, one_pwz) < 5e-6 and rel_l2(pzd, one_pzd) < 5e-6
    # early stop: every rank takes the same decision from the summed log-likelihood
    kw2 = dict(n_iter=300, n_iter_per_test=5, tolerance=1e-4, random_state=13)
    _, _, i2 = plsa.plsa_fit(X, k, sw, devices=[0, 1], return_info=True, **kw2)
    _, _, i1 = plsa.plsa_fit(X, k, sw, device=0, return_info=True, **kw2)
    assert i2["n_iter"] == i1["n_iter"] < 300
    assert np.allclose(i2["ll_trace"], i1["ll_trace"], rtol=1e-7)


@pytest.mark.skipif(_lib.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("k", [10, 20, 128])
def test_two_shot_exchange_is_the_one_shot_exchange(corpus, k, monkeypatch):
    """Both exchanges add the ranks' partial sums in rank order (shard_reduce_kernel /
    shard_slice_reduce_kernel + shard_slice_gather_kernel): same bits, same trace."""
    X = corpus
    sw = np.ones(X.shape[0], dtype=np.float32)
    kw = dict(n_iter=25, n_iter_per_test=5, tolerance=0.0, random_state=5)
    out = []
    for two_shot in ("0", "1"):
        monkeypatch.setenv("ENSTOP_B200_TWO_SHOT", two_shot)
        out.append(plsa.plsa_fit(X, k, sw, devices=[0, 1], return_info=True, **kw))
    (pzd0, pwz0, i0), (pzd1, pwz1, i1) = out
    assert i0["p2p"] and i1["p2p"]
    assert np.array_equal(pzd0, pzd1) and np.array_equal(pwz0, pwz1)
    assert np.array_equal(np.asarray(i0["ll_trace"]), np.asarray(i1["ll_trace"]))


@pytest.mark.skipif(_lib.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_estimator(corpus):
    X = corpus
    m2 = PLSA(n_components=8, n_iter=20, tolerance=0.0, random_state=1, devices=[0, 1]).fit(X)
    m1 = PLSA(n_components=8, n_iter=20, tolerance=0.0, random_state=1, device=0).fit(X)
    assert rel_l2(m2.components_, m1.components_) < 5e-6
    assert rel_l2(m2.embedding_, m1.embedding_) < 5e-6
