"""GPU: the work-item launch orders of the row pass (plsa_set_option "item_order").

Orders 1 ("window") and 2 ("band") only permute the work items (tests/test_host_cpu.py checks the
plan on the CPU); every per-row sum is taken over the same entries in the same order, so the
factors must agree with order 0 to float rounding of the column sums / log-likelihood (whose
per-CTA partial sums are grouped differently) and with the float64 oracle at the parity
tolerance.  Runs last (file name) — it covers an option that is off by default.
"""
import numpy as np
import pytest
from conftest import rel_l2

from enstop_b200 import _lib, plsa, synth
from oracle import oracle

pytestmark = pytest.mark.gpu


def _fit(X, k, order, n_iter, chunk=0, sw=None, use_sw=False):
    n, m = X.shape
    rng = np.random.RandomState(7)
    pzd0, pwz0 = plsa.plsa_init(X, k, "random", rng)
    ctx = _lib.Context(0)
    try:
        ctx.set_option("item_order", order)
        if chunk:
            ctx.set_option("chunk", chunk)
        ctx.upload_csr(X)
        ctx.set_factors(pzd0.astype(np.float32), pwz0.astype(np.float32))
        ctx.set_sample_weight(sw)
        iters, trace = ctx.em(n_iter, n_iter_per_test=5, tolerance=0.0, use_sample_weights=use_sw)
        pzd, pwz = ctx.get_factors()
    finally:
        ctx.close()
    return pzd, pwz, iters, trace, (pzd0.astype(np.float32), pwz0.astype(np.float32))


@pytest.mark.parametrize("k,chunk", [(20, 32), (10, 64), (128, 0), (7, 32)])
def test_window_order_is_the_same_model(k, chunk):
    # planted corpus with rows and columns far longer than the work-item length
    X = synth.make_corpus(1500, 900, 120_000, seed=21, planted=True, k_true=6)
    a = _fit(X, k, 0, 5, chunk)
    b = _fit(X, k, 1, 5, chunk)
    c = _fit(X, k, 2, 5, chunk)
    assert a[2] == b[2] == c[2] == 5
    assert rel_l2(c[1], a[1]) < 5e-6 and rel_l2(c[0], a[0]) < 5e-6
    assert rel_l2(b[1], a[1]) < 5e-6 and rel_l2(b[0], a[0]) < 5e-6
    assert np.allclose(a[3], b[3], rtol=1e-9)
    sw = np.ones(X.shape[0], dtype=np.float32)
    ez, ew = oracle.plsa_fit(X, k, sw, init=a[4], n_iter=5, tolerance=0.0, precision="f64")
    assert rel_l2(b[1], ew) < 2e-5 and rel_l2(b[0], ez) < 2e-5


def test_window_order_with_sample_weights_and_bit_repeatable():
    X = synth.make_corpus(1200, 700, 90_000, seed=22, planted=True, k_true=5)
    sw = (0.5 + np.random.RandomState(3).rand(X.shape[0])).astype(np.float32)
    a = _fit(X, 12, 1, 8, 32, sw=sw, use_sw=True)
    b = _fit(X, 12, 1, 8, 32, sw=sw, use_sw=True)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    c = _fit(X, 12, 0, 8, 32, sw=sw, use_sw=True)
    assert rel_l2(a[1], c[1]) < 5e-6 and rel_l2(a[0], c[0]) < 5e-6


def test_failed_calls_leave_a_consistent_context():
    """A rejected bootstrap keeps the corpus and model the context had; a failed upload leaves
    the context without a corpus (not with half of one)."""
    import types
    X = synth.make_corpus(300, 200, 6_000, seed=4, planted=True, k_true=3)
    n, m = X.shape
    pzd0, pwz0 = plsa.plsa_init(X, 4, "random", np.random.RandomState(1))
    with _lib.Context(0) as ctx:
        ctx.upload_csr(X)
        ctx.set_factors(pzd0.astype(np.float32), pwz0.astype(np.float32))
        with pytest.raises(_lib.PlsaError, match="row index out of range"):
            ctx.bootstrap(np.array([0, n + 5, 2], dtype=np.int32))
        assert ctx.shape == (n, m, X.nnz)
        iters, _ = ctx.em(2, tolerance=0.0)              # corpus and factors are still there
        assert iters == 2
        ctx.bootstrap(np.arange(n - 1, -1, -1, dtype=np.int32))
        assert ctx.shape == (n, m, X.nnz)
        bad = types.SimpleNamespace(indptr=X.indptr, indices=np.where(
            np.arange(X.nnz) == 17, m + 3, X.indices).astype(np.int32), data=X.data, shape=X.shape)
        with pytest.raises(_lib.PlsaError, match="column index out of range"):
            ctx.upload_csr(bad)
        with pytest.raises(_lib.PlsaError, match="no corpus uploaded"):
            ctx.set_factors(pzd0.astype(np.float32), pwz0.astype(np.float32))
        ctx.upload_csr(X)                                 # and the context is reusable
        ctx.set_factors(pzd0.astype(np.float32), pwz0.astype(np.float32))
        assert ctx.em(1, tolerance=0.0)[0] == 1
