"""Time the all-pairs topic distances at the C4 ensemble size (16 members x 20 topics,
50k terms) on the GPU and with the numpy restatement (oracle) on the host."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from enstop_b200 import _lib
from oracle import oracle
rng = np.random.RandomState(0)
topics = rng.dirichlet(np.full(50_000, 0.05), size=320).astype(np.float32)
for kind in ("hellinger", "kl"):
    _lib.topic_distances(topics[:32], kind)
    t0 = time.perf_counter(); D = _lib.topic_distances(topics, kind); dt = time.perf_counter() - t0
    kms = _lib.last_distances_ms()
    t0 = time.perf_counter()
    R = oracle.all_pairs_hellinger_distance(topics) if kind == "hellinger" else oracle.all_pairs_kl_divergence(topics)
    dh = time.perf_counter() - t0
    err = np.abs(D - R).max()
    print("%s 320x50000: GPU %.1f ms wall (incl. 64 MB upload; kernels %.2f ms), host numpy %.1f ms, max abs diff %.2e" % (kind, dt * 1e3, kms, dh * 1e3, err))

# the same on a stack that is already resident on the GPU (what the ensemble does after its gather)
import scipy.sparse as sp
m, k, members = 50_000, 20, 16
X = sp.random(64, m, density=0.01, format="csr", random_state=0, dtype=np.float32)
X.data[:] = 1.0
with _lib.Context(0) as ctx:
    ctx.upload_csr(X)
    pzd = np.full((64, k), 1.0 / k, dtype=np.float32)
    for r in range(members):
        ctx.set_factors(pzd, topics[r * k:(r + 1) * k])
        ctx.stash_topics(r, members)
    stacked = _lib.gather_topics([ctx], [members])
    assert np.array_equal(stacked, topics)
    for kind in ("hellinger", "kl"):
        for _ in range(3):   # the GPU idled while numpy computed the host reference: let it clock up
            _lib.gathered_distances(ctx, kind)
        walls, kernels = [], []
        for _ in range(5):
            t0 = time.perf_counter(); D = _lib.gathered_distances(ctx, kind); walls.append((time.perf_counter() - t0) * 1e3)
            kernels.append(_lib.last_distances_ms())
        assert np.array_equal(D, _lib.topic_distances(topics, kind))
        print("%s 320x50000 from the resident stack: wall ms %s, kernels ms %s" % (
            kind, [round(w, 2) for w in walls], [round(w, 2) for w in kernels]))
