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
    t0 = time.perf_counter()
    R = oracle.all_pairs_hellinger_distance(topics) if kind == "hellinger" else oracle.all_pairs_kl_divergence(topics)
    dh = time.perf_counter() - t0
    err = np.abs(D - R).max()
    print("%s 320x50000: GPU %.1f ms (incl. 64 MB upload), host numpy %.1f ms, max abs diff %.2e" % (kind, dt * 1e3, dh * 1e3, err))
