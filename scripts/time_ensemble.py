"""Wall time of the C4 config: EnsembleTopics(n_components=20, n_starts=16) on the C2 corpus,
on 1..G GPUs; phases of ensemble_fit (fan-out + gather, clustering, refit)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from enstop_b200 import EnsembleTopics, _lib, enstop_, plsa, synth

X = synth.make_config("C2", planted=True).astype(np.float32)   # planted topics: stable clusters exist
G = _lib.device_count()
kw = dict(n_iter=80, n_iter_per_test=10, tolerance=0.001, e_step_thresh=1e-32, random_state=42)
for g in sorted({1, G}):
    enstop_.ensemble_of_topics(X, 20, n_runs=g, n_jobs=g, **kw)            # warm
    t0 = time.perf_counter()
    topics = enstop_.ensemble_of_topics(X, 20, n_runs=16, n_jobs=g, **kw)
    t1 = time.perf_counter()
    stable = enstop_.generate_combined_topics_hellinger(topics, 3, 4)
    t2 = time.perf_counter()
    if stable.shape[0]:
        plsa.plsa_refit(X, stable, np.ones(X.shape[0], dtype=np.float32), e_step_thresh=1e-32,
                        random_state=42)
    t3 = time.perf_counter()
    print("GPUs %d: 16 members %.3f s (%.2e nnz*k*iters/s if all 80 iterations ran), clustering %.3f s "
          "(%d stable topics), refit %.3f s" % (g, t1 - t0, X.nnz * 20 * 80 * 16 / (t1 - t0), t2 - t1,
                                                 stable.shape[0], t3 - t2))
    t0 = time.perf_counter()
    model = EnsembleTopics(n_components=20, n_starts=16, n_jobs=g, topic_combination="hellinger",
                           random_state=42).fit(X)
    print("GPUs %d: EnsembleTopics.fit end to end %.3f s, %d topics" % (g, time.perf_counter() - t0,
                                                                         model.n_components_))
