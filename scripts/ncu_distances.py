"""All-pairs Hellinger and KL at the C4 ensemble size (320 topics x 50 000 terms) through
plsa_topic_distances, for an ncu capture of topic_pairs_kernel (no host reference here)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from enstop_b200 import _lib
rng = np.random.RandomState(0)
topics = rng.dirichlet(np.full(50_000, 0.05), size=320).astype(np.float32)
for kind in ("hellinger", "kl"):
    for _ in range(3):
        _lib.topic_distances(topics, kind)
    print(kind, "kernels ms", round(_lib.last_distances_ms(), 3))
