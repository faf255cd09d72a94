"""Seeded initial factors of C2 (20 x 50000, then 100000 x 20) drawn by csrc/host_init.cpp:
wall time and, with ENSTOP_B200_INIT_PROFILE=1, the time line of the slices."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from enstop_b200 import _lib
n, m, k = 100000, 50000, 20
for rep in range(4):
    rng = np.random.RandomState(42)
    t0 = time.perf_counter()
    w = _lib.random_rows(rng, k, m)
    t1 = time.perf_counter()
    p = _lib.random_rows(rng, n, k)
    t2 = time.perf_counter()
    print("threads %s ratio %s: pwz %.2f ms, pzd %.2f ms" % (os.environ.get("ENSTOP_B200_INIT_THREADS", "auto"),
          os.environ.get("ENSTOP_B200_INIT_RATIO", "default"), 1e3 * (t1 - t0), 1e3 * (t2 - t1)), flush=True)
