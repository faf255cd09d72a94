"""scripts/fit_phases.py on every rank of a torchrun launch at the same time (the ranks share
the host's cores and memory): phases of PLSA(n_iter=20).fit at C2, pinned inputs."""
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch.distributed as dist
from sklearn.utils import check_random_state
from enstop_b200 import _lib, plsa, synth

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = int(os.environ["LOCAL_RANK"])
dist.init_process_group("gloo")
X = _lib.pinned_csr(synth.make_config("C2"))
k, n_iter = 20, 20
if rank == 0:
    print("cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)),
          "LOCAL_WORLD_SIZE", os.environ.get("LOCAL_WORLD_SIZE"), flush=True)
plsa.PLSA(n_components=k, n_iter=n_iter, tolerance=0.0, random_state=42, device=dev).fit(X)
ms = lambda a, b: round(1e3 * (b - a), 2)
for rep in range(4):
    dist.barrier()
    t0 = time.perf_counter()
    ctx = _lib.acquire_context(dev)
    stamps = {}

    def stage():
        a = time.perf_counter()
        ctx.set_option("presort", 1)
        ctx.upload_csr(X)
        b = time.perf_counter()
        ctx.prepare(k, False)
        c = time.perf_counter()
        stamps.update(upload=(a, b), prepare=(b, c))

    th = threading.Thread(target=stage)
    th.start()
    a = time.perf_counter()
    rng = check_random_state(42 + rank)
    p, w = plsa._random_init_f32(X.shape[0], X.shape[1], k, rng, ctx=ctx)
    b = time.perf_counter()
    th.join()
    c = time.perf_counter()
    ctx.set_factors(p, w)
    ctx.set_sample_weight(None)
    d = time.perf_counter()
    ctx.em(n_iter, 10, 0.0, 1e-32)
    e = time.perf_counter()
    ctx.get_factors()
    f = time.perf_counter()
    _lib.release_context(ctx)
    rec = {"rank": rank, "init draw": ms(a, b), "upload done at": ms(t0, stamps["upload"][1]),
           "prepare": ms(*stamps["prepare"]), "join at": ms(t0, c), "set_factors": ms(c, d),
           "em": ms(d, e), "get_factors": ms(e, f), "total": ms(t0, f)}
    out = [None] * world
    dist.all_gather_object(out, rec)
    if rank == 0 and rep > 0:
        for r in out:
            if r["rank"] in (0, world // 2, world - 1):
                print("rep", rep, r, flush=True)
        print("rep", rep, "max total over ranks", max(r["total"] for r in out), flush=True)
