"""Worker for tests/test_multirank_cpu.py::test_sharded_em_two_ranks_gloo: the algebra of the
document-sharded fit (include/plsa_b200.h plsa_set_shard) on CPU, world size 2, gloo.  Every
rank owns a contiguous shard of the documents (plsa.shard_rows), does the E-step and the
P(z|d) update locally, and the ranks add their raw P(w|z) sums and log-likelihoods with an
all-reduce — exactly the exchange the GPU path makes over NCCL.  Checked against the
unsharded oracle (float64 restatement of enstop/plsa.py)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
from enstop_b200 import plsa, synth  # noqa: E402


def em_shard_step(Xd, U, V, thresh, allreduce):
    """plsa.py:91-105 + :182-202 for the rows of this shard; V (k x m) is global."""
    P = U[:, None, :] * V.T[None, :, :]                 # [rows, m, k] products
    P[P <= thresh] = 0.0
    norm = P.sum(axis=-1)
    post = np.divide(P, norm[..., None], out=np.zeros_like(P), where=norm[..., None] > 0)
    S = Xd[..., None] * post
    U_new = S.sum(axis=1)
    rs = U_new.sum(axis=1, keepdims=True)
    U_new = np.divide(U_new, rs, out=U_new.copy(), where=rs > 0)
    V_raw = allreduce(np.ascontiguousarray(S.sum(axis=0).T))   # the one exchange per iteration
    cs = V_raw.sum(axis=1, keepdims=True)
    V_new = np.divide(V_raw, cs, out=V_raw.copy(), where=cs > 0)
    return U_new, V_new


def main():
    import torch
    import torch.distributed as dist
    rank, world, _ = bench.dist_env()
    plumb = bench.Plumbing(rank, world)

    def allreduce(a):
        t = torch.from_numpy(a)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t.numpy()

    n, m, k, n_iter = 240, 150, 5, 12
    X = synth.make_corpus(n, m, 6000, seed=4, planted=True, k_true=4)
    bounds = plsa.shard_rows(X.indptr, world)
    lo, hi = bounds[rank], bounds[rank + 1]
    rng = np.random.RandomState(17)
    U0, V0 = plsa.plsa_init(X, k, "random", rng)         # same stream on every rank
    U, V = U0[lo:hi].copy(), V0.copy()
    Xd = np.asarray(X[lo:hi].todense(), dtype=np.float64)
    for _ in range(n_iter):
        U, V = em_shard_step(Xd, U, V, 1e-32, allreduce)
    ll_local = np.array([np.sum(Xd[Xd > 0] * np.log((U @ V)[Xd > 0]))])
    ll = allreduce(ll_local)[0]
    parts = [None] * world if rank == 0 else None
    dist.gather_object(U, parts, dst=0)
    if rank == 0:
        from oracle import oracle
        sw = np.ones(n, dtype=np.float32)
        ref_U, ref_V = oracle.plsa_fit(X, k, sw, n_iter=n_iter, tolerance=0.0, random_state=17,
                                       precision="f64")
        U_all = np.concatenate(parts)
        err_u = np.linalg.norm(U_all - ref_U) / np.linalg.norm(ref_U)
        err_v = np.linalg.norm(V - ref_V) / np.linalg.norm(ref_V)
        ref_ll = oracle.log_likelihood(X, ref_V, ref_U)
        assert err_u < 1e-6 and err_v < 1e-6, (err_u, err_v)
        assert abs(ll - ref_ll) < 1e-6 * abs(ref_ll), (ll, ref_ll)
        print("SHARDED_EM_OK", world, bounds)
    plumb.barrier()
    plumb.close()


if __name__ == "__main__":
    main()
