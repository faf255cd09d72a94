"""Float64-exact EM outputs for the two C1 golden corpora (tests/golden/c1_*.npz), so that
bench.py can print engine-vs-exact and reference-vs-exact beside engine-vs-reference without
touching the oracle at run time.

    python tests/golden/make_exact.py

Runs the oracle's float64 instantiation (oracle/plsa_oracle_impl.h: the reference algorithm
of enstop/plsa.py:91-105,182-202 with double accumulators) from the goldens' own float32
start, 50 iterations, tolerance 0 — the same run whose reference output the goldens hold as
pzd_50 / pwz_50.  Output: tests/golden/c1_exact.npz (float32 copies of the float64 results).
"""
import os
import sys

import numpy as np
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import oracle  # noqa: E402


def main():
    out = {}
    for tag in ("c1_planted", "c1_zipf"):
        g = np.load(os.path.join(HERE, tag + ".npz"))
        X = sp.csr_matrix((g["data"], g["indices"], g["indptr"]), shape=tuple(g["shape"]))
        sw = np.ones(X.shape[0], dtype=np.float32)
        pzd, pwz = oracle.plsa_fit(X, int(g["k"]), sw, init=(g["pzd0"], g["pwz0"]), n_iter=50,
                                   tolerance=0.0, precision="f64")
        out[tag + "_pzd_50"] = np.asarray(pzd, dtype=np.float32)
        out[tag + "_pwz_50"] = np.asarray(pwz, dtype=np.float32)
        ref_v = np.linalg.norm(g["pwz_50"].astype(np.float64) - pwz) / np.linalg.norm(pwz)
        ref_u = np.linalg.norm(g["pzd_50"].astype(np.float64) - pzd) / np.linalg.norm(pzd)
        print("%s: reference vs exact  components %.3e  embedding %.3e" % (tag, ref_v, ref_u))
    np.savez_compressed(os.path.join(HERE, "c1_exact.npz"), **out)


if __name__ == "__main__":
    main()
