"""Golden all-pairs topic distances from the REAL reference (enstop/enstop_.py:234-263).

Run in the build container only:  python tests/golden/make_golden_distances.py

``enstop/enstop_.py`` cannot be imported here (dask / hdbscan / umap are not installed), so
the three functions on this path — ``kl_divergence`` (:233-241), ``all_pairs_kl_divergence``
(:244-252), ``all_pairs_hellinger_distance`` (:255-263) — are compiled from the reference's
own source text, read where it lies.  ``hellinger`` comes from ``umap.distances`` (umap-learn,
un-pinned ``>=0.3.8``, absent); the reference file carries that function's body as a
commented-out block (:30-46), which is un-commented and compiled the same way.
"""
import ast
import os
import re
import sys

import numba
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/enstop/enstop_.py"


def reference_functions():
    text = open(SRC).read()
    tree = ast.parse(text)
    wanted = ("kl_divergence", "all_pairs_kl_divergence", "all_pairs_hellinger_distance")
    lines = text.splitlines()
    pieces = []
    # umap.distances.hellinger, from the commented-out copy in the reference file
    start = next(i for i, l in enumerate(lines) if l.startswith("# @numba.njit()"))
    block = []
    for l in lines[start:]:
        if not l.startswith("#"):
            break
        block.append(re.sub(r"^# ?", "", l))
    pieces.append("\n".join(block))
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in wanted:
            first = min([node.lineno] + [d.lineno for d in node.decorator_list])
            pieces.append("\n".join(lines[first - 1:node.end_lineno]))
    ns = {"numba": numba, "np": np}
    exec(compile("\n\n".join(pieces), SRC, "exec"), ns)
    return ns


def main():
    ns = reference_functions()
    rng = np.random.RandomState(7)
    base = rng.dirichlet(np.full(90, 0.15), size=5)
    topics = np.vstack([b * (1 + 0.05 * rng.rand(90)) for b in base for _ in range(5)])
    topics[topics < 1e-4] = 0.0                     # exact zeros: the KL mask matters
    topics /= topics.sum(axis=1, keepdims=True)
    topics[3] *= 2.5                                # rows need not sum to one (Hellinger l1 norms)
    topics[11] = 0.0                                # an all-zero row
    topics[12] = topics[13]                         # an exact duplicate
    topics = topics.astype(np.float32)
    kl = ns["all_pairs_kl_divergence"](topics)
    hel = ns["all_pairs_hellinger_distance"](topics)
    np.savez_compressed(os.path.join(HERE, "topic_distances.npz"), topics=topics, kl=kl, hellinger=hel)
    print("topics", topics.shape, "kl", kl.shape, float(kl.max()), "hellinger", float(hel.max()))


if __name__ == "__main__":
    sys.exit(main())
