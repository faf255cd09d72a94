"""Golden for the cluster representatives of the ensemble (enstop/enstop_.py:309-312 and
:397-411): given cluster labels (and, for the UMAP combiner, HDBSCAN membership strengths) the
reference takes the (weighted) mean of sqrt(topic), squares it and renormalises.

Run in the build container only:  python tests/golden/make_golden_combine.py

``enstop/enstop_.py`` cannot be imported (dask / hdbscan / umap absent) and the clustering that
produces the labels is third-party code, so the two statement groups that FOLLOW the clustering
are lifted from the reference's source text where it lies (found by AST: the
``result = np.empty(...)`` assignment and the ``for`` loop after it, in
``generate_combined_topics_kl`` and ``generate_combined_topics_hellinger_umap``) and executed
on fixed labels / strengths.
"""
import ast
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/enstop/enstop_.py"


def tail_of(func_name):
    text = open(SRC).read()
    lines = text.splitlines()
    tree = ast.parse(text)
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == func_name)
    for i, st in enumerate(fn.body):
        if (isinstance(st, ast.Assign) and isinstance(st.targets[0], ast.Name)
                and st.targets[0].id == "result"):
            loop = fn.body[i + 1]
            assert isinstance(loop, ast.For)
            block = lines[st.lineno - 1:loop.end_lineno]
            indent = len(block[0]) - len(block[0].lstrip())
            return "\n".join(l[indent:] for l in block)
    raise RuntimeError("pattern not found in " + func_name)


def main():
    rng = np.random.RandomState(11)
    base = rng.dirichlet(np.full(120, 0.2), size=4)
    topics = np.vstack([b * (1 + 0.1 * rng.rand(120)) for b in base for _ in range(6)])
    topics /= topics.sum(axis=1, keepdims=True)
    topics = topics.astype(np.float32)
    labels = np.repeat(np.arange(4), 6)
    labels[[2, 9, 17]] = -1                      # noise points belong to no cluster
    strengths = rng.uniform(0.2, 1.0, size=labels.shape[0])
    strengths[labels < 0] = 0.0
    ns = {"np": np, "all_topics": topics, "labels": labels}
    exec(compile(tail_of("generate_combined_topics_kl"), SRC, "exec"), ns)
    mean = ns["result"].copy()
    ns = {"np": np, "all_topics": topics, "labels": labels, "membership_strengths": strengths}
    exec(compile(tail_of("generate_combined_topics_hellinger_umap"), SRC, "exec"), ns)
    weighted = ns["result"].copy()
    np.savez_compressed(os.path.join(HERE, "topic_combine.npz"), all_topics=topics, labels=labels,
                        strengths=strengths, combined_mean=mean, combined_weighted=weighted)
    print("combined", mean.shape, weighted.shape, float(np.abs(mean - weighted).max()))


if __name__ == "__main__":
    sys.exit(main())
