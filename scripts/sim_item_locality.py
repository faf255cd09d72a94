#!/usr/bin/env python
"""CPU model of the gather locality of the row pass's work-item order (no GPU needed).

A CTA of the row pass carries 8 warps x (32/G) consecutive work items (48 at k=20); every
stored entry of an item gathers one factor row (one 128-byte line) of the other factor.
The L1 of an SM can only serve a gather whose line another entry of a co-resident item
already pulled in.  For an item order this script counts, per CTA, the distinct gathered rows
against the gathers (the share of gathers that CAN hit in L1 when the CTA's lines all stay
resident), and the lane-steps a warp wastes because its items differ in length.

Orders compared for the split-row chunks (whole rows are left longest-first):
  row      chunks of a row adjacent (the order build_items() produces today)
  window   chunks sorted by the position of their first entry's gathered row (so the chunks a
           CTA carries cover the same window of gathered rows), inside length buckets

The orders are taken from the library's own planner (plsa_plan_items, host-only), so what is
modelled is what the pass launches; --explore adds variants computed here (length buckets).

    python scripts/sim_item_locality.py [C1|C2] [--chunk 256] [--explore --bucket 1,16,32]
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from enstop_b200 import _lib, synth  # noqa: E402


def build_items(indptr, chunk, align):
    """Mirror of build_items() in enstop_b200/csrc/plsa_b200.cu (before the sort)."""
    indptr = indptr.astype(np.int64)
    lead = indptr[:-1] & (align - 1)
    span = np.diff(indptr) + lead
    nc = np.maximum(1, -(-span // chunk))
    per = np.minimum(chunk, -(-(-(-span // nc)) // align) * align)
    per = np.where(span <= chunk, span, per)
    nc = np.where(span <= chunk, 1, -(-span // np.maximum(per, 1)))
    row = np.repeat(np.arange(indptr.shape[0] - 1), nc)
    first = np.cumsum(nc) - nc
    c = np.arange(row.shape[0]) - first[row]
    start = (indptr[:-1] - lead)[row] + c * per[row]
    length = np.minimum(per[row], span[row] - c * per[row])
    skip = np.where(c == 0, lead[row], 0)
    split = nc[row] > 1
    return dict(row=row, start=start, len=length, skip=skip, split=split, c=c, nc=nc[row])


def evaluate(items, order, gat, items_per_cta, items_per_warp, label):
    start = items["start"][order]
    length = items["len"][order]
    skip = items["skip"][order]
    n_items = order.shape[0]
    cta = np.arange(n_items) // items_per_cta
    real = length - skip
    # all gathered indices with their CTA
    offs = np.repeat(start + skip - (np.cumsum(real) - real), real) + np.arange(real.sum())
    g = gat[offs].astype(np.int64)
    key = np.repeat(cta, real) * (gat.max() + 1) + g
    distinct = np.unique(key).shape[0]
    total = g.shape[0]
    # lane-steps: a warp walks max(len) of its items
    pad = (-n_items) % items_per_warp
    L = np.concatenate([length, np.zeros(pad, dtype=length.dtype)]).reshape(-1, items_per_warp)
    steps = (L.max(axis=1) * items_per_warp).sum()
    # only split items
    sp = items["split"][order]
    sp_real = np.repeat(sp, real)
    d_split = np.unique(key[sp_real]).shape[0]
    t_split = int(sp_real.sum())
    print("%-22s gathers %9d  distinct/CTA %9d  reusable %5.1f %%   split-chunk entries %9d "
          "reusable %5.1f %%   warp occupancy of lane-steps %5.1f %%"
          % (label, total, distinct, 100.0 * (1 - distinct / total), t_split,
             100.0 * (1 - d_split / max(1, t_split)), 100.0 * length.sum() / steps))
    return 1 - distinct / total


# ---- a closer model: an LRU of `lines` 128-byte lines per SM, four resident CTAs ------------
def lru_hit_rate(items, order, gat, items_per_cta, items_per_warp, lines=1400, sms=148,
                 sample=(0, 37, 74, 111), resident=4, block=4):
    """Gather hit rate of an LRU line cache shared by the CTAs resident on one SM.  CTA b runs
    on SM b % sms (waves of `resident` CTAs per SM, a slot takes the SM's next CTA when its
    CTA ends); every resident CTA advances one entry block per step, all its items together."""
    from collections import OrderedDict
    start = items["start"][order]
    length = items["len"][order]
    skip = items["skip"][order]
    n_items = order.shape[0]
    n_cta = -(-n_items // items_per_cta)
    hits = total = 0
    for sm in sample:
        queue = list(range(sm, n_cta, sms))
        cache = OrderedDict()
        slots = []

        def open_cta(b):
            lo, hi = b * items_per_cta, min(n_items, (b + 1) * items_per_cta)
            return [lo, hi, 0, int(length[lo:hi].max())]
        while queue and len(slots) < resident:
            slots.append(open_cta(queue.pop(0)))
        while slots:
            for s in list(slots):
                lo, hi, t, mx = s
                for i in range(lo, hi):
                    a = max(t, int(skip[i]))
                    b_ = min(t + block, int(length[i]))
                    for g in gat[start[i] + a:start[i] + b_] if b_ > a else ():
                        g = int(g)
                        total += 1
                        if g in cache:
                            hits += 1
                            cache.move_to_end(g)
                        else:
                            cache[g] = True
                            if len(cache) > lines:
                                cache.popitem(last=False)
                s[2] = t + block
                if s[2] >= mx:
                    slots.remove(s)
                    if queue:
                        slots.append(open_cta(queue.pop(0)))
    return hits / max(1, total), total



def inflight_working_set(items, gat, items_per_cta, ctas_in_flight=592, row_bytes=128):
    """Distinct gathered rows of every run of `ctas_in_flight` consecutive CTAs (what the whole
    GPU has in flight at one time) x 128 bytes: the L2 footprint of the gathered factor."""
    start, length, skip = items["start"], items["len"], items["skip"]
    real = length - skip
    n_items = start.shape[0]
    win = (np.arange(n_items) // items_per_cta) // ctas_in_flight
    offs = np.repeat(start + skip - (np.cumsum(real) - real), real) + np.arange(real.sum())
    key = np.repeat(win, real).astype(np.int64) * (int(gat.max()) + 1) + gat[offs]
    uniq = np.unique(key)
    per_win = np.bincount((uniq // (int(gat.max()) + 1)).astype(np.int64))
    tot_win = np.bincount(np.repeat(win, real))
    return per_win * row_bytes / 1e6, tot_win


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("config", nargs="?", default="C2")
    ap.add_argument("--chunk", type=int, default=256)
    ap.add_argument("--align", type=int, default=4)
    ap.add_argument("--k", type=int, default=None)
    ap.add_argument("--bucket", default="1,8,16,32")
    ap.add_argument("--explore", action="store_true")
    ap.add_argument("--lru-lines", default="700,1400")
    args = ap.parse_args()
    cfg = synth.CONFIGS[args.config]
    k = args.k or cfg["k"]
    G = max(1, -(-k // 4))
    G = G if G <= 8 else (16 if G <= 16 else 32)
    ipw = 32 // G
    ipc = 8 * ipw
    X = synth.make_config(args.config)
    print("%s: %d x %d, %d entries; k=%d, %d items per warp, %d per CTA, chunk %d"
          % (args.config, X.shape[0], X.shape[1], X.nnz, k, ipw, ipc, args.chunk))
    for name, M in (("term pass", X.T.tocsr()), ("doc pass", X)):
        M.sort_indices()
        gat = M.indices
        for order, label in ((0, "row (library, order 0)"), (1, "window (library, order 1)"),
                             (2, "band (library, order 2)")):
            p = _lib.plan_items(M.indptr, args.chunk, align=args.align, order=order)
            it = dict(start=p["start"], len=p["len"].astype(np.int64), skip=p["skip"].astype(np.int64),
                      split=p["slot"] >= 0, row=p["row"])
            ident = np.arange(it["start"].shape[0])
            if order == 0:
                print("--", name, "(%d items, %d of them chunks of split rows)"
                      % (ident.shape[0], int(it["split"].sum())))
            evaluate(it, ident, gat, ipc, ipw, label)
            ws, tw = inflight_working_set(it, gat, ipc)
            print("      gathered rows in flight per 592 CTAs: %s MB (gathers per run: %s k)"
                  % (" ".join("%.1f" % w for w in ws), " ".join("%d" % (t // 1000) for t in tw)))
            for lines in [int(x) for x in args.lru_lines.split(",")]:
                h, tot = lru_hit_rate(it, ident, gat, ipc, ipw, lines=lines)
                print("      LRU of %4d lines per SM (4 resident CTAs, 4 SMs sampled): gather hit "
                      "rate %5.1f %%" % (lines, 100 * h))
        if not args.explore:
            continue
        items = build_items(M.indptr, args.chunk, args.align)
        # today's order: stable counting sort by length, descending
        base = np.argsort(-items["len"], kind="stable")
        evaluate(items, base, gat, ipc, ipw, "row (today)")
        first_gat = gat[np.minimum(items["start"] + items["skip"], gat.shape[0] - 1)]
        for b in [int(x) for x in args.bucket.split(",")]:
            # chunks: length bucket (descending), then first gathered row; whole rows keep
            # their exact length as the key so that they stay longest-first
            lb = np.where(items["split"], -(-items["len"] // b) * b, items["len"])
            sec = np.where(items["split"], first_gat, 0)
            order = np.lexsort((items["row"], sec, -lb))
            evaluate(items, order, gat, ipc, ipw, "window, bucket %d" % b)


if __name__ == "__main__":
    main()
