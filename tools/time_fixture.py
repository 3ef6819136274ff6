#!/usr/bin/env python3
"""Developer helper: time the product on the committed fixtures for several fixed windows."""
import json
import lzma
import os
import shutil
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import sibeliaz_b200 as sb  # noqa: E402

G = os.path.join(ROOT, "tests", "golden", "examples")
d = tempfile.mkdtemp()


def unxz(n):
    out = os.path.join(d, n)
    with lzma.open(os.path.join(G, n + ".xz")) as f, open(out, "wb") as g:
        shutil.copyfileobj(f, g)
    return out


fas = [unxz("genome1.fa"), unxz("genome2.fa")]
for k in (15, 25):
    dbg = unxz("k%d.dbg" % k)
    st = sb.JunctionStorage(dbg, fas, k, 150)
    for W in [int(x) for x in (sys.argv[1:] or ["1024", "4096", "16384", "65536"])]:
        bf = sb.BlocksFinder(st, k, window_init=W, window_max=W, collect_counters=True)
        bf.create(50, 200)
        bf.enumerate_seeds()
        t = time.time()
        b = bf.find_blocks(50, 200)
        s = bf.stats
        print("k=%d W=%6d find %.3fs windows %d rounds %d runs %d+%d t_walk %d trav_ms %.1f blocks %d" % (
            k, W, time.time() - t, s["windows"], s["rounds"], s["traversals_first"], s["traversals_rerun"], s["t_walk"],
            s["ms_traverse_kernels"], len(b)), flush=True)
        bf.close()
