#!/usr/bin/env python3
"""A/B on one GPU: the junction finder's table passes (k_edges: one random 16-byte slot per text position) under the
device's L2 fetch granularity (LCG_L2_FETCH = 32 / 64 / 128 bytes; unset = the driver's default), and the cost of k-mers
wider than one word (k = 31 vs 33 vs 63 vs 127).  Prints one line per setting: mean ms of k_edges and of the whole device
pipeline over --steps builds (sequences uploaded every build; CUDA-event times from lcg_stats).

    python tools/ab_graph_l2fetch.py [--genomes 4] [--length 10000000] [--steps 3]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--genomes", type=int, default=4)
    ap.add_argument("--length", type=int, default=10_000_000)
    ap.add_argument("--steps", type=int, default=3)
    a = ap.parse_args()
    import sibeliaz_b200 as sb
    from tools.gen_synthetic import generate
    d = "/tmp/lcb_ab_graph/%dx%d" % (a.genomes, a.length)
    os.makedirs(d, exist_ok=True)
    fas = generate(d, "star", a.genomes, a.length, 0.05, 1)
    seqs = []
    for f in fas:
        with open(f, "rb") as fh:
            seqs.append(b"".join(line.strip() for line in fh if not line.startswith(b">")))

    def run(k, fetch):
        if fetch is None:
            os.environ.pop("LCG_L2_FETCH", None)
        else:
            os.environ["LCG_L2_FETCH"] = str(fetch)
        sb.JunctionGraph(sequences=seqs, k=k).close()
        e = dv = 0.0
        for _ in range(a.steps):
            g = sb.JunctionGraph(sequences=seqs, k=k)
            e += g.stats["ms_edges"]
            dv += g.stats["ms_device"]
            st = g.stats
            g.close()
        print("k=%-3d l2_fetch=%-7s k_edges %8.3f ms  device pipeline %8.3f ms  (%d k-mer positions, %d distinct, %d junctions, table %d slots)"
              % (k, fetch if fetch else "default", e / a.steps, dv / a.steps, st["n_kmers"], st["n_distinct"], st["n_junctions"], st["table_slots"]), flush=True)

    for fetch in (None, 32, 64, 128):
        run(21, fetch)
    for k in (31, 33, 63, 127):
        run(k, None)
    run(33, 32)


if __name__ == "__main__":
    main()
