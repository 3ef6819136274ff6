#!/usr/bin/env python3
"""Developer helper: junction finder on the GPU vs the CPU restatement on given FASTA files; prints stats and the
first differing records.   python tools/graph_check.py K fasta...  [--no-oracle]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import sibeliaz_b200 as sb  # noqa: E402

REC = np.dtype([("pos", "<u4"), ("id", "<i8")])


def main():
    k = int(sys.argv[1])
    fas = [a for a in sys.argv[2:] if not a.startswith("--")]
    for rep in range(2):
        t = time.time()
        g = sb.JunctionGraph(fas, k)
        print("gpu build %.3f s" % (time.time() - t), {a: (round(b, 2) if isinstance(b, float) else b) for a, b in g.stats.items()}, flush=True)
    out = g.write("/tmp/graph_check_gpu.dbg")
    for a in sys.argv:
        if a.startswith("--ref-dbg="):  # a junction file of the reference twopaco for the same input: normal forms must agree
            from oracle_binding import canonical_junctions
            t = time.time()
            same = canonical_junctions(out) == canonical_junctions(a.split("=", 1)[1], "/tmp/graph_check_ref.canon")
            print("normal form identical to the reference's junction file:", same, "(%.1f s)" % (time.time() - t), flush=True)
    if "--time-ref" in sys.argv:
        from oracle_binding import run_twopaco
        th = min(16, os.cpu_count() or 1)
        t = time.time()
        run_twopaco(fas, k, "/tmp/graph_check_ref.dbg", threads=th, tmpdir="/tmp")
        print("reference twopaco -t %d: %.1f s" % (th, time.time() - t), flush=True)
    if "--cli" in sys.argv:
        import subprocess
        for rep in range(2):
            t = time.time()
            r = subprocess.run([sb.GRAPH_CLI_PATH, "--tmpdir", "/tmp", "-t", "16", "-k", str(k), "--filtermemory", "4", "-o", "/tmp/graph_check_cli.dbg", "--stats"] + fas,
                               capture_output=True, text=True)
            print("twopaco (B200) whole binary %.3f s rc=%d %s" % (time.time() - t, r.returncode, r.stderr.strip()[-700:]), flush=True)
    if "--no-oracle" in sys.argv:
        return
    from oracle_binding import graph_oracle_build
    t = time.time()
    n = graph_oracle_build(fas, k, "/tmp/graph_check_oracle.dbg")
    print("oracle %.1f s, %d records" % (time.time() - t, n))
    a, b = np.fromfile(out, REC), np.fromfile("/tmp/graph_check_oracle.dbg", REC)
    if len(a) == len(b) and (a == b).all():
        print("IDENTICAL", len(a))
        return
    print("DIFFERENT: gpu %d records, oracle %d" % (len(a), len(b)))
    m = min(len(a), len(b))
    bad = np.flatnonzero(a[:m] != b[:m])
    print("first differing indices", bad[:10])
    for i in bad[:10]:
        print(i, "gpu", a[i], "oracle", b[i])
    pa, pb = set(map(tuple, a.tolist())), set(map(tuple, b.tolist()))
    print("only gpu", sorted(pa - pb)[:10], "only oracle", sorted(pb - pa)[:10])


if __name__ == "__main__":
    main()
