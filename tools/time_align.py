#!/usr/bin/env python3
"""Developer helper: the alignment stage (include/sibeliaz_align.h) on the examples' 1350 blocks: wall clock, kernel time,
arena levels, and the check against the reference's shipped alignment.maf (all 1332 golden paragraphs byte for byte)."""
import json
import lzma
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import sibeliaz_b200 as sb
    from oracle_binding import Oracle, maf_paragraphs
    from conftest import _unxz
    g = os.path.join(ROOT, "tests", "golden", "examples")
    d = tempfile.mkdtemp(prefix="align_")
    fas = [_unxz(os.path.join(g, "genome%d.fa.xz" % i), os.path.join(d, "genome%d.fa" % i)) for i in (1, 2)]
    dbg = _unxz(os.path.join(g, "k25.dbg.xz"), os.path.join(d, "k25.dbg"))
    orc = Oracle(dbg, fas, 25, 150)
    orc.find_blocks(50, 200)
    out = os.path.join(d, "lcb")
    orc.generate_output(out, True, 256, 50)
    files = [os.path.join(out, n) for n in os.listdir(out) if n.endswith(".tmp")]
    skip_long = "--short" in sys.argv
    if skip_long:  # leave out the chunk files that hold a block longer than 5 kbp (developer runs on a tight GPU budget)
        files = [f for f in files if max((len(t) for line in open(f) for t in line.split("@")), default=0) < 5000]
    for rep in range(2):
        t = time.time()
        st = sb.global_alignment(files, "genome1.fa genome2.fa", os.path.join(d, "alignment.maf"))
        print("rep %d: %.2fs" % (rep, time.time() - t), json.dumps(st), flush=True)
    mine = maf_paragraphs(os.path.join(d, "alignment.maf"))
    golden = maf_paragraphs(lzma.open(os.path.join(g, "golden_k25_alignment.maf.xz"), "rt").read(), is_text=True)
    hit = sum(1 for k, v in golden.items() if mine.get(k) == v)
    miss = sum(1 for k in golden if k not in mine)
    print("paragraphs: mine %d, golden %d, byte-identical %d, absent from mine %d, different %d" % (
        len(mine), len(golden), hit, miss, len(golden) - hit - miss))


if __name__ == "__main__":
    main()
