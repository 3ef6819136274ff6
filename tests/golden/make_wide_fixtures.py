#!/usr/bin/env python3
"""Regenerates tests/golden/wide_k/* (run in the build container, where oracle/_ref exists): vertex sizes beyond one
64-bit word, made by the compiled, unmodified reference.

star4x200k_k33.canon.xz   junction file of reference twopaco -k 33 on the star 4 x 200 kbp input (tools/gen_synthetic.py
                          star, rate 0.05, seed 7), in the label-free normal form (oracle_binding.canonical_junctions):
                          the reference's own labels differ from run to run
star4x200k_k33.gff.xz     blocks_coords.gff of reference sibeliaz-lcb -k 33 -b 200 -m 50 -a 150 on that junction file
nrich_k{33,63,127}.canon.xz   reference twopaco on the N-rich multi-record input of tests/graph_cases.py, normal form
"""
import lzma
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from graph_cases import write_nrich  # noqa: E402
from oracle_binding import REF_LCB, REF_TWOPACO, canonical_junctions, run_reference_lcb, run_twopaco  # noqa: E402
from tools.gen_synthetic import generate  # noqa: E402


def xz_bytes(data, dst):
    with lzma.open(dst, "wb", preset=9 | lzma.PRESET_EXTREME) as g:
        g.write(data)


def main():
    assert os.path.exists(REF_LCB) and os.path.exists(REF_TWOPACO), "run `make -C oracle ref` first"
    out = os.path.join(HERE, "wide_k")
    os.makedirs(out, exist_ok=True)
    with tempfile.TemporaryDirectory() as tmp:
        fa = generate(os.path.join(tmp, "star"), "star", 4, 200000, 0.05, 7)
        dbg = run_twopaco(fa, 33, os.path.join(tmp, "star33.dbg"), threads=1)
        xz_bytes(canonical_junctions(dbg), os.path.join(out, "star4x200k_k33.canon.xz"))
        lcb = os.path.join(tmp, "lcb33")
        os.makedirs(lcb)
        run_reference_lcb(dbg, fa, 33, lcb, b=200, m=50, a=150, threads=1)
        xz_bytes(open(os.path.join(lcb, "blocks_coords.gff"), "rb").read(), os.path.join(out, "star4x200k_k33.gff.xz"))
        nr = os.path.join(tmp, "nrich")
        os.makedirs(nr)
        fas = write_nrich(nr)
        for k in (33, 63, 127):
            dbg = run_twopaco(fas, k, os.path.join(nr, "k%d.dbg" % k), threads=1)
            xz_bytes(canonical_junctions(dbg), os.path.join(out, "nrich_k%d.canon.xz" % k))
    for f in sorted(os.listdir(out)):
        print("%9d %s" % (os.path.getsize(os.path.join(out, f)), f))


if __name__ == "__main__":
    main()
