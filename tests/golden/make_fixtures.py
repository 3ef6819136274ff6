#!/usr/bin/env python3
"""Regenerates tests/golden/* (run in the build container, where /root/reference and oracle/_ref exist).

examples/    the reference's own fixture: examples/genome{1,2}.fa and the shipped golden
             examples/sibeliaz_out/blocks_coords.gff (defaults k=25 b=200 m=50 a=150), xz-compressed,
             plus junction files from the compiled reference twopaco (k=25, k=15) and the compiled
             reference sibeliaz-lcb's GFF for k=15 (BASELINE configs[0]; no shipped golden exists for it).
star4x200k/  seeded synthetic (tools/gen_synthetic.py star, 4 x 200 kbp, rate 0.05, seed 7), k=21:
             junction file from reference twopaco and GFF (+ .tmp chunks listing) from reference sibeliaz-lcb.
"""
import lzma
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle_binding import REF_LCB, REF_TWOPACO, run_reference_lcb, run_twopaco  # noqa: E402
from tools.gen_synthetic import generate  # noqa: E402

REF = "/root/reference/examples"


def xz(src, dst):
    with open(src, "rb") as f, lzma.open(dst, "wb", preset=9 | lzma.PRESET_EXTREME) as g:
        shutil.copyfileobj(f, g)


def main():
    assert os.path.exists(REF_LCB) and os.path.exists(REF_TWOPACO), "run `make -C oracle ref` first"
    ex = os.path.join(HERE, "examples")
    os.makedirs(ex, exist_ok=True)
    fas = [os.path.join(REF, "genome1.fa"), os.path.join(REF, "genome2.fa")]
    for f in fas:
        xz(f, os.path.join(ex, os.path.basename(f) + ".xz"))
    xz(os.path.join(REF, "sibeliaz_out", "blocks_coords.gff"), os.path.join(ex, "golden_k25_blocks_coords.gff.xz"))
    # the alignment stage's golden: 1332 MAF paragraphs made by the reference pipeline (spoa per block)
    xz(os.path.join(REF, "sibeliaz_out", "alignment.maf"), os.path.join(ex, "golden_k25_alignment.maf.xz"))
    with tempfile.TemporaryDirectory() as tmp:
        for k in (25, 15):
            dbg = run_twopaco(fas, k, os.path.join(tmp, "k%d.dbg" % k), threads=1)
            xz(dbg, os.path.join(ex, "k%d.dbg.xz" % k))
            out = os.path.join(tmp, "out%d" % k)
            os.makedirs(out)
            run_reference_lcb(dbg, fas, k, out, b=200, m=50, a=150, threads=1)
            xz(os.path.join(out, "blocks_coords.gff"), os.path.join(ex, "ref_k%d_blocks_coords.gff.xz" % k))
        st = os.path.join(HERE, "star4x200k")
        os.makedirs(st, exist_ok=True)
        fa = generate(os.path.join(tmp, "star"), "star", 4, 200000, 0.05, 7)
        dbg = run_twopaco(fa, 21, os.path.join(tmp, "star.dbg"), threads=1)
        xz(dbg, os.path.join(st, "k21.dbg.xz"))
        out = os.path.join(tmp, "starout")
        os.makedirs(out)
        run_reference_lcb(dbg, fa, 21, out, b=200, m=50, a=150, threads=1, noseq=False, chunks=4)
        xz(os.path.join(out, "blocks_coords.gff"), os.path.join(st, "ref_blocks_coords.gff.xz"))
        with lzma.open(os.path.join(st, "ref_chunks.tmp.xz"), "wb", preset=9) as g:
            for i in range(4):
                with open(os.path.join(out, "%d.tmp" % i), "rb") as f:
                    g.write(b"== %d.tmp\n" % i)
                    shutil.copyfileobj(f, g)
    for d, _, files in os.walk(HERE):
        for f in sorted(files):
            p = os.path.join(d, f)
            print("%9d %s" % (os.path.getsize(p), os.path.relpath(p, HERE)))


if __name__ == "__main__":
    main()
