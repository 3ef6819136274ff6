#!/usr/bin/env python3
"""Regenerates tests/golden/spoa_sample/* (run in the build container): the data of spoa's own unit tests
(spoa/test/data/sample.fastq.gz, 55 reads of 149 - 515 characters; spoa_test.cpp:20-86) as ONE block of an LCB chunk file,
and the MSA the unmodified reference library makes of it in the mode its `Global` test and the sibeliaz wrapper share
(kNW, 5 / -4 / -8 linear: spoa_test.cpp:245-259, `spoa -l 1 -r 1 -e -8` sibeliaz:66), printed by oracle/_ref/spoa-ref.

block.tmp.xz   the chunk file (headers r<i>;0;<len>;+;<len>)
msa.maf.xz     spoa-ref --chunk block.tmp -l 1 -r 1 -e -8
"""
import gzip
import lzma
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle_binding import REF_SPOA, write_chunk  # noqa: E402


def main():
    lines = gzip.open("/root/reference/spoa/test/data/sample.fastq.gz", "rt").read().splitlines()
    seqs = [lines[i + 1] for i in range(0, len(lines), 4)]
    assert len(seqs) == 55
    out = os.path.join(HERE, "spoa_sample")
    os.makedirs(out, exist_ok=True)
    tmp = os.path.join(out, "block.tmp")
    write_chunk(tmp, [[("r%d;0;%d;+;%d" % (i, len(s), len(s)), s) for i, s in enumerate(seqs)]])
    maf = subprocess.run([REF_SPOA, "--chunk", tmp, "-l", "1", "-r", "1", "-e", "-8"], check=True, stdout=subprocess.PIPE).stdout
    with lzma.open(os.path.join(out, "msa.maf.xz"), "wb", preset=9) as g:
        g.write(maf)
    with open(tmp, "rb") as f, lzma.open(tmp + ".xz", "wb", preset=9) as g:
        g.write(f.read())
    os.remove(tmp)
    for n in sorted(os.listdir(out)):
        print("%8d %s" % (os.path.getsize(os.path.join(out, n)), n))


if __name__ == "__main__":
    main()
