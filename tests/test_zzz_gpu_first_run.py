"""GPU tests written after the last GPU session of round 2 (no GPU minutes were left to run them): each restates a test the
reference itself ships for a neighbouring step.  The CPU suite runs the same checks on the restatements and on the device
code compiled for the host / under the warp emulator.  The file name sorts last, so that nothing depends on them."""
import lzma
import os

import pytest

import sibeliaz_b200 as sb
from conftest import GOLDEN

pytestmark = pytest.mark.gpu


def test_gpu_reproduces_the_reference_msa_of_spoa_sample(tmp_path):
    """spoa's unit-test data (55 reads; tests/test_spoa_sample.py): one block of 55 copies -- more copies than any block of
    the examples, so the optimistic arena levels of the alignment stage are outgrown and the block is run again one level up."""
    from test_spoa_sample import FIX, check_msa_properties
    chunk = str(tmp_path / "block.tmp")
    with lzma.open(os.path.join(FIX, "block.tmp.xz")) as f, open(chunk, "wb") as g:
        g.write(f.read())
    with lzma.open(os.path.join(FIX, "msa.maf.xz"), "rt") as f:
        want = f.read()
    out = str(tmp_path / "sample.maf")
    st = sb.global_alignment([chunk], "sample", out)
    got = open(out).read()
    assert got == "##maf version=1\n# sibeliaz v1.2.7 \n# cmd=sample\n" + want
    check_msa_properties(got, chunk)
    assert st["n_blocks"] == 1 and st["kernel_launches"] >= 1


def test_reference_selftest_of_the_junction_finder_on_the_gpu(tmp_path):
    """`twopaco --test` restated (tests/graph_selftest.py, tests/test_graph_selftest.py) against the GPU junction finder:
    the positions in its junction file are the naively computed ones, for the self-test's k = 3 .. 9 and two wider ones."""
    import random
    from graph_selftest import check, make_case, write_fasta
    chrs = make_case(random.Random(4))
    fa = write_fasta(str(tmp_path / "test.fa"), chrs)
    for k in (3, 5, 7, 9, 15, 33):
        g = sb.JunctionGraph([fa], k)
        assert check(g.write(str(tmp_path / ("gpu%d.dbg" % k))), chrs, k) > 12
        g.close()
