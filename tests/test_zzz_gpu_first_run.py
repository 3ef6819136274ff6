"""GPU tests written after the last GPU session of round 2 (no GPU minutes were left to run them).  Two restate a test the
reference itself ships for a neighbouring step; the third runs the DEFAULT pair of traversal kernels (the common-case kernel
with the general one beside it) against the oracle on every input kind -- tests/test_gpu_parity.py asks for the step counters
(collect_counters=True), which only the general kernel keeps, so there the common-case kernel is covered through the GFF
comparisons of the CLI / fused-pipeline tests and bench.py's parity check, not block by block.  The CPU suite runs the same
checks on the restatements and on the device code compiled for the host / under the warp emulator.  The file name sorts
last, so that nothing depends on these."""
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


def _default_kernels_equal_oracle(case, **kw):
    import numpy as np
    from oracle_binding import Oracle
    orc = Oracle(case.graph, case.fastas, case.k, case.a)
    ob = orc.find_blocks(case.m, case.b)
    st = sb.JunctionStorage(case.graph, case.fastas, case.k, case.a)
    bf = sb.BlocksFinder(st, case.k, **kw)  # no step counters: the common-case kernel takes every work item first
    pb = bf.find_blocks(case.m, case.b)
    assert len(pb) == len(ob["id"]) > 0
    assert np.array_equal(pb["id"], ob["id"]) and np.array_equal(pb["chr"], ob["chr"])
    assert np.array_equal(pb["start"].astype(np.uint64), ob["start"]) and np.array_equal(pb["end"].astype(np.uint64), ob["end"])
    assert bf.stats["lean_runs"] > 0
    stats = dict(bf.stats)
    bf.close()
    orc.close()
    return stats


def test_default_kernels_block_by_block_star_and_examples(star_small, examples):
    from conftest import Case
    for window in (None, 256, 1024):
        kw = dict(window_init=window, window_max=window) if window else {}
        _default_kernels_equal_oracle(star_small, **kw)
    for m, b in ((100, 100), (30, 500), (200, 50)):
        _default_kernels_equal_oracle(Case(star_small.name, star_small.graph, star_small.fastas, star_small.k, b=b, m=m))
    _default_kernels_equal_oracle(Case(star_small.name, star_small.graph, star_small.fastas, star_small.k, a=4))
    _default_kernels_equal_oracle(examples["k25"])
    st = _default_kernels_equal_oracle(examples["k15"])
    assert st["lean_bails"] > 0  # repeat-rich: evaluations handed over to the general kernel in the middle of a path


@pytest.mark.parametrize("kind,genomes,length,k,seed", [("mammal", 8, 1500000, 25, 3), ("pangenome", 16, 500000, 15, 4)])
def test_default_kernels_block_by_block_synthetic_kinds(tmp_path, kind, genomes, length, k, seed):
    from test_gpu_parity import _fresh_case
    _default_kernels_equal_oracle(_fresh_case(tmp_path, kind, genomes, length, k, seed))
