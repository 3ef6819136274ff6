"""Parity tests proper: the CUDA path (through the C ABI) against the CPU restatement oracle and against the
committed reference fixtures.  Bit-exact: integer / byte / index work only."""
import filecmp
import os

import numpy as np
import pytest

from conftest import canonical_gff

pytestmark = pytest.mark.gpu


def _oracle_and_product(case, tmp_path, window=None, gen_seq=False, chunks=0):
    import sibeliaz_b200 as sb
    from oracle_binding import Oracle
    orc = Oracle(case.graph, case.fastas, case.k, case.a)
    st = sb.JunctionStorage(case.graph, case.fastas, case.k, case.a)
    kw = {}
    if window:
        kw = dict(window_init=window, window_max=window)
    bf = sb.BlocksFinder(st, case.k, collect_counters=True, **kw)
    return orc, st, bf


def _check_index(orc, st):
    oi, pi = orc.index(), st.arrays()
    for name in ("chr_off", "pos_id", "pos_bp", "next_ch", "prev_rc", "vtx_off", "occ_g"):
        assert np.array_equal(oi[name], pi[name]), name


def _check_seeds(orc, bf):
    os_, ps = orc.seeds(), bf.seeds()
    assert len(os_["vid"]) == len(ps["vid"])
    for name in ("vid", "ch", "count", "rank", "res_pos", "res_chr"):
        assert np.array_equal(os_[name], ps[name]), "seed field %s differs" % name


def _check_blocks(orc, bf, case):
    ob = orc.find_blocks(case.m, case.b)
    pb = bf.find_blocks(case.m, case.b)
    assert len(pb) == len(ob["id"]), "block instances: product %d oracle %d" % (len(pb), len(ob["id"]))
    assert np.array_equal(pb["id"], ob["id"])
    assert np.array_equal(pb["chr"], ob["chr"])
    assert np.array_equal(pb["start"].astype(np.uint64), ob["start"])
    assert np.array_equal(pb["end"].astype(np.uint64), ob["end"])
    assert bf.stats["kernel_launches"] > 0 and bf.stats["traversals_first"] >= bf.stats["n_seeds"]


def _check_gff(bf, case, tmp_path):
    out = str(tmp_path / ("out_" + case.name))
    bf.generate_output(out, False, 0)
    gff = os.path.join(out, "blocks_coords.gff")
    assert canonical_gff(gff) == canonical_gff(case.ref_gff)
    assert filecmp.cmp(gff, case.ref_gff, shallow=False), "GFF is set-equal but not byte-identical"


def test_star_small_all_stages(star_small, tmp_path):
    orc, st, bf = _oracle_and_product(star_small, tmp_path)
    _check_index(orc, st)
    bf.create(star_small.m, star_small.b)
    _check_seeds(orc, bf)
    _check_blocks(orc, bf, star_small)
    _check_gff(bf, star_small, tmp_path)


@pytest.mark.parametrize("window", [256, 1024, 8192])
def test_star_small_window_independent(star_small, tmp_path, window):
    """The result must not depend on the speculation window (only on the reference's phase size 256)."""
    orc, st, bf = _oracle_and_product(star_small, tmp_path, window=window)
    _check_blocks(orc, bf, star_small)


def test_star_small_block_sequences(star_small, tmp_path):
    import sibeliaz_b200 as sb
    st = sb.JunctionStorage(star_small.graph, star_small.fastas, star_small.k, star_small.a)
    bf = sb.BlocksFinder(st, star_small.k)
    bf.find_blocks(star_small.m, star_small.b)
    out = str(tmp_path / "seq")
    bf.generate_output(out, True, 4)
    got = b""
    for i in range(4):
        got += b"== %d.tmp\n" % i + open(os.path.join(out, "%d.tmp" % i), "rb").read()
    assert got == open(star_small.ref_chunks, "rb").read()


def test_examples_k15(examples, tmp_path):
    """BASELINE configs[0]: examples/genome1.fa + genome2.fa, k=15."""
    case = examples["k15"]
    orc, st, bf = _oracle_and_product(case, tmp_path)
    _check_index(orc, st)
    bf.create(case.m, case.b)
    _check_seeds(orc, bf)
    _check_blocks(orc, bf, case)
    _check_gff(bf, case, tmp_path)


def test_examples_k25_golden(examples, tmp_path):
    """The reference's own golden: examples/sibeliaz_out/blocks_coords.gff at defaults."""
    case = examples["k25"]
    orc, st, bf = _oracle_and_product(case, tmp_path)
    bf.create(case.m, case.b)
    _check_seeds(orc, bf)
    _check_blocks(orc, bf, case)
    _check_gff(bf, case, tmp_path)
