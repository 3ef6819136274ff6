"""Pins oracle/graph_oracle.cpp -- the CPU restatement of the junction set TwoPaCo computes -- against junction files
made by the reference itself: the committed fixtures (examples k=15 / k=25, star 4 x 200 kbp k=21, all produced by the
compiled reference twopaco) and, where oracle/_ref/twopaco is present, fresh runs on an N-rich multi-record input.
Comparison is in the label-free normal form (the reference's vertex ids and orientations are seeded from /dev/urandom
and differ from run to run, see the header of graph_oracle.cpp)."""
import lzma
import os
import subprocess

import pytest

from conftest import GOLDEN
from graph_cases import write_nrich
from oracle_binding import REF_TWOPACO, canonical_junctions, graph_oracle_build, run_twopaco


@pytest.mark.parametrize("which", ["k15", "k25"])
def test_graph_oracle_matches_reference_fixture_examples(examples, which, tmp_path):
    case = examples[which]
    mine = str(tmp_path / "mine.dbg")
    n = graph_oracle_build(case.fastas, case.k, mine)
    assert n > 100000 and os.path.getsize(mine) == os.path.getsize(case.graph)
    assert canonical_junctions(mine) == canonical_junctions(case.graph, str(tmp_path / "ref.canon"))


def test_graph_oracle_matches_reference_fixture_star(star_small, tmp_path):
    mine = str(tmp_path / "mine.dbg")
    graph_oracle_build(star_small.fastas, star_small.k, mine)
    assert canonical_junctions(mine) == canonical_junctions(star_small.graph, str(tmp_path / "ref.canon"))


@pytest.mark.skipif(not os.path.exists(REF_TWOPACO), reason="compiled reference twopaco not present")
@pytest.mark.parametrize("k", [15, 21])
def test_graph_oracle_matches_compiled_reference_on_nrich_input(tmp_path, k):
    fas = write_nrich(str(tmp_path))
    ref = run_twopaco(fas, k, str(tmp_path / "ref.dbg"), threads=4)
    mine = str(tmp_path / "mine.dbg")
    graph_oracle_build(fas, k, mine)
    assert canonical_junctions(mine) == canonical_junctions(ref)


def test_string_keyed_restatement_equals_word_keyed(tmp_path, monkeypatch):
    """k > 31 uses the k-mer strings themselves as keys; at a small k both variants must write the same file."""
    fas = write_nrich(str(tmp_path))
    a, b = str(tmp_path / "words.dbg"), str(tmp_path / "strings.dbg")
    graph_oracle_build(fas, 21, a)
    monkeypatch.setenv("GRO_STRING_KEYS", "1")
    graph_oracle_build(fas, 21, b)
    assert open(a, "rb").read() == open(b, "rb").read()


@pytest.mark.parametrize("k", [33, 63, 127])
def test_graph_oracle_wide_k_matches_reference_fixture_nrich(tmp_path, k):
    """Vertex sizes beyond one 64-bit word against the compiled reference's junction files (tests/golden/wide_k,
    made by tests/golden/make_wide_fixtures.py)."""
    mine = str(tmp_path / "mine.dbg")
    graph_oracle_build(write_nrich(str(tmp_path)), k, mine)
    with lzma.open(os.path.join(GOLDEN, "wide_k", "nrich_k%d.canon.xz" % k)) as f:
        assert canonical_junctions(mine) == f.read()


def test_wide_k_pipeline_matches_reference_fixture_star(star_small, tmp_path):
    """k = 33 end to end on the CPU checkers: restated junction file == the reference twopaco's (normal form), and the
    LCB restatement on it writes the reference sibeliaz-lcb's blocks_coords.gff byte for byte."""
    from oracle_binding import Oracle
    mine = str(tmp_path / "mine.dbg")
    graph_oracle_build(star_small.fastas, 33, mine)
    with lzma.open(os.path.join(GOLDEN, "wide_k", "star4x200k_k33.canon.xz")) as f:
        assert canonical_junctions(mine) == f.read()
    o = Oracle(mine, star_small.fastas, 33, star_small.a)
    o.find_blocks(star_small.m, star_small.b)
    o.generate_output(str(tmp_path / "out"), False, 0, star_small.m)
    with lzma.open(os.path.join(GOLDEN, "wide_k", "star4x200k_k33.gff.xz")) as f:
        assert open(tmp_path / "out" / "blocks_coords.gff", "rb").read() == f.read()


@pytest.mark.skipif(not os.path.exists(REF_TWOPACO), reason="compiled reference twopaco not present")
@pytest.mark.parametrize("k", [45, 65, 201])
def test_graph_oracle_wide_k_matches_compiled_reference(tmp_path, k):
    fas = write_nrich(str(tmp_path))
    ref = run_twopaco(fas, k, str(tmp_path / "ref.dbg"), threads=4)
    mine = str(tmp_path / "mine.dbg")
    graph_oracle_build(fas, k, mine)
    assert canonical_junctions(mine) == canonical_junctions(ref)


def test_normal_form_is_invariant_under_relabelling(star_small, tmp_path):
    """Flip orientations and permute ids of a junction file: the normal form must not change (and must change when a
    position moves)."""
    import numpy as np
    raw = np.fromfile(star_small.graph, dtype=np.dtype([("pos", "<u4"), ("id", "<i8")]))
    sep = (raw["pos"] == 0xFFFFFFFF) | (raw["id"] == np.iinfo(np.int64).max)
    rng = np.random.default_rng(3)
    ids = np.abs(raw["id"][~sep])
    uniq = np.unique(ids)
    perm = dict(zip(uniq.tolist(), (rng.permutation(len(uniq)) + 7).tolist()))
    flip = dict(zip(uniq.tolist(), rng.integers(0, 2, len(uniq)).tolist()))
    out = raw.copy()
    new = [(-1 if (i < 0) != bool(flip[abs(i)]) else 1) * perm[abs(i)] for i in raw["id"][~sep].tolist()]
    out["id"][~sep] = new
    p = str(tmp_path / "relabelled.dbg")
    out.tofile(p)
    base = canonical_junctions(star_small.graph, str(tmp_path / "a.canon"))
    assert canonical_junctions(p) == base
    out["pos"][np.flatnonzero(~sep)[5]] += 1
    out.tofile(p)
    assert canonical_junctions(p) != base


def test_lcb_output_does_not_depend_on_labels(star_small, tmp_path):
    """The reason the normal form is the right parity criterion: sibeliaz-lcb's blocks are the same on the reference's
    junction file and on the oracle's differently labelled one."""
    from oracle_binding import Oracle
    mine = str(tmp_path / "mine.dbg")
    graph_oracle_build(star_small.fastas, star_small.k, mine)
    a = Oracle(star_small.graph, star_small.fastas, star_small.k, star_small.a)
    b = Oracle(mine, star_small.fastas, star_small.k, star_small.a)
    a.find_blocks(star_small.m, star_small.b)
    b.find_blocks(star_small.m, star_small.b)
    a.generate_output(str(tmp_path / "a"), False, 0, star_small.m)
    b.generate_output(str(tmp_path / "b"), False, 0, star_small.m)
    assert open(tmp_path / "a" / "blocks_coords.gff", "rb").read() == open(tmp_path / "b" / "blocks_coords.gff", "rb").read()
    assert open(tmp_path / "a" / "blocks_coords.gff", "rb").read() == open(star_small.ref_gff, "rb").read()
