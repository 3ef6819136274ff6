"""Pins the CPU restatement oracle: (1) the reference's own golden GFF (examples, defaults k=25),
(2) fixtures produced by the unmodified reference compiled into oracle/_ref (examples k=15, star4x200k incl.
.tmp block sequences), and -- when oracle/_ref is present -- (3) a fresh differential run on a seeded synthetic."""
import filecmp
import os

import numpy as np
import pytest

from conftest import canonical_gff
from oracle_binding import REF_LCB, REF_TWOPACO, Oracle, run_reference_lcb, run_twopaco


def _oracle_gff(case, tmp_path, gen_seq=False, chunks=0):
    orc = Oracle(case.graph, case.fastas, case.k, case.a)
    orc.find_blocks(case.m, case.b)
    out = str(tmp_path / ("oracle_" + case.name))
    found, cov = orc.generate_output(out, gen_seq, chunks, case.m)
    return orc, out, found, cov


def test_oracle_matches_reference_golden_k25(examples, tmp_path):
    case = examples["k25"]
    orc, out, found, cov = _oracle_gff(case, tmp_path)
    assert found == 1350  # "Blocks found: 1350" for the shipped golden
    assert filecmp.cmp(os.path.join(out, "blocks_coords.gff"), case.ref_gff, shallow=False)
    # SURVEY.md Appendix B, measured on an instrumented copy of the reference
    c = orc.counters
    assert (c["t_walk"], c["t_occ"], c["t_scan"], c["t_score"]) == (152102880, 38173200, 51744448, 29219757)
    assert (c["process"], c["reruns"], c["mpv"], c["pushes"]) == (137356, 13009, 2423421, 5674170)


def test_oracle_matches_compiled_reference_k15(examples, tmp_path):
    case = examples["k15"]
    orc, out, found, cov = _oracle_gff(case, tmp_path)
    assert found == 478
    assert filecmp.cmp(os.path.join(out, "blocks_coords.gff"), case.ref_gff, shallow=False)
    assert orc.counters["t_walk"] == 19482410 and orc.counters["reruns"] == 1137


def test_oracle_star_small_gff_and_sequences(star_small, tmp_path):
    orc, out, found, cov = _oracle_gff(star_small, tmp_path, gen_seq=True, chunks=4)
    assert filecmp.cmp(os.path.join(out, "blocks_coords.gff"), star_small.ref_gff, shallow=False)
    got = b""
    for i in range(4):
        got += b"== %d.tmp\n" % i + open(os.path.join(out, "%d.tmp" % i), "rb").read()
    assert got == open(star_small.ref_chunks, "rb").read()


def test_oracle_seed_order_is_total(star_small):
    s = Oracle(star_small.graph, star_small.fastas, star_small.k, star_small.a).seeds()
    n = len(s["vid"])
    assert n > 1000
    key = list(zip((-s["count"].astype(np.int64)).tolist(), s["rank"].tolist(), s["res_pos"].tolist(), s["res_chr"].tolist()))
    assert key == sorted(key) and len(set(key)) == n  # Bundle::operator< is a total order (resolve is unique)
    assert (s["count"] > 1).all()


def test_oracle_phase_size_is_semantic(star_small):
    """The 256-seed phase is part of the output definition: other widths may differ, 256 is pinned."""
    a = Oracle(star_small.graph, star_small.fastas, star_small.k, star_small.a)
    b256 = a.find_blocks(star_small.m, star_small.b, phase=256)
    b1 = a.find_blocks(star_small.m, star_small.b, phase=1)
    assert len(b256["id"]) > 0 and len(b1["id"]) > 0
    a2 = Oracle(star_small.graph, star_small.fastas, star_small.k, star_small.a)
    again = a2.find_blocks(star_small.m, star_small.b, phase=256)
    for f in ("id", "chr", "start", "end"):
        assert np.array_equal(b256[f], again[f])  # deterministic


@pytest.mark.skipif(not (os.path.exists(REF_LCB) and os.path.exists(REF_TWOPACO)), reason="oracle/_ref not built")
@pytest.mark.parametrize("kind,k,seed", [("star", 15, 11), ("pangenome", 21, 4), ("mammal", 25, 3)])
def test_oracle_differential_vs_compiled_reference(tmp_path, kind, k, seed):
    from tools.gen_synthetic import generate
    d = str(tmp_path)
    n = 6 if kind == "pangenome" else 3
    fas = generate(d, kind, n, 150000, 0.04, seed)
    dbg = run_twopaco(fas, k, os.path.join(d, "g.dbg"), threads=2)
    ref_out = os.path.join(d, "ref")
    os.makedirs(ref_out)
    run_reference_lcb(dbg, fas, k, ref_out, threads=3)
    orc = Oracle(dbg, fas, k, 150)
    orc.find_blocks(50, 200)
    out = os.path.join(d, "orc")
    orc.generate_output(out, False, 0, 50)
    assert canonical_gff(os.path.join(out, "blocks_coords.gff")) == canonical_gff(os.path.join(ref_out, "blocks_coords.gff"))
    assert filecmp.cmp(os.path.join(out, "blocks_coords.gff"), os.path.join(ref_out, "blocks_coords.gff"), shallow=False)
