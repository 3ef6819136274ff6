"""GPU parity of the junction finder (include/sibeliaz_graph.h) through the C ABI: its junction file must equal the CPU
restatement's BYTE FOR BYTE (both use the deterministic labelling of graph_oracle.cpp), hence the compiled reference's in
the label-free normal form, and sibeliaz-lcb must produce the reference's blocks from it."""
import os
import subprocess

import numpy as np
import pytest

import sibeliaz_b200 as sb
from graph_cases import write_nrich
from oracle_binding import canonical_junctions, graph_oracle_build

pytestmark = pytest.mark.gpu


def _same_as_oracle(fastas, k, tmp_path):
    g = sb.JunctionGraph(fastas, k)
    mine = g.write(str(tmp_path / "gpu.dbg"))
    orc = str(tmp_path / "oracle.dbg")
    n = graph_oracle_build(fastas, k, orc)
    assert g.stats["n_junctions"] == n and g.stats["kernel_launches"] > 0
    assert open(mine, "rb").read() == open(orc, "rb").read()
    return g, mine


@pytest.mark.parametrize("which", ["k15", "k25"])
def test_examples_equal_oracle_and_reference_fixture(examples, which, tmp_path):
    case = examples[which]
    g, mine = _same_as_oracle(case.fastas, case.k, tmp_path)
    assert canonical_junctions(mine) == canonical_junctions(case.graph, str(tmp_path / "ref.canon"))
    j = g.junctions()
    assert len(j["id"]) == g.stats["n_junctions"] and (np.diff(j["chr"].astype(np.int64)) >= 0).all()


def test_star_equal_oracle_and_reference_fixture(star_small, tmp_path):
    g, mine = _same_as_oracle(star_small.fastas, star_small.k, tmp_path)
    assert canonical_junctions(mine) == canonical_junctions(star_small.graph, str(tmp_path / "ref.canon"))


@pytest.mark.parametrize("k", [9, 15, 21, 31])
def test_nrich_multi_record_input_equals_oracle(tmp_path, k):
    _same_as_oracle(write_nrich(str(tmp_path)), k, tmp_path)


def test_in_memory_records_and_edge_cases(tmp_path):
    recs = [b"ACGTTGCATGTCAGTNACGTTGCATGTCAGT", b"", b"acgtt", b"ACGTTGCATGTCAGT", b"NNNNNNNNNNNNNNNNNNNN"]
    fa = str(tmp_path / "m.fa")
    with open(fa, "w") as f:
        for i, r in enumerate(recs):
            f.write(">r%d\n%s\n" % (i, r.decode()))
    g = sb.JunctionGraph(sequences=recs, k=5)
    orc = str(tmp_path / "oracle.dbg")
    graph_oracle_build([fa], 5, orc)
    assert open(g.write(str(tmp_path / "gpu.dbg")), "rb").read() == open(orc, "rb").read()
    with pytest.raises(sb.LcbError):
        sb.JunctionGraph(sequences=recs, k=6)     # even k
    with pytest.raises(sb.LcbError):
        sb.JunctionGraph(sequences=recs, k=257)   # beyond eight 64-bit words (wider k-mers: tests/test_zz_gpu_wide_k.py)


def test_finite_abundance_threshold(star_small, tmp_path):
    g = sb.JunctionGraph(star_small.fastas, star_small.k, abundance=3)
    orc = str(tmp_path / "oracle.dbg")
    graph_oracle_build(star_small.fastas, star_small.k, orc, abundance=3)
    assert open(g.write(str(tmp_path / "gpu.dbg")), "rb").read() == open(orc, "rb").read()


def test_pipeline_twopaco_then_lcb_reproduces_reference_blocks(star_small, tmp_path):
    """Both drop-in binaries in sequence, as the sibeliaz wrapper runs them (sibeliaz:145-146)."""
    dbg = str(tmp_path / "de_bruijn_graph.dbg")
    r = subprocess.run([sb.GRAPH_CLI_PATH, "--tmpdir", str(tmp_path), "-t", "4", "-k", str(star_small.k), "--filtermemory", "4", "-o", dbg]
                       + star_small.fastas, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = str(tmp_path / "out")
    r = subprocess.run([sb.CLI_PATH, "--graph", dbg] + star_small.fastas + ["-k", str(star_small.k), "-b", "200", "-o", out, "-m", "50",
                        "-t", "4", "--abundance", "150", "--noseq"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert open(os.path.join(out, "blocks_coords.gff"), "rb").read() == open(star_small.ref_gff, "rb").read()


# ---- fused pipeline: FASTA -> junctions -> junction index -> blocks, all on the device -------------------------------
def _fused_vs_file_path(fastas, k, a, m, b, tmp_path):
    """Blocks of the fused pipeline == blocks of the file path on the junction file of the same graph (same vertex ids),
    seed list included: the device-built index is the host loader's index."""
    g = sb.JunctionGraph(fastas, k)
    dbg = g.write(str(tmp_path / "g.dbg"))
    st = sb.JunctionStorage(dbg, fastas, k, a)
    bf = sb.BlocksFinder(st, k)
    blocks = bf.find_blocks(m, b)
    seeds = bf.seeds()
    fs = sb.FusedStorage(fastas, k, a)
    ff = sb.BlocksFinder(fs, k)
    fblocks = ff.find_blocks(m, b)
    fseeds = ff.seeds()
    assert ff.stats["n_records"] == st.n_records and ff.stats["n_vertices"] == st.n_vertices
    for key in ("vid", "ch", "count", "rank", "res_pos", "res_chr"):
        assert np.array_equal(seeds[key], fseeds[key]), key
    assert len(blocks) == len(fblocks) > 0 and blocks.tobytes() == fblocks.tobytes()
    return ff, fs


def test_fused_pipeline_star(star_small, tmp_path):
    ff, fs = _fused_vs_file_path(star_small.fastas, star_small.k, star_small.a, star_small.m, star_small.b, tmp_path)
    out = str(tmp_path / "out")
    ff.generate_output(out, False, 0)
    assert open(os.path.join(out, "blocks_coords.gff"), "rb").read() == open(star_small.ref_gff, "rb").read()


@pytest.mark.parametrize("which", ["k15", "k25"])
def test_fused_pipeline_examples(examples, which, tmp_path):
    case = examples[which]
    ff, fs = _fused_vs_file_path(case.fastas, case.k, case.a, case.m, case.b, tmp_path)
    out = str(tmp_path / "out")
    ff.generate_output(out, False, 0)
    assert open(os.path.join(out, "blocks_coords.gff"), "rb").read() == open(case.ref_gff, "rb").read()


def test_fused_pipeline_abundance_and_nrich(star_small, tmp_path):
    _fused_vs_file_path(star_small.fastas, star_small.k, 4, star_small.m, star_small.b, tmp_path)   # filter bites
    (tmp_path / "n").mkdir()
    _fused_vs_file_path(write_nrich(str(tmp_path / "n")), 15, 150, 50, 200, tmp_path / "n")


def test_fused_cli(star_small, tmp_path):
    out = str(tmp_path / "out")
    r = subprocess.run([sb.CLI_PATH, "--construct"] + star_small.fastas + ["-k", str(star_small.k), "-b", "200", "-o", out, "-m", "50",
                        "--abundance", "150", "--noseq", "--stats"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert open(os.path.join(out, "blocks_coords.gff"), "rb").read() == open(star_small.ref_gff, "rb").read()
