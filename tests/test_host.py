"""CPU-only tests of the product's host side: the C-ABI library loads and exports every declared symbol, the
loader reproduces the oracle's index, the output stage reproduces the reference's files, errors map to codes."""
import ctypes
import filecmp
import os
import re
import subprocess

import numpy as np
import pytest

import sibeliaz_b200 as sb
from oracle_binding import Oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "sibeliaz_lcb.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(lcb_[a-z_0-9]+)\s*\(", hdr))
    assert len(declared) >= 18
    lib = ctypes.CDLL(sb.LIB_PATH)
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, missing
    assert set(sb.EXPORTS) <= declared
    ghdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "sibeliaz_graph.h")).read(), flags=re.S)
    gdecl = set(re.findall(r"\b(lcg_[a-z_0-9]+)\s*\(", ghdr))
    assert len(gdecl) >= 7
    assert not [s for s in sorted(gdecl) if not hasattr(lib, s)]
    ahdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "sibeliaz_align.h")).read(), flags=re.S)
    adecl = set(re.findall(r"\b(lca_[a-z_0-9]+)\s*\(", ahdr))
    assert len(adecl) >= 7
    assert not [s for s in sorted(adecl) if not hasattr(lib, s)]
    assert b"1.2.7" in ctypes.cast(sb.load_library().lcb_version(), ctypes.c_char_p).value


def test_no_torch_types_or_oracle_in_product():
    for root, _, files in os.walk(os.path.join(ROOT, "sibeliaz_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(root, f), errors="replace").read()
                assert "oracle" not in src.replace("no oracle", ""), "%s mentions the oracle" % f
                assert "torch" not in src or f == "build.py", f


@pytest.mark.parametrize("which", ["k15", "k25"])
def test_loader_matches_oracle_index(examples, which):
    case = examples[which]
    st = sb.JunctionStorage(case.graph, case.fastas, case.k, case.a)
    orc = Oracle(case.graph, case.fastas, case.k, case.a)
    oi, pi = orc.index(), st.arrays()
    for name in ("chr_off", "pos_id", "pos_bp", "next_ch", "prev_rc", "vtx_off", "occ_g"):
        assert np.array_equal(oi[name], pi[name]), name
    assert st.get_chr_number() == 8 and st.get_chr_description(0) == "Genome1.Chr1"
    assert [st.get_chr_length(c) for c in range(8)] == oi["chr_len"].tolist()


def test_abundance_filter_is_strict(star_small):
    full = sb.JunctionStorage(star_small.graph, star_small.fastas, star_small.k, 150)
    cut = sb.JunctionStorage(star_small.graph, star_small.fastas, star_small.k, 4)  # keeps vertices with < 4 occurrences
    a = cut.arrays()
    deg = np.diff(a["vtx_off"])
    assert cut.n_records < full.n_records and deg.max() == 3
    o = Oracle(star_small.graph, star_small.fastas, star_small.k, 4).index()
    assert np.array_equal(o["pos_id"], a["pos_id"]) and np.array_equal(o["occ_g"], a["occ_g"])


def test_output_stage_reproduces_reference_files(star_small, tmp_path):
    """Feed the oracle's raw block instances through the product's GenerateOutput equivalent."""
    orc = Oracle(star_small.graph, star_small.fastas, star_small.k, star_small.a)
    ob = orc.find_blocks(star_small.m, star_small.b)
    st = sb.JunctionStorage(star_small.graph, star_small.fastas, star_small.k, star_small.a)
    bf = sb.BlocksFinder(st, star_small.k)
    blocks = np.zeros(len(ob["id"]), sb.BLOCK_DTYPE)
    for f in ("id", "chr", "start", "end"):
        blocks[f] = ob[f]
    bf.blocks = blocks
    out = str(tmp_path / "out")
    found, cov = bf.generate_output(out, True, 4, min_block=star_small.m)
    assert filecmp.cmp(os.path.join(out, "blocks_coords.gff"), star_small.ref_gff, shallow=False)
    got = b""
    for i in range(4):
        got += b"== %d.tmp\n" % i + open(os.path.join(out, "%d.tmp" % i), "rb").read()
    assert got == open(star_small.ref_chunks, "rb").read()
    o_found, o_cov = orc.generate_output(str(tmp_path / "o"), False, 0, star_small.m)
    assert (found, round(cov, 9)) == (o_found, round(o_cov, 9))


def test_loader_errors(tmp_path, star_small):
    with pytest.raises(sb.LcbError) as e:
        sb.JunctionStorage(str(tmp_path / "missing.dbg"), star_small.fastas, 21)
    assert e.value.code == 2 and "Can't read the input file" in str(e.value)
    bad = tmp_path / "bad.fa"
    bad.write_text(">x\nACGTJ\n")
    with pytest.raises(sb.LcbError) as e:
        sb.JunctionStorage(star_small.graph, [str(bad)], 21)
    assert e.value.code == 3 and "invalid character 'J'" in str(e.value)
    nohdr = tmp_path / "nohdr.fa"
    nohdr.write_text("ACGT\n")
    with pytest.raises(sb.LcbError) as e:
        sb.JunctionStorage(star_small.graph, [str(nohdr)], 21)
    assert "should start with a '>'" in str(e.value)


def test_fasta_rules(tmp_path):
    """lower case is upper-cased, whitespace skipped, IUPAC kept, header = first token; empty junction file is fine."""
    fa = tmp_path / "a.fa"
    fa.write_text(">chrA some description\nacgtn\nRYK M\n>chrB\nTTTT")
    dbg = tmp_path / "empty.dbg"
    dbg.write_bytes(b"")
    st = sb.JunctionStorage(str(dbg), [str(fa)], 3)
    assert st.n_records == 0 and st.get_chr_number() == 0
    o = Oracle(str(dbg), [str(fa)], 3)
    assert o.N == 0


def test_cli_contract(star_small, tmp_path):
    cli = sb.CLI_PATH
    r = subprocess.run([cli, "-k", "24", "--graph", "x", "y.fa"], capture_output=True, text=True)
    assert r.returncode == 1 and "error:" in r.stderr and "odd" in r.stderr
    r = subprocess.run([cli, "y.fa"], capture_output=True, text=True)
    assert r.returncode == 1 and "--graph" in r.stderr
    r = subprocess.run([cli, "--graph", str(tmp_path / "nope.dbg"), star_small.fastas[0]], capture_output=True, text=True)
    assert r.returncode == 1 and r.stdout.startswith("Loading the graph...") and "error: Can't read the input file" in r.stderr
    r = subprocess.run([cli, "--version"], capture_output=True, text=True)
    assert r.returncode == 0 and "1.2.7" in r.stdout


def test_product_fails_loudly_without_gpu(star_small):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    st = sb.JunctionStorage(star_small.graph, star_small.fastas, star_small.k, star_small.a)
    with pytest.raises(sb.LcbError) as e:
        sb.BlocksFinder(st, star_small.k).find_blocks(50, 200)
    assert e.value.code == 4 and "no CPU fallback" in str(e.value)


def test_align_cli_error_contract(tmp_path):
    """sibeliaz-align: rc 1 + `error: <message>` on stderr (like sibeliaz.cpp:145-154); without a GPU it must refuse, not
    fall back to a CPU path."""
    import subprocess
    r = subprocess.run([sb.ALIGN_CLI_PATH, "--cmd", "x"], capture_output=True, text=True)
    assert r.returncode == 1 and r.stderr.startswith("error: missing -o")
    import torch
    if not torch.cuda.is_available():
        chunk = tmp_path / "0.tmp"
        chunk.write_text("> a;0;4;+;9@ACGT@> b;0;4;+;9@ACGA@\n")
        r = subprocess.run([sb.ALIGN_CLI_PATH, "--cmd", "x", "-o", str(tmp_path / "a.maf"), str(chunk)], capture_output=True, text=True)
        assert r.returncode == 1 and "no CUDA device" in r.stderr and not (tmp_path / "a.maf").exists()


def test_cli_rejects_malformed_numbers(tmp_path):
    """TCLAP (the reference's parser) refuses values that are not entirely a number of the argument's type
    (sibeliaz.cpp:37-111): so do the three drop-in binaries, before any CUDA call."""
    import subprocess
    import sibeliaz_b200 as sb
    fa = tmp_path / "a.fa"
    fa.write_text(">a\nACGT\n")
    for args in (["-k", "abc"], ["-k", " 25"], ["-k", "+25"], ["-k", "4294967297"], ["-b", "-3"], ["-t", "1x"]):
        r = subprocess.run([sb.CLI_PATH, "--graph", str(tmp_path / "g.dbg"), str(fa)] + args, capture_output=True, text=True)
        assert r.returncode == 1 and "Couldn't read argument value" in r.stderr, (args, r.stderr)
    r = subprocess.run([sb.CLI_PATH, "--graph", "x", str(fa), "-k", "24"], capture_output=True, text=True)
    assert r.returncode == 1 and "must be odd" in r.stderr
    r = subprocess.run([sb.GRAPH_CLI_PATH, "-k", "abc", "-f", "3", str(fa)], capture_output=True, text=True)
    assert r.returncode == 1 and "Couldn't read argument value" in r.stderr
    r = subprocess.run([sb.ALIGN_CLI_PATH, "--gpu", "x", "-o", str(tmp_path / "o.maf")], capture_output=True, text=True)
    assert r.returncode == 1 and "Couldn't read argument value" in r.stderr
