"""Consumer compatibility: the shipped downstream tool maf2synteny, compiled unmodified from
the reference's sources (oracle/Makefile -> oracle/_ref/maf2synteny), must ingest blocks_coords.gff
(maf2synteny/src/maf_tools.cpp:122-198: `##sequence-region` lengths, 9 tab-separated columns, `ID=<n>` attribute)."""
import filecmp
import os
import subprocess

import pytest

from oracle_binding import ORACLE_DIR

M2S = os.path.join(ORACLE_DIR, "_ref", "maf2synteny")
needs_m2s = pytest.mark.skipif(not os.path.exists(M2S), reason="oracle/_ref/maf2synteny not built (needs /root/reference)")


def run_m2s(gff, out):
    os.makedirs(out, exist_ok=True)
    r = subprocess.run([M2S, "-o", out, "-b", "1000,5000", gff], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stderr
    files = []
    for scale in ("1000", "5000"):
        for name in ("blocks_coords.txt", "genomes_permutations.txt", "coverage_report.txt"):
            p = os.path.join(out, scale, name)
            assert os.path.getsize(p) > 0, p
            files.append(p)
    return files


@needs_m2s
def test_maf2synteny_ingests_golden_gff(examples, tmp_path):
    """The format contract itself, on the reference's own golden file (which the product reproduces byte for byte)."""
    files = run_m2s(examples["k25"].ref_gff, str(tmp_path / "golden"))
    perms = open(files[1]).read()
    assert perms.count(">") == 8  # every chromosome of both genomes (2 x 4) carries synteny blocks


@needs_m2s
@pytest.mark.gpu
def test_maf2synteny_on_product_gff(star_small, tmp_path):
    """maf2synteny's result on the GPU path's GFF equals its result on the reference's GFF."""
    import sibeliaz_b200 as sb
    st = sb.JunctionStorage(star_small.graph, star_small.fastas, star_small.k, star_small.a)
    bf = sb.BlocksFinder(st, star_small.k)
    bf.find_blocks(star_small.m, star_small.b)
    out = str(tmp_path / "lcb")
    bf.generate_output(out, False, 0)
    bf.close()
    mine = run_m2s(os.path.join(out, "blocks_coords.gff"), str(tmp_path / "mine"))
    ref = run_m2s(star_small.ref_gff, str(tmp_path / "ref"))
    for a, b in zip(mine, ref):
        assert filecmp.cmp(a, b, shallow=False), a
