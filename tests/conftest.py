import lzma
import os
import shutil
import sys
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def _unxz(src, dst):
    if not os.path.exists(dst):
        with lzma.open(src, "rb") as f, open(dst + ".part", "wb") as g:
            shutil.copyfileobj(f, g)
        os.replace(dst + ".part", dst)
    return dst


class Case:
    """One parity case: junction file + FASTA files + parameters (+ reference GFF when a fixture holds one)."""

    def __init__(self, name, graph, fastas, k, b=200, m=50, a=150, ref_gff=None):
        self.name, self.graph, self.fastas, self.k, self.b, self.m, self.a, self.ref_gff = name, graph, fastas, k, b, m, a, ref_gff


@pytest.fixture(scope="session")
def data_dir():
    d = os.path.join(tempfile.gettempdir(), "sibeliaz_b200_testdata")
    os.makedirs(d, exist_ok=True)
    return d


@pytest.fixture(scope="session")
def examples(data_dir):
    ex = os.path.join(GOLDEN, "examples")
    out = os.path.join(data_dir, "examples")
    os.makedirs(out, exist_ok=True)
    fas = [_unxz(os.path.join(ex, "genome%d.fa.xz" % i), os.path.join(out, "genome%d.fa" % i)) for i in (1, 2)]
    files = {n: _unxz(os.path.join(ex, n + ".xz"), os.path.join(out, n))
             for n in ("k25.dbg", "k15.dbg", "golden_k25_blocks_coords.gff", "ref_k25_blocks_coords.gff", "ref_k15_blocks_coords.gff")}
    return {
        "k25": Case("examples_k25", files["k25.dbg"], fas, 25, ref_gff=files["golden_k25_blocks_coords.gff"]),
        "k15": Case("examples_k15", files["k15.dbg"], fas, 15, ref_gff=files["ref_k15_blocks_coords.gff"]),
    }


@pytest.fixture(scope="session")
def star_small(data_dir):
    from tools.gen_synthetic import generate
    st = os.path.join(GOLDEN, "star4x200k")
    out = os.path.join(data_dir, "star4x200k")
    os.makedirs(out, exist_ok=True)
    fas = generate(out, "star", 4, 200000, 0.05, 7)
    dbg = _unxz(os.path.join(st, "k21.dbg.xz"), os.path.join(out, "k21.dbg"))
    gff = _unxz(os.path.join(st, "ref_blocks_coords.gff.xz"), os.path.join(out, "ref_blocks_coords.gff"))
    chunks = _unxz(os.path.join(st, "ref_chunks.tmp.xz"), os.path.join(out, "ref_chunks.tmp"))
    c = Case("star4x200k", dbg, fas, 21, ref_gff=gff)
    c.ref_chunks = chunks
    return c


def canonical_gff(path):
    """(seq, start, end, strand, id-class) rows with block ids renamed by first appearance of their sorted
    member set -- the north star's set-equality criterion."""
    rows = {}
    with open(path) as f:
        for line in f:
            if line.startswith("#"):
                continue
            p = line.rstrip("\n").split("\t")
            rows.setdefault(p[8], []).append((p[0], int(p[3]), int(p[4]), p[6]))
    return sorted(tuple(sorted(v)) for v in rows.values())
