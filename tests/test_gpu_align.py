"""Alignment stage on the GPU (include/sibeliaz_align.h) through the C ABI against the CPU
restatement (oracle/poa_oracle.cpp) and the reference's shipped golden alignment.maf.  Byte-exact."""
import lzma
import os

import pytest

from oracle_binding import Oracle, maf_paragraphs, poa_oracle_text, write_chunk

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "examples", "golden_k25_alignment.maf.xz")
HEAD = "##maf version=1\n# sibeliaz v1.2.7 \n# cmd=%s\n"


@pytest.fixture(scope="module")
def example_chunks(examples, tmp_path_factory):
    case = examples["k25"]
    orc = Oracle(case.graph, case.fastas, case.k, case.a)
    orc.find_blocks(case.m, case.b)
    out = str(tmp_path_factory.mktemp("lcb_chunks"))
    orc.generate_output(out, True, 256, case.m)
    return out


def test_edge_cases_equal_restatement(tmp_path):
    import sibeliaz_b200 as sb
    blocks = [
        [("a;0;4;+;9", "ACGT")],
        [("a;0;1;+;9", "A"), ("b;0;1;+;9", "C"), ("c;0;1;+;9", "A")],
        [("a;0;8;+;9", "ACGTACGT"), ("b;0;3;+;9", "CGT"), ("c;0;12;-;30", "TTACGTACGTTT"), ("d;0;8;+;9", "ACGAACGT")],
        [("a;0;6;+;9", "acgtNN"), ("b;0;6;+;9", "ACGTNN"), ("c;0;7;+;9", "acgRtNN")],
        [("a;0;5;+;9", "AAAAA"), ("b;0;5;+;9", "TTTTT"), ("c;0;5;+;9", "AATTA"), ("d;0;5;+;9", "TTAAT"), ("e;0;5;+;9", "ATATA")],
        [("a;0;70;+;99", "ACGTTGCA" * 8 + "ACGTTG"), ("b;0;64;+;99", "ACGTTGCA" * 8), ("c;0;33;+;99", "ACGTTGCAA" * 3 + "ACGTTG")],
    ]
    f = write_chunk(str(tmp_path / "edge.tmp"), blocks)
    out = str(tmp_path / "edge.maf")
    st = sb.global_alignment([f], "edge", out)
    assert open(out).read() == HEAD % "edge" + poa_oracle_text(f)
    assert st["n_blocks"] == len(blocks) and st["kernel_launches"] >= 1
    rows, _ = sb.align_blocks([[s.encode() for _, s in b] for b in blocks])  # the array entry point gives the same rows
    want = [line.rsplit(" ", 1)[1] for line in poa_oracle_text(f).splitlines() if line.startswith("s ")]
    assert [r.decode() for b in rows for r in b] == want


def test_chunk_files_equal_restatement(example_chunks, tmp_path):
    import sibeliaz_b200 as sb
    names = ["%d.tmp" % i for i in (3, 7, 11, 42, 100, 200, 255)]
    files = [os.path.join(example_chunks, n) for n in names]
    out = str(tmp_path / "slice.maf")
    sb.global_alignment(files, "genome1.fa genome2.fa", out)
    want = HEAD % "genome1.fa genome2.fa" + "".join(poa_oracle_text(os.path.join(example_chunks, n)) for n in sorted(names))
    assert open(out).read() == want


def _filtered_chunks(src_dir, dst_dir, longest):
    """Copies of the chunk files without the blocks that hold a copy of `longest` characters or more."""
    os.makedirs(dst_dir, exist_ok=True)
    files, kept = [], 0
    for n in os.listdir(src_dir):
        if not n.endswith(".tmp"):
            continue
        lines = [ln for ln in open(os.path.join(src_dir, n)) if max(len(t) for t in ln.split("@")) < longest]
        kept += len(lines)
        with open(os.path.join(dst_dir, n), "w") as f:
            f.writelines(lines)
        files.append(os.path.join(dst_dir, n))
    return files, kept


def test_examples_alignment_maf_contains_the_golden(example_chunks, tmp_path):
    """The 256 chunk files of the examples: every paragraph of the reference's shipped alignment.maf must come out byte for
    byte, in the golden's order: all 1350 blocks, all 1332 golden paragraphs (the 18 longest blocks, 8 x 10 - 27.6 kbp, are
    the ones the golden lacks; they run one block per CTA).  LCA_TEST_FULL=0 leaves out blocks with a copy of 6 kbp or more."""
    import sibeliaz_b200 as sb
    full = os.environ.get("LCA_TEST_FULL", "1") == "1"
    files, kept = _filtered_chunks(example_chunks, str(tmp_path / "chunks"), 10 ** 9 if full else 6000)
    out = str(tmp_path / "alignment.maf")
    st = sb.global_alignment(files, "genome1.fa genome2.fa", out)
    mine = maf_paragraphs(out)
    golden = maf_paragraphs(lzma.open(GOLDEN, "rt").read(), is_text=True)
    assert len(golden) == 1332 and st["n_blocks"] == kept == len(mine) and kept >= (1350 if full else 1200)
    common = [k for k in mine if k in golden]
    assert len(common) >= (1332 if full else 1200)
    assert all(mine[k] == golden[k] for k in common)
    assert common == [k for k in golden if k in mine]  # same relative order as the wrapper's sorted concatenation


def test_align_cli(example_chunks, tmp_path):
    """sibeliaz-align, the one command that replaces the wrapper's global_alignment(): same file as the library call."""
    import subprocess
    import sibeliaz_b200 as sb
    names = ["%d.tmp" % i for i in (3, 7, 11)]
    files = [os.path.join(example_chunks, n) for n in names]
    out = str(tmp_path / "cli.maf")
    r = subprocess.run([sb.ALIGN_CLI_PATH, "--cmd", "genome1.fa genome2.fa", "-o", out, "--stats"] + files, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    want = HEAD % "genome1.fa genome2.fa" + "".join(poa_oracle_text(os.path.join(example_chunks, n)) for n in sorted(names))
    assert open(out).read() == want
    assert '"blocks": ' in r.stderr
