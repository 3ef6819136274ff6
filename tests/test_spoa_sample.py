"""The data of spoa's own unit tests (spoa/test/data/sample.fastq.gz: 55 reads, spoa_test.cpp:20-86) through the alignment
stage's checkers, in the mode spoa's `Global` test and the sibeliaz wrapper share (kNW, 5 / -4 / -8 linear gaps,
spoa_test.cpp:245-259; sibeliaz:66).  Fixture tests/golden/spoa_sample (made by tests/golden/make_spoa_fixture.py from the
unmodified reference library): the restatement and the product's core compiled for the host must print the same MSA byte
for byte, and the MSA must have the properties spoa's Check() asserts (spoa_test.cpp:58-80); the GPU path does the same in
tests/test_zzz_gpu_first_run.py.  The consensus that test also
compares is not part of this path: the wrapper only uses the MSA rows (sibeliaz:64-100)."""
import lzma
import os
import subprocess

import pytest

from conftest import GOLDEN
from oracle_binding import REF_SPOA, poa_oracle_text

HERE = os.path.dirname(os.path.abspath(__file__))
FIX = os.path.join(GOLDEN, "spoa_sample")


@pytest.fixture(scope="module")
def sample(tmp_path_factory):
    d = tmp_path_factory.mktemp("spoa_sample")
    chunk = str(d / "block.tmp")
    with lzma.open(os.path.join(FIX, "block.tmp.xz")) as f, open(chunk, "wb") as g:
        g.write(f.read())
    with lzma.open(os.path.join(FIX, "msa.maf.xz"), "rt") as f:
        return chunk, f.read()


def check_msa_properties(maf_text, chunk):
    """spoa_test.cpp:58-80: one row per sequence, all rows of one length, no column of gaps only, rows without their gaps
    are the inputs."""
    seqs = [t for t in open(chunk).read().strip().split("@") if t and not t.startswith(">")]
    rows = [line.rsplit(" ", 1)[1] for line in maf_text.splitlines() if line.startswith("s ")]
    assert len(rows) == len(seqs) == 55
    width = len(rows[0])
    assert all(len(r) == width for r in rows)
    assert all(any(r[i] != "-" for r in rows) for i in range(width))
    assert [r.replace("-", "") for r in rows] == seqs


def test_restatement_reproduces_the_reference_msa(sample):
    chunk, want = sample
    assert poa_oracle_text(chunk) == want
    check_msa_properties(want, chunk)


@pytest.mark.skipif(not os.path.exists(REF_SPOA), reason="oracle/_ref/spoa-ref not built (needs /root/reference)")
def test_fixture_is_what_the_reference_library_prints(sample):
    chunk, want = sample
    assert subprocess.run([REF_SPOA, "--chunk", chunk, "-l", "1", "-r", "1", "-e", "-8"], check=True, stdout=subprocess.PIPE, text=True).stdout == want


def test_product_core_on_host_reproduces_the_reference_msa(sample, tmp_path):
    chunk, want = sample
    exe = str(tmp_path / "poa_core_host")
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(HERE, "poa_core_host.cpp")], check=True)
    for level in (0, 2):
        assert subprocess.run([exe, "--chunk", chunk, "--level", str(level)], check=True, stdout=subprocess.PIPE, text=True).stdout == want
