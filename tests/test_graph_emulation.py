"""The junction finder's per-position device code (sibeliaz_b200/csrc/graph_kmer.cuh) on the CPU: tests/graph_emu.cpp
compiles the header for the host exactly as written and runs the passes of graph_device.cu position by position from
eight free-running threads (atomics are real atomics).  Its junction file must equal the CPU restatement's byte for byte,
for one-word k-mers (k <= 31: the path the GPU suite has always covered -- here it validates the harness) and for every
wider table width (31 < k <= 255, two to eight words), whose slots name a k-mer by the text position of one occurrence."""
import lzma
import os
import subprocess

import pytest

from conftest import GOLDEN, ROOT
from graph_cases import write_nrich
from oracle_binding import canonical_junctions, graph_oracle_build

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="session")
def graph_emu(data_dir):
    exe = os.path.join(data_dir, "graph_emu")
    srcs = [os.path.join(HERE, "graph_emu.cpp"), os.path.join(ROOT, "sibeliaz_b200", "csrc", "graph_kmer.cuh")]
    if not os.path.exists(exe) or any(os.path.getmtime(s) > os.path.getmtime(exe) for s in srcs):
        subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", "-Wno-unknown-pragmas", "-o", exe, srcs[0]], check=True)
    return exe


def _emu(exe, fastas, k, out, abundance=0, threads=8):
    subprocess.run([exe, str(k), str(threads), str(abundance), out] + list(fastas), check=True, stdout=subprocess.PIPE, timeout=300)
    return open(out, "rb").read()


@pytest.mark.parametrize("k", [15, 31, 33, 63, 65, 95, 97, 127, 129, 161, 191, 193, 255])
def test_device_code_on_host_equals_oracle_nrich(graph_emu, tmp_path, k):
    """Every table width: 1 word (k = 15, 31), 2 (33, 63), 3 (65, 95), 4 (97, 127), 5 (129), 6 (161, 191), 7 (193), 8 (255);
    k = 63, 95, 127, 191, 255 fill their top word (no masking), the others do not."""
    fas = write_nrich(str(tmp_path))
    orc = str(tmp_path / "oracle.dbg")
    n = graph_oracle_build(fas, k, orc)
    assert n > 0
    assert _emu(graph_emu, fas, k, str(tmp_path / "emu.dbg")) == open(orc, "rb").read()


@pytest.mark.parametrize("k", [21, 33])
def test_device_code_on_host_finite_abundance(graph_emu, tmp_path, k):
    fas = write_nrich(str(tmp_path))
    orc = str(tmp_path / "oracle.dbg")
    graph_oracle_build(fas, k, orc, abundance=2)
    full = str(tmp_path / "full.dbg")
    graph_oracle_build(fas, k, full)
    assert open(orc, "rb").read() != open(full, "rb").read()  # the threshold bites on this input
    assert _emu(graph_emu, fas, k, str(tmp_path / "emu.dbg"), abundance=2) == open(orc, "rb").read()


def test_device_code_on_host_star_k33_equals_oracle_and_reference(graph_emu, star_small, tmp_path):
    """800 kbp, 14 k junctions: device code == restatement byte for byte, restatement == the compiled reference's junction
    file (committed fixture) in the label-free normal form."""
    orc = str(tmp_path / "oracle.dbg")
    graph_oracle_build(star_small.fastas, 33, orc)
    emu = str(tmp_path / "emu.dbg")
    assert _emu(graph_emu, star_small.fastas, 33, emu) == open(orc, "rb").read()
    with lzma.open(os.path.join(GOLDEN, "wide_k", "star4x200k_k33.canon.xz")) as f:
        assert canonical_junctions(emu) == f.read()


def test_single_thread_and_many_threads_agree(graph_emu, tmp_path):
    """The result must not depend on which occurrence claims a slot (the representative of a wide k-mer)."""
    fas = write_nrich(str(tmp_path))
    a = _emu(graph_emu, fas, 65, str(tmp_path / "t1.dbg"), threads=1)
    b = _emu(graph_emu, fas, 65, str(tmp_path / "t16.dbg"), threads=16)
    assert a == b
