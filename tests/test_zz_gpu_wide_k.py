"""GPU parity of the junction finder for vertex sizes beyond one 64-bit word (31 < k <= 255; graph_kmer.cuh, Kmer<W> with
W = 2 .. 8) and of the LCB path behind it at k = 33, through the drop-in binaries (which call the C ABI), each run in its
own process with a time limit.  Checkers: the CPU restatement (byte for byte) and fixtures made by the compiled reference
(tests/golden/wide_k).  The same device code runs on the CPU in tests/test_graph_emulation.py.  Green on a B200 in round-2
session 25 (profiles/gpu_suite_widek_r2.log); the file name sorts last because every case starts its own process."""
import lzma
import os
import subprocess

import pytest

import sibeliaz_b200 as sb
from conftest import GOLDEN
from graph_cases import write_nrich
from oracle_binding import canonical_junctions, graph_oracle_build

pytestmark = pytest.mark.gpu


def _twopaco(fastas, k, out, tmp, extra=()):
    r = subprocess.run([sb.GRAPH_CLI_PATH, "--tmpdir", str(tmp), "-t", "4", "-k", str(k), "--filtermemory", "4", "-o", out] + list(extra) + list(fastas),
                       capture_output=True, text=True, timeout=180)
    assert r.returncode == 0, r.stderr[-2000:]
    return open(out, "rb").read()


@pytest.mark.parametrize("k", [33, 63, 65, 97, 127, 129, 191, 193, 255])
def test_wide_k_nrich_equals_oracle(tmp_path, k):
    fas = write_nrich(str(tmp_path))
    orc = str(tmp_path / "oracle.dbg")
    assert graph_oracle_build(fas, k, orc) > 0
    assert _twopaco(fas, k, str(tmp_path / "gpu.dbg"), tmp_path) == open(orc, "rb").read()


def test_wide_k_finite_abundance(tmp_path):
    fas = write_nrich(str(tmp_path))
    orc = str(tmp_path / "oracle.dbg")
    graph_oracle_build(fas, 33, orc, abundance=2)
    assert _twopaco(fas, 33, str(tmp_path / "gpu.dbg"), tmp_path, ["-a", "2"]) == open(orc, "rb").read()


def test_k33_star_graph_then_blocks_equal_reference(star_small, tmp_path):
    """twopaco drop-in at k = 33 == restatement (bytes) == compiled reference (normal form); sibeliaz-lcb drop-in on that
    junction file == the reference sibeliaz-lcb's GFF; and the fused binary (--construct) writes the same GFF."""
    orc = str(tmp_path / "oracle.dbg")
    graph_oracle_build(star_small.fastas, 33, orc)
    dbg = str(tmp_path / "gpu.dbg")
    assert _twopaco(star_small.fastas, 33, dbg, tmp_path) == open(orc, "rb").read()
    with lzma.open(os.path.join(GOLDEN, "wide_k", "star4x200k_k33.canon.xz")) as f:
        assert canonical_junctions(dbg) == f.read()
    with lzma.open(os.path.join(GOLDEN, "wide_k", "star4x200k_k33.gff.xz")) as f:
        want = f.read()
    for mode, args in (("file", ["--graph", dbg]), ("fused", ["--construct"])):
        out = str(tmp_path / mode)
        r = subprocess.run([sb.CLI_PATH] + args + star_small.fastas + ["-k", "33", "-b", "200", "-o", out, "-m", "50", "-t", "4",
                            "--abundance", "150", "--noseq"], capture_output=True, text=True, timeout=180)
        assert r.returncode == 0, r.stderr[-2000:]
        assert open(os.path.join(out, "blocks_coords.gff"), "rb").read() == want, mode
