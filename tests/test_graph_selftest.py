"""`twopaco --test` restated (tests/graph_selftest.py; TwoPaCo/src/graphconstructor/test.cpp:163-254 with the parameters of
constructor.cpp:147: chromosomes of 9000 characters, 6 of them, k = 3, 5, 7, 9, change rate 0.05, indel rate 0.1): the
positions in the junction file must be the naively computed ones.  Run against the compiled reference (this pins the
restated naive definition itself), the CPU restatement, and the device code compiled for the host."""
import os
import random
import subprocess

import pytest

from graph_selftest import check, make_case, write_fasta
from oracle_binding import REF_TWOPACO, graph_oracle_build, run_twopaco
from test_graph_emulation import graph_emu  # noqa: F401  (fixture)

KS = (3, 5, 7, 9)


@pytest.mark.skipif(not os.path.exists(REF_TWOPACO), reason="compiled reference twopaco not present")
def test_naive_definition_holds_for_the_compiled_reference(tmp_path):
    rnd = random.Random(11)
    chrs = make_case(rnd)
    fa = write_fasta(str(tmp_path / "test.fa"), chrs)
    for k in KS:
        assert check(run_twopaco([fa], k, str(tmp_path / ("ref%d.dbg" % k)), threads=4), chrs, k) > 12


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_selftest_restatement_and_device_code(graph_emu, tmp_path, seed):  # noqa: F811
    rnd = random.Random(seed)
    chrs = make_case(rnd)
    fa = write_fasta(str(tmp_path / "test.fa"), chrs)
    for k in KS + (15, 33):
        orc = str(tmp_path / ("o%d.dbg" % k))
        graph_oracle_build([fa], k, orc)
        assert check(orc, chrs, k) > 12
        emu = str(tmp_path / ("e%d.dbg" % k))
        subprocess.run([graph_emu, str(k), "4", "0", emu, fa], check=True, stdout=subprocess.PIPE, timeout=300)
        assert open(emu, "rb").read() == open(orc, "rb").read()
