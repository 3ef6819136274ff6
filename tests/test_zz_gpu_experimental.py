"""Variants of GPU paths that are not the default configuration, each in its own process with a time limit so that it can
neither break nor wedge the rest of the suite.  The file name sorts last."""
import os
import subprocess
import sys

import pytest

from oracle_binding import Oracle, poa_oracle_text

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
HEAD = "##maf version=1\n# sibeliaz v1.2.7 \n# cmd=%s\n"

CHILD = """
import sys
sys.path.insert(0, %r)
import sibeliaz_b200 as sb
st = sb.global_alignment(sys.argv[2:], "x", sys.argv[1])
print(st)
"""


@pytest.mark.parametrize("threshold", ["256", "0"])
def test_cta_rows_threshold_does_not_change_the_alignment(examples, tmp_path, threshold):
    """One block per CTA for every block whose longest copy has >= 256 characters (default: 2048), and never (0)."""
    case = examples["k25"]
    orc = Oracle(case.graph, case.fastas, case.k, case.a)
    orc.find_blocks(case.m, case.b)
    out = str(tmp_path / "lcb")
    orc.generate_output(out, True, 256, case.m)
    names = ["%d.tmp" % i for i in (3, 7, 11)]
    files = [os.path.join(out, n) for n in names]
    maf = str(tmp_path / "cta.maf")
    r = subprocess.run([sys.executable, "-c", CHILD % ROOT, maf] + files, env=dict(os.environ, LCA_CTA_ROWS=threshold),
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=240)
    assert r.returncode == 0, r.stderr[-2000:]
    want = HEAD % "x" + "".join(poa_oracle_text(os.path.join(out, n)) for n in sorted(names))
    assert open(maf).read() == want
