#!/usr/bin/env python3
"""Developer helper: whole-binary wall clock of `sibeliaz-lcb --construct` (FASTA -> blocks in one process) on the
4x100 Mbp k=25 headline input, next to the two separate drop-in binaries."""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools.gen_synthetic import generate  # noqa: E402
import sibeliaz_b200 as sb  # noqa: E402

length = int(sys.argv[1]) if len(sys.argv) > 1 else 100000000
k = sys.argv[2] if len(sys.argv) > 2 else "25"
d = "/tmp/hf"
os.makedirs(d, exist_ok=True)
fas = generate(d, "star", 4, length, 0.05, 1)
common = ["-k", k, "-b", "200", "-m", "50", "-t", str(min(32, os.cpu_count() or 1)), "--abundance", "150", "--noseq", "--stats"]
for rep in range(3):
    t = time.time()
    r = subprocess.run([sb.CLI_PATH, "--construct"] + fas + common + ["-o", d + "/fused"], capture_output=True, text=True, env=dict(os.environ, LCB_LOAD_TRACE="1"))
    print("fused whole binary %.3f s rc=%d" % (time.time() - t, r.returncode))
    print(r.stderr[-2600:] if rep == 2 else r.stderr.strip().splitlines()[-1][:1000], flush=True)
t = time.time()
subprocess.run([sb.GRAPH_CLI_PATH, "--tmpdir", d, "-t", "16", "-k", k, "--filtermemory", "4", "-o", d + "/g.dbg"] + fas, check=True, capture_output=True)
t1 = time.time()
subprocess.run([sb.CLI_PATH, "--graph", d + "/g.dbg"] + fas + common + ["-o", d + "/two"], check=True, capture_output=True)
t2 = time.time()
print("separate binaries: twopaco %.3f s + sibeliaz-lcb %.3f s = %.3f s" % (t1 - t, t2 - t1, t2 - t))
import filecmp
print("fused GFF == two-step GFF:", filecmp.cmp(d + "/fused/blocks_coords.gff", d + "/two/blocks_coords.gff", shallow=False))
