#!/usr/bin/env python3
"""Developer helper: whole-binary wall clock of the drop-in CLI on the 4x100 Mbp k=25 headline input (the junction file
is made once with the reference twopaco and passed in), optionally next to the compiled reference."""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from tools.gen_synthetic import generate  # noqa: E402
import sibeliaz_b200 as sb  # noqa: E402

dbg = sys.argv[1]
d = "/tmp/h"
os.makedirs(d, exist_ok=True)
fas = generate(d, "star", 4, 100000000, 0.05, 1)
args = ["--graph", dbg] + fas + ["-k", "25", "-b", "200", "-m", "50", "-t", str(min(32, os.cpu_count() or 1)), "--abundance", "150", "--noseq"]
extra = [a for a in sys.argv[2:] if a.startswith("--") and a != "--ref"]
for rep in range(3):
    t = time.time()
    r = subprocess.run([sb.CLI_PATH] + args + ["-o", d + "/out", "--stats"] + extra, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                       env=dict(os.environ, LCB_LOAD_TRACE="1"))
    print("whole binary %.3fs rc=%d" % (time.time() - t, r.returncode))
    print(r.stderr[-3400:] if rep == 2 else r.stderr.strip().splitlines()[-1][:1200])
if "--ref" in sys.argv:
    from oracle_binding import REF_LCB
    import filecmp
    th = min(32, os.cpu_count() or 1)
    t = time.time()
    subprocess.run([REF_LCB, "--graph", dbg] + fas + ["-k", "25", "-b", "200", "-m", "50", "-t", str(th), "-a", "150", "--noseq", "-o", d + "/ref"], stdout=subprocess.DEVNULL)
    print("reference -t %d whole binary %.2fs" % (th, time.time() - t))
    print("GFF byte-identical:", filecmp.cmp(d + "/out/blocks_coords.gff", d + "/ref/blocks_coords.gff", shallow=False))
