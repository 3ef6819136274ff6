#!/usr/bin/env python3
"""Developer helper: wall-clock of the drop-in CLI vs the compiled reference on a cached bench workload."""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
import sibeliaz_b200 as sb  # noqa: E402
from oracle_binding import REF_LCB  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "star4x10M_k21"
dbg, fas, k = bench.prepare_workload(name, 0)
args = ["--graph", dbg] + fas + ["-k", str(k), "-b", "200", "-m", "50", "-a", "150", "--noseq"]
for rep in range(3):
    t = time.perf_counter()
    r = subprocess.run([sb.CLI_PATH] + args + ["-o", "/tmp/cli_out", "--stats"], capture_output=True, text=True, env=dict(os.environ, LCB_LOAD_TRACE="1"))
    dt = time.perf_counter() - t
    print("B200 CLI rep %d: %.3f s rc=%d" % (rep, dt, r.returncode))
    print("   ", r.stderr.strip().splitlines()[-1][:900])
if "--ref" in sys.argv:
    for th in (32,):
        t = time.perf_counter()
        subprocess.run([REF_LCB] + args + ["-o", "/tmp/ref_out", "-t", str(th)], capture_output=True)
        print("reference -t %d: %.3f s" % (th, time.perf_counter() - t))
    import filecmp
    print("GFF identical:", filecmp.cmp("/tmp/cli_out/blocks_coords.gff", "/tmp/ref_out/blocks_coords.gff", shallow=False))
