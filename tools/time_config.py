#!/usr/bin/env python3
"""Developer helper: one BASELINE configuration at full size through the drop-in binaries, next to the compiled
reference (oracle/_ref) on the same box: wall clocks and byte comparison of blocks_coords.gff.

    python tools/time_config.py --kind mammal --genomes 8 --length 100000000 --rate 0.03 --seed 3 --k 25
"""
import argparse
import filecmp
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from tools.gen_synthetic import generate  # noqa: E402
import sibeliaz_b200 as sb  # noqa: E402
from oracle_binding import REF_LCB, REF_TWOPACO  # noqa: E402


def run(cmd, limit, **kw):
    t = time.time()
    try:
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=limit, **kw)
        return time.time() - t, r.returncode, r
    except subprocess.TimeoutExpired:
        return time.time() - t, -9, None


def same(a, b):
    return os.path.exists(a) and os.path.exists(b) and filecmp.cmp(a, b, shallow=False)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--kind", default="mammal")
    ap.add_argument("--genomes", type=int, default=8)
    ap.add_argument("--length", type=int, default=100000000)
    ap.add_argument("--rate", type=float, default=0.03)
    ap.add_argument("--seed", type=int, default=3)
    ap.add_argument("--k", type=int, default=25)
    ap.add_argument("--dir", default="/tmp/cfg")
    ap.add_argument("--ref-limit", type=int, default=400)
    ap.add_argument("--no-ref", action="store_true")
    ap.add_argument("--ref-twopaco", action="store_true")
    a = ap.parse_args()
    d = os.path.join(a.dir, "%s_%dx%d_k%d" % (a.kind, a.genomes, a.length, a.k))
    os.makedirs(d, exist_ok=True)
    t = time.time()
    fas = generate(d, a.kind, a.genomes, a.length, a.rate, a.seed)
    print("generated %d x %d bp (%s) in %.1fs" % (a.genomes, a.length, a.kind, time.time() - t), flush=True)
    threads = min(32, os.cpu_count() or 1)
    common = ["-k", str(a.k), "-b", "200", "-m", "50", "--abundance", "150", "--noseq"]
    for rep in range(2):
        dt, rc, r = run([sb.CLI_PATH, "--construct"] + fas + common + ["-t", str(threads), "-o", d + "/fused", "--stats"], 600,
                        env=dict(os.environ, LCB_LOAD_TRACE="1"))
        print("B200 fused binary (FASTA -> blocks_coords.gff): %.2fs rc=%d" % (dt, rc), flush=True)
        if r is not None:
            print((r.stderr.strip().splitlines() or [""])[-1][:1500], flush=True)
            print(r.stdout.strip().splitlines()[-2:], flush=True)
    dt, rc, r = run([sb.GRAPH_CLI_PATH, "--tmpdir", d, "-t", str(threads), "-k", str(a.k), "--filtermemory", "8", "-o", d + "/b200.dbg"] + fas, 600)
    print("B200 twopaco: %.2fs rc=%d" % (dt, rc), flush=True)
    dt, rc, r = run([sb.CLI_PATH, "--graph", d + "/b200.dbg"] + fas + common + ["-t", str(threads), "-o", d + "/two"], 600)
    print("B200 sibeliaz-lcb on that junction file: %.2fs rc=%d" % (dt, rc), flush=True)
    print("fused GFF == two-step GFF:", same(d + "/fused/blocks_coords.gff", d + "/two/blocks_coords.gff"), flush=True)
    if a.no_ref:
        return
    ref_dbg = d + "/b200.dbg"  # same junction file for both (SURVEY 8c caveat 1) unless --ref-twopaco
    if a.ref_twopaco:
        dt, rc, r = run([REF_TWOPACO, "--tmpdir", d, "-t", str(threads), "-k", str(a.k), "--filtermemory", "8", "-o", d + "/ref.dbg"] + fas, a.ref_limit)
        print("reference twopaco -t %d: %.2fs rc=%d" % (threads, dt, rc), flush=True)
        if rc:
            return
        ref_dbg = d + "/ref.dbg"
    dt, rc, r = run([REF_LCB, "--graph", ref_dbg] + fas + common + ["-t", str(threads), "-o", d + "/ref"], a.ref_limit)
    print("reference sibeliaz-lcb -t %d: %.2fs rc=%d" % (threads, dt, rc), flush=True)
    if rc:
        return
    print((r.stdout.strip().splitlines() or [""])[-2:], flush=True)
    print("GFF byte-identical to the reference's:", same(d + "/fused/blocks_coords.gff", d + "/ref/blocks_coords.gff"), flush=True)


if __name__ == "__main__":
    main()
