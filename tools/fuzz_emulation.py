#!/usr/bin/env python3
"""Developer tool (CPU only): randomised differential test of the traversal DEVICE CODE against the oracle.

Every case draws an input kind (star / mammal / pangenome, tools/gen_synthetic.py), genome count, length, divergence, k,
-b, -m, abundance and a pattern of used edges, builds the junction file with the CPU restatement of the junction finder,
and runs `process_seed` of lcb_lean.cuh / lcb_traverse.cuh exactly as written under the lockstep warp emulator
(tests/cuda_emu.h, tests/trav_emu.cpp) on a sample of the seeds.  Each evaluation's bestInstance list must equal the
oracle's epoch-threshold Process (tests/test_traversal_emulation.py::emulate_and_compare).  One line per case.

    python tools/fuzz_emulation.py --cases 40 --seed 1 [--evals 250] [--log profiles/fuzz_emulation_r2.log]
"""
import argparse
import os
import random
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=20)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--evals", type=int, default=250)
    ap.add_argument("--log", default=None)
    a = ap.parse_args()
    from conftest import Case
    from oracle_binding import graph_oracle_build
    from test_traversal_emulation import emulate_and_compare
    from tools.gen_synthetic import generate
    tests = os.path.join(ROOT, "tests")
    work = tempfile.mkdtemp(prefix="lcb_fuzz_")
    exe = os.path.join(work, "trav_emu")
    subprocess.run(["g++", "-O2", "-std=c++20", "-pthread", "-I", os.path.join(tests, "emu_include"), "-o", exe, os.path.join(tests, "trav_emu.cpp")], check=True)
    rnd = random.Random(a.seed)
    log = open(a.log, "a") if a.log else None

    def say(msg):
        print(msg, flush=True)
        if log:
            log.write(msg + "\n")
            log.flush()

    say("# tools/fuzz_emulation.py --cases %d --seed %d --evals %d" % (a.cases, a.seed, a.evals))
    total = bad = 0
    for c in range(a.cases):
        kind = rnd.choice(["star", "star", "mammal", "mammal", "pangenome"])
        genomes = rnd.choice([2, 3, 4, 6, 8, 12] if kind != "pangenome" else [4, 8, 16])
        length = rnd.choice([20000, 40000, 80000]) if kind != "pangenome" else rnd.choice([150000, 250000])
        rate = rnd.choice([0.01, 0.03, 0.05, 0.1])
        k = rnd.choice([11, 15, 21, 25, 31, 35])
        b = rnd.choice([50, 100, 200, 200, 400])
        m = rnd.choice([20, 50, 50, 100])
        ab = rnd.choice([150, 150, 150, 8, 3])
        period, run = rnd.choice([(150, 40), (60, 10), (400, 200), (1000, 3), (10 ** 9, 0)])
        gseed = rnd.randrange(1, 10 ** 6)
        mode = rnd.choice(["--lean", "--lean", "general"])
        d = os.path.join(work, "case%d" % c)
        os.makedirs(d)
        t0 = time.time()
        fas = generate(d, kind, genomes, length, rate, gseed)
        dbg = os.path.join(d, "g.dbg")
        graph_oracle_build(fas, k, dbg)
        case = Case("fuzz%d" % c, dbg, fas, k, b=b, m=m, a=ab)
        desc = "case %3d  %-9s %2d x %6d bp rate %.2f seed %6d  k=%2d b=%3d m=%3d a=%3d used %s/%s  %-7s" % (
            c, kind, genomes, length, rate, gseed, k, b, m, ab, run, period if period < 10 ** 9 else "-", mode)
        try:
            n, nonempty = emulate_and_compare(case, lambda S: list(range(0, S, max(1, S // a.evals))), exe, d, max_instances=48,
                                              extra_args=["--lean"] if mode == "--lean" else [], used_period=period, used_run=run)
            say("%s  %4d evaluations (%4d non-empty) equal the oracle  %5.1f s" % (desc, n, nonempty, time.time() - t0))
            total += n
        except AssertionError as e:
            bad += 1
            say("%s  MISMATCH %s" % (desc, str(e)[:400]))
        except Exception as e:  # e.g. a per-seed capacity of the emulated arena
            say("%s  skipped: %s" % (desc, str(e)[:200]))
    say("# %d evaluations over %d cases, %d cases with a mismatch" % (total, a.cases, bad))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
