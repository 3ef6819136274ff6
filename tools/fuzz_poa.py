#!/usr/bin/env python3
"""Developer tool (CPU only): randomised differential test of the alignment stage's core (sibeliaz_b200/csrc/poa_core.cuh:
graph update, topological sort, traceback, MSA -- compiled for the host, tests/poa_core_host.cpp) and of its device row
kernels run under the lockstep warp emulator (tests/poa_warp_emu.cpp; one warp per block and one block per CTA) against
the CPU restatement of spoa (oracle/poa_oracle.cpp), byte for byte.  Blocks: 1 - 12 copies of an ancestor of 1 - 700
characters with substitutions / insertions / deletions at 0 - 40 %, truncated copies, lower case and non-ACGT characters.

    python tools/fuzz_poa.py --cases 300 --seed 1 [--emulated 40] [--log profiles/fuzz_poa_r2.log]
"""
import argparse
import os
import random
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def make_block(rnd, max_len):
    length = rnd.choice([1, 2, 5, 31, 32, 33, 127, 128, 129, rnd.randrange(1, max_len)])
    alphabet = rnd.choice(["ACGT", "ACGT", "ACGT", "AC", "ACGTN", "acgtACGTRY"])
    anc = "".join(rnd.choice(alphabet) for _ in range(length))
    rate = rnd.choice([0.0, 0.02, 0.08, 0.2, 0.4])
    block = []
    for i in range(rnd.randrange(1, 13)):
        out = []
        for ch in anc:
            r = rnd.random()
            if r < rate / 3:
                continue
            if r < 2 * rate / 3:
                out += [rnd.choice(alphabet), ch]
            elif r < rate:
                out.append(rnd.choice(alphabet))
            else:
                out.append(ch)
        s = "".join(out) or rnd.choice(alphabet)
        if rnd.random() < 0.15:  # a truncated copy
            a = rnd.randrange(len(s))
            s = s[a:a + rnd.randrange(1, len(s) - a + 1)]
        block.append(("c%d;0;%d;%s;9999" % (i, len(s), rnd.choice("+-")), s))
    return block


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=100)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--emulated", type=int, default=20, help="of the cases, how many (short ones) also run under the warp emulator")
    ap.add_argument("--log", default=None)
    a = ap.parse_args()
    from oracle_binding import poa_oracle_text, write_chunk
    tests = os.path.join(ROOT, "tests")
    work = tempfile.mkdtemp(prefix="poa_fuzz_")
    host, emu = os.path.join(work, "poa_core_host"), os.path.join(work, "poa_warp_emu")
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", host, os.path.join(tests, "poa_core_host.cpp")], check=True)
    subprocess.run(["g++", "-O2", "-std=c++20", "-pthread", "-o", emu, os.path.join(tests, "poa_warp_emu.cpp")], check=True)
    rnd = random.Random(a.seed)
    bad = blocks = emulated = 0
    lines = ["# tools/fuzz_poa.py --cases %d --seed %d --emulated %d" % (a.cases, a.seed, a.emulated)]
    for c in range(a.cases):
        with_emu = c < a.emulated
        bl = [make_block(rnd, 260 if with_emu else 700) for _ in range(rnd.randrange(1, 4))]
        f = write_chunk(os.path.join(work, "c%d.tmp" % c), bl)
        want = poa_oracle_text(f)
        blocks += len(bl)
        runs = [([host, "--chunk", f, "--level", str(rnd.choice([0, 1, 2]))], "host core")]
        if with_emu:
            runs.append(([emu, "--chunk", f, "--cta", "0"], "warp rows"))
            runs.append(([emu, "--chunk", f, "--cta", str(rnd.choice([1, 2, 4]))], "cta rows"))
            emulated += 1
        for cmd, what in runs:
            got = subprocess.run(cmd, check=True, stdout=subprocess.PIPE, text=True, timeout=1200).stdout
            if got != want:
                bad += 1
                lines.append("MISMATCH case %d (%s): %s" % (c, what, f))
                print(lines[-1], flush=True)
    lines.append("# %d cases, %d blocks (host build of the core); %d cases also under the warp emulator (warp rows and CTA rows); %d mismatches" % (
        a.cases, blocks, emulated, bad))
    print(lines[-1])
    if a.log:
        with open(a.log, "a") as fh:
            fh.write("\n".join(lines) + "\n")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
