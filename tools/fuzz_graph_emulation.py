#!/usr/bin/env python3
"""Developer tool (CPU only): randomised differential test of the junction finder's DEVICE CODE (graph_kmer.cuh compiled for
the host, tests/graph_emu.cpp) against the CPU restatement (oracle/graph_oracle.cpp), byte for byte.  Inputs: a few records
derived from a common ancestor (substitutions, an inversion, repeats) with runs of N, IUPAC codes, lower case, records shorter
than / exactly / just above k, empty records; k over every table width (1 .. 8 words), finite and infinite abundance.

    python tools/fuzz_graph_emulation.py --cases 200 --seed 1 [--log profiles/fuzz_graph_emulation_r2.log]
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
COMP = {"A": "T", "C": "G", "G": "C", "T": "A", "N": "N"}


def make_case(rnd, d, k):
    def dna(n):
        return "".join(rnd.choice("ACGT") for _ in range(n))

    L = rnd.choice([300, 2000, 8000])
    anc = dna(L)
    if rnd.random() < 0.5:  # a repeat: junctions in the middle of nowhere
        unit = dna(rnd.choice([k + 3, 2 * k, 5 * k]))
        for _ in range(rnd.randrange(2, 6)):
            p = rnd.randrange(max(1, L - len(unit)))
            anc = anc[:p] + unit + anc[p + len(unit):]
    recs = []
    for g in range(rnd.randrange(1, 6)):
        s = list(anc)
        rate = rnd.choice([0.0, 0.01, 0.05])
        for i in range(len(s)):
            if rnd.random() < rate:
                s[i] = rnd.choice("ACGT")
        for _ in range(rnd.randrange(0, 4)):
            p = rnd.randrange(len(s))
            for i in range(p, min(len(s), p + rnd.choice([1, 2, k - 1, k, k + 1, 3 * k]))):
                s[i] = rnd.choice(["N", "N", "R", "Y", "n"])
        s = "".join(s)
        if rnd.random() < 0.3:  # reverse complement of a stretch
            a = rnd.randrange(len(s) // 2)
            b = a + rnd.randrange(1, len(s) // 2)
            s = s[:a] + "".join(COMP.get(c.upper(), "N") for c in reversed(s[a:b])) + s[b:]
        if rnd.random() < 0.3:
            s = s.lower()
        cut = rnd.randrange(1, len(s))
        recs += [s[:cut], s[cut:]] if rnd.random() < 0.5 else [s]
    for extra in (k - 1, k, k + 1, k + 2, 0):
        if rnd.random() < 0.4:
            p = rnd.randrange(max(1, L - extra))
            recs.insert(rnd.randrange(len(recs) + 1), anc[p:p + extra])
    files, per = [], max(1, len(recs) // rnd.randrange(1, 3))
    for fi in range(0, len(recs), per):
        path = os.path.join(d, "f%d.fa" % len(files))
        with open(path, "w") as f:
            for ri, s in enumerate(recs[fi:fi + per]):
                f.write(">r%d_%d some text\n" % (fi, ri))
                w = rnd.choice([60, 70, 10 ** 9])
                for i in range(0, len(s), w):
                    f.write(s[i:i + w] + "\n")
        files.append(path)
    return files


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=100)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--log", default=None)
    a = ap.parse_args()
    from oracle_binding import graph_oracle_build
    work = tempfile.mkdtemp(prefix="lcg_fuzz_")
    exe = os.path.join(work, "graph_emu")
    subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", "-Wno-unknown-pragmas", "-o", exe, os.path.join(ROOT, "tests", "graph_emu.cpp")], check=True)
    rnd = random.Random(a.seed)
    bad, records, by_width = 0, 0, {}
    lines = ["# tools/fuzz_graph_emulation.py --cases %d --seed %d" % (a.cases, a.seed)]
    for c in range(a.cases):
        k = rnd.choice([3, 5, 9, 15, 21, 25, 31, 33, 35, 47, 63, 65, 95, 97, 127, 129, 159, 161, 191, 193, 223, 225, 255])
        ab = rnd.choice([0, 0, 0, 1, 2, 3, 10])
        d = os.path.join(work, "c%d" % c)
        os.makedirs(d)
        files = make_case(rnd, d, k)
        orc, emu = os.path.join(d, "o.dbg"), os.path.join(d, "e.dbg")
        n = graph_oracle_build(files, k, orc, abundance=ab if ab else 2 ** 64 - 1)
        subprocess.run([exe, str(k), str(rnd.choice([1, 4, 8])), str(ab), emu] + files, check=True, stdout=subprocess.PIPE, timeout=300)
        same = open(orc, "rb").read() == open(emu, "rb").read()
        records += n
        w = (2 * k + 63) // 64
        by_width[w] = by_width.get(w, 0) + 1
        if not same:
            bad += 1
            lines.append("MISMATCH case %d k=%d abundance=%d files=%s" % (c, k, ab, files))
            print(lines[-1], flush=True)
    lines.append("# %d cases (by k-mer width in words: %s), %d junction records compared, %d mismatches" % (
        a.cases, ", ".join("%d: %d" % kv for kv in sorted(by_width.items())), records, bad))
    print("\n".join(lines[-1:]))
    if a.log:
        with open(a.log, "a") as f:
            f.write("\n".join(lines) + "\n")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
