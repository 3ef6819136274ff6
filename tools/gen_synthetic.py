#!/usr/bin/env python3
"""Seeded synthetic multi-genome FASTA generator.

Test/bench input generator; not part of the product path.  All sequences are upper-case ACGT,
80-column FASTA, one file per genome named g<i>.fa with header ``>g<i>.chr1``.

kinds
  star      ancestor uniform i.i.d. ACGT; each genome = ancestor with independent substitutions at
            `rate` per site (BASELINE configs[1] "4x10 Mbp, 0.05 subs/site", the 4x100 Mbp headline
            and configs[4]).
  mammal    configs[2]: ancestor = i.i.d. + `repeat_frac` of bases covered by interspersed repeats from
            200 families (log-uniform 300-6000 bp, each copy diverged 10-20 % from its consensus);
            per genome `rate` subs/site, 1 indel / 2 kb (len 1-10), 20 inversions of 10-500 kb.
  pangenome configs[3]: core + 40 accessory islands of 25 kb present with p=0.5 per genome,
            `rate` subs/site, 5 inversions per genome.
"""
import argparse
import os
import sys

import numpy as np

ALPHABET = np.frombuffer(b"ACGT", dtype=np.uint8)
COMP = np.array([3, 2, 1, 0], dtype=np.uint8)


def mutate(rng, s, rate):
    m = rng.random(s.size) < rate
    n = int(m.sum())
    s = s.copy()
    s[m] = (s[m] + rng.integers(1, 4, n, dtype=np.uint8)) % 4
    return s


def invert(rng, s, count, lo, hi):
    for _ in range(count):
        ln = int(rng.integers(lo, hi + 1))
        if ln >= s.size:
            continue
        a = int(rng.integers(0, s.size - ln))
        s[a:a + ln] = COMP[s[a:a + ln][::-1]]
    return s


def indels(rng, s, per_bp, max_len):
    n = int(s.size * per_bp)
    if n == 0:
        return s
    pos = np.sort(rng.integers(0, s.size, n))
    out, prev = [], 0
    for p in pos:
        p = int(p)
        if p < prev:
            continue
        out.append(s[prev:p])
        ln = int(rng.integers(1, max_len + 1))
        if rng.random() < 0.5:
            out.append(rng.integers(0, 4, ln, dtype=np.uint8))
            prev = p
        else:
            prev = min(s.size, p + ln)
    out.append(s[prev:])
    return np.concatenate(out)


def gen_star(rng, n, length, rate, **_):
    anc = rng.integers(0, 4, length, dtype=np.uint8)
    return [mutate(rng, anc, rate) for _ in range(n)]


def gen_mammal(rng, n, length, rate, repeat_frac=0.05, **_):
    anc = rng.integers(0, 4, length, dtype=np.uint8)
    fams = []
    for _ in range(200):
        ln = int(np.exp(rng.uniform(np.log(300), np.log(6000))))
        fams.append(rng.integers(0, 4, ln, dtype=np.uint8))
    covered, target = 0, int(length * repeat_frac)
    while covered < target:
        f = fams[int(rng.integers(0, len(fams)))]
        copy = mutate(rng, f, rng.uniform(0.10, 0.20))
        if copy.size >= length:
            break
        a = int(rng.integers(0, length - copy.size))
        anc[a:a + copy.size] = copy
        covered += copy.size
    out = []
    for _ in range(n):
        s = mutate(rng, anc, rate)
        s = indels(rng, s, 1.0 / 2000, 10)
        s = invert(rng, s, 20, min(10000, max(1, s.size // 100)), min(500000, max(2, s.size // 10)))
        out.append(s)
    return out


def gen_pangenome(rng, n, length, rate, **_):
    core_len = int(length * 0.8)
    core = rng.integers(0, 4, core_len, dtype=np.uint8)
    isl_len = max(1, min(25000, length // 200))
    islands = [rng.integers(0, 4, isl_len, dtype=np.uint8) for _ in range(40)]
    sites = np.sort(rng.integers(0, core_len, 40))
    out = []
    for _ in range(n):
        present = rng.random(40) < 0.5
        parts, prev = [], 0
        for j, p in enumerate(sites):
            p = int(p)
            parts.append(core[prev:p])
            if present[j]:
                parts.append(islands[j])
            prev = p
        parts.append(core[prev:])
        s = mutate(rng, np.concatenate(parts), rate)
        s = invert(rng, s, 5, max(1, s.size // 100), max(2, s.size // 10))
        out.append(s)
    return out


KINDS = {"star": gen_star, "mammal": gen_mammal, "pangenome": gen_pangenome}


def write_fasta(path, header, codes):
    seq = ALPHABET[codes]
    n = seq.size
    full = (n // 80) * 80
    with open(path, "wb") as f:
        f.write(b">" + header.encode() + b"\n")
        if full:
            rows = np.empty((full // 80, 81), dtype=np.uint8)
            rows[:, :80] = seq[:full].reshape(-1, 80)
            rows[:, 80] = 10
            f.write(rows.tobytes())
        if n > full:
            f.write(seq[full:].tobytes() + b"\n")


def generate(outdir, kind="star", genomes=4, length=1000000, rate=0.05, seed=1, **kw):
    """Write g0.fa .. g<n-1>.fa into outdir; returns the list of paths."""
    os.makedirs(outdir, exist_ok=True)
    rng = np.random.default_rng(seed)
    seqs = KINDS[kind](rng, genomes, length, rate, **kw)
    paths = []
    for i, s in enumerate(seqs):
        p = os.path.join(outdir, "g%d.fa" % i)
        write_fasta(p, "g%d.chr1" % i, s)
        paths.append(p)
    return paths


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--kind", default="star", choices=sorted(KINDS))
    ap.add_argument("--genomes", type=int, default=4)
    ap.add_argument("--length", type=int, default=1000000)
    ap.add_argument("--rate", type=float, default=0.05)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("-o", "--outdir", required=True)
    a = ap.parse_args()
    for p in generate(a.outdir, a.kind, a.genomes, a.length, a.rate, a.seed):
        print(p)


if __name__ == "__main__":
    sys.exit(main())
