"""Shared inputs of the junction-finder tests: an N-rich, multi-record, mixed-case pair of FASTA files that exercises
the rules of oracle/graph_oracle.cpp (dummy edges next to 'N', records shorter than / exactly k, recurring record
starts, IUPAC codes)."""
import os
import random


def write_nrich(dirname):
    rnd = random.Random(5)

    def dna(n):
        return "".join(rnd.choice("ACGT") for _ in range(n))

    anc = dna(30000)
    recs = []
    for g in range(3):
        s = list(anc)
        for i in range(len(s)):
            if rnd.random() < 0.03:
                s[i] = rnd.choice("ACGT")
        for _ in range(12):
            p = rnd.randrange(len(s) - 50)
            for i in range(p, p + rnd.choice([1, 1, 2, 7, 40])):
                s[i] = "N"
        s = "".join(s)
        if g == 1:
            s = s.lower()
        recs.append((">g%d.a some description" % g, s[:20000]))
        recs.append((">g%d.b" % g, s[20000:]))
    recs.append((">tiny", "ACGTACGTACG"))           # shorter than any k used
    recs.append((">exact15", anc[100:115]))         # exactly k for k = 15
    recs.append((">dup_start", anc[:60]))           # starts like g*.a: an 'N'-adjacent k-mer that recurs
    recs.append((">iupac", anc[500:560] + "RYKM" + anc[560:640]))
    paths = [os.path.join(dirname, "nrich1.fa"), os.path.join(dirname, "nrich2.fa")]
    for path, part, width in ((paths[0], recs[:5], 70), (paths[1], recs[5:], 61)):
        with open(path, "w") as f:
            for h, s in part:
                f.write(h + "\n")
                for i in range(0, len(s), width):
                    f.write(s[i:i + width] + "\n")
    return paths
