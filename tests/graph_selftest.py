"""The reference's own known-answer test of the junction finder, restated (TwoPaCo/src/graphconstructor/test.cpp:19-254,
run by `twopaco --test`, constructor.cpp:147): random chromosomes with an 'N' every ~500 characters, five mutated copies
(substitutions, insertions, deletions), and a NAIVE definition of the junction positions -- a k-mer is a junction iff, over
both strands, it is followed (or preceded) by more than one distinct character, every non-ACGT character and every
sequence end counting as a character of its own; the first and the last k-mer of a chromosome are always marked.  The
reference compares the positions its junction file holds with these marks; so do the tests here, for every implementation
of the step (restatement, device code on the host, GPU)."""
import numpy as np

COMP = {"A": "T", "C": "G", "G": "C", "T": "A"}


def generate_sequence(rnd, length):  # test.cpp:19-37
    return "".join("N" if rnd.randrange(500) == 0 else rnd.choice("ACGT") for _ in range(length))


def mutate_sequence(rnd, chrom, change_rate, mutation_rate):  # test.cpp:39-67
    out = []
    for ch in chrom:
        if rnd.random() <= change_rate:
            if rnd.random() <= mutation_rate:
                out.append(rnd.choice("ACGT"))
            elif rnd.random() <= 0.5:
                out.append(ch)
                out.append(rnd.choice("ACGT"))
            # else: the character is dropped
        else:
            out.append(ch)
    return "".join(out)


def find_junctions_naively(chrs, k):  # test.cpp:71-160
    """-> (set of junction k-mers (both orientations), marks[i] = bool array over the positions of chromosome i)"""
    unknown = [1000]

    def fresh():
        unknown[0] += 1
        return unknown[0]

    genome = []
    for s in chrs:
        g = [fresh()] + [c if c in COMP else fresh() for c in s] + [fresh()]
        genome.append(g)
        genome.append([COMP[c] if c in COMP else fresh() for c in reversed(g)])
    ins, outs = {}, {}
    for g in genome:
        definite = np.array([c in COMP for c in g], dtype=np.int64)
        run = np.concatenate([[0], np.cumsum(definite)])
        for i in range(0, len(g) - k + 1):
            if run[i + k] - run[i] != k:
                continue
            v = "".join(g[i:i + k])
            if i + k < len(g):
                outs.setdefault(v, set()).add(g[i + k])
            if i > 0:
                ins.setdefault(v, set()).add(g[i - 1])
    junction = set()
    for e in (ins, outs):
        for v, chars in e.items():
            if len(chars) > 1:
                junction.add(v)
                junction.add("".join(COMP[c] for c in reversed(v)))
    marks = []
    for s in chrs:
        m = np.zeros(len(s), dtype=bool)
        for pos in range(len(s)):
            if pos == 0 or pos == len(s) - k or s[pos:pos + k] in junction:
                m[pos] = True
        marks.append(m)
    return junction, marks


def marks_of_junction_file(path, chrs):
    """What JunctionPositionReader::RestoreAllVectors yields (junctionapi.h:72-98)."""
    raw = np.fromfile(path, dtype=np.dtype([("pos", "<u4"), ("id", "<i8")]))
    marks = [np.zeros(len(s), dtype=bool) for s in chrs]
    ids = [dict() for _ in chrs]
    chrom = 0
    for pos, vid in zip(raw["pos"].tolist(), raw["id"].tolist()):
        if pos == 0xFFFFFFFF or vid == np.iinfo(np.int64).max:
            chrom += 1
            continue
        marks[chrom][pos] = True
        ids[chrom][pos] = vid
    return marks, ids


def make_case(rnd, length=9000, chr_number=6, change_rate=0.05, indel_rate=0.1):
    chrs = [generate_sequence(rnd, length)]
    for _ in range(1, chr_number):
        chrs.append(mutate_sequence(rnd, chrs[0], change_rate, indel_rate))
    return chrs


def write_fasta(path, chrs):
    with open(path, "w") as f:
        for j, s in enumerate(chrs):
            f.write(">%d\n%s\n" % (j, s))
    return path


def check(path, chrs, k):
    """The reference's two assertions (test.cpp:214-241): marks equal the naive ones, every junction k-mer has an id --
    here: the occurrences of one k-mer (either orientation) carry one |id|, signed by orientation."""
    junction, naive = find_junctions_naively(chrs, k)
    fast, ids = marks_of_junction_file(path, chrs)
    for i in range(len(chrs)):
        diff = np.flatnonzero(naive[i] != fast[i])
        assert diff.size == 0, "chr %d pos %d: %s != %s (k=%d)" % (i, diff[0], fast[i][diff[0]], naive[i][diff[0]], k)
    label = {}
    for i, s in enumerate(chrs):
        for pos, vid in ids[i].items():
            v = s[pos:pos + k]
            if v not in junction:
                continue  # a first / last k-mer that is not a junction: unique stub id
            rc = "".join(COMP[c] for c in reversed(v))
            canon, sign = (v, 1) if v < rc else (rc, -1)
            want = label.setdefault(canon, vid * sign)
            assert want == vid * sign, "k-mer %s carries ids %d and %d" % (canon, want, vid * sign)
    assert len(set(abs(x) for x in label.values())) == len(label)
    return sum(int(m.sum()) for m in fast)
