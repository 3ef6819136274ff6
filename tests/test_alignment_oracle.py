"""The stage after the path: `spoa` per block -> alignment.maf.  This file pins the CHECKER of
that stage: the reference's spoa library (oracle/_ref/spoa-ref, built by oracle/Makefile from the sources where they lie)
behind the restated wrapper logic (tests/oracle_binding.reference_global_alignment) must reproduce the paragraphs of the
reference's shipped golden examples/sibeliaz_out/alignment.maf byte for byte.

The golden holds 1332 of the 1350 blocks: the 18 longest ones (9.9 - 27.6 kbp x 8 copies) are missing because the wrapper
drops a block whose spoa run printed nothing (sibeliaz:68-72), which is what happens when spoa runs out of memory; a full
run here reproduces all 1332 byte for byte (4 minutes, done once by hand); the test below runs a slice."""
import lzma
import os

import pytest

from oracle_binding import REF_SPOA, Oracle, maf_paragraphs, reference_global_alignment

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "examples", "golden_k25_alignment.maf.xz")
needs_spoa = pytest.mark.skipif(not os.path.exists(REF_SPOA), reason="oracle/_ref/spoa-ref not built (needs /root/reference)")


@needs_spoa
def test_spoa_reference_reproduces_golden_paragraphs(examples, tmp_path):
    case = examples["k25"]
    orc = Oracle(case.graph, case.fastas, case.k, case.a)
    orc.find_blocks(case.m, case.b)
    out = str(tmp_path / "lcb")
    orc.generate_output(out, True, 256, case.m)  # `--chunks 256` as the wrapper passes it (sibeliaz:146)
    chunks = ["%d.tmp" % i for i in (3, 7, 11, 42, 100, 200, 255)]
    maf = reference_global_alignment(out, "genome1.fa genome2.fa", chunks)
    assert maf.startswith("##maf version=1\n# sibeliaz v1.2.7 \n# cmd=genome1.fa genome2.fa\n\na\ns ")
    mine = maf_paragraphs(maf, is_text=True)
    golden = maf_paragraphs(lzma.open(GOLDEN, "rt").read(), is_text=True)
    assert len(golden) == 1332
    checked = 0
    for key, rows in mine.items():
        if key in golden:  # (a block the golden lacks is one of the 18 long ones)
            assert rows == golden[key]
            checked += 1
    assert checked >= 30 and checked >= len(mine) - 2


SLICE = (3, 7, 11, 42, 100, 200, 255)


@pytest.fixture(scope="module")
def example_chunks(examples, tmp_path_factory):
    case = examples["k25"]
    orc = Oracle(case.graph, case.fastas, case.k, case.a)
    orc.find_blocks(case.m, case.b)
    out = str(tmp_path_factory.mktemp("lcb_chunks"))
    orc.generate_output(out, True, 256, case.m)
    return out


@needs_spoa
def test_poa_restatement_equals_reference_spoa(example_chunks):
    """oracle/poa_oracle.cpp (the restatement a device kernel is checked against) == the reference's spoa library,
    byte for byte (all 1350 blocks of the examples agree; the test runs a slice)."""
    from oracle_binding import poa_oracle_text
    import subprocess
    for i in SLICE:
        f = os.path.join(example_chunks, "%d.tmp" % i)
        ref = subprocess.run([REF_SPOA, "--chunk", f, "-l", "1", "-r", "1", "-e", "-8"], check=True, stdout=subprocess.PIPE, text=True).stdout
        assert poa_oracle_text(f) == ref and ref.count("\na\n") >= 4


def test_poa_core_host_build_equals_restatement(example_chunks, tmp_path):
    """The product's POA core (sibeliaz_b200/csrc/poa_core.cuh: graph update, topological sort, traceback, MSA -- the code the
    kernel runs, row recurrence in scalar form) compiled for the host == the restatement, at every arena level, and on edge
    cases (one copy, one character, unequal lengths, other alphabets)."""
    import subprocess
    from oracle_binding import poa_oracle_text, write_chunk
    here = os.path.dirname(os.path.abspath(__file__))
    exe = str(tmp_path / "poa_core_host")
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(here, "poa_core_host.cpp")], check=True)
    edge = write_chunk(str(tmp_path / "edge.tmp"), [
        [("a;0;4;+;9", "ACGT")],
        [("a;0;1;+;9", "A"), ("b;0;1;+;9", "C"), ("c;0;1;+;9", "A")],
        [("a;0;8;+;9", "ACGTACGT"), ("b;0;3;+;9", "CGT"), ("c;0;12;-;30", "TTACGTACGTTT"), ("d;0;8;+;9", "ACGAACGT")],
        [("a;0;6;+;9", "acgtNN"), ("b;0;6;+;9", "ACGTNN"), ("c;0;7;+;9", "acgRtNN")],
        [("a;0;5;+;9", "AAAAA"), ("b;0;5;+;9", "TTTTT"), ("c;0;5;+;9", "AATTA"), ("d;0;5;+;9", "TTAAT"), ("e;0;5;+;9", "ATATA")],
    ])
    files = [edge] + [os.path.join(example_chunks, "%d.tmp" % i) for i in SLICE[:4]]
    for f in files:
        want = poa_oracle_text(f)
        for level in (0, 2):
            got = subprocess.run([exe, "--chunk", f, "--level", str(level)], check=True, stdout=subprocess.PIPE, text=True).stdout
            assert got == want, (f, level)


def test_warp_row_kernel_under_emulation(tmp_path):
    """The device row kernel AS WRITTEN (poa_core.cuh: four 32-column chunks per step, predecessor rows as pointers, the
    max-scan with its carry through __shfl_up_sync / __shfl_sync, the __syncwarp placement) run by 32 host threads in
    lockstep (tests/poa_warp_emu.cpp) == the restatement: edge cases and blocks of 130-333 bp with substitutions and indels
    (nodes with several predecessors, columns beyond one 128-column step)."""
    import random
    import subprocess
    from oracle_binding import poa_oracle_text, write_chunk
    here = os.path.dirname(os.path.abspath(__file__))
    exe = str(tmp_path / "poa_warp_emu")
    subprocess.run(["g++", "-O2", "-std=c++20", "-pthread", "-o", exe, os.path.join(here, "poa_warp_emu.cpp")], check=True)
    rnd = random.Random(5)

    def mutate(s, rate):
        out = []
        for ch in s:
            r = rnd.random()
            if r < rate / 3:
                continue
            if r < 2 * rate / 3:
                out += [rnd.choice("ACGT"), ch]
            elif r < rate:
                out.append(rnd.choice("ACGT"))
            else:
                out.append(ch)
        return "".join(out)

    blocks = [
        [("a;0;4;+;9", "ACGT")],
        [("a;0;1;+;9", "A"), ("b;0;1;+;9", "C"), ("c;0;1;+;9", "A")],
        [("a;0;8;+;9", "ACGTACGT"), ("b;0;3;+;9", "CGT"), ("c;0;12;-;30", "TTACGTACGTTT"), ("d;0;8;+;9", "ACGAACGT")],
        [("a;0;6;+;9", "acgtNN"), ("b;0;6;+;9", "ACGTNN"), ("c;0;7;+;9", "acgRtNN")],
    ]
    for n, length, rate in ((5, 200, 0.08), (3, 333, 0.15), (8, 130, 0.05), (4, 257, 0.3)):
        anc = "".join(rnd.choice("ACGT") for _ in range(length))
        blocks.append([("c%d;0;%d;+;999" % (i, length), mutate(anc, rate)) for i in range(n)])
    def emulate(name, bl, cta):
        f = write_chunk(str(tmp_path / (name + ".tmp")), bl)
        got = subprocess.run([exe, "--chunk", f, "--cta", str(cta)], check=True, stdout=subprocess.PIPE, text=True, timeout=900).stdout
        assert got == poa_oracle_text(f), (name, cta)

    # one warp per block (run_block / dp_row): the kernel validated on the GPU
    emulate("warp", blocks[:5] + [blocks[7]], 0)
    # one block per CTA for long blocks (run_block_cta / dp_row_cta: the warps of a CTA share every row; two-pass scan, the
    # pieces' maxima exchanged through shared memory, double-buffered): several warps in one step; one warp over three steps;
    # two warps over two steps.  Kept small: emulated threads on a few cores spend their time in barriers.
    emulate("cta4", blocks[:4], 4)
    emulate("cta1", [blocks[5]], 1)
    emulate("cta2", [[(h, s2[:300]) for h, s2 in blocks[5]]], 2)
    # a block that outgrows the optimistic arena level 0 (30 noisy copies of a 60-character ancestor: more graph nodes than
    # 2 x the longest copy): every lane must leave the block together with err = 1, and the run one level up -- the device
    # driver's retry ladder -- must give the restatement's rows
    noisy = [[("n%d;0;9;+;99" % i, mutate(blocks[4][0][1][:60], 0.3)) for i in range(30)]]
    f = write_chunk(str(tmp_path / "level0.tmp"), noisy)
    for cta in (0, 2):
        r = subprocess.run([exe, "--chunk", f, "--cta", str(cta), "--level", "0"], check=True, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=900)
        assert "block retried at level 1" in r.stderr and r.stdout == poa_oracle_text(f), cta
