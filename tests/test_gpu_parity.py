"""Parity tests proper: the CUDA path (through the C ABI) against the CPU restatement oracle and against the
committed reference fixtures.  Bit-exact: integer / byte / index work only."""
import filecmp
import os

import numpy as np
import pytest

from conftest import canonical_gff

pytestmark = pytest.mark.gpu


def _oracle_and_product(case, tmp_path, window=None, gen_seq=False, chunks=0):
    import sibeliaz_b200 as sb
    from oracle_binding import Oracle
    orc = Oracle(case.graph, case.fastas, case.k, case.a)
    st = sb.JunctionStorage(case.graph, case.fastas, case.k, case.a)
    kw = {}
    if window:
        kw = dict(window_init=window, window_max=window)
    bf = sb.BlocksFinder(st, case.k, collect_counters=True, **kw)
    return orc, st, bf


def _check_index(orc, st):
    oi, pi = orc.index(), st.arrays()
    for name in ("chr_off", "pos_id", "pos_bp", "next_ch", "prev_rc", "vtx_off", "occ_g"):
        assert np.array_equal(oi[name], pi[name]), name


def _check_seeds(orc, bf):
    os_, ps = orc.seeds(), bf.seeds()
    assert len(os_["vid"]) == len(ps["vid"])
    for name in ("vid", "ch", "count", "rank", "res_pos", "res_chr"):
        assert np.array_equal(os_[name], ps[name]), "seed field %s differs" % name


def _check_blocks(orc, bf, case):
    ob = orc.find_blocks(case.m, case.b)
    pb = bf.find_blocks(case.m, case.b)
    assert len(pb) == len(ob["id"]), "block instances: product %d oracle %d" % (len(pb), len(ob["id"]))
    assert np.array_equal(pb["id"], ob["id"])
    assert np.array_equal(pb["chr"], ob["chr"])
    assert np.array_equal(pb["start"].astype(np.uint64), ob["start"])
    assert np.array_equal(pb["end"].astype(np.uint64), ob["end"])
    assert bf.stats["kernel_launches"] > 0 and bf.stats["traversals_first"] >= bf.stats["n_seeds"]


def _check_gff(bf, case, tmp_path):
    out = str(tmp_path / ("out_" + case.name))
    bf.generate_output(out, False, 0)
    gff = os.path.join(out, "blocks_coords.gff")
    assert canonical_gff(gff) == canonical_gff(case.ref_gff)
    assert filecmp.cmp(gff, case.ref_gff, shallow=False), "GFF is set-equal but not byte-identical"


def test_star_small_all_stages(star_small, tmp_path):
    orc, st, bf = _oracle_and_product(star_small, tmp_path)
    _check_index(orc, st)
    bf.create(star_small.m, star_small.b)
    _check_seeds(orc, bf)
    _check_blocks(orc, bf, star_small)
    _check_gff(bf, star_small, tmp_path)


@pytest.mark.parametrize("window", [256, 1024, 8192])
def test_star_small_window_independent(star_small, tmp_path, window):
    """The result must not depend on the speculation window (only on the reference's phase size 256)."""
    orc, st, bf = _oracle_and_product(star_small, tmp_path, window=window)
    _check_blocks(orc, bf, star_small)


def test_star_small_block_sequences(star_small, tmp_path):
    import sibeliaz_b200 as sb
    st = sb.JunctionStorage(star_small.graph, star_small.fastas, star_small.k, star_small.a)
    bf = sb.BlocksFinder(st, star_small.k)
    bf.find_blocks(star_small.m, star_small.b)
    out = str(tmp_path / "seq")
    bf.generate_output(out, True, 4)
    got = b""
    for i in range(4):
        got += b"== %d.tmp\n" % i + open(os.path.join(out, "%d.tmp" % i), "rb").read()
    assert got == open(star_small.ref_chunks, "rb").read()


def test_examples_k15(examples, tmp_path):
    """BASELINE configs[0]: examples/genome1.fa + genome2.fa, k=15."""
    case = examples["k15"]
    orc, st, bf = _oracle_and_product(case, tmp_path)
    _check_index(orc, st)
    bf.create(case.m, case.b)
    _check_seeds(orc, bf)
    _check_blocks(orc, bf, case)
    _check_gff(bf, case, tmp_path)


def test_examples_k25_golden(examples, tmp_path):
    """The reference's own golden: examples/sibeliaz_out/blocks_coords.gff at defaults."""
    case = examples["k25"]
    orc, st, bf = _oracle_and_product(case, tmp_path)
    bf.create(case.m, case.b)
    _check_seeds(orc, bf)
    _check_blocks(orc, bf, case)
    _check_gff(bf, case, tmp_path)


def _fresh_case(tmp_path, kind, genomes, length, k, seed, rate=0.03):
    from conftest import Case
    from oracle_binding import REF_TWOPACO, run_twopaco
    from tools.gen_synthetic import generate
    if not os.path.exists(REF_TWOPACO):
        pytest.skip("oracle/_ref/twopaco (input producer) not built")
    d = str(tmp_path / kind)
    fas = generate(d, kind, genomes, length, rate, seed)
    dbg = run_twopaco(fas, k, os.path.join(d, "g.dbg"), threads=8)
    return Case("%s_%dx%d_k%d" % (kind, genomes, length, k), dbg, fas, k)


@pytest.mark.parametrize("kind,genomes,length,k,seed", [("mammal", 8, 1500000, 25, 3), ("pangenome", 16, 500000, 15, 4)])
def test_synthetic_kinds_blocks_and_gff(tmp_path, kind, genomes, length, k, seed):
    """Scaled-down BASELINE configs[2] (mammalian-like with repeats, indels, inversions) and configs[3] (pan-genome)."""
    case = _fresh_case(tmp_path, kind, genomes, length, k, seed)
    orc, st, bf = _oracle_and_product(case, tmp_path)
    _check_index(orc, st)
    bf.create(case.m, case.b)
    _check_seeds(orc, bf)
    _check_blocks(orc, bf, case)
    out_o, out_p = str(tmp_path / "o"), str(tmp_path / "p")
    orc.generate_output(out_o, False, 0, case.m)
    bf.generate_output(out_p, False, 0)
    assert filecmp.cmp(os.path.join(out_o, "blocks_coords.gff"), os.path.join(out_p, "blocks_coords.gff"), shallow=False)


def test_cli_drop_in(star_small, tmp_path):
    """The sibeliaz-lcb binary with the wrapper's flags (SibeliaZ-LCB/sibeliaz:146) writes the reference's files."""
    import subprocess
    import sibeliaz_b200 as sb
    out = str(tmp_path / "cli")
    cmd = [sb.CLI_PATH, "--graph", star_small.graph] + star_small.fastas + ["-k", "21", "-b", "200", "-o", out, "-m", "50", "-t", "4",
                                                                          "--abundance", "150", "--chunks", "4"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.splitlines()
    assert lines[0] == "Loading the graph..." and lines[1] == "Analyzing the graph..." and lines[2].startswith("[") and lines[2].endswith("]")
    assert lines[3] == "Generating the output..." and lines[4].startswith("Blocks found: ") and lines[5].startswith("Coverage: ")
    assert filecmp.cmp(os.path.join(out, "blocks_coords.gff"), star_small.ref_gff, shallow=False)
    got = b""
    for i in range(4):
        got += b"== %d.tmp\n" % i + open(os.path.join(out, "%d.tmp" % i), "rb").read()
    assert got == open(star_small.ref_chunks, "rb").read()


def test_two_gpus_match_one(star_small, tmp_path):
    """Seed-sharded 2-GPU run (NCCL min-allreduce of the epoch claims) must return the identical commit-ordered list."""
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    script = tmp_path / "two.py"
    script.write_text('''
import os, sys, ctypes, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, %r); sys.path.insert(0, %r)
import sibeliaz_b200 as sb
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
st = sb.JunctionStorage(%r, %r, 21, 150)
bf = sb.BlocksFinder(st, 21, device=rank, window_init=2048, window_max=2048).create(50, 200)
idb = torch.zeros(128, dtype=torch.uint8, device="cuda")
if rank == 0:
    buf = ctypes.create_string_buffer(128); sb.load_library().lcb_comm_unique_id(buf)
    idb = torch.tensor(list(buf.raw), dtype=torch.uint8, device="cuda")
dist.broadcast(idb, 0)
bf.comm_init(rank, world, bytes(idb.cpu().tolist()))
b = bf.find_blocks(50, 200)
np.save(%r + "/blocks_%%d.npy" %% rank, b)
dist.destroy_process_group()
''' % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.dirname(os.path.abspath(__file__)), star_small.graph,
       star_small.fastas, str(tmp_path)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29533", str(script)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    import sibeliaz_b200 as sb
    from oracle_binding import Oracle
    ob = Oracle(star_small.graph, star_small.fastas, 21, 150).find_blocks(50, 200)
    for rank in range(2):
        b = np.load(str(tmp_path / ("blocks_%d.npy" % rank)))
        assert len(b) == len(ob["id"]) and np.array_equal(b["id"], ob["id"]) and np.array_equal(b["start"].astype(np.uint64), ob["start"])
        assert np.array_equal(b["end"].astype(np.uint64), ob["end"]) and np.array_equal(b["chr"], ob["chr"])


@pytest.mark.parametrize("m,b", [(100, 100), (30, 500), (200, 50)])
def test_star_small_other_parameters(star_small, tmp_path, m, b):
    """-m / -b other than the wrapper's defaults (min block, max branch = max flank)."""
    from conftest import Case
    case = Case(star_small.name, star_small.graph, star_small.fastas, star_small.k, b=b, m=m)
    orc, st, bf = _oracle_and_product(case, tmp_path)
    _check_blocks(orc, bf, case)


def test_abundance_threshold(star_small, tmp_path):
    from conftest import Case
    case = Case(star_small.name, star_small.graph, star_small.fastas, star_small.k, a=4)
    orc, st, bf = _oracle_and_product(case, tmp_path)
    bf.create(case.m, case.b)
    _check_seeds(orc, bf)
    _check_blocks(orc, bf, case)


def test_pool_overflow_retries_with_smaller_window(star_small, tmp_path, monkeypatch):
    """Result pools too small for the active set: the driver must abandon it, halve it and still return the exact result."""
    monkeypatch.setenv("LCB_TEST_POOL_ENTRIES", "4096")
    orc, st, bf = _oracle_and_product(star_small, tmp_path, window=8192)
    _check_blocks(orc, bf, star_small)
    assert bf.stats["pool_restarts"] > 0 and bf.stats["windows"] > 2


@pytest.mark.parametrize("which", ["star", "examples_k15"])
def test_big_arena_rerun(star_small, examples, which):
    """Seeds that outgrow the per-warp scratch arena are evaluated again in one of the big arena slots.  The
    -DLCB_TINY_ARENA build of the same sources (64 instances / 512 path vertices / 512 read-set intervals per warp) makes
    ordinary fixtures take that route all the time; the result must still be the oracle's."""
    import json
    import subprocess
    import sys
    from sibeliaz_b200.build import LIB_TINY
    assert os.path.exists(LIB_TINY), "run __graft_entry__.build() first"
    case = star_small if which == "star" else examples["k15"]
    here = os.path.dirname(os.path.abspath(__file__))
    r = subprocess.run([sys.executable, os.path.join(here, "variant_runner.py"), case.graph, str(case.k), str(case.a), str(case.m),
                        str(case.b)] + list(case.fastas), env=dict(os.environ, LCB_LIB_PATH=LIB_TINY), stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    v = json.loads(r.stdout.strip().splitlines()[-1])
    assert v["library"] == LIB_TINY
    assert v["same"] and v["n"] > 0
    if which == "examples_k15":
        assert v["big_arena_runs"] > 0
