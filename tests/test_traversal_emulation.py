"""The traversal kernel's device code AS WRITTEN (sibeliaz_b200/csrc/lcb_traverse.cuh: process_seed with mpv_fast / mpv_mid /
the general vote, push_parallel / push_group, path_score, shadow state, spills) executed on the CPU by 32 host threads in
lockstep (tests/cuda_emu.h, tests/trav_emu.cpp) against the oracle's epoch-threshold Process (oracle/liblcb_oracle_epoch.so):
same seeds, same epoch array (runs of edges claimed by seed 0, i.e. used for everybody), same bestInstance lists.
This is how a missing __syncwarp between the lanes' probes of the path hash and lane 0's insert was found (ThreadSanitizer
on the emulator) although the GPU, which keeps the warp converged there, never showed it."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import sibeliaz_b200 as sb

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _epoch_oracle():
    from oracle_binding import build_oracle
    path = os.path.join(ROOT, "oracle", "liblcb_oracle_epoch.so")
    if not os.path.exists(path):
        build_oracle()
    L = C.CDLL(path)
    L.lcbo_load.restype = C.c_void_p
    L.lcbo_load.argtypes = [C.c_char_p, C.POINTER(C.c_char_p), C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_int]
    L.lcbo_enumerate_seeds.argtypes = [C.c_void_p]
    L.lcbo_enumerate_seeds.restype = C.c_int64
    L.lcbo_get_seeds.argtypes = [C.c_void_p] + [C.c_void_p] * 6
    L.lcbo_epoch_prepare.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
    L.lcbo_epoch_process.argtypes = [C.c_void_p, C.c_int64, C.c_uint32, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_int)]
    L.lcbo_free.argtypes = [C.c_void_p]
    return L


def emulate_and_compare(case, sample_of, exe, workdir, thresh=1000, max_instances=10 ** 9, extra_args=(), used_period=150, used_run=40):
    """Returns (evaluations, non-empty results); asserts that the emulated device code and the oracle agree on every one.
    Every `used_period` records a run of `used_run` edges counts as used (claimed by seed 0)."""
    st = sb.JunctionStorage(case.graph, case.fastas, case.k, case.a)
    lib = sb.load_library()
    lib.lcb_index_pack.argtypes = [C.c_void_p]
    assert lib.lcb_index_pack(st._h) == 0  # the product's own device record layout
    v = sb.IndexView()
    lib.lcb_index_get_view(st._h, C.byref(v))
    N, V, Cn = v.n_records, v.n_vertices, v.n_chr
    rec = np.ctypeslib.as_array(C.cast(v.packed_rec, C.POINTER(C.c_int32)), shape=(N * 4,)).copy()
    occ = np.ctypeslib.as_array(C.cast(v.packed_occ, C.POINTER(C.c_int32)), shape=(N * 2,)).copy()
    vo = np.ctypeslib.as_array(v.vtx_off, shape=(V + 1,)).astype(np.uint32)
    vo = np.concatenate([vo, vo[-1:]])
    co = np.ctypeslib.as_array(v.chr_off, shape=(Cn + 1,)).astype(np.uint32)
    E = np.full(N + 32, 0xFFFFFFFF, np.uint32)
    for g in range(0, N, used_period):  # runs of used edges
        E[g:g + used_run] = 0
    L = _epoch_oracle()
    err = C.create_string_buffer(512)
    files = (C.c_char_p * len(case.fastas))(*[f.encode() for f in case.fastas])
    h = L.lcbo_load(case.graph.encode(), files, len(case.fastas), case.k, case.a, err, 512)
    assert h, err.value
    S = L.lcbo_enumerate_seeds(h)
    vid, ch = np.zeros(S, np.int64), np.zeros(S, np.uint8)
    other = [np.zeros(S, np.uint64) for _ in range(4)]
    L.lcbo_get_seeds(h, vid.ctypes.data, ch.ctypes.data, *[o.ctypes.data for o in other])
    L.lcbo_epoch_prepare(h, case.m, case.b, case.b, 8)
    inst, rs, nrs = np.zeros(2 * 8192, np.int64), np.zeros(2 * 65536, np.int64), C.c_int()
    sample, want = [], []
    for i in sample_of(S):
        n = L.lcbo_epoch_process(h, i, thresh, E.ctypes.data, inst.ctypes.data, 8192, rs.ctypes.data, 65536, C.byref(nrs))
        if n > max_instances:
            continue  # (an emulated evaluation costs a barrier per collective: keep the big ones out of the CPU suite)
        sample.append(i)
        want.append(" ".join([str(n)] + [str(x) for x in inst[:2 * n]]))
    L.lcbo_free(h)
    inp, out = os.path.join(workdir, "trav_in.bin"), os.path.join(workdir, "trav_out.txt")
    with open(inp, "wb") as f:
        np.array([N, V, Cn, len(sample)], np.int64).tofile(f)
        np.array([case.k, case.b, case.m, case.b, 8], np.int32).tofile(f)
        for a in (rec, occ, vo, co, E):
            a.tofile(f)
        jobs = np.zeros((len(sample), 3), np.uint32)
        jobs[:, 0] = vid[sample].astype(np.int32).view(np.uint32)
        jobs[:, 1] = ch[sample]
        jobs[:, 2] = thresh
        jobs.tofile(f)
    subprocess.run([exe, inp, out] + list(extra_args), check=True, timeout=1500)
    got = open(out).read().splitlines()
    bad = [(sample[j], want[j][:120], got[j][:120]) for j in range(len(sample)) if got[j] != want[j]]
    assert not bad, bad[:3]
    return len(sample), sum(1 for w in want if not w.startswith("0"))


@pytest.fixture(scope="module")
def trav_emu(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("emu") / "trav_emu")
    subprocess.run(["g++", "-O2", "-std=c++20", "-pthread", "-I", os.path.join(HERE, "emu_include"), "-o", exe,
                    os.path.join(HERE, "trav_emu.cpp")], check=True)
    return exe


def test_device_traversal_code_equals_oracle_star(star_small, trav_emu, tmp_path):
    n, nonempty = emulate_and_compare(star_small, lambda S: list(range(0, 24)) + list(range(24, S, max(1, S // 40))), trav_emu, str(tmp_path))
    assert n >= 60 and nonempty >= 40


def test_common_case_kernel_code_equals_oracle_star(star_small, trav_emu, tmp_path):
    """lcb_lean.cuh (the common-case traversal kernel's body) first, the general code for what it hands back -- the way the
    two kernels share a round on the device.  Wider runs by hand: every third seed of this fixture (3982 evaluations, 74
    handed back) and 2073 evaluations of examples k=25 (257 handed back) agree with the oracle."""
    n, nonempty = emulate_and_compare(star_small, lambda S: list(range(0, 24)) + list(range(30, S, max(1, S // 60))), trav_emu, str(tmp_path),
                                      extra_args=["--lean"])
    assert n >= 80 and nonempty >= 60


def test_common_case_kernel_long_paths(examples, trav_emu, tmp_path):
    """examples at k=25 (duplicate-rich, paths of up to 1180 edges): paths beyond the 128 vertices of the shared-memory hash
    go on in the per-warp second-level table (undo log across both levels), multi-pass votes, up to 32 instances; what
    does not fit is handed back in the middle of a path."""
    n, nonempty = emulate_and_compare(examples["k25"], lambda S: list(range(3, S, max(1, S // 70))), trav_emu, str(tmp_path), max_instances=40,
                                      extra_args=["--lean"])
    assert n >= 50 and nonempty >= 45


def test_common_case_kernel_hands_back_what_it_cannot_do(examples, trav_emu, tmp_path):
    """examples at k=15: nearly every evaluation meets a vertex that occurs several times on one chromosome, i.e. leaves the
    common case in the middle of a path; the shared state must be left clean for the general code every time."""
    n, nonempty = emulate_and_compare(examples["k15"], lambda S: list(range(6, S, max(1, S // 20))), trav_emu, str(tmp_path), max_instances=32,
                                      extra_args=["--lean"])
    assert n >= 12 and nonempty >= 10


def test_common_case_kernel_deep_look_ahead(tmp_path_factory, trav_emu, tmp_path):
    """8 x 80 kbp mammalian-like input (interspersed repeats, indels, inversions) at k=25: look-ahead walks longer than the
    24 junctions of one pass and winners deeper than a warp's 32 lanes (the push loop then reloads), 8 walks per vote."""
    from conftest import Case
    from oracle_binding import run_twopaco
    from tools.gen_synthetic import generate
    d = str(tmp_path_factory.mktemp("mammal8x80k"))
    fas = generate(d, "mammal", 8, 80000, 0.03, 3)
    dbg = os.path.join(d, "g.dbg")
    run_twopaco(fas, 25, dbg, threads=4)
    n, nonempty = emulate_and_compare(Case("mammal", dbg, fas, 25), lambda S: list(range(0, S, max(1, S // 160))), trav_emu, str(tmp_path),
                                      extra_args=["--lean"])
    assert n >= 120 and nonempty >= 100


def test_device_traversal_code_equals_oracle_repeat_rich(examples, trav_emu, tmp_path):
    """examples at k=15: repeat-rich, so the pushes meet vertices that occur several times on a chromosome (push_group), more
    than 32 instances (bisected order, spill arena) and the general vote.  Evaluations whose result has more than 32
    instances are left to manual runs (a barrier per collective)."""
    n, nonempty = emulate_and_compare(examples["k15"], lambda S: list(range(6, S, max(1, S // 45))), trav_emu, str(tmp_path), max_instances=32)
    assert n >= 30 and nonempty >= 25
