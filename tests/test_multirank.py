"""World-size-2 (gloo, CPU) test of the multi-rank protocol of DESIGN.md section 7, the one lcb_device.cu runs over the peers'
mailboxes: the seeds of the rolling active set are dealt round-robin over ranks; every rank evaluates only its own seeds
and keeps their read-sets; the RESULTS of a round (speculative result + conflict flag, commit-time re-run) go to every rank
(here: all_gather_object; on the device: stores into the peers' mailboxes from the traversal kernel); then every rank
rebuilds the same epochs from all results, recomputes every seed's conflict status (replicated data only), validates its
own seeds' read-sets, and one MIN over the ranks' first dirty seed fixes the commit frontier; the clean prefix is committed
by every rank from its replica.  Seed evaluation itself is the oracle's epoch-threshold Process
(oracle/liblcb_oracle_epoch.so), so the test checks the PROTOCOL -- sharding, exchange, replicated state, termination,
ordered emit -- independently of CUDA.  Every rank's output must equal the sequential oracle's blocksInstance_ list."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INF = 0xFFFFFFFF
PHASE = 256


def _worker(rank, world, port, graph, fastas, k, W, out_queue):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    L = C.CDLL(os.path.join(ROOT, "oracle", "liblcb_oracle_epoch.so"))
    L.lcbo_load.restype = C.c_void_p
    L.lcbo_load.argtypes = [C.c_char_p, C.POINTER(C.c_char_p), C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_int]
    L.lcbo_num_records.restype = C.c_int64
    L.lcbo_num_records.argtypes = [C.c_void_p]
    L.lcbo_enumerate_seeds.restype = C.c_int64
    L.lcbo_enumerate_seeds.argtypes = [C.c_void_p]
    L.lcbo_epoch_prepare.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
    L.lcbo_epoch_process.argtypes = [C.c_void_p, C.c_int64, C.c_uint32, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_int)]
    L.lcbo_get_index.argtypes = [C.c_void_p] + [C.c_void_p] * 8
    err = C.create_string_buffer(256)
    files = (C.c_char_p * len(fastas))(*[f.encode() for f in fastas])
    h = L.lcbo_load(graph.encode(), files, len(fastas), k, 150, err, 256)
    N = L.lcbo_num_records(h)
    S = L.lcbo_enumerate_seeds(h)
    L.lcbo_epoch_prepare(h, 50, 200, 200, 8)
    pos_bp = np.zeros(N, np.uint32)
    L.lcbo_get_index(h, None, None, pos_bp.ctypes.data, None, None, None, None, None)
    inst_buf, rs_buf, n_rs = np.zeros(2 * 4096, np.int64), np.zeros(2 * 65536, np.int64), C.c_int()

    def process(i, thresh, E):
        n = L.lcbo_epoch_process(h, i, thresh, E.ctypes.data, inst_buf.ctypes.data, 4096, rs_buf.ctypes.data, 65536, C.byref(n_rs))
        assert n <= 4096 and n_rs.value <= 65536
        return inst_buf[:2 * n].reshape(-1, 2).copy(), rs_buf[:2 * n_rs.value].reshape(-1, 2).copy()

    def edges(inst):
        fg = inst[:, 0] & ((1 << 62) - 1)
        lo, hi = np.minimum(fg, inst[:, 1]), np.maximum(fg, inst[:, 1]) - 1
        return [np.arange(a, b + 1) for a, b in zip(lo, hi)]

    def conflicts(inst, E, limit):
        return len(inst) > 1 and any((E[e] < limit).any() for e in edges(inst))

    def changed(rs, Ea, Eb, limit):
        for lo, hi in rs:
            if ((Ea[lo:hi + 1] < limit) != (Eb[lo:hi + 1] < limit)).any():
                return True
        return False

    # rolling active set [c0, c1) as in find_blocks_device_loop: admit `W` seeds per round (at most 4 W active)
    Ecur = np.full(N, INF, np.uint32)
    out, blocks_before, rounds_total = [], 0, 0
    c0 = c1 = 0
    r0, r1, conf = {}, {}, {}  # replicated: results and conflict flags of ALL active seeds
    R0, R1, has1 = {}, {}, {}  # owner only: read-sets, "the commit-time re-run is up to date"
    need0, need1 = set(), set()
    while c0 < S:
        admit = min(W, S - c1, max(0, 4 * W - (c1 - c0)))
        for i in range(c1, c1 + admit):  # k_admit: fresh state on every rank, the owner queues the evaluation
            r0[i], conf[i] = np.zeros((0, 2), np.int64), False
            r1.pop(i, None)
            if i % world == rank:
                need0.add(i)
                R0[i], has1[i] = np.zeros((0, 2), np.int64), False
        c1 += admit
        rounds_total += 1
        # traversal of the own work items; every published result is also a message to the peers
        msgs = []
        for i in sorted(need0):
            r0[i], R0[i] = process(i, i // PHASE * PHASE, Ecur)
            conf[i] = has1[i] = conflicts(r0[i], Ecur, i)
            msgs.append((i, 0, conf[i], r0[i]))
            if conf[i]:
                r1[i], R1[i] = process(i, i, Ecur)
                msgs.append((i, 1, False, r1[i]))
        for i in sorted(need1):
            r1[i], R1[i] = process(i, i, Ecur)
            msgs.append((i, 1, False, r1[i]))
        need0, need1 = set(), set()
        everyone = [None] * world
        dist.all_gather_object(everyone, msgs)  # <- the exchange step (peer mailboxes on the device)
        for src, lst in enumerate(everyone):
            if src == rank:
                continue
            for i, slot, cf, res in lst:  # k_xapply
                if slot == 0:
                    r0[i], conf[i] = res, cf
                else:
                    r1[i] = res

        def final(i):
            f = r1.get(i, np.zeros((0, 2), np.int64)) if conf[i] else r0[i]
            return f if len(f) > 1 else f[:0]

        Enew = np.where(Ecur < c0, Ecur, INF).astype(np.int64)  # k_rebase: committed claims only
        for i in range(c0, c1):  # k_claim: every rank, every active seed
            for e in edges(final(i)):
                Enew[e] = np.minimum(Enew[e], i)
        Enew = Enew.astype(np.uint32)
        first_dirty = INF
        for i in range(c0, c1):  # k_validate
            own = i % world == rank
            rs0 = own and changed(R0[i], Ecur, Enew, i // PHASE * PHASE)
            was = conf[i]
            conf[i] = conflicts(r0[i], Enew, i)  # replicated data only: the same on every rank
            if not own:
                continue
            dirty = False
            if rs0:
                need0.add(i)
                has1[i] = False
                dirty = True
            else:
                dirty = conf[i] != was
                if conf[i] and (not has1[i] or changed(R1[i], Ecur, Enew, i)):
                    need1.add(i)
                    has1[i] = True
                    dirty = True
                if not conf[i]:
                    has1[i] = False
            if dirty:
                first_dirty = min(first_dirty, i)
        d = torch.tensor([first_dirty], dtype=torch.int64)
        dist.all_reduce(d, op=dist.ReduceOp.MIN)  # the 16-word control exchange of the device
        fd = min(int(d.item()), c1)
        if fd > c0:  # every rank commits the prefix from its replica
            for i in range(c0, fd):
                f = final(i)
                if len(f):
                    blocks_before += 1
                    for fgs, bg in f:
                        pos, fg = bool(fgs >> 62), int(fgs & ((1 << 62) - 1))
                        if pos:
                            out.append((i, blocks_before, int(pos_bp[fg]), int(pos_bp[bg]) + k))
                        else:
                            out.append((i, -blocks_before, int(pos_bp[bg]), int(pos_bp[fg]) + k))
                for tab in (r0, R0, r1, R1, conf, has1):
                    tab.pop(i, None)
            c0 = fd
        Ecur = Enew
    gathered = [None] * world
    dist.all_gather_object(gathered, out)
    if rank == 0:
        out_queue.put((gathered, rounds_total))
    dist.destroy_process_group()


@pytest.mark.parametrize("window", [1024, 4096])
def test_two_rank_protocol_matches_sequential_oracle(star_small, window):
    import torch.multiprocessing as mp
    from oracle_binding import Oracle
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29620 + window % 97
    procs = [ctx.Process(target=_worker, args=(r, 2, port, star_small.graph, star_small.fastas, star_small.k, window, q)) for r in range(2)]
    for p in procs:
        p.start()
    per_rank, rounds = q.get(timeout=600)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ob = Oracle(star_small.graph, star_small.fastas, star_small.k, star_small.a).find_blocks(star_small.m, star_small.b)
    assert len(per_rank) == 2
    for rows in per_rank:  # every rank holds the whole commit-ordered list
        assert len(rows) == len(ob["id"]) > 1000
        assert [r[1] for r in rows] == ob["id"].tolist()
        assert [r[2] for r in rows] == ob["start"].tolist()
        assert [r[3] for r in rows] == ob["end"].tolist()
    assert rounds >= 2
