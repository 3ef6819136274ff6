#!/usr/bin/env python3
"""bench.py -- junctions traversed/sec of the sibeliaz-lcb hot loop (BASELINE.json metric).

    python bench.py --gpus 1 --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus 1 --steps K ...   # the reference's own CPU path (oracle/_ref)
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

A *step* is one pass of the FindBlocks equivalent (seed enumeration + sort + carving-path traversal +
ordered commit; blocksfinder.h:453-530) over the workload, J/s = kept junction records / step time.
Default workload = the north-star input of BASELINE.json: synthetic 4 x 100 Mbp star phylogeny (0.05 subs/site, seed 1),
k=25, -b 200 -m 50 -a 150 (14.97 M junction records; inputs: tools/gen_synthetic.py, junction file by the reference
twopaco).  `--workload star4x10M_k21` is BASELINE configs[1] (the round-1 bench workload).

  value      step timed with CUDA events on the library's stream, index already resident in HBM
  e2e        the same through the C ABI from HOST arrays: lcb_create (pinned staging + H2D) + enumerate +
             find_blocks (+ D2H of the block instances) + lcb_destroy, every step
  roofline   k_traverse: algorithmic bytes (13*T_walk + 17*T_occ + T_scan + 32*T_score, step counts from the
             oracle on this workload) / summed CUDA-event time of its launches, vs the
             measured HBM copy bandwidth in MEASURED_PEAKS.json
  cpu_baseline  oracle/_ref/sibeliaz-lcb-ref (the unmodified reference) on this box's host cores
  wall_clock    whole-binary seconds (N=1 only): the drop-in CLI sibeliaz_b200/bin/sibeliaz-lcb on the same junction file,
             default (one process, exits when the CUDA context is released) and --detach, next to the reference binary
             at -t min(cores, 32); GFF compared byte for byte

The reference arm (--impl reference) times at most REF_MAX_STEPS whole runs of the reference binary (a pass over the
north-star input takes ~20 s) and says so in its line.

Only the cpu_baseline / --impl reference legs and the one-off parity check touch oracle/.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "junctions traversed/sec (sibeliaz-lcb hot loop)"
UNIT = "junctions/s"
WORKLOADS = {
    # name: (kind, genomes, length, rate, seed, k)
    "star4x10M_k21": ("star", 4, 10_000_000, 0.05, 1, 21),
    "star4x1M_k21": ("star", 4, 1_000_000, 0.05, 1, 21),
    "star4x100M_k25": ("star", 4, 100_000_000, 0.05, 1, 25),
}
B, M, A = 200, 50, 150
DEFAULT_WORKLOAD = "star4x100M_k25"
REF_MAX_STEPS = 3
DESCRIPTIONS = {
    "star4x100M_k25": "synthetic star 4x100 Mbp, 0.05 subs/site, seed 1, k=25, -b 200 -m 50 -a 150 (BASELINE north-star input)",
    "star4x10M_k21": "synthetic star 4x10 Mbp, 0.05 subs/site, seed 1, k=21, -b 200 -m 50 -a 150 (BASELINE configs[1])",
}


def prepare_workload(name, rank):
    """FASTA + junction file, cached under /tmp; rank 0 builds, others wait."""
    from tools.gen_synthetic import generate
    from oracle_binding import REF_TWOPACO, run_twopaco
    kind, g, length, rate, seed, k = WORKLOADS[name]
    d = os.path.join(os.environ.get("LCB_BENCH_DIR", "/tmp/sibeliaz_b200_bench"), name)
    dbg, done = os.path.join(d, "g.dbg"), os.path.join(d, ".done")
    fas = [os.path.join(d, "g%d.fa" % i) for i in range(g)]
    if rank == 0 and not os.path.exists(done):
        os.makedirs(d, exist_ok=True)
        generate(d, kind, g, length, rate, seed)
        if not os.path.exists(REF_TWOPACO):
            raise RuntimeError("oracle/_ref/twopaco (input producer, built by __graft_entry__.build()) is missing")
        run_twopaco(fas, k, dbg, threads=min(16, os.cpu_count() or 1))
        open(done, "w").close()
    while not os.path.exists(done):
        time.sleep(0.2)
    return dbg, fas, k


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    FIELDS = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits"],
                                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        self.stop_flag = True
        self.join(timeout=6)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": int(self.rows[0][1]) if self.rows[0][1].isdigit() else None,
                "reasons": reasons, "samples": len(self.rows)}


def median(xs):
    xs = sorted(xs)
    return xs[len(xs) // 2] if xs else None


def time_cli(dbg, fas, k, threads, out, extra=()):
    """Whole-binary wall clock of the drop-in CLI (fork+exec to exit), same flags as the reference run."""
    import sibeliaz_b200 as sb
    os.makedirs(out, exist_ok=True)
    cmd = [sb.CLI_PATH, "--graph", dbg] + fas + ["-k", str(k), "-b", str(B), "-o", out, "-m", str(M), "-t", str(threads),
                                                "--abundance", str(A), "--noseq"] + list(extra)
    t0 = time.perf_counter()
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    t1 = time.perf_counter()
    if r.returncode:
        raise RuntimeError("sibeliaz-lcb failed: " + r.stderr.decode(errors="replace")[-300:])
    return t1 - t0


def time_reference(dbg, fas, k, threads):
    """Runs oracle/_ref/sibeliaz-lcb-ref and returns (t_find_s, t_total_s): FindBlocks is the span between the
    reference's own 'Analyzing the graph...' and 'Generating the output...' stdout lines (sibeliaz.cpp:133,142)."""
    from oracle_binding import REF_LCB
    out = os.path.join(os.path.dirname(dbg), "ref_out")
    os.makedirs(out, exist_ok=True)
    cmd = [REF_LCB, "--graph", dbg] + fas + ["-k", str(k), "-b", str(B), "-o", out, "-m", str(M), "-t", str(threads), "--abundance", str(A), "--noseq"]
    t0 = time.perf_counter()
    p = subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, bufsize=0)
    marks, buf = {}, b""
    while True:
        c = p.stdout.read(1)
        if not c:
            break
        buf += c
        if c == b"\n":
            line = buf.decode(errors="replace")
            buf = b""
            for key in ("Analyzing the graph", "Generating the output"):
                if key in line and key not in marks:
                    marks[key] = time.perf_counter()
    p.wait()
    t1 = time.perf_counter()
    if p.returncode or len(marks) < 2:
        raise RuntimeError("reference sibeliaz-lcb failed")
    return marks["Generating the output"] - marks["Analyzing the graph"], t1 - t0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--no-wall-clock", action="store_true")
    ap.add_argument("--window", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    steps, warmup = max(1, a.steps), max(0, a.warmup)
    host_threads = min(32, os.cpu_count() or 1)  # the sibeliaz wrapper caps sibeliaz-lcb at 32 threads (sibeliaz:139)
    config = {"workload": DESCRIPTIONS.get(a.workload, a.workload), "name": a.workload,
              "l2_flush": "a 256 MiB device buffer is overwritten before every timed step (outside the timed region)"}

    # ------------------------------------------------------------------ reference arm
    if a.impl == "reference":
        if rank != 0:
            return 0
        dbg, fas, k = prepare_workload(a.workload, 0)
        import sibeliaz_b200 as sb
        n_records = sb.JunctionStorage(dbg, fas, k, A).n_records
        big = n_records > 5_000_000  # a pass of the reference over the north-star input takes ~20 s: bounded sample
        ref_steps = min(steps, REF_MAX_STEPS) if big else steps
        ref_warmup = 0 if big else min(warmup, 1)
        for _ in range(ref_warmup):
            time_reference(dbg, fas, k, host_threads)
        t_find, totals = 0.0, []
        for _ in range(ref_steps):
            tf, tt = time_reference(dbg, fas, k, host_threads)
            t_find += tf
            totals.append(tt)
        value = ref_steps * n_records / t_find
        line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": ref_steps, "steps_requested": steps,
                "warmup": ref_warmup, "ms_per_step": 1000.0 * t_find / ref_steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "int32/int64", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": value, "unit": UNIT, "cores": host_threads, "kind": "reference",
                                 "sample": "%d whole passes of the unmodified reference sibeliaz-lcb -t %d over the workload (of %d steps requested: one pass "
                                           "takes %.1f s); value = records / FindBlocks span (its 'Analyzing the graph...' to 'Generating the output...' lines)"
                                           % (ref_steps, host_threads, steps, median(totals))},
                "wall_clock": {"reference_s": median(totals), "threads": host_threads, "runs": len(totals)},
                "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ our arm
    import numpy as np
    import torch
    import sibeliaz_b200 as sb
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dbg, fas, k = prepare_workload(a.workload, rank)
    storage = sb.JunctionStorage(dbg, fas, k, A)
    n_records = storage.n_records
    win = dict(window_init=a.window, window_max=a.window) if a.window else {}

    comm_id = [None]

    def make_finder(collect=False):
        """One context per call.  N > 1: lcb_create_shared -- rank 0 uploads the index (one PCIe copy), the other ranks
        receive it over NVLink; the communicator and the peer mailboxes are made once per process and reused."""
        if world == 1:
            bf = sb.BlocksFinder(storage, k, device=local_rank, collect_counters=collect, **win)
            bf.create(M, B)
            return bf
        if comm_id[0] is None:
            idb = torch.zeros(128, dtype=torch.uint8, device="cuda")
            if rank == 0:
                import ctypes
                buf = ctypes.create_string_buffer(128)
                sb.load_library().lcb_comm_unique_id(buf)
                idb = torch.tensor(list(buf.raw), dtype=torch.uint8, device="cuda")
            dist.broadcast(idb, 0)
            comm_id[0] = bytes(idb.cpu().tolist())
        bf = sb.BlocksFinder(storage, k, device=local_rank, collect_counters=collect, shared=(rank, world, comm_id[0]), **win)
        bf.create(M, B)
        return bf

    def sync():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    # one-off parity check against the oracle (the checker), which also yields the algorithmic step counts
    parity, alg_bytes, counters = None, None, None
    bf = make_finder()
    blocks = bf.find_blocks(M, B)
    if rank == 0:
        from oracle_binding import Oracle
        orc = Oracle(dbg, fas, k, A)
        ob = orc.find_blocks(M, B)
        parity = bool(len(ob["id"]) == len(blocks) and np.array_equal(ob["id"], blocks["id"]) and np.array_equal(ob["chr"], blocks["chr"])
                      and np.array_equal(ob["start"], blocks["start"].astype(np.uint64)) and np.array_equal(ob["end"], blocks["end"].astype(np.uint64)))
        counters = orc.counters
        alg_bytes = 13 * counters["t_walk"] + 17 * counters["t_occ"] + counters["t_scan"] + 32 * counters["t_score"]
        orc.close()

    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def flush_l2():
        flush_buf.add_(1)  # read+write of 256 MiB > 126 MB L2
        torch.cuda.synchronize()

    # ---- value: index resident in HBM, whole FindBlocks equivalent per step, device-timed
    for _ in range(warmup):
        bf.reset_seeds()
        bf.find_blocks(M, B)
    sync()
    sampler = ClockSampler(local_rank)
    sampler.start()
    step_ms, trav_ms, trav_launches, launches = [], 0.0, 0, 0
    t0 = time.perf_counter()
    for _ in range(steps):
        flush_l2()
        bf.reset_seeds()
        bf.find_blocks(M, B)
        step_ms.append(bf.stats["ms_step_device"])
        trav_ms += bf.stats["ms_traverse_kernels"]
        trav_launches += bf.stats["traverse_launches"]
        launches += bf.stats["kernel_launches"]
    sync()
    wall_ms = 1000.0 * (time.perf_counter() - t0)
    clocks = sampler.summary()
    dev_ms = float(sum(step_ms))
    if dist is not None:
        t = torch.tensor([dev_ms, wall_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, wall_ms = t.tolist()
    total_records = n_records * steps  # every rank works on the same junction set; the seeds are what is sharded
    value = total_records / (dev_ms / 1000.0)
    stats = dict(bf.stats)
    bf.close()

    # ---- e2e: host arrays -> lcb_create (H2D) -> enumerate -> find (D2H) -> destroy, every step
    e2e_t, h2d, d2h = 0.0, 0, 0
    for it in range(1 + steps):
        flush_l2()
        sync()
        t0 = time.perf_counter()
        f = make_finder()
        t1 = time.perf_counter()
        blk = f.find_blocks(M, B)
        t2 = time.perf_counter()
        h2d, d2h = f.stats["h2d_bytes"], f.stats["d2h_bytes"]
        e2e_parts = {"create_ms": 1000 * (t1 - t0), "pack_h2d_ms": f.stats["ms_h2d"], "find_call_ms": 1000 * (t2 - t1),
                     "enumerate_ms": f.stats["ms_enumerate"], "find_ms": f.stats["ms_find"], "d2h_ms": f.stats["ms_d2h"]}
        f.close()
        sync()
        e2e_parts["destroy_ms"] = 1000 * (time.perf_counter() - t2)
        if it > 0:
            e2e_t += time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([e2e_t], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_t = t.item()
    e2e_value = total_records / e2e_t

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    traffic, traffic_src = None, None
    for cand in ("ncu_traffic_r2.json", "ncu_traffic_r1.json"):
        # dram__bytes_read+write of this kernel on this workload from the committed ncu capture.  Round 2 measured a whole pass
        # (all traversal launches); the tool's schedule has more, smaller launches than a live pass, so the per-launch figure
        # is that total over the launches of the live pass -- the same denominator `achieved` uses.
        try:
            t = json.load(open(os.path.join(ROOT, "profiles", cand)))
            t = t.get(a.workload, t if (a.workload == "star4x10M_k21" and "mean_dram_bytes_per_launch" in t) else None)
            if t and "dram_bytes_per_pass" in t and trav_launches:
                traffic, traffic_src = t["dram_bytes_per_pass"] * steps / trav_launches, cand
                break
            if t and "mean_dram_bytes_per_launch" in t:
                traffic, traffic_src = t["mean_dram_bytes_per_launch"], cand
                break
        except Exception:
            pass
    achieved = (alg_bytes * steps / 1e9) / (trav_ms / 1000.0) if trav_ms > 0 else 0.0
    roofline = {"bound": "hbm", "kernel": "k_traverse_lean", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_unit": "bytes per launch (ncu --set full, profiles/%s)" % traffic_src, "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6.65 TB/s (B200_PROFILING.md)",
                "algorithmic_bytes_per_step": alg_bytes, "algorithmic_bytes_per_launch": alg_bytes * steps / max(1, trav_launches), "launches_per_step": trav_launches / steps,
                "kernel_ms_per_step": trav_ms / steps, "kernel_share_of_step": trav_ms / dev_ms if dev_ms else None,
                "kernels": "k_traverse_lean (common case) + k_traverse (the rest), one CUDA-event pair around both per round",
                "note": "latency-bound dependent random walk: see DESIGN.md section 5"}
    cpu, ref_total = None, None
    if not a.no_cpu_baseline and world == 1:
        try:
            tf, ref_total = time_reference(dbg, fas, k, host_threads)
            cpu = {"value": n_records / tf, "unit": UNIT, "cores": host_threads, "kind": "reference",
                   "sample": "whole workload once: unmodified reference sibeliaz-lcb -t %d; FindBlocks span %.3f s, whole binary %.3f s" % (host_threads, tf, ref_total)}
        except Exception as e:  # the baseline is reported, never required for the GPU numbers
            cpu = {"value": None, "unit": UNIT, "cores": host_threads, "kind": "reference", "sample": "failed: %s" % e}
    wall = None
    if not a.no_wall_clock and world == 1:
        # whole-binary wall clock (the north-star target is stated on it): same junction file, same flags, same -t
        try:
            import filecmp
            d = os.path.dirname(dbg)
            sync = [time_cli(dbg, fas, k, host_threads, os.path.join(d, "our_out")) for _ in range(3)]
            det = [time_cli(dbg, fas, k, host_threads, os.path.join(d, "our_out_detach"), ["--detach"]) for _ in range(3)]
            time.sleep(1.0)  # the last detached worker is still releasing its context
            ref_gff = os.path.join(d, "ref_out", "blocks_coords.gff")
            same = filecmp.cmp(os.path.join(d, "our_out", "blocks_coords.gff"), ref_gff, shallow=False) if os.path.exists(ref_gff) else None
            wall = {"ours_s": median(sync), "ours_runs_s": [round(x, 3) for x in sync], "ours_detach_s": median(det),
                    "ours_detach_runs_s": [round(x, 3) for x in det], "reference_s": ref_total, "threads": host_threads,
                    "ratio": (ref_total / median(sync)) if ref_total else None, "ratio_detach": (ref_total / median(det)) if ref_total else None,
                    "gff_identical_to_reference": same,
                    "note": "fork+exec to process exit, CUDA context creation and teardown included; ours_s is the default (one process), "
                            "ours_detach_s returns when the outputs are complete and a worker releases the GPU afterwards"}
        except Exception as e:
            wall = {"error": str(e)}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": dev_ms / steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int32/uint32",
            "data": "synthetic", "config": config, "parity_vs_oracle": parity,
            "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "ms_per_step": 1000.0 * e2e_t / steps},
            "roofline": roofline, "cpu_baseline": cpu, "wall_clock": wall,
            "detail": {"records": int(n_records), "seeds": int(stats["n_seeds"]), "block_instances": int(stats["n_block_instances"]),
                       "windows": int(stats["windows"]), "rounds": int(stats["rounds"]), "traversals": int(stats["traversals_first"] + stats["traversals_rerun"]),
                       "wall_ms_per_step": wall_ms / steps, "e2e_last_step_parts": e2e_parts, "ms_enumerate": stats["ms_enumerate"], "oracle_counters": counters}}
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
