#!/usr/bin/env python3
"""Measurement of the junction finder (not the north-star metric, which bench.py reports): one JSON line.

    python tools/bench_graph.py [--workload star4x10M_k21|star4x100M_k25] [--steps 5] [--warmup 2] [--no-cpu-baseline]

value   = k-mer positions per second, CUDA-event time of the whole device pipeline (sequences resident in HBM)
e2e     = the same through lcg_build from HOST sequences (H2D of the bases and D2H of the junction list inside the timed region)
roofline= the table-building kernel k_edges: algorithmic bytes = one 16-byte slot read per k-mer position + one 16-byte
          slot write per distinct k-mer, over its CUDA-event duration, against the measured HBM copy bandwidth
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

WORKLOADS = {"star4x10M_k21": (4, 10000000, 21, 1), "star4x100M_k25": (4, 100000000, 25, 1)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="star4x10M_k21", choices=sorted(WORKLOADS))
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    a = ap.parse_args()
    import sibeliaz_b200 as sb
    from tools.gen_synthetic import generate
    n, length, k, seed = WORKLOADS[a.workload]
    d = os.path.join("/tmp/lcb_bench_graph", a.workload)
    os.makedirs(d, exist_ok=True)
    fas = generate(d, "star", n, length, 0.05, seed)
    seqs = []
    for f in fas:
        with open(f, "rb") as fh:
            seqs.append(b"".join(line.strip() for line in fh if not line.startswith(b">")))
    for _ in range(a.warmup):
        sb.JunctionGraph(sequences=seqs, k=k).close()
    dev_ms = edges_ms = e2e_s = 0.0
    stats = None
    for _ in range(a.steps):
        t = time.perf_counter()
        g = sb.JunctionGraph(sequences=seqs, k=k)
        e2e_s += time.perf_counter() - t
        stats = g.stats
        dev_ms += stats["ms_device"]
        edges_ms += stats["ms_edges"]
        g.close()
    kmers = stats["n_kmers"]
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    alg = 16 * kmers + 16 * stats["n_distinct"]
    achieved = alg * a.steps / 1e9 / (edges_ms / 1e3)
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_graph_r1.json")))["k_edges"]["dram_bytes"]
    except Exception:
        pass
    cpu = None
    if not a.no_cpu_baseline:
        from oracle_binding import run_twopaco
        th = min(16, os.cpu_count() or 1)  # "will not use more than 16 threads" (sibeliaz:138)
        t = time.perf_counter()
        run_twopaco(fas, k, os.path.join(d, "ref.dbg"), threads=th, tmpdir=d)
        dt = time.perf_counter() - t
        cpu = {"value": kmers / dt, "unit": "k-mer positions/s", "cores": th, "kind": "reference",
               "sample": "whole workload once: unmodified reference twopaco -t %d --filtermemory 4, whole binary %.1f s" % (th, dt)}
    print(json.dumps({
        "metric": "k-mer positions/sec (junction finder, the twopaco step)", "value": kmers * a.steps / (dev_ms / 1e3), "unit": "k-mer positions/s",
        "n_gpus": 1, "steps": a.steps, "warmup": a.warmup, "ms_per_step": dev_ms / a.steps, "higher_is_better": True, "dtype": "u64", "data": "synthetic",
        "config": {"workload": a.workload, "k": k},
        "e2e": {"value": kmers * a.steps / e2e_s, "unit": "k-mer positions/s", "ms_per_step": 1e3 * e2e_s / a.steps,
                "h2d_bytes_per_step": stats["n_bases"], "d2h_bytes_per_step": 12 * stats["n_candidates"]},
        "roofline": {"bound": "hbm", "kernel": "k_edges", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "algorithmic_bytes_per_launch": alg, "kernel_ms": edges_ms / a.steps,
                     "kernel_share_of_step": edges_ms / dev_ms, "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback"},
        "cpu_baseline": cpu, "detail": {kk: (round(v, 3) if isinstance(v, float) else v) for kk, v in stats.items()}}))


if __name__ == "__main__":
    main()
