#!/usr/bin/env python3
"""Developer timing helper: generate a star synthetic, build its junction file with the reference twopaco,
run the product (and optionally the oracle / the compiled reference) and print stats.  Not a test."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--genomes", type=int, default=4)
    ap.add_argument("--length", type=int, default=10000000)
    ap.add_argument("--k", type=int, default=21)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--kind", default="star")
    ap.add_argument("--rate", type=float, default=0.05)
    ap.add_argument("--window", type=int, default=0)
    ap.add_argument("--wmax", type=int, default=0)
    ap.add_argument("--oracle", action="store_true")
    ap.add_argument("--ref", action="store_true")
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--dir", default="/tmp/lcb_time")
    ap.add_argument("--dbg", default=None, help="junction file made earlier by the reference twopaco for exactly this synthetic")
    ap.add_argument("--no-counters", action="store_true", help="timed configuration: the kernel without step counters")
    ap.add_argument("--construct", action="store_true", help="fused pipeline: junctions found on the GPU, no reference twopaco run")
    a = ap.parse_args()
    import numpy as np
    import sibeliaz_b200 as sb
    from oracle_binding import Oracle, run_reference_lcb, run_twopaco
    from tools.gen_synthetic import generate
    d = os.path.join(a.dir, "%s_%dx%d_k%d_s%d" % (a.kind, a.genomes, a.length, a.k, a.seed))
    os.makedirs(d, exist_ok=True)
    dbg = os.path.join(d, "g.dbg")
    t = time.time()
    fas = generate(d, a.kind, a.genomes, a.length, a.rate, a.seed)
    if a.dbg:
        dbg = a.dbg
    if a.construct:
        t = time.time()
        g = sb.JunctionGraph(fas, a.k)
        g.write(dbg)
        print("junction finder (GPU) %.2fs" % (time.time() - t), {k_: (round(v, 1) if isinstance(v, float) else v) for k_, v in g.stats.items()}, flush=True)
        g.close()
    if not os.path.exists(dbg):
        run_twopaco(fas, a.k, dbg, threads=min(16, os.cpu_count() or 1))
    print("input ready in %.1fs" % (time.time() - t), flush=True)
    t = time.time()
    if a.construct:
        st = sb.FusedStorage(fas, a.k, 150)
        print("fused storage %.2fs" % (time.time() - t), flush=True)
    else:
        st = sb.JunctionStorage(dbg, fas, a.k, 150)
        print("load %.2fs  records %d vertices %d" % (time.time() - t, st.n_records, st.n_vertices), flush=True)
    for rep in range(a.reps):
        bf = sb.BlocksFinder(st, a.k, window_init=a.window, window_max=a.wmax or a.window, collect_counters=(False if a.no_counters else (2 if os.environ.get('LCB_TRACE_ROUNDS') else True)))
        t = time.time()
        bf.create(50, 200)
        t_create = time.time() - t
        t = time.time()
        bf.enumerate_seeds()
        t_enum = time.time() - t
        t = time.time()
        blocks = bf.find_blocks(50, 200)
        t_find = time.time() - t
        s = bf.stats
        print(json.dumps(dict(rep=rep, create_s=round(t_create, 3), enum_s=round(t_enum, 3), find_s=round(t_find, 3),
                              jps=round(s["n_records"] / (t_enum + t_find)), **{k: (round(v, 2) if isinstance(v, float) else v) for k, v in s.items()})), flush=True)
        bf.close()
    if a.oracle:
        t = time.time()
        orc = Oracle(dbg, fas, a.k, 150)
        ob = orc.find_blocks(50, 200)
        print("oracle %.2fs" % (time.time() - t), orc.counters)
        ok = len(ob["id"]) == len(blocks) and np.array_equal(ob["id"], blocks["id"]) and np.array_equal(ob["start"], blocks["start"].astype(np.uint64)) \
            and np.array_equal(ob["end"], blocks["end"].astype(np.uint64)) and np.array_equal(ob["chr"], blocks["chr"])
        print("PARITY", ok, len(ob["id"]), len(blocks))
    if a.ref:
        import filecmp
        import subprocess
        out = os.path.join(d, "refout")
        os.makedirs(out, exist_ok=True)
        for th in ((1,) if a.length <= 20000000 else ()) + (min(32, os.cpu_count() or 1),):
            t = time.time()
            log = run_reference_lcb(dbg, fas, a.k, out, threads=th)
            print("reference -t %d total %.2fs" % (th, time.time() - t), flush=True)
        # whole-binary wall clock of the drop-in CLI on the same input, and byte comparison of the GFF
        mine = os.path.join(d, "myout")
        t = time.time()
        r = subprocess.run([sb.CLI_PATH, "--graph", dbg] + fas + ["-k", str(a.k), "-b", "200", "-o", mine, "-m", "50", "-t", str(min(32, os.cpu_count() or 1)),
                            "--abundance", "150", "--noseq", "--stats"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        print("sibeliaz-lcb (B200) whole binary %.2fs rc=%d" % (time.time() - t, r.returncode))
        print(r.stdout.strip().splitlines()[-2:], r.stderr.strip()[-900:])
        print("GFF byte-identical to the reference:", filecmp.cmp(os.path.join(mine, "blocks_coords.gff"), os.path.join(out, "blocks_coords.gff"), shallow=False))


if __name__ == "__main__":
    main()
