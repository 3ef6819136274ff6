#!/usr/bin/env python3
"""Developer tool: attribute the stall samples / executed instructions of an ncu report's SASS page to the device
functions of lcb_traverse.cuh.  ncu's CSV export of the CUDA-source view carries no metrics, so the SASS rows are
joined (by instruction order) with `nvdisasm -g` line info of the same build.

    python tools/ncu_by_function.py gpurun_out/prof.ncu-rep [kernel-substring]
"""
import collections
import csv
import glob
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def function_ranges(path):
    out, cur = [], None
    for no, line in enumerate(open(path), 1):
        m = re.match(r"^(?:template\s*<[^>]*>\s*)?(?:__device__|__global__|static|inline|__forceinline__|__noinline__|\s)+[\w:<>\*& ]*?\b(\w+)\s*\(", line)
        if m and not line.startswith(" ") and "(" in line and not line.strip().startswith("//"):
            cur = m.group(1)
            out.append((no, cur))
    return out


def func_of(ranges, line):
    name = "?"
    for start, n in ranges:
        if start <= line:
            name = n
        else:
            break
    return name


def main():
    rep = sys.argv[1]
    want = sys.argv[2] if len(sys.argv) > 2 else "k_traverseILb1E"
    lib = os.path.join(ROOT, "sibeliaz_b200", "lib", "libsibeliaz_lcb.so")
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
    cub = max(glob.glob(os.path.join(tmp, "*.cubin")), key=os.path.getsize)
    sass = subprocess.run(["nvdisasm", "-g", "-c", cub], stdout=subprocess.PIPE, text=True).stdout.splitlines()
    ranges = {os.path.basename(p): function_ranges(p) for p in glob.glob(os.path.join(ROOT, "sibeliaz_b200", "csrc", "*"))}
    instr = []  # (file, line) per instruction of the wanted function
    on, cur = False, ("?", 0)
    for ln in sass:
        if ln.startswith(".text."):
            on = want in ln
            continue
        if not on:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
            instr.append(cur)
    rows = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE, text=True).stdout.splitlines()))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    h = rows[hdr]
    body = rows[hdr + 1:]
    ci = {n: h.index(n) for n in ("# Samples", "Instructions Executed", "stall_long_sb", "stall_wait", "stall_short_sb", "stall_branch_resolving", "stall_selected", "stall_no_inst", "stall_not_selected", "stall_lg", "stall_mio", "stall_math", "stall_dispatch", "stall_barrier")}
    body = [r for r in body if r and r[0].startswith("0x")]
    if instr and len(body) % len(instr) == 0 and len(body) > len(instr):
        print("%d launches in the report: aggregated" % (len(body) // len(instr)))
        instr = instr * (len(body) // len(instr))
    elif len(body) != len(instr):
        print("warning: %d SASS rows in the report vs %d instructions in this build (stale build?)" % (len(body), len(instr)))
    agg = collections.defaultdict(lambda: collections.Counter())
    lines = collections.Counter()
    tot = collections.Counter()
    for r, (f, l) in zip(body, instr):
        fn = func_of(ranges.get(f, []), l) if f in ranges else f
        for n, i in ci.items():
            v = int(float(r[i] or 0))
            agg[fn][n] += v
            tot[n] += v
        lines[(f, l)] += int(float(r[ci["# Samples"]] or 0))
    S, I = tot["# Samples"], tot["Instructions Executed"]
    print("total samples %d, warp instructions %d" % (S, I))
    print("stalls: " + ", ".join("%s %.1f%%" % (n[6:], 100.0 * tot[n] / max(S, 1)) for n in ci if n.startswith("stall_") and tot[n] * 50 > S))
    print("%-26s %8s %8s   %s" % ("function", "samples%", "instr%", "top stalls"))
    for fn, c in sorted(agg.items(), key=lambda kv: -kv[1]["# Samples"])[:22]:
        st = sorted(((c[n], n[6:]) for n in ci if n.startswith("stall_")), reverse=True)[:3]
        print("%-26s %8.1f %8.1f   %s" % (fn, 100.0 * c["# Samples"] / max(S, 1), 100.0 * c["Instructions Executed"] / max(I, 1),
                                        ", ".join("%s %.0f%%" % (n, 100.0 * v / max(c["# Samples"], 1)) for v, n in st)))
    print("hottest lines:")
    for (f, l), v in lines.most_common(25):
        src = ""
        p = os.path.join(ROOT, "sibeliaz_b200", "csrc", f)
        if os.path.exists(p):
            src = open(p).read().splitlines()[l - 1].strip()[:110]
        print("  %5.1f%%  %s:%d  %s" % (100.0 * v / max(S, 1), f, l, src))


if __name__ == "__main__":
    main()
