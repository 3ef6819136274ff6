"""Run one LCB case in THIS process against the oracle and print a JSON verdict.  Used by GPU tests that need a
different build of the library (LCB_LIB_PATH is read when the library is first loaded, so they start a subprocess).

    python tests/variant_runner.py <graph> <k> <a> <m> <b> <fasta...>
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    graph, k, a, m, b = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
    fastas = sys.argv[6:]
    import sibeliaz_b200 as sb
    from oracle_binding import Oracle
    ob = Oracle(graph, fastas, k, a).find_blocks(m, b)
    st = sb.JunctionStorage(graph, fastas, k, a)
    bf = sb.BlocksFinder(st, k)
    pb = bf.find_blocks(m, b)
    same = len(pb) == len(ob["id"]) and np.array_equal(pb["id"], ob["id"]) and np.array_equal(pb["chr"], ob["chr"]) and \
        np.array_equal(pb["start"].astype(np.uint64), ob["start"]) and np.array_equal(pb["end"].astype(np.uint64), ob["end"])
    print(json.dumps(dict(same=bool(same), n=int(len(pb)), library=sb.load_library()._name,
                          big_arena_runs=int(bf.stats["big_arena_runs"]), traversals=int(bf.stats["traversals_first"]))))
    bf.close()


if __name__ == "__main__":
    main()
