#!/bin/bash
# round-2 session 6: lane-per-seed validation filter, multi-pass look-ahead in the common-case kernel, beside/behind policy
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2s6_pytest.log 2>&1; tail -3 gpurun_out/r2s6_pytest.log
run() {
  label=$1; shift
  echo "== $label H" >> gpurun_out/r2s6_ab.log
  env "$@" timeout 300 python tools/time_case.py --length 100000000 --k 25 --reps 3 --construct --no-counters 2>&1 | grep '"rep"' >> gpurun_out/r2s6_ab.log
  echo "== $label C2" >> gpurun_out/r2s6_ab.log
  env "$@" timeout 300 python tools/time_case.py --length 10000000 --k 21 --reps 3 --construct --no-counters --oracle 2>&1 | grep '"rep"\|PARITY' >> gpurun_out/r2s6_ab.log
}
run default X=1
run nodiff LCB_NO_DIFF=1
run behind LCB_HEAVY_BEHIND=1
echo "== pangenome 16x5M k15" >> gpurun_out/r2s6_ab.log
timeout 900 python tools/time_case.py --kind pangenome --genomes 16 --length 5000000 --k 15 --seed 4 --rate 0.02 --reps 2 --construct --no-counters --oracle 2>&1 | grep '"rep"\|PARITY' >> gpurun_out/r2s6_ab.log
echo "== mammal 8x10M k25" >> gpurun_out/r2s6_ab.log
timeout 900 python tools/time_case.py --kind mammal --genomes 8 --length 10000000 --k 25 --seed 3 --rate 0.03 --reps 2 --construct --no-counters --oracle 2>&1 | grep '"rep"\|PARITY' >> gpurun_out/r2s6_ab.log
echo "== mammal 8x10M k25 nolean" >> gpurun_out/r2s6_ab.log
LCB_NO_LEAN=1 timeout 900 python tools/time_case.py --kind mammal --genomes 8 --length 10000000 --k 25 --seed 3 --rate 0.03 --reps 2 --construct --no-counters 2>&1 | grep '"rep"\|PARITY' >> gpurun_out/r2s6_ab.log
LCB_TRACE_ROUNDS=1 timeout 300 python tools/time_case.py --length 100000000 --k 25 --reps 1 --construct --no-counters > gpurun_out/r2s6_h_trace.log 2>&1
python - <<'P'
import json
cur=None
for l in open('gpurun_out/r2s6_ab.log'):
    if l.startswith('=='): cur=l.strip(); continue
    if l.startswith('PARITY'): print(cur, l.strip()); continue
    try: d=json.loads(l)
    except Exception: continue
    print(cur, d['rep'], 'find_ms',d['ms_find'],'trav_ms',d['ms_traverse_kernels'],'rounds',d['rounds'],'lean',d.get('lean_runs'),d.get('lean_bails'),'why',d.get('lean_bail_why'))
P
