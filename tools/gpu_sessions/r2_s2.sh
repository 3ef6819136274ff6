#!/bin/bash
# round-2 session 2: the common-case traversal kernel (lcb_lean.cuh): parity suite, A/B against the general kernel alone
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2s2_pytest.log 2>&1; tail -3 gpurun_out/r2s2_pytest.log
for mode in lean nolean; do
  [ $mode = nolean ] && export LCB_NO_LEAN=1 || unset LCB_NO_LEAN
  echo "== $mode H" >> gpurun_out/r2s2_ab.log
  timeout 300 python tools/time_case.py --length 100000000 --k 25 --reps 3 --construct --no-counters 2>&1 | grep '"rep"' >> gpurun_out/r2s2_ab.log
  echo "== $mode C2" >> gpurun_out/r2s2_ab.log
  timeout 300 python tools/time_case.py --length 10000000 --k 21 --reps 3 --construct --no-counters --oracle 2>&1 | grep '"rep"\|PARITY' >> gpurun_out/r2s2_ab.log
done
unset LCB_NO_LEAN
LCB_TRACE_ROUNDS=1 timeout 300 python tools/time_case.py --length 100000000 --k 25 --reps 1 --construct --no-counters > gpurun_out/r2s2_h_trace.log 2>&1
LCB_TRACE_ROUNDS=1 timeout 300 python tools/time_case.py --length 10000000 --k 21 --reps 1 --construct --no-counters > gpurun_out/r2s2_c2_trace.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none --kernel-name regex:k_traverse_lean --launch-skip 8 --launch-count 3 \
  -o gpurun_out/r2s2_h_lean -f python tools/time_case.py --length 100000000 --k 25 --reps 1 --construct --no-counters > gpurun_out/r2s2_ncu.log 2>&1
python - <<'P'
import json
cur=None
for l in open('gpurun_out/r2s2_ab.log'):
    if l.startswith('=='): cur=l.strip(); continue
    if l.startswith('PARITY'): print(cur, l.strip()); continue
    try: d=json.loads(l)
    except Exception: continue
    print(cur, d['rep'], 'find_ms',d['ms_find'],'trav_ms',d['ms_traverse_kernels'],'rounds',d['rounds'],'lean',d.get('lean_runs'),d.get('lean_bails'))
P
