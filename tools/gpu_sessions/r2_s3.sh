#!/bin/bash
# round-2 session 3: device-driven round loop + graph tail (A/B), TMA window staging (A/B), occupancy of the common-case kernel
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2s3_pytest.log 2>&1; tail -3 gpurun_out/r2s3_pytest.log
run() { # label, env..., -- lib
  label=$1; shift
  echo "== $label H" >> gpurun_out/r2s3_ab.log
  env "$@" timeout 300 python tools/time_case.py --length 100000000 --k 25 --reps 3 --construct --no-counters 2>&1 | grep '"rep"' >> gpurun_out/r2s3_ab.log
  echo "== $label C2" >> gpurun_out/r2s3_ab.log
  env "$@" timeout 300 python tools/time_case.py --length 10000000 --k 21 --reps 3 --construct --no-counters --oracle 2>&1 | grep '"rep"\|PARITY' >> gpurun_out/r2s3_ab.log
}
L=$PWD/sibeliaz_b200/lib
run devloop_graph X=1
run devloop_nograph LCB_NO_GRAPH=1
run hostloop LCB_HOST_LOOP=1
run l5 LCB_LIB_PATH=$L/libsibeliaz_lcb_l5.so
run tma LCB_LIB_PATH=$L/libsibeliaz_lcb_tma.so
run l8 LCB_LIB_PATH=$L/libsibeliaz_lcb_l8.so
LCB_TRACE_ROUNDS=1 timeout 300 python tools/time_case.py --length 100000000 --k 25 --reps 1 --construct --no-counters > gpurun_out/r2s3_h_trace.log 2>&1
python - <<'P'
import json
cur=None
for l in open('gpurun_out/r2s3_ab.log'):
    if l.startswith('=='): cur=l.strip(); continue
    if l.startswith('PARITY'): print(cur, l.strip()); continue
    try: d=json.loads(l)
    except Exception: continue
    print(cur, d['rep'], 'find_ms',d['ms_find'],'trav_ms',d['ms_traverse_kernels'],'rounds',d['rounds'],'launches',d['kernel_launches'],'lean',d.get('lean_runs'),d.get('lean_bails'))
P
