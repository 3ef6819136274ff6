#!/bin/bash
# round-2 session 8: which change slowed H between sessions 5 and 7?  single-pass vote (mpv1) vs multi-pass vote
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
run() {
  label=$1; shift
  echo "== $label H" >> gpurun_out/r2s8_ab.log
  env "$@" timeout 300 python tools/time_case.py --length 100000000 --k 25 --reps 3 --construct --no-counters 2>&1 | grep '"rep"' >> gpurun_out/r2s8_ab.log
}
L=$PWD/sibeliaz_b200/lib
run default X=1
run mpv1 LCB_LIB_PATH=$L/libsibeliaz_lcb_mpv1.so
run default_behind LCB_HEAVY_BEHIND=1
run mpv1_behind LCB_LIB_PATH=$L/libsibeliaz_lcb_mpv1.so LCB_HEAVY_BEHIND=1
timeout 900 ncu --set full --import-source on --clock-control none --kernel-name regex:k_traverse_lean --launch-skip 70 --launch-count 3 \
  -o gpurun_out/r2s8_h_lean -f python tools/time_case.py --length 100000000 --k 25 --reps 1 --construct --no-counters > gpurun_out/r2s8_ncu.log 2>&1
python - <<'P'
import json
cur=None
for l in open('gpurun_out/r2s8_ab.log'):
    if l.startswith('=='): cur=l.strip(); continue
    try: d=json.loads(l)
    except Exception: continue
    print(cur, d['rep'], 'find_ms',d['ms_find'],'trav_ms',d['ms_traverse_kernels'],'rounds',d['rounds'],'lean',d.get('lean_runs'),d.get('lean_bails'))
P
