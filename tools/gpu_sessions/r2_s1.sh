#!/bin/bash
# round-2 session 1: baseline numbers of the round-1 build on the north-star input (H = star 4x100 Mbp k=25),
# occupancy variants A/B, ncu capture on H, alignment stage with CTA rows
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
nproc > gpurun_out/r2s1_host.txt; lscpu | head -20 >> gpurun_out/r2s1_host.txt; free -g >> gpurun_out/r2s1_host.txt
# 1. H: per-round trace (fused index)
LCB_TRACE_ROUNDS=1 timeout 600 python tools/time_case.py --length 100000000 --k 25 --reps 1 --construct --no-counters > gpurun_out/r2s1_h_trace.log 2>&1
# 2. occupancy variants on H and on configs[1]
for v in base c5 c4s c6s c8s; do
  lib=sibeliaz_b200/lib/libsibeliaz_lcb_$v.so; [ $v = base ] && lib=sibeliaz_b200/lib/libsibeliaz_lcb.so
  [ -f $lib ] || continue
  echo "== $v H" >> gpurun_out/r2s1_variants.log
  LCB_LIB_PATH=$PWD/$lib timeout 300 python tools/time_case.py --length 100000000 --k 25 --reps 3 --construct --no-counters 2>&1 | grep '"rep"' >> gpurun_out/r2s1_variants.log
  echo "== $v C2" >> gpurun_out/r2s1_variants.log
  LCB_LIB_PATH=$PWD/$lib timeout 300 python tools/time_case.py --length 10000000 --k 21 --reps 3 --construct --no-counters 2>&1 | grep '"rep"' >> gpurun_out/r2s1_variants.log
done
# 3. ncu --set full of four k_traverse launches on H
timeout 900 ncu --set full --import-source on --clock-control none --kernel-name regex:k_traverse --launch-skip 8 --launch-count 4 \
  -o gpurun_out/r2s1_h_traverse -f python tools/time_case.py --length 100000000 --k 25 --reps 1 --construct --no-counters > gpurun_out/r2s1_ncu.log 2>&1
# 4. alignment stage, all 1350 example blocks, CTA rows on
LCA_CTA_ROWS=2048 timeout 600 python tools/time_align.py > gpurun_out/r2s1_align_cta.log 2>&1
tail -3 gpurun_out/r2s1_align_cta.log
grep -c round gpurun_out/r2s1_h_trace.log
cat gpurun_out/r2s1_variants.log | cut -c1-400
