#!/bin/bash
# round-2 session 14: BASELINE configs[2] and configs[4] inputs at FULL size on one GPU against the unmodified reference binary
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1500 python tools/time_config.py --kind mammal --genomes 8 --length 100000000 --rate 0.03 --seed 3 --k 25 --ref-limit 900 > gpurun_out/r2s14_c3_mammal8x100M.log 2>&1
tail -8 gpurun_out/r2s14_c3_mammal8x100M.log | cut -c1-300
rm -rf /tmp/cfg
timeout 2400 python tools/time_config.py --kind star --genomes 4 --length 500000000 --rate 0.05 --seed 5 --k 25 --ref-limit 1500 > gpurun_out/r2s14_c5_star4x500M.log 2>&1
tail -8 gpurun_out/r2s14_c5_star4x500M.log | cut -c1-300
