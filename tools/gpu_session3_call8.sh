#!/bin/bash
# Developer helper: BASELINE configs[4] input (4 x 500 Mbp star, k=25) on ONE B200: capacity check of the fused pipeline
# (64 GB k-mer table, ~75 M junction records); fused vs two-step self-consistency, no reference run at this size.
mkdir -p gpurun_out
timeout 900 python tools/time_config.py --kind star --genomes 4 --length 500000000 --rate 0.05 --seed 5 --k 25 --no-ref > gpurun_out/r1s3h_c5_star4x500M.log 2>&1; echo "rc=$?"
cut -c1-900 gpurun_out/r1s3h_c5_star4x500M.log
nvidia-smi --query-gpu=memory.used,memory.total --format=csv
