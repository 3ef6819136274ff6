#!/bin/bash
# Developer helper: BASELINE configs[2] at full size (8 x 100 Mbp mammalian-like, k=25, -a 150) through the drop-in binaries
# next to the compiled reference sibeliaz-lcb on the same junction file.
mkdir -p gpurun_out
nproc; free -g | head -2
timeout 1000 python tools/time_config.py --kind mammal --genomes 8 --length 100000000 --rate 0.03 --seed 3 --k 25 --ref-limit 420 > gpurun_out/r1s3g_c3_mammal8x100M.log 2>&1; echo "rc=$?"
cut -c1-700 gpurun_out/r1s3g_c3_mammal8x100M.log
