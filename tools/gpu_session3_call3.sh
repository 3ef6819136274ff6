#!/bin/bash
# Developer helper: round trace of the 16 x 5 Mbp k=15 pan-genome (BASELINE configs[3]) after the 32-way multiset search
mkdir -p gpurun_out
O=gpurun_out/r1s3c
LCB_TRACE_ROUNDS=1 timeout 600 python tools/time_case.py --kind pangenome --genomes 16 --length 5000000 --k 15 --rate 0.02 --seed 4 --construct --reps 1 > ${O}_pangenome_trace.log 2>&1; echo "pangenome rc=$?"; grep -c "\[round\]" ${O}_pangenome_trace.log; grep "\[round\]" ${O}_pangenome_trace.log | cut -c1-260 | head -40; tail -2 ${O}_pangenome_trace.log | cut -c1-700
