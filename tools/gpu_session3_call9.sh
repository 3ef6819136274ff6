#!/bin/bash
# Developer helper: first run of the alignment stage on the GPU
mkdir -p gpurun_out
O=gpurun_out/r1s3i
timeout 150 python -m pytest tests/test_gpu_align.py -x -q -k "edge or chunk_files" > ${O}_align_small.log 2>&1; echo "small rc=$?"; tail -25 ${O}_align_small.log | cut -c1-300
timeout 100 python tools/time_align.py --short > ${O}_align_short.log 2>&1; echo "short rc=$?"; tail -5 ${O}_align_short.log | cut -c1-600
timeout 170 python tools/time_align.py > ${O}_align_full.log 2>&1; echo "full rc=$?"; tail -5 ${O}_align_full.log | cut -c1-600
