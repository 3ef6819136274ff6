#!/bin/bash
# round-2 session 25: first GPU run of the wide k-mer path (31 < k <= 255) and of the host's stream watchdog; parity + graph
# suites as regression guard for the graph_kmer.cuh refactor; A/B of the L2 fetch granularity for the table passes
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
( time timeout 60 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/r2s25_smoke.log 2>&1; tail -4 gpurun_out/r2s25_smoke.log
timeout 170 python -m pytest tests/test_zz_gpu_wide_k.py tests/test_gpu_graph.py tests/test_gpu_parity.py -m gpu -q > gpurun_out/r2s25_pytest.log 2>&1; tail -15 gpurun_out/r2s25_pytest.log
timeout 60 python tools/ab_graph_l2fetch.py --steps 3 > gpurun_out/r2s25_ab_l2fetch.log 2>&1; cat gpurun_out/r2s25_ab_l2fetch.log
