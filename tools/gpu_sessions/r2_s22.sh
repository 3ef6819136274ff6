#!/bin/bash
# round-2 session 22: what the driver does at round end: build check, smoke(), GPU suite
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
( time python -c "import __graft_entry__ as g; g.build(); g.smoke()" ) > gpurun_out/r2s22_smoke.log 2>&1; tail -5 gpurun_out/r2s22_smoke.log
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2s22_pytest.log 2>&1; tail -3 gpurun_out/r2s22_pytest.log
