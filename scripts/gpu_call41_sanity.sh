# sanity of the final binary (after removing two unused helpers): k1_tc5 parity + smoke
set -u
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 14 python -m pytest tests/test_gpu_tc5.py -x -q 2>&1 | tail -1
timeout 12 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-120
