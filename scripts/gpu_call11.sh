set -u
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/r2c11_gpu_tests.log 2>&1
echo "gpu tests exit $?" >> gpurun_out/r2c11_gpu_tests.log; tail -6 gpurun_out/r2c11_gpu_tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
