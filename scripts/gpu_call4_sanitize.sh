set -u
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
TOOL=$1
timeout 200 python scripts/sanitize_case.py > gpurun_out/r2_sanitize_plain_$TOOL.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/r2_sanitize_plain_$TOOL.log; exit 1; }
timeout 900 compute-sanitizer --tool $TOOL --print-limit 20 python scripts/sanitize_case.py > gpurun_out/r2_sanitize_$TOOL.log 2>&1
echo "exit $?" >> gpurun_out/r2_sanitize_$TOOL.log
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|SANITIZE_CASE_DONE|ok: rss|exit|=========.*(Invalid|Race|hazard)" gpurun_out/r2_sanitize_$TOOL.log | head -20
