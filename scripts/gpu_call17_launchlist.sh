# round 2: ncu launch list of the bench command (cfg3s = 1/10 of the branches of cfg3), after the same command exited 0 without ncu
set -u
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 300 python bench.py --workload cfg3s --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2c17_plain.json 2> gpurun_out/r2c17_plain.err
rc=$?; echo "plain exit $rc"; [ $rc = 0 ] || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_cfg3s.csv \
  python bench.py --workload cfg3s --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2c17_ncu.log 2>&1
echo "ncu exit $?"; wc -l gpurun_out/r2_launches_cfg3s.csv
