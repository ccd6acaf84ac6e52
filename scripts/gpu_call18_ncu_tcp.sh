# round 2: one ncu --set full capture of the persistent per-branch HMC kernel (after the same command exited 0 without ncu)
set -u
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 120 python scripts/seq_rate.py 100000 4 50 100 > gpurun_out/r2c18_plain.log 2>&1
rc=$?; echo "plain exit $rc"; tail -1 gpurun_out/r2c18_plain.log | cut -c1-150; [ $rc = 0 ] || exit 1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_hmc_persistent -s 5 -c 1 -o gpurun_out/r2_prof_tcp \
  python scripts/seq_rate.py 100000 4 50 100 > gpurun_out/r2c18_ncu.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/r2c18_ncu.log | cut -c1-200; ls -la gpurun_out/r2_prof_tcp.ncu-rep
