set -u
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29641 tests/multirank_worker.py --rate > gpurun_out/r2c14_worker8.log 2>&1
echo "worker exit $?" >> gpurun_out/r2c14_worker8.log
grep -E "MULTIRANK|exit" gpurun_out/r2c14_worker8.log | cut -c1-4000 | tr '|' '\n' | grep -v "^dtheta$"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29651 scripts/full_chain.py --iterations 6 --group-size 1 > gpurun_out/r2c14_fullchain_g1.log 2>&1
echo "full_chain g=1 exit $?"; grep '^{' gpurun_out/r2c14_fullchain_g1.log | tail -1
