set -u
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29641 tests/multirank_worker.py > gpurun_out/r2c13_worker.log 2>&1
echo "worker exit $?" >> gpurun_out/r2c13_worker.log
grep -E "MULTIRANK|exit|Error|error" gpurun_out/r2c13_worker.log | cut -c1-3000 | tr '|' '\n' | grep -v "^dtheta$"
for g in 64 256; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29651 scripts/full_chain.py --individuals 125000 --iterations 4 --group-size $g 2>&1 | grep '^{' | tail -1 | cut -c 300-700
done
