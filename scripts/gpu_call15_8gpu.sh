set -u
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29651 scripts/full_chain.py --iterations 50 --group-size 1 > gpurun_out/r2c15_fullchain_g1_50.log 2>&1
echo "full_chain g=1 x50 exit $?"; grep '^{' gpurun_out/r2c15_fullchain_g1_50.log | tail -1
