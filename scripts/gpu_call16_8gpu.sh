set -u
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29651 scripts/full_chain.py --iterations 50 --group-size 64 > gpurun_out/r2c16_fullchain_g64.log 2>&1
echo "full_chain g=64 exit $?"; grep '^{' gpurun_out/r2c16_fullchain_g64.log | tail -1
