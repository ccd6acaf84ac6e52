set -u
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29641 tests/multirank_worker.py > gpurun_out/r2c13_worker.log 2>&1
echo "worker exit $?" >> gpurun_out/r2c13_worker.log
grep -E "MULTIRANK|exit|Error|error" gpurun_out/r2c13_worker.log | cut -c1-4000 | tr '|' '\n' | grep -v "^dtheta$" | tail -4
timeout 600 python -m pytest tests/test_gpu_sharded_chain.py tests/test_cli_gpu.py -q -m gpu 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29643 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2_bench_cfg3_2gpu.json 2>gpurun_out/r2_bench_cfg3_2gpu.err
echo "bench exit $?"; python -c "import json;d=json.loads(open('gpurun_out/r2_bench_cfg3_2gpu.json').read().strip().splitlines()[-1]);print(d['value'],d['e2e']['value'],d['sequential_schedule'])"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29644 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 | cut -c1-300
