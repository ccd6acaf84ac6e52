# 2 GPUs with k1_tc5 as the default: one-process-per-GPU worker (sharded chain, peer-memory sums), the torchrun tests, bench line
set -u
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29641 tests/multirank_worker.py > gpurun_out/r2c33_worker.log 2>&1
echo "worker exit $?" >> gpurun_out/r2c33_worker.log
grep -E "MULTIRANK|exit|Error|error" gpurun_out/r2c33_worker.log | cut -c1-3000 | tr '|' '\n' | tail -6
timeout 600 python -m pytest tests/test_gpu_sharded_chain.py tests/test_cli_gpu.py -q -m gpu 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29643 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2_bench_cfg3_2gpu_k1tc5.json 2>gpurun_out/r2c33_bench.err
echo "bench exit $?"; python -c "import json;d=json.loads(open('gpurun_out/r2_bench_cfg3_2gpu_k1tc5.json').read().strip().splitlines()[-1]);print(d['n_gpus'],d['value'],d['ms_per_step'],d['k1_ms'],d['e2e']['value'],d['sequential_schedule']['visits_per_s'])"
