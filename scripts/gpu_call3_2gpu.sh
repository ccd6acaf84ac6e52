# round 2, call 3 (2 GPUs): one process per GPU -- sharded sequential chain (incl. early rejections), block-Jacobi sweeps over the
# bulk exchange, sliced Net.gradient, the CLI under torchrun, the sequential rate, and bench at N = 2
set -u
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2c3_smi.log 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29641 tests/multirank_worker.py --rate > gpurun_out/r2c3_worker.log 2>&1
echo "worker exit $?" >> gpurun_out/r2c3_worker.log
grep -E "MULTIRANK|exit" gpurun_out/r2c3_worker.log | cut -c1-1500
timeout 900 python -m pytest tests/test_gpu_sharded_chain.py -q -m gpu > gpurun_out/r2c3_sharded_tests.log 2>&1
echo "sharded tests exit $?" >> gpurun_out/r2c3_sharded_tests.log
tail -5 gpurun_out/r2c3_sharded_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29643 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2c3_bench_2gpu.json 2>gpurun_out/r2c3_bench_2gpu.err
echo "bench exit $?"; tail -c 1200 gpurun_out/r2c3_bench_2gpu.json; tail -3 gpurun_out/r2c3_bench_2gpu.err
