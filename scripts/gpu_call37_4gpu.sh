# 4 GPUs with k1_tc5 as the default: bench line
set -u
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29643 bench.py --gpus 4 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_cfg3_4gpu_k1tc5.json 2>gpurun_out/r2c37_bench.err
echo "bench exit $?"; python -c "import json;d=json.loads(open('gpurun_out/r2_bench_cfg3_4gpu_k1tc5.json').read().strip().splitlines()[-1]);print(d['n_gpus'],d['value'],d['ms_per_step'],d['k1_ms'],d['e2e']['value'],d['sequential_schedule']['visits_per_s'])"
