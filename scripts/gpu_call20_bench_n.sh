set -u
N=$1
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29643 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r2_bench_cfg3_${N}gpu.json 2>gpurun_out/r2_bench_cfg3_${N}gpu.err
echo "bench exit $?"; python -c "import json;d=json.loads(open('gpurun_out/r2_bench_cfg3_${N}gpu.json').read().strip().splitlines()[-1]);print(d['n_gpus'],d['value'],d['ms_per_step'],d['k1_ms'],d['e2e']['value'],d['sequential_schedule'])"
