# round 2, call 6 (8 GPUs): parity gate on one GPU, bench at N = 8 / 4, the peer-memory checks with 8 ranks, configs[4] shortened
set -u
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tensor_core or grouped" > gpurun_out/r2c6_gate.log 2>&1 || { echo "parity gate failed"; tail -20 gpurun_out/r2c6_gate.log; exit 1; }
tail -1 gpurun_out/r2c6_gate.log
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for n in 8 4; do
  timeout 600 $TR --nproc-per-node $n --master-port $((29660+n)) bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/r2c6_bench_${n}gpu.json 2>gpurun_out/r2c6_bench_${n}gpu.err
  echo "bench $n exit $?"; python -c "
import json; d=json.loads(open('gpurun_out/r2c6_bench_${n}gpu.json').read().strip().splitlines()[-1]); print({k:d[k] for k in ('n_gpus','value','ms_per_step','k1_ms')}, 'e2e', d['e2e'])"
done
timeout 600 $TR --nproc-per-node 8 --master-port 29671 tests/multirank_worker.py > gpurun_out/r2c6_worker8.log 2>&1
echo "worker8 exit $?"; grep -E "MULTIRANK" gpurun_out/r2c6_worker8.log | cut -c1-900
for g in 64 0; do
  timeout 900 $TR --nproc-per-node 8 --master-port $((29680+g)) scripts/full_chain.py --iterations 50 --group-size $g > gpurun_out/r2c6_fullchain_g$g.log 2>&1
  echo "full_chain g=$g exit $?"; grep '"what"' gpurun_out/r2c6_fullchain_g$g.log | cut -c1-1300
done
timeout 400 $TR --nproc-per-node 8 --master-port 29690 scripts/full_chain.py --iterations 3 --group-size 1 > gpurun_out/r2c6_fullchain_g1.log 2>&1
echo "full_chain g=1 exit $?"; grep '"what"' gpurun_out/r2c6_fullchain_g1.log | cut -c1-1300
