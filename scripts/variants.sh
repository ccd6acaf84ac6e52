for v in b200 v1 v3 v5 v7; do
  export BANN_LIB_PATH=/root/repo/rs-bann_b200/libbann_$v.so
  echo "== $v"
  timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tensor_core or hmc_step or train_visits or grouped" 2>&1 | tail -2
  timeout 200 python bench.py --workload cfg3s --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('k1_ms',d['k1_ms'],'value',d['value'],'frac',d['roofline']['frac'])"
done
