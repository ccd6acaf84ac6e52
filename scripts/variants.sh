# usage: bash scripts/variants.sh <workload> <lib suffix> ...   (kernel-variant A/B on the GPU box: parity subset + bench)
wl=$1; shift
for v in "$@"; do
  export BANN_LIB_PATH=/root/repo/rs-bann_b200/libbann_$v.so
  echo "== $v"
  timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tensor_core or hmc_step or train_visits or grouped" 2>&1 | tail -1
  timeout 200 python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('k1_ms',d['k1_ms'],'value',d['value'],'frac',d['roofline']['frac'])"
done
