# round 2, call 1: parity of the restructured tensor-core tails + activations, A/B against the round-1 library, one ncu capture
set -u
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2c1_smi.log 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tensor_core or activations or fwd_bwd_parity or hmc_step or train_visits or grouped or net_gradient or golden" > gpurun_out/r2c1_parity.log 2>&1
echo "parity exit $?" >> gpurun_out/r2c1_parity.log
timeout 600 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_wide.py -x -q -m gpu > gpurun_out/r2c1_full.log 2>&1
echo "fullsize exit $?" >> gpurun_out/r2c1_full.log
for v in r1 b200 r1 b200; do
  BANN_LIB_PATH=$GRAFT_REPO_ROOT/rs-bann_b200/libbann_$v.so timeout 300 python bench.py --workload cfg3s --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2c1_cfg3s_$v.json 2>gpurun_out/r2c1_cfg3s_$v.err
  python -c "import sys,json; d=json.loads(open('gpurun_out/r2c1_cfg3s_$v.json').read().strip().splitlines()[-1]); print('$v cfg3s k1_ms',d['k1_ms'],'value',d['value'],'frac',d['roofline']['frac'])" >> gpurun_out/r2c1_ab.log 2>&1
done
for v in r1 b200; do
  BANN_LIB_PATH=$GRAFT_REPO_ROOT/rs-bann_b200/libbann_$v.so timeout 300 python bench.py --workload cfg2 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2c1_cfg2_$v.json 2>gpurun_out/r2c1_cfg2_$v.err
  python -c "import sys,json; d=json.loads(open('gpurun_out/r2c1_cfg2_$v.json').read().strip().splitlines()[-1]); print('$v cfg2 k1_ms',d['k1_ms'],'value',d['value'],'frac',d['roofline']['frac'])" >> gpurun_out/r2c1_ab.log 2>&1
done
timeout 400 python bench.py --workload cfg3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2c1_cfg3.json 2>gpurun_out/r2c1_cfg3.err
python bench.py --workload cfg3s --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2c1_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k1_tc -s 4 -c 1 -o gpurun_out/r2c1_prof python bench.py --workload cfg3s --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2c1_ncu.log 2>&1
cat gpurun_out/r2c1_ab.log; tail -3 gpurun_out/r2c1_parity.log; tail -3 gpurun_out/r2c1_full.log; cat gpurun_out/r2c1_cfg3.json | cut -c1-400
