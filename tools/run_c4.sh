# quick hang guard first
timeout 120 python -m pytest tests/test_gpu_rsgm.py -m gpu -x -q -k "compute_rsgm_vs_oracle or sweep_cluster_strips" 2>&1 | tail -5 > gpurun_out/r2_c4_quick.log; rc=$?
tail -3 gpurun_out/r2_c4_quick.log
if ! grep -q " passed" gpurun_out/r2_c4_quick.log || grep -q "failed\|error" gpurun_out/r2_c4_quick.log; then echo QUICK_FAILED; exit 1; fi
timeout 400 python -m pytest tests/test_gpu_rsgm.py tests/test_gpu_benchpath.py -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2_c4_pytest.log; tail -3 gpurun_out/r2_c4_pytest.log
for v in default red bytes; do
  export VPPB200_V_RED=0 VPPB200_BYTE_SUMS=0
  [ $v = red ] && export VPPB200_V_RED=1
  [ $v = bytes ] && export VPPB200_BYTE_SUMS=1
  timeout 120 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_c4_bench_$v.json 2> gpurun_out/r2_c4_bench_$v.err
  tail -c 300 gpurun_out/r2_c4_bench_$v.err
  python -c "
import json;d=json.load(open('gpurun_out/r2_c4_bench_$v.json'));print('$v',d['ms_per_step'],d['config']['stage_ms_per_step_serial'],d['parity_probe']['ok'])"
done
