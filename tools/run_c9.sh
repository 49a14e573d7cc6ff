timeout 300 python -m pytest tests/test_gpu_rsgm.py tests/test_gpu_benchpath.py -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2_c9_pytest.log
tail -4 gpurun_out/r2_c9_pytest.log
for v in plain red bytes; do
  export VPPB200_V_RED=0 VPPB200_BYTE_SUMS=0
  case $v in red) export VPPB200_V_RED=1;; bytes) export VPPB200_BYTE_SUMS=1;; esac
  timeout 120 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_c9_bench_$v.json 2> gpurun_out/r2_c9_bench_$v.err
  tail -c 300 gpurun_out/r2_c9_bench_$v.err
  python -c "
import json;d=json.load(open('gpurun_out/r2_c9_bench_$v.json'));print('$v',d['ms_per_step'],d['config']['stage_ms_per_step_serial'],d['parity_probe']['ok'])"
done
VPPB200_LIB_SUFFIX=_trace VPPB200_V_RED=1 timeout 120 python tools/vtrace.py
