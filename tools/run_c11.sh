export VPPB200_V_RED=1
for v in base r96 r80; do
  export VPPB200_LIB_SUFFIX=
  case $v in r96) export VPPB200_LIB_SUFFIX=_r96;; r80) export VPPB200_LIB_SUFFIX=_r80;; esac
  timeout 120 python bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/r2_c11_bench_$v.json 2> gpurun_out/r2_c11_bench_$v.err
  tail -c 300 gpurun_out/r2_c11_bench_$v.err
  python -c "
import json;d=json.load(open('gpurun_out/r2_c11_bench_$v.json'));print('$v',d['ms_per_step'],d['e2e']['value'],d['config']['stage_ms_per_step_serial'],d['parity_probe']['ok'])"
done
