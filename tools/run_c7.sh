export VPPB200_LIB_SUFFIX=_vp4
timeout 200 python -m pytest tests/test_gpu_rsgm.py tests/test_gpu_benchpath.py -m gpu -x -q -k "compute_rsgm_vs_oracle or sweep_cluster_strips or wide_strips or random_shapes" 2>&1 | tail -5 > gpurun_out/r2_c7_quick.log
tail -3 gpurun_out/r2_c7_quick.log
for v in vp3 vp3red vp4red vp4; do
  export VPPB200_V_RED=0 VPPB200_LIB_SUFFIX=
  case $v in vp3red) export VPPB200_V_RED=1;; vp4red) export VPPB200_V_RED=1 VPPB200_LIB_SUFFIX=_vp4;; vp4) export VPPB200_LIB_SUFFIX=_vp4;; esac
  timeout 120 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_c7_bench_$v.json 2> gpurun_out/r2_c7_bench_$v.err
  tail -c 300 gpurun_out/r2_c7_bench_$v.err
  python -c "
import json;d=json.load(open('gpurun_out/r2_c7_bench_$v.json'));print('$v',d['ms_per_step'],d['config']['stage_ms_per_step_serial'],d['parity_probe']['ok'])"
done
